// banded8_conv.cu -- the banded8 kernel's converting build: the same kernels with a run-time destination channel
// count, i.e. DynamicImage::to_rgb8() / to_rgba8() (/root/reference/src/transform.rs:123,131,140) fused into the store.
#define IKC_BANDED8_CONV 1
#include "banded8.cu"
