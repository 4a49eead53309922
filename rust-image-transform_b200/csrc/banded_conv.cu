// banded_conv.cu -- the banded kernel's converting build: the same kernels with a run-time destination channel
// count, i.e. DynamicImage::to_rgb8() / to_rgba8() (/root/reference/src/transform.rs:123,131,140) fused into the store.
#define IKC_BANDED_CONV 1
#include "banded.cu"
