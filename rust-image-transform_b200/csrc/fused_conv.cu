// fused_conv.cu -- the ring kernels of fused.cu once more, with a run-time destination channel count
// (DevJob::out_channels: DynamicImage::to_rgb8() / to_rgba8() fused into the store).  A separate
// translation unit so that the common case keeps its compile-time pixel stride and the two halves
// compile in parallel.
#define IKC_FUSED_CONV 1
#include "fused.cu"
