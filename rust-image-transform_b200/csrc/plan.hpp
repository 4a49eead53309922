// plan.hpp -- host-side planning for the resize hot path: the dims rule and the per-pass
// window/weight tables.  Product code (compiled into libimagekit_cuda.so).
//
// Follows /root/reference/src/transform.rs:62-90 (target size in f32) and the published
// algorithm of crate image 0.25.8 (math/utils.rs resize_dimensions; imageops/sample.rs
// window + weight generation and filter kernels).  Weights are built on the host with glibc
// sinf/expf -- the same libm calls Rust's f32::sin/f32::exp lower to on linux-gnu -- so the
// tables are bit-identical to what the reference computes; the GPU only consumes them.
#pragma once
#include <cstdint>
#include <memory>
#include <vector>

namespace ikc {

enum Filter : int { kNearest = 0, kTriangle = 1, kCatmullRom = 2, kGaussian = 3, kLanczos3 = 4 };

// a1-a3: returns an ikc_dims_code (0 resample, 1 passthrough, 2 clone, 3 copy).
int target_dims(uint32_t ow, uint32_t oh, bool has_w, uint32_t w, bool has_h, uint32_t h,
                uint32_t* tw, uint32_t* th);

// One separable pass n_in -> n_out with a given filter.
struct PassPlan {
    int filter = 0;
    uint32_t n_in = 0, n_out = 0;
    uint32_t stride = 0;                // taps slots per output in `w` (>= max count)
    std::vector<int32_t> left;          // [n_out] first source index of the window
    std::vector<int32_t> count;         // [n_out] taps in the window (right - left)
    std::vector<float> w;               // [n_out * stride] normalised weights, zero padded
    // Ring form, used by the fused kernels.  ring_k = max number of windows that contain any
    // one source index.  ring[(y * ring_stride + j) * 2 + {0,1}] = weight of source index y for the
    // output o (o mod ring_k == j) whose window contains y, else 0; each weight is stored twice
    // so that one 64-bit load feeds a packed fma.rn.f32x2.
    int ring_k = 0;
    int ring_stride = 0;                // ring_k rounded up to even
    std::vector<float> ring_v, ring_h;  // [n_in * ring_stride * 2] each, scaled by kRingScaleV / kRingScaleH
                                        // (device_types.hpp: the kernel feeds bytes as denormals)
    std::vector<int32_t> right;         // [n_out] left + count
    uint32_t max_count = 0;
    // Uniform stretch: outputs o in [uni_lo, uni_hi) all end exactly uni_step source indices after their
    // predecessor (right[o] - right[o-1] == uni_step), as every interior output of an integer-ratio
    // downscale does.  uni_step == 0: no stretch long enough to be worth a specialised loop.
    int uni_step = 0, uni_lo = 0, uni_hi = 0;
    // Exact 2x upscale (n_out == 2 * n_in): every window fits a frame of up2_taps source indices that
    // starts at (o >> 1) + up2_off.  up2_pairs[(k * up2_taps + t) * 2 + p] = weight of source index
    // k + up2_off + t for output 2k + p (0 outside its window): the two outputs of a source index
    // share their taps, so one packed FMA feeds both (csrc/up2.cu).  up2_taps == 0: not applicable.
    int up2_off = 0, up2_taps = 0;
    std::vector<float> up2_pairs;       // [n_in * up2_taps * 2]
    int up2_uni_lo = 0, up2_uni_hi = 0; // source indices [lo, hi) whose pairs are bit-identical (the interior)
};

std::shared_ptr<const PassPlan> build_pass(int filter, uint32_t n_in, uint32_t n_out);

float filter_support(int filter);

}  // namespace ikc
