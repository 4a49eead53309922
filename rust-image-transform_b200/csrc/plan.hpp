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

// Band form, 8-bit (csrc/banded8.cu: the vertical pass as an integer matrix product, tcgen05.mma kind::i8, the
// source bytes used as they are).  Chunks of kBand8Chunk source indices (the K of one i8 MMA); chunk k touches at
// most kBand8Window consecutive outputs starting at output 8 * band8_gbase[k] (groups of 8; band8_gbase[n_chunks]
// = number of groups).  Every weight is the integer W = round(w * 2^band8_shift) (each output's weights are
// nudged to sum to exactly 2^band8_shift), split into band8_limbs signed 8-bit digits, base 256, low digits in
// [-128, 127].  band8_tiles: per chunk one K-major s8 operand tile of (band8_limbs * 32) rows x 32 indices in the
// shared-memory layout the MMA reads; row = (output - 8 * band8_gbase[k]) * band8_limbs + digit, most significant
// digit first.
// band8_limbs == 0: not applicable (upscale, a chunk window wider than 32 outputs, or too many taps).
struct Band8 {
    int limbs = 0, shift = 0;
    std::vector<int32_t> gbase;   // [n_chunks + 1]
    std::vector<int8_t> tiles;    // [n_chunks][limbs * 32 * 32]
};
constexpr int kBand8Chunk = 32, kBand8Group = 8, kBand8Window = 32;
constexpr int kBand8TRowsMax = 128;
constexpr int kBand8Base = 256;   // digit base: W = hi * 256 + lo, lo in [-128, 127], |hi| <= 127

// Row-band form of the same integer weights, for the kernel whose accumulator lanes are OUTPUT rows (banded8t.cu): the
// weights are the A operand.  Band r = outputs [rows r, rows r + rows), rows <= 128; its chunks (of kBand8Chunk source indices) start at
// the band's first source index k_lo[r]: chunk c = indices k_lo[r] + 32 c .. + 31.  tiles: per band, chunk and digit (most significant first) one K-major s8 operand
// tile of 128 rows x 32 indices in the shared-memory layout the MMA reads (8 x 16-byte core matrices: rows 128 bytes
// apart, the two halves of the 32 indices 2048 bytes apart).  chunks == 0: not applicable (a band needs more than
// kBand8TMaxChunks chunks, i.e. the ratio is well above 2).
struct Band8T {
    int chunks = 0;
    int rows = kBand8TRowsMax;   // outputs per band (<= 128 = the MMA's M; rows of a tile beyond it are zero): chosen to minimise
                                  // bands x chunks, e.g. 120 at exactly 2:1 (8 chunks of 32 source indices instead of 9)
    std::vector<int32_t> k_lo;    // [n_bands] first source index of each band
    std::vector<int8_t> tiles;    // [n_bands][chunks][2][128 * 32]
};
constexpr int kBand8TRows = 128, kBand8TMaxChunks = 10;


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
    // Band form (downscales; csrc/banded.cu runs the vertical pass as a banded matrix product on the
    // tensor cores).  The source indices are cut into chunks of kBandChunk (the K of one f16 MMA); chunk k
    // only touches the outputs of band_n / 16 consecutive groups of 16, starting at group band_gbase[k]
    // (band_gbase[n_chunks] = number of groups; groups below band_gbase[k + 1] are final after chunk k).
    // band_tiles holds, per chunk, two K-major f16 operand tiles [band_n outputs x 16 indices] in the
    // shared-memory layout the MMA reads (no swizzle: 8x8 core matrices, index-group stride band_n * 16 B,
    // output-group stride 128 B): the weights * 2^14 rounded to f16 ("hi"), then the f16 of what that
    // rounding lost ("lo"); hi + lo carries 22 bits of every weight.  band_n == 0: not applicable.
    int band_n = 0;
    std::vector<int32_t> band_gbase;    // [n_chunks + 1]
    std::vector<uint16_t> band_tiles;   // [n_chunks][2][band_n * 16] f16 bit patterns
    Band8 band8;                        // the 8-bit band form (see above)
    Band8T band8t;                      // its row-band form (shares band8.shift; two digits)
    // Exact 2:1 pass whose every window lies inside the 12 source indices [2 o - 5, 2 o + 7) and whose uniform stretch has
    // exactly those 12 taps: what banded8t.cu's register-resident horizontal filter is written for.
    bool h2_12 = false;
};

constexpr int kBandChunk = 16;          // source indices per chunk (K of tcgen05.mma kind::f16)
constexpr int kBandGroup = 16;          // outputs per group
constexpr int kBandMaxN = 48;           // widest chunk window the kernel's accumulator ring takes
constexpr float kBandScaleW = 16384.0f; // 2^14: weights are stored as f16 of w * 2^14
// A byte dropped into an f16 word is the denormal b * 2^-24, so the f32 intermediate of the band kernel is the
// reference's times 2^-10; its horizontal weights carry the compensating 2^10 (exact).
constexpr float kBandScaleH = 1024.0f;

uint16_t f32_to_f16_rn(float f);        // IEEE binary16, round to nearest even (host)
float f16_to_f32(uint16_t h);

// vertical_forms: also build the tensor-core operand tiles (band_*, band8, band8t), which only the vertical use of a pass
// needs (they are most of a pass's bytes).
std::shared_ptr<const PassPlan> build_pass(int filter, uint32_t n_in, uint32_t n_out, bool vertical_forms = true);
constexpr int kBand8DefaultLimbs = 2;

float filter_support(int filter);

}  // namespace ikc
