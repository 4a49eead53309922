// device_types.hpp -- plain structs shared by the host planner and the CUDA kernels.
#pragma once
#include <cstdint>

#include <vector_types.h>

namespace ikc {

// The fused ring kernel turns a byte into a float by dropping it into an all-zero word: the denormal
// b * 2^-149 (one PRMT, no arithmetic; FFMA2 takes denormal operands at full rate on sm_100a).  The ring
// weights carry the compensating powers of two, which is exact: vertical weights * 2^75, so the f32
// intermediate is the reference's times 2^-74, and horizontal weights * 2^74.
constexpr float kRingScaleV = 37778931862957161709568.0f;   // 2^75
constexpr float kRingScaleH = 18889465931478580854784.0f;   // 2^74

// Device-resident tables of one separable pass (see plan.hpp for the host form).
struct DevPass {
    const int32_t* left;   // [n_out]
    const int32_t* right;  // [n_out]
    const float* w;        // [n_out * stride]
    const float* ring_v;   // [n_in * ring_stride * 2] duplicated ring weights * kRingScaleV (pass used vertically), or nullptr
    const float* ring_h;   // the same * kRingScaleH (pass used horizontally)
    int32_t stride;
    int32_t ring_k;        // windows covering any source index (ring size)
    int32_t ring_stride;   // ring_k rounded up to even (row stride of `ring`, in weight pairs)
    int32_t n_in, n_out;
    int32_t max_count;
    int32_t uni_step;      // outputs [uni_lo, uni_hi) end uni_step source indices after their predecessor
    int32_t uni_lo, uni_hi;
    const float2* up2_pairs_v;  // exact 2x upscale: [n_in][up2_taps] (weight for output 2k, for output 2k+1) * kRingScaleV,
    const float2* up2_pairs_h;  // and * kRingScaleH (up2.cu feeds bytes as denormals too); nullptr if not a 2x upscale
    int32_t up2_off, up2_taps;
    int32_t up2_uni_lo, up2_uni_hi;  // source indices with bit-identical pairs
    // Band form (plan.hpp: PassPlan::band_*), used when the pass runs vertically on the tensor cores (banded.cu).
    const uint16_t* band_tiles;      // [n_chunks][2][band_n * 16] f16, shared-memory operand layout; nullptr if none
    const int32_t* band_gbase;       // [n_chunks + 1]
    int32_t band_n;                  // output rows per operand tile (32 or 48); 0 if none
    // 8-bit band form (plan.hpp: Band8), used when the pass runs vertically as an integer product (banded8.cu).
    const int8_t* band8_tiles;       // [n_chunks][limbs * 32 x 32] s8, shared-memory operand layout; nullptr if none
    const int32_t* band8_gbase;      // [n_chunks + 1], groups of 8 outputs
    int32_t band8_limbs;             // base-256 digits per weight (2); 0 if none
    int32_t band8_shift;             // weights are round(w * 2^shift)
    // Row-band form of the same digits (plan.hpp: Band8T), used by the kernel whose accumulator lanes are output rows
    // (banded8t.cu).
    const int8_t* band8t_tiles;      // [n_bands][band8t_chunks][2][128 x 32] s8, operand layout; nullptr if none
    const int32_t* band8t_klo;       // [n_bands] first source index of each band
    int32_t band8t_chunks;           // operand tiles per band and digit (<= 10); 0 if none
    int32_t band8t_rows;             // outputs per band (<= 128)
};

// One image resize, device pointers.
struct DevJob {
    const uint8_t* src;
    uint8_t* dst;
    float* tmp;            // generic path only: f32 intermediate [dh][sw*channels]
    uint64_t src_pitch;    // bytes
    uint64_t dst_pitch;    // bytes
    uint32_t sw, sh, dw, dh;
    int32_t channels;
    int32_t out_channels;  // interleaved samples per destination pixel: == channels, or 3 / 4 when the store
                           // applies DynamicImage::to_rgb8() / to_rgba8() (8-bit only)
    int32_t bps;           // bytes per sample: 1 or 2
    DevPass v, h;
    // Banded kernel only: a 2-D TMA tensor map (CUtensorMap, 128 bytes) of the source raster seen as
    // [sh rows][src_pitch / 4 words]; box = 16 rows x 512 bytes, out-of-bounds words read as zero.
    alignas(64) uint8_t src_map[128];
    // banded8 kernel: the same raster as [sh rows][src_pitch bytes] of u8, box = 32 rows x 128 bytes, 128-byte swizzle
    // (the box lands in shared memory as the MN-major operand tile the integer MMA reads).
    alignas(64) uint8_t src_map8[128];
};

// One CTA's share of a job in the fused kernel: output columns [ox0,ox1) x rows [oy0,oy1).
struct WorkItem {
    int32_t job;
    int32_t ox0, ox1;
    int32_t oy0, oy1;
};

// Launch-wide shared-memory geometry of the fused kernel (max over the batch's items).
struct FusedGeom {
    int32_t tmp_px;        // tmp row capacity in pixels (float4 each); odd
    int32_t n_items;
};

// Launch-wide shared-memory geometry of the banded (tensor-core) kernel (max over the batch's items).
struct BandGeom {
    int32_t band_n;        // output rows per weight tile of the vertical pass (all jobs of a launch share it)
    int32_t max_out;       // outputs of the widest strip (rows of the staged left/right + weight table)
    int32_t hw_pairs;      // staged horizontal weights (one duplicated pair per output and tap) of the largest strip
    int32_t n_items;
};

// Launch-wide geometry of the banded8 (integer tensor-core) kernel.
struct Band8Geom {
    int32_t limbs;         // digits per weight: all jobs of a launch share it
    int32_t max_out;
    int32_t hw_pairs;
    int32_t n_items;
};

// Launch-wide geometry of the banded8t (row-band integer tensor-core) kernel.
struct Band8TGeom {
    int32_t chunks;        // weight tiles per band and digit: max over the launch's jobs (sizes shared memory)
    int32_t n_items;
};

// Launch-wide shared-memory geometry of the tile kernel (max over the batch's tiles).
struct TileGeom {
    int32_t pitch_f;        // floats per staged row: footprint columns * channels, rounded up to 4
    int32_t max_src_rows;   // source rows of the largest tile footprint
    int32_t max_tile_rows;  // output rows of the tallest tile
    int32_t max_tile_cols;  // output columns of the widest tile
    int32_t vstride;        // taps per output row in the staged vertical weights
    int32_t hstride;        // taps per output column in the staged horizontal weights
    int32_t out_pitch_b;    // bytes per row of the staged output tile (multiple of 4, + one spare word)
    int32_t n_items;
};

}  // namespace ikc
