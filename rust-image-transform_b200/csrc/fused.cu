// fused.cu -- fused single-launch "ring" resize kernel for 8-bit downscales (sm_100a).
//
// This is the product's main path for the reference's actual workload: Lanczos3 downscales of
// 8-bit rasters -- Rgb8/Rgba8, and Luma8/LumaA8 on the general loops -- (resize_image, /root/reference/src/transform.rs:62-90, whose arithmetic is
// image 0.25.8 imageops::resize = vertical_sample then horizontal_sample).
//
// One CTA owns an output tile (column strip x row chunk) of one image of the batch.  It streams
// the source rows it needs through a TMA-fed (cp.async.bulk -> UBLKCP) shared-memory ring exactly
// once, keeps the unclamped f32 intermediate of the vertical pass in shared memory (never in HBM),
// and writes finished u8 rows.  Both passes are input-stationary "ring" filters: every source
// sample is converted once and FMA'd into the <= K outputs whose windows contain it, using packed
// fma.rn.f32x2 (FFMA2) with host-built weights.  Pass order (vertical first), f32 unclamped
// intermediate, clamp + round-half-away at the very end are the reference's; only the summation
// uses FMA and a different association, hence max |delta| <= 1 LSB instead of bit equality.
#include <cuda_runtime.h>

#include <cstdint>

#include "device_types.hpp"
#include "launch.hpp"

// This file is compiled twice: as itself (IKC_FUSED_CONV=0) -> the kernels that store the source's own
// channel count (compile-time pixel stride) plus the host-side helpers and the dispatcher; and through
// fused_conv.cu (IKC_FUSED_CONV=1) -> the same kernels with a run-time destination channel count
// (to_rgb8()/to_rgba8() fused into the store).
#ifndef IKC_FUSED_CONV
#define IKC_FUSED_CONV 0
#endif

namespace ikc {
namespace {

constexpr bool kConv = IKC_FUSED_CONV != 0;

// A CTA is four compute warps plus one producer warp that keeps the source ring full.  A compute
// thread owns two 32-bit source words of every staged row.
constexpr int kComputeWarps = 4;
constexpr int kComputeThreads = 32 * kComputeWarps;
constexpr int kThreads = kComputeThreads + 32;
constexpr int kMaxSegs = 2 * kComputeWarps;      // horizontal segments: one per half warp
constexpr int kStageRows = 4;                    // source rows per ring stage (one mbarrier pair)
constexpr int kStages = 4;
constexpr int kRingRows = kStageRows * kStages;
constexpr int kSrcRowBytes = 1024;               // staged bytes per source row
constexpr int kHalfRowBytes = kSrcRowBytes / 2;  // offset of a thread's second word
constexpr int kTmpRows = 16;                     // f32 intermediate rows per group
constexpr int kHeaderBytes = 512;                // mbarriers, then the uniform stretches' tap weights
constexpr int kTapOffsetV = 128, kTapOffsetH = 320;  // byte offsets of the (duplicated) tap weights, <= 24 pairs each
constexpr int kMaxStages = 8;
constexpr int kMaxStripOut = 272;                // outputs of one strip (256) + ring pre-roll, in the left/right table
constexpr size_t kFusedMaxSmem = 113 * 1024;     // the planner's budget (context.cpp): two CTAs per SM

constexpr int max_src_bytes(int channels) {
    return channels == 4 ? 1024 : channels == 3 ? 864 : channels == 2 ? 512 : 256;
}

__device__ __forceinline__ uint32_t smem_addr(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}
// The producer's wait: try_wait with a suspend-time hint parks the thread in hardware until the phase
// completes (or the hint, in ns, runs out), so its polling does not take issue slots from the compute warps.
__device__ __forceinline__ void mbar_wait_parked(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity), "r"(20000u)
        : "memory");
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_addr(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar))
        : "memory");
}
// Barrier over the compute warps only (the producer warp never joins it).
__device__ __forceinline__ void compute_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kComputeThreads) : "memory"); }

// u8 -> f32 with one PRMT and no arithmetic: the byte alone in a zero word is the denormal b * 2^-149;
// the ring weights carry the compensating power of two (device_types.hpp: kRingScaleV / kRingScaleH).
template <int BYTE>
__device__ __forceinline__ float byte_to_float(uint32_t word) {
    return __uint_as_float(__byte_perm(word, 0u, 0x4440u + BYTE));
}

// Quantise one finished pixel: clamp to [0,255] and round half away from zero (f32::round), as
// trunc(v + 0.5) with saturation.  The + 0.5 is already in `v`: every horizontal accumulator starts at
// 0.5 instead of 0.  trunc(v + 0.5) equals round(v) for every non-negative float except
// v = 0.5 - 2^-25; this path's sums differ from the reference's by FMA/association rounding anyway, and
// the EXACT path (generic.cu) has no such case.  float -> s32 (rz) followed by the saturating u8 pack
// compiles to two F2IP.U8.F32.TRUNC for the whole pixel.
__device__ __forceinline__ uint32_t pack_pixel(float4 v_plus_half) {
    const int r = __float2int_rz(v_plus_half.x), g = __float2int_rz(v_plus_half.y);
    const int b = __float2int_rz(v_plus_half.z), a = __float2int_rz(v_plus_half.w);
    uint32_t hi, px;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(a), "r"(b), "r"(0));   // bytes: b, a
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(px) : "r"(g), "r"(r), "r"(hi));  // bytes: r, g, b, a
    return px;
}
constexpr float kRoundBias = 0.5f;  // initial value of every horizontal accumulator

// Store one finished pixel straight to the destination raster.  Lanes of a half warp hold 16 different
// rows of the same column, so these are scattered 4-byte (or 1-byte) stores; the sectors are completed
// in L2 by the same thread's next pixels before they reach HBM.
// `co` = destination channels: C, or the other of {3, 4} when the store applies DynamicImage::to_rgb8()
// (drop alpha) / to_rgba8() (alpha = 255) -- src/transform.rs:123,131,140 does that on the CPU before
// encoding.  4-channel destinations are word aligned: the planner only sends those here.
template <int C>
__device__ __forceinline__ void store_pixel(uint8_t* dst_px, float4 v_plus_half, int co) {
    uint32_t w = pack_pixel(v_plus_half);  // bytes: the pixel's C channels, then don't-care lanes
    if (C <= 2 && co >= 3) {               // grey (+ alpha) -> r, g, b (, a): the grey replicated
        const uint32_t grey = w & 0xffu;
        const uint32_t alpha = C == 2 ? (w >> 8) & 0xffu : 0xffu;
        w = grey * 0x010101u | (alpha << 24);
    } else if (C == 3) {
        w |= 0xff000000u;                  // rgb -> rgba: opaque
    }
    if (co == 4) {
        *reinterpret_cast<uint32_t*>(dst_px) = w;
    } else {
        dst_px[0] = uint8_t(w);
        if (co >= 2) dst_px[1] = uint8_t(w >> 8);
        if (co >= 3) dst_px[2] = uint8_t(w >> 16);
    }
}

}  // namespace

// Shared memory: [mbarriers][source ring: 16 rows x 1024 B][vertical weight ring: 16 rows x KSV
//                pairs][horizontal weights of the strip: tmp_px x KSH pairs][(left,right) of the
//                strip's outputs][tmp: 16 rows x tmp_px x float4].  Weights travel with TMA bulk
//                copies too, so the hot loops only read shared memory.
//
// Vertical phase: compute thread t owns source byte columns [4t,4t+4) and [512+4t,512+4t+4) of the
// strip and marches down the source rows; its KV ring slots hold the partial sums of the <= KV output
// rows currently open (slot = output row mod KV; the slot loop is unrolled, so every accumulator has a
// fixed register).  A finished row goes to tmp as one float4 per pixel (fewer than 4 channels are
// padded to 4 lanes).  Every 16 finished rows the horizontal phase runs: lane & 15 = tmp row, half
// warp = x segment of the strip; the same ring march along x, started one window early so that each
// segment is self-contained.  Finished pixels are quantised and stored straight to HBM.
//
// SV / SH > 0 add loops specialised for a uniform stretch of the pass (PassPlan::uni_step == SV: every
// output ends exactly SV source rows after its predecessor, the interior of an integer-ratio resize):
// a whole ring revolution (K outputs, K*S source rows) is straight-line code with no window lookups.
// The general loops still handle the image borders, the ring pre-roll and every non-uniform pass.
//
// The ring is addressed by "virtual rows": chunk row r lives at virtual row rr = r + p0, ring row
// rr % 16, stage (rr / 4) % 4.  p0 in [0,4) shifts the stage boundaries so that the first uniform
// revolution starts on one; the first stage then simply holds 4 - p0 rows.
template <int C, int KV, int KH, int SV, int SH, bool CONV>
__global__ void __launch_bounds__(kThreads, 2)
fused_ring_kernel(const DevJob* __restrict__ jobs, const WorkItem* __restrict__ items, const FusedGeom geom) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int KSV = (KV + 1) & ~1;  // ring row stride (weight pairs), even
    constexpr int KSH = (KH + 1) & ~1;
    // Uniform stretches whose whole window (K * S taps) fits a few registers run from tap weights held
    // in registers instead of the per-row ring weights in shared memory.
    constexpr int LV = KV * SV, LH = KH * SH;
    constexpr bool RWV = SV > 0 && LV <= 12, RWH = SH > 0 && LH <= 12;

    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty_bar = full_bar + kMaxStages;   // one arrival per compute warp that has drained the stage
    uint64_t* hw_bar = empty_bar + kMaxStages;
    uint8_t* src_ring = smem + kHeaderBytes;
    float4* vw_ring = reinterpret_cast<float4*>(src_ring + kRingRows * kSrcRowBytes);  // [row][KSV/2]
    float4* hw_smem = vw_ring + kRingRows * (KSV / 2);                                 // [px][KSH/2]
    int2* hlr = reinterpret_cast<int2*>(hw_smem + size_t(geom.tmp_px) * (KSH / 2));    // (left, right)
    float4* tmp = reinterpret_cast<float4*>(hlr + kMaxStripOut);
    float2* const vtap = reinterpret_cast<float2*>(smem + kTapOffsetV);  // [LV] (w, w)
    float2* const htap = reinterpret_cast<float2*>(smem + kTapOffsetH);  // [LH]

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;

    const WorkItem it = items[blockIdx.x];
    const DevJob* __restrict__ J = jobs + it.job;
    const int ox0 = it.ox0, ox1 = it.ox1, oy0 = it.oy0, oy1 = it.oy1;

    const int32_t* __restrict__ vleft = J->v.left;
    const int32_t* __restrict__ vright = J->v.right;
    const int32_t* __restrict__ hleft = J->h.left;
    const int32_t* __restrict__ hright = J->h.right;
    const float4* __restrict__ vring = reinterpret_cast<const float4*>(J->v.ring_v);
    const float4* __restrict__ hring = reinterpret_cast<const float4*>(J->h.ring_h);

    // Strip geometry along x: source pixels [xl, xr), source bytes [b0, b0 + nb) (16-byte aligned).
    const int xl = __ldg(hleft + ox0);
    const int xr = __ldg(hright + ox1 - 1);
    const int row_bytes = int(J->sw) * C;
    const int b0 = (xl * C) & ~15;
    const int b1 = min((xr * C + 15) & ~15, (row_bytes + 15) & ~15);
    const int nb = b1 - b0;
    const int pxb = b0 / C;  // first (possibly partial) pixel held in tmp column 0
    // Chunk geometry along y: source rows [y_first, y_last).
    const int y_first = __ldg(vleft + oy0);
    const int y_last = __ldg(vright + oy1 - 1);
    const int nrows = y_last - y_first;

    // Uniform vertical stretch of this chunk: revolutions that start at a multiple of KV inside
    // [v_fast_lo, v_fast_hi) run the specialised loop.  p0 puts the first of them on a stage boundary.
    int v_fast_lo = 0, v_fast_hi = 0, p0 = 0;
    if (SV > 0 && J->v.uni_step == SV) {
        const int lo = (max(max(oy0, J->v.uni_lo), 1) + KV - 1) / KV * KV;
        // With tap registers the outputs that enter the ring during a revolution (the next KV) must belong
        // to the stretch too, unless they lie below the chunk and are never emitted.
        const int hi = (!RWV || oy1 <= J->v.uni_hi) ? min(oy1, J->v.uni_hi) : J->v.uni_hi - KV;
        if (lo + KV <= hi) {
            v_fast_lo = lo;
            v_fast_hi = hi;
            const int before = max(0, __ldg(vright + lo - 1) - y_first);  // chunk rows consumed before it
            p0 = (4 - (before & 3)) & 3;
        }
    }

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_bar + s, 1);
            mbar_init(empty_bar + s, kComputeWarps);
        }
        mbar_init(hw_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    // tap weights of the uniform stretches (every output of a stretch has the same ones)
    if (RWV && tid < LV && J->v.uni_step == SV) {
        const float w = __ldg(J->v.w + size_t(J->v.uni_lo) * J->v.stride + tid) * kRingScaleV;
        vtap[tid] = make_float2(w, w);
    }
    if (RWH && tid >= 32 && tid < 32 + LH && J->h.uni_step == SH) {
        const float w = __ldg(J->h.w + size_t(J->h.uni_lo) * J->h.stride + (tid - 32)) * kRingScaleH;
        htap[tid - 32] = make_float2(w, w);
    }
    // (left, right) of the strip's outputs (plus the few before ox0 whose windows reach into the strip)
    const int o_lo = max(0, ox0 - KH + 1);
    for (int i = tid; i < ox1 - o_lo; i += kThreads) hlr[i] = make_int2(__ldg(hleft + o_lo + i), __ldg(hright + o_lo + i));
    __syncthreads();

    // ---------------------------------------------------------------- producer warp
    if (warp == kComputeWarps) {
        if (lane == 0) {
            const uint8_t* const gsrc = J->src + size_t(y_first) * J->src_pitch + b0;
            const size_t src_pitch = J->src_pitch;
            const int v_end = p0 + nrows;  // virtual rows [p0, v_end) exist
            auto issue_fill = [&](int f) {   // virtual rows [4f, 4f+4) -> stage f % 4
                const int stage = f % kStages;
                const int v0 = max(f * kStageRows, p0), v1 = min((f + 1) * kStageRows, v_end);
                const int n = v1 - v0;
                const uint32_t wbytes = uint32_t(n) * (KSV * 8);
                mbar_expect_tx(full_bar + stage, uint32_t(n) * uint32_t(nb) + wbytes);
                for (int i = 0; i < n; ++i)
                    bulk_load(src_ring + ((v0 + i) % kRingRows) * kSrcRowBytes, gsrc + size_t(v0 + i - p0) * src_pitch,
                              uint32_t(nb), full_bar + stage);
                bulk_load(vw_ring + (v0 % kRingRows) * (KSV / 2), vring + size_t(y_first + v0 - p0) * (KSV / 2), wbytes,
                          full_bar + stage);
            };
            const int n_fill = (v_end + kStageRows - 1) / kStageRows;
            int f = 0;
            for (; f < min(n_fill, kStages); ++f) issue_fill(f);
            // horizontal ring weights of the strip's source pixels [xl, xr): one bulk copy, used by every group
            const uint32_t hbytes = uint32_t(xr - xl) * (KSH * 8);
            mbar_expect_tx(hw_bar, hbytes);
            bulk_load(hw_smem, hring + size_t(xl) * (KSH / 2), hbytes, hw_bar);
            for (; f < n_fill; ++f) {
                mbar_wait_parked(empty_bar + f % kStages, uint32_t(f / kStages - 1) & 1);  // every compute warp has drained it
                issue_fill(f);
            }
        }
        return;
    }

    const int2* const lr_tab = hlr - o_lo;  // indexed by absolute output column

    // ---------------------------------------------------------------- horizontal segmentation
    // The strip's outputs [ox0, ox1) are cut into runs of `per` outputs (a multiple of KH, so every
    // segment starts on the same ring slot and the half warps of a warp walk the unrolled slot code in
    // lock step).  Segment s computes outputs [os, oe) from the source pixels [left[os], right[oe-1]):
    // neighbouring segments overlap by one window, which costs a few extra FMAs per row but leaves
    // every output to exactly one thread (no partial sums to exchange between segments).
    const int n_out = ox1 - ox0;
    const int per = ((n_out + kMaxSegs - 1) / kMaxSegs + KH - 1) / KH * KH;
    const int hrow = lane & 15;
    const int sidx = 2 * warp + (lane >> 4);
    const int os = ox0 + sidx * per;
    const bool h_active = os < ox1;
    const int oe = min(os + per, ox1);
    const int seg_lo = h_active ? lr_tab[os].x : xr;
    const int h_slot0 = ox0 % KH;  // ring slot of every segment's first output
    const float4* const my_row = tmp + size_t(hrow) * geom.tmp_px - pxb;  // indexed by absolute source pixel
    uint8_t* const dst_base = J->dst;
    const size_t dst_pitch = J->dst_pitch;
    const int CO = CONV ? J->out_channels : C;  // destination channels (3 or 4)
    bool hw_ready = false;
    // [h_fast_lo, h_fast_hi): whole revolutions of this segment inside the pass's uniform stretch.
    int h_fast_lo = 0, h_fast_hi = 0, h_pre = 0;
    if (SH > 0 && h_active && J->h.uni_step == SH) {
        const int lo = (max(max(os, J->h.uni_lo), 1) + KH - 1) / KH * KH;
        const int hi = (!RWH || oe <= J->h.uni_hi) ? min(oe, J->h.uni_hi) : J->h.uni_hi - KH;  // as for the rows
        if (lo + KH <= hi) {
            h_fast_lo = lo;
            h_fast_hi = lo + (hi - lo) / KH * KH;
            // The pre-roll revolution is uniform too when the walk starts SH pixels before the first
            // window (so that the first pre-roll output also takes SH pixels): those pixels only reach
            // slots of outputs that are never stored.
            if (lo == os && h_slot0 == 0 && os - KH >= J->h.uni_lo) {
                h_fast_lo = os - KH;
                h_pre = SH;
            }
        }
    }

    // ---------------------------------------------------------------- vertical state
    float2 vacc[KV][4];  // [slot][byte pair]: bytes 0-1, 2-3 of the first word, then of the second
#pragma unroll
    for (int j = 0; j < KV; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) vacc[j][q] = make_float2(0.0f, 0.0f);
    const uint8_t* const my_src = src_ring + 4 * tid;

    // Where this thread's eight vertical results land in a tmp row (float index within the row).
    int emit_off[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int in_row = (i >> 2) * kHalfRowBytes + 4 * tid;  // this word's offset in the staged row
        const int byte = b0 + in_row + (i & 3);
        // words beyond the strip's staged bytes go to the spare last column of the tmp row
        emit_off[i] = in_row < nb ? (byte / C - pxb) * 4 + (byte % C) : (geom.tmp_px - 1) * 4 + (i & 3);
    }

    // `right` of 32 consecutive outputs lives one per lane and is broadcast with a shuffle one output
    // ahead of its use, so no global load sits on the row loop's critical path.
    int vr_base = max(0, oy0 - KV);
    auto vr_load = [&](int base) {
        const int o = base + lane;
        return (o < oy1) ? __ldg(vright + o) : 0x7fffffff;
    };
    int vr_val = vr_load(vr_base);
    auto v_end_of = [&](int o) {  // right[o] for o >= vr_base (uniform)
        while (o - vr_base >= 32) {
            vr_base += 32;
            vr_val = vr_load(vr_base);
        }
        return __shfl_sync(0xffffffffu, vr_val, o - vr_base);
    };
    // Ring order starts at the first output whose window reaches row y_first: outputs above the chunk
    // that are still open there hold ring slots until they close (their sums are never emitted).
    int ov = vr_base;
    int yend_next = v_end_of(ov);
    while (yend_next <= y_first) yend_next = v_end_of(++ov);
    int c_start = ov % KV;  // slot of output ov is ov mod KV; the unrolled slot loop is entered here

    // All ring bookkeeping derives from rr = virtual rows consumed so far: ring row rr % 16, stage
    // (rr / 4) % 4, fill parity (rr / 16) & 1.
    uint32_t rr = uint32_t(p0);
    const int y_virt0 = y_first - p0;  // source row of virtual row 0
    int g0 = oy0;           // first output row of the group being assembled in tmp
    int emitted = 0;        // rows of that group already in tmp
    if (p0 != 0) mbar_wait(full_bar + 0, 0);  // the first stage is entered in its middle

    // (source-pair major: consecutive FFMA2s share their source operand, which the register operand reuse
    //  cache serves -- measured 127 vs 108 FMA lanes/clk/SM against the slot-major order)
    // fresh >= 0: slot `fresh` starts a new output with this row (its previous output was emitted just
    // before), so its old contents are ignored instead of being zeroed separately.
    auto fma_row = [&](uint32_t d0, uint32_t d1, const float4 (&w)[KSV / 2], int fresh) {
        float2 s[4];
        s[0] = make_float2(byte_to_float<0>(d0), byte_to_float<1>(d0));
        s[1] = make_float2(byte_to_float<2>(d0), byte_to_float<3>(d0));
        s[2] = make_float2(byte_to_float<0>(d1), byte_to_float<1>(d1));
        s[3] = make_float2(byte_to_float<2>(d1), byte_to_float<3>(d1));
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int j = 0; j < KV; ++j) {
                const float4 ww = w[j >> 1];
                const float2 wj = (j & 1) ? make_float2(ww.z, ww.w) : make_float2(ww.x, ww.y);
                vacc[j][q] = __ffma2_rn(wj, s[q], j == fresh ? make_float2(0.0f, 0.0f) : vacc[j][q]);
            }
        }
    };
    // The same for row m of a uniform revolution, weights from the tap registers: slot j is at tap
    // (m - SV*(j+1)) mod LV of its window (the slot's previous output closed SV*(j+1) rows into the revolution).
    auto fma_row_taps = [&](uint32_t d0, uint32_t d1, const float2* wt, int m) {
        float2 s[4];
        s[0] = make_float2(byte_to_float<0>(d0), byte_to_float<1>(d0));
        s[1] = make_float2(byte_to_float<2>(d0), byte_to_float<3>(d0));
        s[2] = make_float2(byte_to_float<0>(d1), byte_to_float<1>(d1));
        s[3] = make_float2(byte_to_float<2>(d1), byte_to_float<3>(d1));
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int j = 0; j < KV; ++j) {
                const int t = (m - SV * (j + 1) + 2 * LV) % (LV > 0 ? LV : 1);  // t == 0: first row of a new output
                vacc[j][q] = __ffma2_rn(wt[t], s[q], t == 0 ? make_float2(0.0f, 0.0f) : vacc[j][q]);
            }
        }
    };
    // Called after the row that closes a stage: this warp has drained it (all its loads have returned:
    // their values were consumed).  The producer refills the stage once all four warps have arrived.
    auto stage_drained = [&](uint32_t stage) {
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar + stage);
    };
    // Virtual rows [rr, rend) into every open ring slot; two rows per trip when both lie in the same
    // ring stage (all loads first, then both rows' conversions and FMAs), otherwise one.
    auto consume_rows = [&](uint32_t rend) {
        while (rr < rend) {
            const uint32_t stage = (rr / kStageRows) % kStages;
            const uint32_t in_stage = rr % kStageRows;
            if (in_stage == 0) mbar_wait(full_bar + stage, (rr / kRingRows) & 1);  // the stage has landed
            const uint32_t ring_row = rr % kRingRows;
            const uint8_t* src = my_src + ring_row * kSrcRowBytes;
            const float4* wrow = vw_ring + ring_row * (KSV / 2);
            if (rr + 1 < rend && in_stage != kStageRows - 1) {
                const uint32_t a0 = *reinterpret_cast<const uint32_t*>(src);
                const uint32_t a1 = *reinterpret_cast<const uint32_t*>(src + kHalfRowBytes);
                const uint32_t c0 = *reinterpret_cast<const uint32_t*>(src + kSrcRowBytes);
                const uint32_t c1 = *reinterpret_cast<const uint32_t*>(src + kSrcRowBytes + kHalfRowBytes);
                float4 wa[KSV / 2], wc[KSV / 2];
#pragma unroll
                for (int jj = 0; jj < KSV / 2; ++jj) { wa[jj] = wrow[jj]; wc[jj] = wrow[KSV / 2 + jj]; }
                fma_row(a0, a1, wa, -1);
                fma_row(c0, c1, wc, -1);
                if (in_stage == kStageRows - 2) stage_drained(stage);
                rr += 2;
            } else {
                const uint32_t a0 = *reinterpret_cast<const uint32_t*>(src);
                const uint32_t a1 = *reinterpret_cast<const uint32_t*>(src + kHalfRowBytes);
                float4 wa[KSV / 2];
#pragma unroll
                for (int jj = 0; jj < KSV / 2; ++jj) wa[jj] = wrow[jj];
                fma_row(a0, a1, wa, -1);
                if (in_stage == kStageRows - 1) stage_drained(stage);
                rr += 1;
            }
        }
    };
    // Slot c's finished row -> tmp row `emitted`; the slot restarts from zero.
    auto emit_slot = [&](float2 (&a)[4]) {
        float* trow = reinterpret_cast<float*>(tmp + size_t(emitted) * geom.tmp_px);
        if (C == 4) {
            *reinterpret_cast<float4*>(trow + emit_off[0]) = make_float4(a[0].x, a[0].y, a[1].x, a[1].y);
            *reinterpret_cast<float4*>(trow + emit_off[4]) = make_float4(a[2].x, a[2].y, a[3].x, a[3].y);
        } else {
            trow[emit_off[0]] = a[0].x; trow[emit_off[1]] = a[0].y;
            trow[emit_off[2]] = a[1].x; trow[emit_off[3]] = a[1].y;
            trow[emit_off[4]] = a[2].x; trow[emit_off[5]] = a[2].y;
            trow[emit_off[6]] = a[3].x; trow[emit_off[7]] = a[3].y;
        }
        ++emitted;
    };

    // State of the specialised vertical loop: whether the revolution in progress is a uniform one, and
    // the ring stage its current rows live in (a stage spans several slots' rows, and a revolution may
    // be interrupted by the horizontal phase after any slot).
    bool v_fast = false, yend_stale = false;
    const uint8_t* f_src = my_src;
    const float4* f_w = vw_ring;
    uint32_t f_stage = 0;

    while (ov < oy1) {
        // ============================ vertical phase: fill tmp until the group is complete
        for (;;) {
            if (SV > 0 && c_start == 0) {
                v_fast = ov >= v_fast_lo && ov + KV <= v_fast_hi && (rr & 3) == 0 &&
                         (yend_stale || uint32_t(yend_next - y_virt0) == rr + SV);
            }
            if (SV > 0 && v_fast) {
                yend_stale = true;
                float2 wt[RWV ? LV : 1];
                if (RWV) {
#pragma unroll
                    for (int t = 0; t < LV; ++t) wt[t] = vtap[t];
                }
#pragma unroll
                for (int c = 0; c < KV; ++c) {
                    if (c >= c_start) {
#pragma unroll
                        for (int i = 0; i < SV; ++i) {
                            const int q = (c * SV + i) & 3;  // row within its stage (static after unrolling)
                            if (q == 0) {
                                f_stage = (rr / kStageRows) % kStages;
                                mbar_wait(full_bar + f_stage, (rr / kRingRows) & 1);
                                f_src = my_src + (rr % kRingRows) * kSrcRowBytes;
                                f_w = vw_ring + (rr % kRingRows) * (KSV / 2);
                            }
                            const uint32_t a0 = *reinterpret_cast<const uint32_t*>(f_src + q * kSrcRowBytes);
                            const uint32_t a1 = *reinterpret_cast<const uint32_t*>(f_src + q * kSrcRowBytes + kHalfRowBytes);
                            if (RWV) {
                                fma_row_taps(a0, a1, wt, c * SV + i);
                            } else {
                                float4 wa[KSV / 2];
#pragma unroll
                                for (int jj = 0; jj < KSV / 2; ++jj) wa[jj] = f_w[q * (KSV / 2) + jj];
                                // row i == 0 of slot c's rows is the first row of the output that re-uses slot c - 1
                                fma_row(a0, a1, wa, i == 0 ? (c + KV - 1) % KV : -1);
                            }
                            if (q == 3) stage_drained(f_stage);
                            rr += 1;
                        }
                        emit_slot(vacc[c]);  // (the slot is not zeroed: its next row starts it afresh)
                        ++ov;
                        if (emitted == kTmpRows) {
                            c_start = (c + 1) % KV;
                            goto vertical_done;
                        }
                    }
                }
                c_start = 0;
                if (ov == oy1) goto vertical_done;
                continue;
            }
            if (SV > 0 && yend_stale) {  // back from uniform revolutions: look the next window end up again,
                yend_next = v_end_of(ov);  // and clear the slot the last revolution emitted last
                yend_stale = false;
#pragma unroll
                for (int q = 0; q < 4; ++q) vacc[KV - 1][q] = make_float2(0.0f, 0.0f);
            }
#pragma unroll
            for (int c = 0; c < KV; ++c) {
                if (c >= c_start) {  // output ov accumulates in slot c
                    const uint32_t rend = uint32_t(yend_next - y_virt0);  // virtual rows output ov needs
                    yend_next = (ov + 1 < oy1) ? v_end_of(ov + 1) : 0x7fffffff;
                    consume_rows(rend);
                    if (ov >= oy0) emit_slot(vacc[c]);
#pragma unroll
                    for (int q = 0; q < 4; ++q) vacc[c][q] = make_float2(0.0f, 0.0f);
                    ++ov;
                    if (emitted == kTmpRows || ov == oy1) {
                        c_start = (c + 1) % KV;
                        goto vertical_done;
                    }
                }
            }
            c_start = 0;
        }
    vertical_done:
        compute_barrier();  // tmp rows [0, emitted) are complete

        // ============================ horizontal phase
        if (!hw_ready) {  // the strip's horizontal weights were requested at kernel start
            mbar_wait(hw_bar, 0);
            hw_ready = true;
        }
        uint8_t* const my_dst = dst_base + size_t(g0 + hrow) * dst_pitch;  // this lane's output row
        const bool row_live = hrow < emitted;
        if (h_active) {
            float2 hacc[KH][2];
#pragma unroll
            for (int j = 0; j < KH; ++j) hacc[j][0] = hacc[j][1] = make_float2(kRoundBias, kRoundBias);
            int x = seg_lo - h_pre;
            const float4* px = my_row + x;
            const float4* wh = hw_smem + (x - xl) * (KSH / 2);
            // Start one ring revolution early: the outputs before `os` whose windows are still open at
            // seg_lo hold their slots until they close; they are walked like any other output (consuming
            // the pixels up to their window end) but never stored.
            int oh = os - KH;
            int hc_start = h_slot0;
            auto window_of = [&](int o) { return (o >= o_lo && o < oe) ? lr_tab[o] : make_int2(0, 0); };
            int2 lr_next = window_of(oh);

            auto accumulate = [&](const float4& p, const float4* w, int fresh) {
                const float2 ph[2] = {make_float2(p.x, p.y), make_float2(p.z, p.w)};
#pragma unroll
                for (int q = 0; q < 2; ++q) {
#pragma unroll
                    for (int j = 0; j < KH; ++j) {
                        const float4 ww = w[j >> 1];
                        const float2 wj = (j & 1) ? make_float2(ww.z, ww.w) : make_float2(ww.x, ww.y);
                        hacc[j][q] = __ffma2_rn(wj, ph[q], j == fresh ? make_float2(kRoundBias, kRoundBias) : hacc[j][q]);
                    }
                }
            };
            for (;;) {
                if (SH > 0 && hc_start == 0 && oh >= h_fast_lo && oh + KH <= h_fast_hi && (oh < os || x + SH == lr_next.y)) {
                    // ---- uniform stretch: every slot finishes after exactly SH more pixels
                    uint8_t* d = my_dst + ptrdiff_t(oh) * CO;
                    float2 ht[RWH ? LH : 1];
                    if (RWH) {
#pragma unroll
                        for (int t = 0; t < LH; ++t) ht[t] = htap[t];
                    }
                    if (RWH && oh < os) {
                                // ---- pre-roll revolution: nothing is stored, and pixel m only matters to the slots
                                // whose real output has started by then (slot j restarts at pixel SH * (j + 1))
#pragma unroll
                        for (int c = 1; c < KH; ++c) {
#pragma unroll
                            for (int i = 0; i < SH; ++i) {
                                const float4 p = px[c * SH + i];
                                const float2 ph[2] = {make_float2(p.x, p.y), make_float2(p.z, p.w)};
#pragma unroll
                                for (int q = 0; q < 2; ++q) {
#pragma unroll
                                    for (int j = 0; j < c; ++j) {
                                        const int t = (c * SH + i - SH * (j + 1) + 2 * LH) % (LH > 0 ? LH : 1);
                                        hacc[j][q] = __ffma2_rn(ht[t], ph[q], t == 0 ? make_float2(kRoundBias, kRoundBias) : hacc[j][q]);
                                    }
                                }
                            }
                        }
                        oh += KH;
                        x += KH * SH;
                        px += KH * SH;
                        wh += KH * SH * (KSH / 2);
                        d += KH * CO;
                    }
                    do {
#pragma unroll
                        for (int c = 0; c < KH; ++c) {
#pragma unroll
                            for (int i = 0; i < SH; ++i) {
                                const float4 p = px[c * SH + i];
                                if (RWH) {  // pixel m of the revolution: slot j is at tap (m - SH*(j+1)) mod LH
                                    const float2 ph[2] = {make_float2(p.x, p.y), make_float2(p.z, p.w)};
#pragma unroll
                                    for (int q = 0; q < 2; ++q) {
#pragma unroll
                                        for (int j = 0; j < KH; ++j) {
                                            const int t = (c * SH + i - SH * (j + 1) + 2 * LH) % (LH > 0 ? LH : 1);
                                            hacc[j][q] = __ffma2_rn(ht[t], ph[q], t == 0 ? make_float2(kRoundBias, kRoundBias) : hacc[j][q]);
                                        }
                                    }
                                } else {
                                    float4 w[KSH / 2];
#pragma unroll
                                    for (int jj = 0; jj < KSH / 2; ++jj) w[jj] = wh[(c * SH + i) * (KSH / 2) + jj];
                                    accumulate(p, w, i == 0 ? (c + KH - 1) % KH : -1);
                                }
                            }
                            const float4 v = make_float4(hacc[c][0].x, hacc[c][0].y, hacc[c][1].x, hacc[c][1].y);
                            if (row_live && oh + c >= os) store_pixel<C>(d, v, CO);
                            d += CO;
                        }
                        oh += KH;
                        x += KH * SH;
                        px += KH * SH;
                        wh += KH * SH * (KSH / 2);
                    } while (oh + KH <= h_fast_hi);
                    hacc[KH - 1][0] = hacc[KH - 1][1] = make_float2(kRoundBias, kRoundBias);  // the other slots restarted themselves
                    lr_next = window_of(oh);
                }
#pragma unroll
                for (int c = 0; c < KH; ++c) {
                    if (c >= hc_start) {  // output oh accumulates in slot c
                        if (oh >= oe) goto horizontal_done;
                        const int xend = lr_next.y;
                        lr_next = window_of(oh + 1);
                        while (x + 1 < xend) {  // two pixels per trip: all loads are issued first
                            const float4 p0v = px[0], p1v = px[1];
                            float4 w0[KSH / 2], w1[KSH / 2];
#pragma unroll
                            for (int jj = 0; jj < KSH / 2; ++jj) { w0[jj] = wh[jj]; w1[jj] = wh[KSH / 2 + jj]; }
                            accumulate(p0v, w0, -1);
                            accumulate(p1v, w1, -1);
                            x += 2; px += 2; wh += KSH;
                        }
                        if (x < xend) {
                            const float4 p0v = px[0];
                            float4 w0[KSH / 2];
#pragma unroll
                            for (int jj = 0; jj < KSH / 2; ++jj) w0[jj] = wh[jj];
                            accumulate(p0v, w0, -1);
                            x += 1; px += 1; wh += KSH / 2;
                        }
                        if (oh >= os && row_live) {
                            const float4 v = make_float4(hacc[c][0].x, hacc[c][0].y, hacc[c][1].x, hacc[c][1].y);
                            store_pixel<C>(my_dst + size_t(oh) * CO, v, CO);
                        }
                        hacc[c][0] = hacc[c][1] = make_float2(kRoundBias, kRoundBias);
                        ++oh;
                    }
                }
                hc_start = 0;
            }
        horizontal_done:;
        }
        compute_barrier();  // tmp is free for the next vertical rows
        g0 += emitted;
        emitted = 0;
    }
}

// ---- launcher ---------------------------------------------------------------------------------

#if !IKC_FUSED_CONV
size_t fused_smem_bytes(int /*channels*/, int kv, int kh, const FusedGeom& g) {
    const size_t ksv = size_t((kv + 1) & ~1), ksh = size_t((kh + 1) & ~1);
    return size_t(kHeaderBytes) + size_t(kRingRows) * kSrcRowBytes + size_t(kRingRows) * ksv * 8 +
           size_t(g.tmp_px) * ksh * 8 + size_t(kMaxStripOut) * sizeof(int2) + size_t(kTmpRows) * g.tmp_px * sizeof(float4);
}

int fused_max_src_bytes(int channels) { return (channels >= 1 && channels <= 4) ? max_src_bytes(channels) : 0; }

int fused_group_rows() { return kTmpRows; }

bool fused_supported(int channels, int kv, int kh) {
    return channels >= 1 && channels <= 4 && kv >= 6 && kv <= 7 && kh >= 6 && kh <= 7;
}

// Specialised uniform loops exist for the integer ratios 2 and 4 (ring size 6, a whole number of ring
// stages per revolution) in both passes at once.
bool fused_has_uniform(int channels, int kv, int kh, int step_v, int step_h) {
    return fused_supported(channels, kv, kh) && channels >= 3 && kv == 6 && kh == 6 && step_v == step_h &&
           (step_v == 2 || step_v == 4);
}

#endif  // !IKC_FUSED_CONV

template <int C, int KV, int KH, int SV, int SH>
static cudaError_t launch_one(const DevJob* jobs, const WorkItem* items, const FusedGeom& geom, cudaStream_t stream) {
    const size_t smem = fused_smem_bytes(C, KV, KH, geom);
    if (smem > kFusedMaxSmem) return cudaErrorInvalidValue;
    // Opt in to > 48 KB dynamic shared memory.  The attribute belongs to the (kernel, device) pair, which
    // concurrent callers share: it is always set to the same planner-wide maximum, never to this launch's
    // own size, so two threads launching different geometries cannot lower it under each other.
    cudaError_t e = cudaFuncSetAttribute(fused_ring_kernel<C, KV, KH, SV, SH, kConv>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, int(kFusedMaxSmem));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(fused_ring_kernel<C, KV, KH, SV, SH, kConv>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    fused_ring_kernel<C, KV, KH, SV, SH, kConv><<<geom.n_items, kThreads, smem, stream>>>(jobs, items, geom);
    return cudaGetLastError();
}

#if IKC_FUSED_CONV
cudaError_t launch_fused_conv(int channels, int kv, int kh, int sv, int sh, const DevJob* jobs, const WorkItem* items,
                              const FusedGeom& geom, cudaStream_t stream) {
#else
cudaError_t launch_fused_conv(int channels, int kv, int kh, int sv, int sh, const DevJob* jobs, const WorkItem* items,
                              const FusedGeom& geom, cudaStream_t stream);  // fused_conv.cu

cudaError_t launch_fused(int channels, int kv, int kh, int sv, int sh, bool convert, const DevJob* jobs,
                         const WorkItem* items, const FusedGeom& geom, cudaStream_t stream) {
    if (convert) return launch_fused_conv(channels, kv, kh, sv, sh, jobs, items, geom, stream);
#endif
#define IKC_CASE(C_, KV_, KH_, SV_, SH_)                                    \
    if (channels == C_ && kv == KV_ && kh == KH_ && sv == SV_ && sh == SH_) \
        return launch_one<C_, KV_, KH_, SV_, SH_>(jobs, items, geom, stream);
    IKC_CASE(4, 6, 6, 2, 2) IKC_CASE(4, 6, 6, 4, 4) IKC_CASE(3, 6, 6, 2, 2) IKC_CASE(3, 6, 6, 4, 4)
    IKC_CASE(4, 6, 6, 0, 0) IKC_CASE(4, 6, 7, 0, 0) IKC_CASE(4, 7, 6, 0, 0) IKC_CASE(4, 7, 7, 0, 0)
    IKC_CASE(3, 6, 6, 0, 0) IKC_CASE(3, 6, 7, 0, 0) IKC_CASE(3, 7, 6, 0, 0) IKC_CASE(3, 7, 7, 0, 0)
    IKC_CASE(2, 6, 6, 0, 0) IKC_CASE(2, 6, 7, 0, 0) IKC_CASE(2, 7, 6, 0, 0) IKC_CASE(2, 7, 7, 0, 0)
    IKC_CASE(1, 6, 6, 0, 0) IKC_CASE(1, 6, 7, 0, 0) IKC_CASE(1, 7, 6, 0, 0) IKC_CASE(1, 7, 7, 0, 0)
#undef IKC_CASE
    return cudaErrorInvalidValue;
}

}  // namespace ikc
