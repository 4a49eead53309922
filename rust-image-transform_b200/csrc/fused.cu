// fused.cu -- fused single-launch "ring" resize kernel for 8-bit downscales (sm_100a).
//
// This is the product's main path for the reference's actual workload: Lanczos3 downscales of
// Rgb8/Rgba8 rasters (resize_image, /root/reference/src/transform.rs:62-90, whose arithmetic is
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

namespace ikc {
namespace {

// A CTA has 8 / WPT warps (WPT = 32-bit source words per thread per row); no dedicated producer warp, so
// the resident warps stay balanced over the 4 SMSPs.
                                                // over the 4 SMSPs and keep a 255-register budget
constexpr int kStageRows = 4;                   // source rows per ring stage (one mbarrier pair)
constexpr int kSrcRowBytes = 1024;              // staged bytes per source row
constexpr int kHalfRowBytes = kSrcRowBytes / 2; // offset of a thread's second word (WPT == 2)
constexpr int kTmpRows = 16;                    // f32 intermediate rows per group
constexpr int kHeaderBytes = 256;               // mbarriers
constexpr int kMaxStages = 8;
constexpr int kMaxStripOut = 272;                // outputs of one strip (256) + ring pre-roll, in the left/right table

template <int C>
struct Layout;
template <>
struct Layout<4> {
    static constexpr int kMaxSrcBytes = 1024;
    static constexpr int kRingRows = 16;
};
template <>
struct Layout<3> {
    static constexpr int kMaxSrcBytes = 864;
    static constexpr int kRingRows = 16;
};
template <>
struct Layout<2> {
    static constexpr int kMaxSrcBytes = 512;
    static constexpr int kRingRows = 16;
};
template <>
struct Layout<1> {
    static constexpr int kMaxSrcBytes = 256;
    static constexpr int kRingRows = 16;
};

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
// 1-D TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_addr(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar))
        : "memory");
}
__device__ __forceinline__ void compute_barrier() { __syncthreads(); }

// u8 -> f32, exact: PRMT drops the byte into the mantissa of 2^23, one FADD removes the bias.
template <int BYTE>
__device__ __forceinline__ float byte_to_float(uint32_t word) {
    return __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7440u + BYTE)) - 8388608.0f;
}

// clamp to [0,255] and round half away from zero (f32::round).  trunc(v + 0.5) equals round(v) for
// every non-negative float except v = 0.5 - 2^-25 (the add rounds up to 1.0); this path's sums differ
// from the reference's by FMA/association rounding anyway, and the EXACT path (generic.cu) has no such case.
__device__ __forceinline__ uint32_t quantize_u8(float v) {
    v = fminf(fmaxf(v, 0.0f), 255.0f);
    return __float2uint_rz(v + 0.5f);
}

// Quantise one finished pixel and store it straight to the destination raster.  Lanes of a half
// warp hold 16 different rows of the same column, so these are scattered 4-byte (or 1-byte) stores;
// the sectors are completed in L2 by the same thread's next pixels before they reach HBM.
template <int C>
__device__ __forceinline__ void store_pixel(uint8_t* dst_px, bool word_ok, float4 v) {
    const uint32_t r = quantize_u8(v.x), g = quantize_u8(v.y), b = quantize_u8(v.z), a = quantize_u8(v.w);
    if (C == 4 && word_ok) {
        *reinterpret_cast<uint32_t*>(dst_px) = r | (g << 8) | (b << 16) | (a << 24);
    } else {
        dst_px[0] = uint8_t(r);
        if (C > 1) dst_px[1] = uint8_t(g);
        if (C > 2) dst_px[2] = uint8_t(b);
        if (C > 3) dst_px[3] = uint8_t(a);
    }
}

// shared-memory atomic add issued by one lane (kept as inline PTX so the compiler does not wrap it
// in its warp-aggregation sequence)
__device__ __forceinline__ int smem_atomic_inc(int* p) {
    int old;
    asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_addr(p)) : "memory");
    return old;
}

}  // namespace

// Block = 4 warps; two CTAs are resident per SM.
//
// Shared memory: [mbarriers][source ring: kRingRows x 1024 B][vertical weight ring: kRingRows x KSV
//                pairs][horizontal weights of the strip: tmp_px x KSH pairs][(left,right) of the
//                strip's outputs][tmp: 16 rows x tmp_px x float4].  Weights travel with TMA bulk
//                copies too, so the hot loops only read shared memory.
//
// Vertical phase: thread t owns source byte columns [4t,4t+4) and [512+4t,512+4t+4) of the strip and
// marches down the source rows; its KV ring slots hold the partial sums of the <= KV output rows
// currently open (slot = output row mod KV).  A finished row goes to tmp as one float4 per pixel
// (fewer than 4 channels are padded to 4 lanes).  Every 16 finished rows the horizontal phase runs:
// lane & 15 = tmp row, half warp = x segment of the strip; the same ring march along x.  Outputs whose
// window straddles a segment boundary are completed from head/tail partial sums parked in tmp columns
// the thread has already consumed.  Finished pixels are quantised and stored straight to HBM.
template <int C, int KV, int KH, int WPT>
__global__ void __launch_bounds__(256 / WPT, 2)
fused_ring_kernel(const DevJob* __restrict__ jobs, const WorkItem* __restrict__ items, const FusedGeom geom) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int KSV = (KV + 1) & ~1;  // ring row stride (weight pairs), even
    constexpr int KSH = (KH + 1) & ~1;
    constexpr int kRingRows = Layout<C>::kRingRows;
    constexpr int kComputeWarps = 8 / WPT;
    constexpr int kThreads = 32 * kComputeWarps;
    constexpr int kMaxSegs = 2 * kComputeWarps;      // horizontal segments: one per half warp
    constexpr int kStages = kRingRows / kStageRows;
    static_assert(kStages <= kComputeWarps && kStageRows == 4 && kRingRows == 16, "ring geometry: warp s refills stage s");

    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty_bar = full_bar + kMaxStages;   // one arrival per warp that has drained the stage
    uint64_t* hw_bar = empty_bar + kMaxStages;
    uint8_t* src_ring = smem + kHeaderBytes;
    float4* vw_ring = reinterpret_cast<float4*>(src_ring + kRingRows * kSrcRowBytes);  // [row][KSV/2]
    float4* hw_smem = vw_ring + kRingRows * (KSV / 2);                                 // [px][KSH/2]
    int2* hlr = reinterpret_cast<int2*>(hw_smem + size_t(geom.tmp_px) * (KSH / 2));    // (left, right)
    float4* tmp = reinterpret_cast<float4*>(hlr + kMaxStripOut);

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
    const float4* __restrict__ vring = reinterpret_cast<const float4*>(J->v.ring);
    const float4* __restrict__ hring = reinterpret_cast<const float4*>(J->h.ring);

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

    // The source ring is refilled by whichever warp is last to finish a stage (see the row loop);
    // the first fill of every stage is issued here.
    const uint8_t* const gsrc = J->src + size_t(y_first) * J->src_pitch + b0;
    const size_t src_pitch = J->src_pitch;
    auto issue_fill = [&](int stage, int r0) {  // rows [r0, r0 + kStageRows) of the chunk -> stage
        const int n = min(kStageRows, nrows - r0);
        const uint32_t wbytes = uint32_t(n) * (KSV * 8);
        mbar_expect_tx(full_bar + stage, uint32_t(n) * uint32_t(nb) + wbytes);
        for (int i = 0; i < n; ++i)
            bulk_load(src_ring + (stage * kStageRows + i) * kSrcRowBytes, gsrc + size_t(r0 + i) * src_pitch,
                      uint32_t(nb), full_bar + stage);
        bulk_load(vw_ring + stage * kStageRows * (KSV / 2), vring + size_t(y_first + r0) * (KSV / 2), wbytes,
                  full_bar + stage);
    };
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_bar + s, 1);
            mbar_init(empty_bar + s, kComputeWarps);
        }
        mbar_init(hw_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int s = 0; s < kStages && s * kStageRows < nrows; ++s) issue_fill(s, s * kStageRows);
        // horizontal ring weights of the strip's source pixels [xl, xr): one bulk copy, used by every group
        const uint32_t hbytes = uint32_t(xr - xl) * (KSH * 8);
        mbar_expect_tx(hw_bar, hbytes);
        bulk_load(hw_smem, hring + size_t(xl) * (KSH / 2), hbytes, hw_bar);
    }
    // (left, right) of the strip's outputs (plus the few before ox0 whose windows reach into the strip)
    const int o_lo = max(0, ox0 - KH + 1);
    for (int i = tid; i < ox1 - o_lo; i += kThreads) hlr[i] = make_int2(__ldg(hleft + o_lo + i), __ldg(hright + o_lo + i));
    __syncthreads();
    const int2* const lr_tab = hlr - o_lo;  // indexed by absolute output column

    // ---------------------------------------------------------------- horizontal segmentation
    // The strip's outputs [ox0, ox1) are cut into n_seg runs of `per` outputs (a multiple of KH, so
    // every segment starts on the same ring slot and the half warps of a warp walk the unrolled slot
    // code in lock step).  Segment s owns outputs [os, oe) and the source pixels [seg_lo, seg_hi)
    // between the end of the previous segment's last window and the end of its own last window.
    const int span = xr - xl;
    const int n_out = ox1 - ox0;
    // A segment only needs 2*KH consumed pixels to park its head and tail partial sums in; windows may
    // span several segments (the fix-up pass adds the tails of all earlier segments they touch).
    const int want_seg = max(1, min(kMaxSegs, span / (2 * KH + 2)));
    const int per_min = (2 * KH * n_out + span - 1) / span + 1;  // outputs that cover >= 2*KH source pixels
    const int per = (max((n_out + want_seg - 1) / want_seg, per_min) + KH - 1) / KH * KH;
    const int hrow = lane & 15;
    const int sidx = 2 * warp + (lane >> 4);
    // A short last run (< KH outputs) is merged into the run before it: near the right edge several
    // windows end on the same (clamped) pixel, and a segment must own at least one pixel per head.
    int n_act = (n_out + per - 1) / per;
    if (n_act > 1 && n_out - (n_act - 1) * per < KH) --n_act;
    const int os = ox0 + sidx * per;
    const bool h_active = sidx < n_act;
    const bool has_next_seg = sidx + 1 < n_act;
    const int oe = has_next_seg ? os + per : ox1;
    int seg_lo = xr, seg_hi = xr;
    if (h_active) {
        seg_lo = (sidx == 0) ? xl : lr_tab[os - 1].y;
        seg_hi = lr_tab[oe - 1].y;
    }
    const int h_slot0 = ox0 % KH;  // ring slot of every segment's first output
    float4* const my_row = tmp + size_t(hrow) * geom.tmp_px - pxb;  // indexed by absolute source pixel
    uint8_t* const dst_base = J->dst;
    const size_t dst_pitch = J->dst_pitch;
    const bool word_ok = C == 4 && ((reinterpret_cast<uintptr_t>(dst_base) | dst_pitch) & 3) == 0;
    bool hw_ready = false;
    // Uniform stretch of this segment's outputs: whole ring revolutions [fast_lo, fast_hi) of outputs
    // whose windows lie inside the segment and end exactly fast_s pixels after the previous one (every
    // interior output of an integer-ratio resize).  The horizontal loop walks them without any per-output
    // window lookups or branches.
    int fast_lo = 0, fast_hi = 0, fast_s = 0;
    if (h_active) {
        auto prev_end = [&](int o) { return o > os ? lr_tab[o - 1].y : seg_lo; };
        int o = os;
        while (o < oe && lr_tab[o].x < seg_lo) ++o;  // heads are finished in the fix-up pass
        for (int cand = (o + KH - 1) / KH * KH; cand + KH <= oe; cand += KH) {
            const int s_px = lr_tab[cand].y - prev_end(cand);
            int hi = cand;
            while (hi < oe && lr_tab[hi].y - prev_end(hi) == s_px) ++hi;
            hi = cand + (hi - cand) / KH * KH;
            if (hi > cand && s_px >= 1) {
                fast_lo = cand; fast_hi = hi; fast_s = s_px;
                break;
            }
        }
    }

    // ---------------------------------------------------------------- vertical state
    float2 vacc[KV][2 * WPT];  // [slot][byte pair]: bytes 0-1, 2-3 of the first word (and of the second)
#pragma unroll
    for (int j = 0; j < KV; ++j)
#pragma unroll
        for (int q = 0; q < 2 * WPT; ++q) vacc[j][q] = make_float2(0.0f, 0.0f);
    const bool v_active0 = 4 * tid < nb;
    const bool v_active1 = WPT == 2 && kHalfRowBytes + 4 * tid < nb;
    const uint8_t* my_src = src_ring + 4 * tid;

    // Where this thread's eight vertical results land in a tmp row (float index within the row).
    int emit_off[4 * WPT];
#pragma unroll
    for (int i = 0; i < 4 * WPT; ++i) {
        const int byte = b0 + (i >> 2) * kHalfRowBytes + 4 * tid + (i & 3);
        emit_off[i] = (byte / C - pxb) * 4 + (byte % C);
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

    // All ring bookkeeping derives from r = rows of the chunk consumed so far: ring row r % 16, stage
    // (r / 4) % 4, fill parity (r / 16) & 1.
    uint32_t r = 0;
    int g0 = oy0;           // first output row of the group being assembled in tmp
    int emitted = 0;        // rows of that group already in tmp

    auto fma_row = [&](uint32_t d0, uint32_t d1, const float4 (&w)[KSV / 2]) {
        const float2 s0 = make_float2(byte_to_float<0>(d0), byte_to_float<1>(d0));
        const float2 s1 = make_float2(byte_to_float<2>(d0), byte_to_float<3>(d0));
        const float2 s2 = make_float2(byte_to_float<0>(d1), byte_to_float<1>(d1));
        const float2 s3 = make_float2(byte_to_float<2>(d1), byte_to_float<3>(d1));
#pragma unroll
        for (int j = 0; j < KV; ++j) {
            const float4 ww = w[j >> 1];
            const float2 wj = (j & 1) ? make_float2(ww.z, ww.w) : make_float2(ww.x, ww.y);
            vacc[j][0] = __ffma2_rn(wj, s0, vacc[j][0]);
            vacc[j][1] = __ffma2_rn(wj, s1, vacc[j][1]);
            if (WPT == 2) {
                vacc[j][2 * WPT - 2] = __ffma2_rn(wj, s2, vacc[j][2 * WPT - 2]);
                vacc[j][2 * WPT - 1] = __ffma2_rn(wj, s3, vacc[j][2 * WPT - 1]);
            }
        }
    };
    // Called after the row that closes a stage: this warp has drained `stage` (all its loads have
    // returned: their values were consumed).  Warp w refills stage w, one stage late: by then the other
    // warps have drained it too, so the wait normally falls through and nobody stalls on the slowest warp.
    auto stage_drained = [&](uint32_t stage, uint32_t r_last) {
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(empty_bar + stage);
            const uint32_t prev = (stage + kStages - 1) % kStages;
            if (uint32_t(warp) == prev && r_last >= 2 * kStageRows - 1) {
                const uint32_t prev_r0 = r_last - (2 * kStageRows - 1);  // first row the previous stage held
                if (prev_r0 + kRingRows < uint32_t(nrows)) {
                    mbar_wait(empty_bar + prev, (prev_r0 / kRingRows) & 1);
                    issue_fill(int(prev), int(prev_r0 + kRingRows));
                }
            }
        }
    };
    // Source rows [r, rend) of the chunk into every open ring slot; two rows per trip when both lie in
    // the same ring stage (all loads first, then both rows' conversions and FMAs), otherwise one.
    auto consume_rows = [&](uint32_t rend) {
        while (r < rend) {
            const uint32_t stage = (r / kStageRows) % kStages;
            const uint32_t in_stage = r % kStageRows;
            if (in_stage == 0) mbar_wait(full_bar + stage, (r / kRingRows) & 1);  // the stage has landed
            const uint32_t ring_row = r % kRingRows;
            const uint8_t* src = my_src + ring_row * kSrcRowBytes;
            const float4* wrow = vw_ring + ring_row * (KSV / 2);
            if (r + 1 < rend && in_stage != kStageRows - 1) {
                const uint32_t a0 = *reinterpret_cast<const uint32_t*>(src);
                const uint32_t a1 = WPT == 2 ? *reinterpret_cast<const uint32_t*>(src + kHalfRowBytes) : 0u;
                const uint32_t c0 = *reinterpret_cast<const uint32_t*>(src + kSrcRowBytes);
                const uint32_t c1 = WPT == 2 ? *reinterpret_cast<const uint32_t*>(src + kSrcRowBytes + kHalfRowBytes) : 0u;
                float4 wa[KSV / 2], wc[KSV / 2];
#pragma unroll
                for (int jj = 0; jj < KSV / 2; ++jj) { wa[jj] = wrow[jj]; wc[jj] = wrow[KSV / 2 + jj]; }
                fma_row(a0, a1, wa);
                fma_row(c0, c1, wc);
                if (in_stage == kStageRows - 2) stage_drained(stage, r + 1);
                r += 2;
            } else {
                const uint32_t a0 = *reinterpret_cast<const uint32_t*>(src);
                const uint32_t a1 = WPT == 2 ? *reinterpret_cast<const uint32_t*>(src + kHalfRowBytes) : 0u;
                float4 wa[KSV / 2];
#pragma unroll
                for (int jj = 0; jj < KSV / 2; ++jj) wa[jj] = wrow[jj];
                fma_row(a0, a1, wa);
                if (in_stage == kStageRows - 1) stage_drained(stage, r);
                r += 1;
            }
        }
    };

    while (ov < oy1) {
        // ============================ vertical phase: fill tmp until the group is complete
        for (;;) {
#pragma unroll
            for (int c = 0; c < KV; ++c) {
                if (c >= c_start) {  // output ov accumulates in slot c
                    const uint32_t rend = uint32_t(yend_next - y_first);  // rows of the chunk output ov needs
                    yend_next = (ov + 1 < oy1) ? v_end_of(ov + 1) : 0x7fffffff;
                    consume_rows(rend);
                    if (ov >= oy0) {
                        float* trow = reinterpret_cast<float*>(tmp + size_t(emitted) * geom.tmp_px);
                        if (C == 4) {
                            if (v_active0)
                                *reinterpret_cast<float4*>(trow + emit_off[0]) =
                                    make_float4(vacc[c][0].x, vacc[c][0].y, vacc[c][1].x, vacc[c][1].y);
                            if (WPT == 2 && v_active1)
                                *reinterpret_cast<float4*>(trow + emit_off[4 * WPT - 4]) =
                                    make_float4(vacc[c][2 * WPT - 2].x, vacc[c][2 * WPT - 2].y, vacc[c][2 * WPT - 1].x,
                                                vacc[c][2 * WPT - 1].y);
                        } else {
                            if (v_active0) {
                                trow[emit_off[0]] = vacc[c][0].x; trow[emit_off[1]] = vacc[c][0].y;
                                trow[emit_off[2]] = vacc[c][1].x; trow[emit_off[3]] = vacc[c][1].y;
                            }
                            if (WPT == 2 && v_active1) {
                                trow[emit_off[4 * WPT - 4]] = vacc[c][2 * WPT - 2].x; trow[emit_off[4 * WPT - 3]] = vacc[c][2 * WPT - 2].y;
                                trow[emit_off[4 * WPT - 2]] = vacc[c][2 * WPT - 1].x; trow[emit_off[4 * WPT - 1]] = vacc[c][2 * WPT - 1].y;
                            }
                        }
                        ++emitted;
                    }
#pragma unroll
                    for (int q = 0; q < 2 * WPT; ++q) vacc[c][q] = make_float2(0.0f, 0.0f);
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
        __syncthreads();  // tmp rows [0, emitted) are complete

        // ============================ horizontal phase
        if (!hw_ready) {  // the strip's horizontal weights were requested at kernel start
            mbar_wait(hw_bar, 0);
            hw_ready = true;
        }
        uint8_t* const my_dst = dst_base + size_t(g0 + hrow) * dst_pitch;  // this lane's output row
        const bool row_live = hrow < emitted;
        int n_heads = 0;
        if (h_active) {
            float2 hacc[KH][2];
#pragma unroll
            for (int j = 0; j < KH; ++j) hacc[j][0] = hacc[j][1] = make_float2(0.0f, 0.0f);
            int x = seg_lo;
            const float4* px = my_row + seg_lo;
            const float4* wh = hw_smem + size_t(seg_lo - xl) * (KSH / 2);
            // Start one ring revolution early: outputs before `os` that are still open at seg_lo (only
            // possible for the strip's first segment) hold their slots until their windows close; they
            // are walked like any other output but never emitted.  Every segment pre-rolls the same KH
            // outputs so that all half warps stay on the same unrolled slot.
            int oh = os - KH;
            int hc_start = h_slot0;
            auto window_of = [&](int o) { return (o >= o_lo && o < oe) ? lr_tab[o] : make_int2(0, 0); };
            int2 lr_next = window_of(oh);

            auto accumulate = [&](const float4& p, const float4* w) {
                const float2 plo = make_float2(p.x, p.y), phi = make_float2(p.z, p.w);
#pragma unroll
                for (int j = 0; j < KH; ++j) {
                    const float4 ww = w[j >> 1];
                    const float2 wj = (j & 1) ? make_float2(ww.z, ww.w) : make_float2(ww.x, ww.y);
                    hacc[j][0] = __ffma2_rn(wj, plo, hacc[j][0]);
                    hacc[j][1] = __ffma2_rn(wj, phi, hacc[j][1]);
                }
            };
            for (;;) {
                if (hc_start == 0 && oh >= fast_lo && oh + KH <= fast_hi) {
                    // ---- uniform stretch: slot c finishes after exactly fast_s more pixels
                    do {
#pragma unroll
                        for (int c = 0; c < KH; ++c) {
                            int left_px = fast_s;
                            for (; left_px >= 2; left_px -= 2) {  // two pixels per trip: all loads first
                                const float4 p0 = px[0], p1 = px[1];
                                float4 w0[KSH / 2], w1[KSH / 2];
#pragma unroll
                                for (int jj = 0; jj < KSH / 2; ++jj) { w0[jj] = wh[jj]; w1[jj] = wh[KSH / 2 + jj]; }
                                accumulate(p0, w0);
                                accumulate(p1, w1);
                                px += 2; wh += KSH;
                            }
                            if (left_px) {
                                const float4 p0 = px[0];
                                float4 w0[KSH / 2];
#pragma unroll
                                for (int jj = 0; jj < KSH / 2; ++jj) w0[jj] = wh[jj];
                                accumulate(p0, w0);
                                px += 1; wh += KSH / 2;
                            }
                            const float4 v = make_float4(hacc[c][0].x, hacc[c][0].y, hacc[c][1].x, hacc[c][1].y);
                            if (row_live) store_pixel<C>(my_dst + size_t(oh + c) * C, word_ok, v);
                            hacc[c][0] = hacc[c][1] = make_float2(0.0f, 0.0f);
                        }
                        oh += KH;
                        x += KH * fast_s;
                    } while (oh + KH <= fast_hi);
                    lr_next = window_of(oh);
                }
#pragma unroll
                for (int c = 0; c < KH; ++c) {
                    if (c >= hc_start) {  // output oh accumulates in slot c
                        if (oh >= oe) goto horizontal_done;
                        const int2 lr = lr_next;
                        lr_next = window_of(oh + 1);
                        const int xend = lr.y;  // <= seg_hi: every owned window ends inside the segment
                        while (x + 1 < xend) {  // two pixels per trip: all loads are issued first
                            const float4 p0 = px[0], p1 = px[1];
                            float4 w0[KSH / 2], w1[KSH / 2];
#pragma unroll
                            for (int jj = 0; jj < KSH / 2; ++jj) { w0[jj] = wh[jj]; w1[jj] = wh[KSH / 2 + jj]; }
                            accumulate(p0, w0);
                            accumulate(p1, w1);
                            x += 2; px += 2; wh += KSH;
                        }
                        if (x < xend) {
                            const float4 p0 = px[0];
                            float4 w0[KSH / 2];
#pragma unroll
                            for (int jj = 0; jj < KSH / 2; ++jj) w0[jj] = wh[jj];
                            accumulate(p0, w0);
                            x += 1; px += 1; wh += KSH / 2;
                        }
                        const float4 v = make_float4(hacc[c][0].x, hacc[c][0].y, hacc[c][1].x, hacc[c][1].y);
                        if (oh >= os) {
                            if (lr.x >= seg_lo) {  // the whole window lies in this segment: finished pixel
                                if (row_live) store_pixel<C>(my_dst + size_t(oh) * C, word_ok, v);
                            } else {               // head: the window started in an earlier segment; park the
                                my_row[seg_lo + n_heads] = v;  // partial sum in a tmp column already consumed
                                ++n_heads;
                            }
                        }
                        hacc[c][0] = hacc[c][1] = make_float2(0.0f, 0.0f);
                        ++oh;
                    }
                }
                hc_start = 0;
            }
        horizontal_done:
            // tails: partial sums of the windows still open at seg_hi (slot = output index mod KH)
            if (has_next_seg) {
#pragma unroll
                for (int j = 0; j < KH; ++j)
                    my_row[seg_lo + KH + j] = make_float4(hacc[j][0].x, hacc[j][0].y, hacc[j][1].x, hacc[j][1].y);
            }
        }
        __syncthreads();
        // fix-up: heads (the first n_heads outputs of the segment) + tails of earlier segments
        for (int i = 0; i < n_heads; ++i) {
            const int oh = os + i;
            float4 v = my_row[seg_lo + i];
            const int first = lr_tab[oh].x;
            const int slot = oh % KH;
            for (int sg = sidx - 1; sg >= 0; --sg) {
                const int sg_hi = lr_tab[ox0 + (sg + 1) * per - 1].y;
                if (sg_hi <= first) break;  // the window starts after that segment
                const int sg_lo = (sg == 0) ? xl : lr_tab[ox0 + sg * per - 1].y;
                const float4 t = my_row[sg_lo + KH + slot];
                v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
            }
            if (row_live) store_pixel<C>(my_dst + size_t(oh) * C, word_ok, v);
        }
        __syncthreads();  // tmp (incl. parked partial sums) is free for the next vertical rows
        g0 += emitted;
        emitted = 0;
    }
}

// ---- launcher ---------------------------------------------------------------------------------

static int ring_rows_for(int channels) {
    switch (channels) {
        case 4: return Layout<4>::kRingRows;
        case 3: return Layout<3>::kRingRows;
        case 2: return Layout<2>::kRingRows;
        default: return Layout<1>::kRingRows;
    }
}

size_t fused_smem_bytes(int channels, int kv, int kh, const FusedGeom& g) {
    const size_t ksv = size_t((kv + 1) & ~1), ksh = size_t((kh + 1) & ~1);
    const size_t ring = size_t(ring_rows_for(channels));
    return size_t(kHeaderBytes) + ring * kSrcRowBytes + ring * ksv * 8 + size_t(g.tmp_px) * ksh * 8 +
           size_t(kMaxStripOut) * sizeof(int2) + size_t(kTmpRows) * g.tmp_px * sizeof(float4);
}

int fused_max_src_bytes(int channels) {
    switch (channels) {
        case 4: return Layout<4>::kMaxSrcBytes;
        case 3: return Layout<3>::kMaxSrcBytes;
        case 2: return Layout<2>::kMaxSrcBytes;
        case 1: return Layout<1>::kMaxSrcBytes;
    }
    return 0;
}

int fused_group_rows() { return kTmpRows; }

bool fused_supported(int channels, int kv, int kh) {
    return (channels == 3 || channels == 4) && kv >= 6 && kv <= 7 && kh >= 6 && kh <= 7;
}

template <int C, int KV, int KH, int WPT>
static cudaError_t launch_one(const DevJob* jobs, const WorkItem* items, const FusedGeom& geom,
                              cudaStream_t stream) {
    const size_t smem = fused_smem_bytes(C, KV, KH, geom);
    // Opt in to > 48 KB dynamic shared memory (per device; cheap, so done on every launch).
    cudaError_t e = cudaFuncSetAttribute(fused_ring_kernel<C, KV, KH, WPT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         int(smem));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(fused_ring_kernel<C, KV, KH, WPT>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    fused_ring_kernel<C, KV, KH, WPT><<<geom.n_items, 256 / WPT, smem, stream>>>(jobs, items, geom);
    return cudaGetLastError();
}

cudaError_t launch_fused(int channels, int kv, int kh, int wpt, const DevJob* jobs, const WorkItem* items,
                         const FusedGeom& geom, cudaStream_t stream) {
#define IKC_CASE(C_, KV_, KH_)                                                                              \
    if (channels == C_ && kv == KV_ && kh == KH_)                                                           \
        return wpt == 1 ? launch_one<C_, KV_, KH_, 1>(jobs, items, geom, stream)                            \
                        : launch_one<C_, KV_, KH_, 2>(jobs, items, geom, stream);
    IKC_CASE(4, 6, 6) IKC_CASE(4, 6, 7) IKC_CASE(4, 7, 6) IKC_CASE(4, 7, 7)
    IKC_CASE(3, 6, 6) IKC_CASE(3, 6, 7) IKC_CASE(3, 7, 6) IKC_CASE(3, 7, 7)
#undef IKC_CASE
    return cudaErrorInvalidValue;
}

}  // namespace ikc
