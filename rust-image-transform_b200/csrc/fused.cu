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

constexpr int kComputeWarps = 4;
constexpr int kComputeThreads = kComputeWarps * 32;
constexpr int kThreads = kComputeThreads;       // no dedicated producer warp: 8 warps/SM stay balanced
                                                // over the 4 SMSPs and keep a 255-register budget
constexpr int kStageRows = 4;                   // source rows per ring stage (one mbarrier pair)
constexpr int kSrcRowBytes = 1024;              // staged bytes per source row: 128 threads x 2 x 4
constexpr int kHalfRowBytes = kSrcRowBytes / 2;
constexpr int kTmpRows = 16;                    // f32 intermediate rows per group
constexpr int kMaxSegs = 2 * kComputeWarps;     // horizontal segments: one per half warp
constexpr int kHeaderBytes = 256;               // mbarriers
constexpr int kMaxStages = 8;

template <int C>
struct Layout;
template <>
struct Layout<4> {
    static constexpr int kMaxSrcBytes = 1024;
    static constexpr int kRingRows = 24;
};
template <>
struct Layout<3> {
    static constexpr int kMaxSrcBytes = 864;
    static constexpr int kRingRows = 20;
};
template <>
struct Layout<2> {
    static constexpr int kMaxSrcBytes = 512;
    static constexpr int kRingRows = 24;
};
template <>
struct Layout<1> {
    static constexpr int kMaxSrcBytes = 256;
    static constexpr int kRingRows = 24;
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

// clamp to [0,255] and round half away from zero (f32::round), exact for every float
__device__ __forceinline__ uint32_t quantize_u8(float v) {
    v = fminf(fmaxf(v, 0.0f), 255.0f);
    const float t = truncf(v);
    return uint32_t((v - t >= 0.5f) ? t + 1.0f : t);
}

template <int C>
__device__ __forceinline__ void store_pixel(uint32_t* out_row, int op, float4 v) {
    if (C == 4) {
        out_row[op] = quantize_u8(v.x) | (quantize_u8(v.y) << 8) | (quantize_u8(v.z) << 16) |
                      (quantize_u8(v.w) << 24);
    } else {
        uint8_t* ob = reinterpret_cast<uint8_t*>(out_row) + op * C;
        ob[0] = uint8_t(quantize_u8(v.x));
        if (C > 1) ob[1] = uint8_t(quantize_u8(v.y));
        if (C > 2) ob[2] = uint8_t(quantize_u8(v.z));
    }
}

}  // namespace

// Block = 4 warps; two CTAs are resident per SM.
//
// Shared memory: [mbarriers][source ring: kRingRows x 1024 B][tmp: 16 rows x tmp_px x float4]
//                [out stage: 16 rows x out_pitch_w words].
//
// Vertical phase: compute thread t owns source byte columns [4t,4t+4) and [512+4t,512+4t+4) of the
// strip and marches down the source rows; its KV ring slots hold the partial sums of the <= KV
// output rows currently open.  A finished row goes to tmp as one float4 per pixel (fewer than 4
// channels are padded to 4 lanes).  Horizontal phase (per group of 16 tmp rows): lane & 15 = tmp
// row, half warp = x segment of the strip; the same ring march along x.  Outputs whose window
// straddles a segment boundary are completed from head/tail partial sums parked in tmp columns the
// thread has already consumed.
template <int C, int KV, int KH>
__global__ void __launch_bounds__(kThreads, 2)
fused_ring_kernel(const DevJob* __restrict__ jobs, const WorkItem* __restrict__ items, const FusedGeom geom) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int KSV = (KV + 1) & ~1;  // ring row stride (weight pairs), even
    constexpr int KSH = (KH + 1) & ~1;
    constexpr int kRingRows = Layout<C>::kRingRows;
    constexpr int kStages = kRingRows / kStageRows;
    static_assert(kStages <= kMaxStages && kRingRows % kStageRows == 0, "ring geometry");

    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
    int* rel_count = reinterpret_cast<int*>(full_bar + kMaxStages);  // warps done with each stage
    uint8_t* src_ring = smem + kHeaderBytes;
    float4* tmp = reinterpret_cast<float4*>(src_ring + kRingRows * kSrcRowBytes);
    uint32_t* out_stage = reinterpret_cast<uint32_t*>(tmp + size_t(kTmpRows) * geom.tmp_px);

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

    // The source ring is refilled by whichever warp is last to finish a stage (see the vertical
    // loop); the first fill of every stage is issued here.
    const uint8_t* const gsrc = J->src + size_t(y_first) * J->src_pitch + b0;
    const size_t src_pitch = J->src_pitch;
    auto issue_fill = [&](int stage, int r0) {  // rows [r0, r0 + kStageRows) of the chunk -> stage
        const int n = min(kStageRows, nrows - r0);
        mbar_expect_tx(full_bar + stage, uint32_t(n) * uint32_t(nb));
        for (int i = 0; i < n; ++i)
            bulk_load(src_ring + (stage * kStageRows + i) * kSrcRowBytes, gsrc + size_t(r0 + i) * src_pitch,
                      uint32_t(nb), full_bar + stage);
    };
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_bar + s, 1);
            rel_count[s] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int s = 0; s < kStages && s * kStageRows < nrows; ++s) issue_fill(s, s * kStageRows);
    }
    __syncthreads();

    // ------------------------------------------------------------------ compute warps
    // ---- vertical state (lives across groups)
    float2 vacc[KV][4];  // [slot][byte pair]: bytes 0-1, 2-3 of the first word; 0-1, 2-3 of the second
#pragma unroll
    for (int j = 0; j < KV; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) vacc[j][q] = make_float2(0.0f, 0.0f);
    const bool v_active0 = 4 * tid < nb;
    const bool v_active1 = kHalfRowBytes + 4 * tid < nb;
    const uint8_t* my_src = src_ring + 4 * tid;
    const float4* wv = vring + size_t(y_first) * (KSV / 2);  // ring weights of the prefetched row
    int y = y_first;         // next source row to consume (its data is in the nxt_* registers)
    int ready = y_first;     // rows below `ready` have landed in the ring
    int fill_stage = 0;      // next stage to wait for
    uint32_t fill_phase = 0;
    int rel_stage = 0;       // stage being drained
    int rel_base = 0;        // chunk-relative index of the first row held by that stage
    int rows_in_stage = 0;   // rows consumed from the stage being drained
    int ring_row = 0;        // ring row of source row y
    uint32_t nxt0 = 0, nxt1 = 0;
    float4 nxtw[KSV / 2];

    // Where this thread's eight vertical results land in a tmp row (float index within the row).
    int emit_off[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int byte = b0 + (i >> 2) * kHalfRowBytes + 4 * tid + (i & 3);
        emit_off[i] = (byte / C - pxb) * 4 + (byte % C);
    }

    // prefetch row y_first
    if (nrows > 0) {
        mbar_wait(full_bar + 0, 0);
        fill_stage = (kStages > 1) ? 1 : 0;
        fill_phase = (kStages > 1) ? 0 : 1;
        ready = min(y_first + kStageRows, y_last);
        nxt0 = *reinterpret_cast<const uint32_t*>(my_src);
        nxt1 = *reinterpret_cast<const uint32_t*>(my_src + kHalfRowBytes);
#pragma unroll
        for (int jj = 0; jj < KSV / 2; ++jj) nxtw[jj] = __ldg(wv + jj);
    }

    // ---- horizontal segmentation of the strip: segment sidx owns source pixels [seg_lo, seg_hi)
    const int span = xr - xl;
    const int max_count_h = J->h.max_count;
    const int n_seg = max(1, min(kMaxSegs, span / (max_count_h + 2 * KH + 2)));
    const int seg_len = (span + n_seg - 1) / n_seg;
    const int hrow = lane & 15;
    const int sidx = 2 * warp + (lane >> 4);
    const bool h_active = sidx < n_seg;
    const int seg_lo = min(xl + sidx * seg_len, xr);
    const int seg_hi = min(seg_lo + seg_len, xr);
    // First output whose window ends after seg_lo (binary search over the monotone `right`).
    int o_first = max(0, ox0 - KH + 1);
    if (h_active) {
        int lo = o_first, hi = ox1;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(hright + mid) > seg_lo) hi = mid; else lo = mid + 1;
        }
        o_first = lo;
    }
    const int o_ring0 = (o_first / KH) * KH;  // ring slot of output o is o % KH

    const int out_bytes = (ox1 - ox0) * C;
    float4* const my_row = tmp + size_t(hrow) * geom.tmp_px - pxb;  // indexed by absolute source pixel
    uint32_t* const my_out = out_stage + size_t(hrow) * geom.out_pitch_w;

    // Advance to the next source row and start loading it (data word pair + ring weights).
    auto v_prefetch = [&](uint32_t& d0, uint32_t& d1, float4 (&w)[KSV / 2]) {
        ++y;
        ring_row = (ring_row + 1 == kRingRows) ? 0 : ring_row + 1;
        wv += KSV / 2;
        if (y < y_last) {
            if (y == ready) {  // the next ring stage must have landed
                mbar_wait(full_bar + fill_stage, fill_phase);
                if (++fill_stage == kStages) { fill_stage = 0; fill_phase ^= 1; }
                ready = min(ready + kStageRows, y_last);
            }
            d0 = *reinterpret_cast<const uint32_t*>(my_src + ring_row * kSrcRowBytes);
            d1 = *reinterpret_cast<const uint32_t*>(my_src + ring_row * kSrcRowBytes + kHalfRowBytes);
#pragma unroll
            for (int jj = 0; jj < KSV / 2; ++jj) w[jj] = __ldg(wv + jj);
        }
    };
    // Accumulate one source row (8 byte columns) into every open ring slot.
    auto v_accumulate = [&](uint32_t d0, uint32_t d1, const float4 (&w)[KSV / 2]) {
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
            vacc[j][2] = __ffma2_rn(wj, s2, vacc[j][2]);
            vacc[j][3] = __ffma2_rn(wj, s3, vacc[j][3]);
        }
        if (++rows_in_stage == kStageRows) {
            // This warp is done with the stage; the last of the 4 warps refills it with the rows one
            // ring revolution further down (TMA bulk copies).
            rows_in_stage = 0;
            __syncwarp();
            if (lane == 0) {
                __threadfence_block();
                if (atomicAdd(rel_count + rel_stage, 1) == kComputeWarps - 1) {
                    rel_count[rel_stage] = 0;
                    __threadfence_block();
                    if (rel_base + kRingRows < nrows) issue_fill(rel_stage, rel_base + kRingRows);
                }
            }
            rel_base += kStageRows;
            if (++rel_stage == kStages) rel_stage = 0;
        }
    };

    // Ring order starts KV outputs above the chunk: rows from y_first on also belong to those
    // outputs; they are consumed but never emitted, and their slots clear as their windows close.
    int ov = oy0 - KV;
    int c_start = ((ov % KV) + KV) % KV;  // slot of output ov is ov mod KV

    for (int g0 = oy0; g0 < oy1; g0 += kTmpRows) {
        const int n_rows = min(kTmpRows, oy1 - g0);
        int remaining = n_rows;

        // ============================ vertical phase: fill tmp rows [0, n_rows)
        for (;;) {
#pragma unroll
            for (int c = 0; c < KV; ++c) {
                if (c >= c_start) {
                    const int yend = (ov >= 0) ? __ldg(vright + ov) : y;
                    while (y < yend) {
                        // Two rows per trip, ping-ponging between the nxt_* registers and a second
                        // set, so each row's loads are issued one row ahead without register copies.
                        uint32_t alt0, alt1;
                        float4 altw[KSV / 2];
                        if (y + 1 < yend) {
                            v_prefetch(alt0, alt1, altw);
                            v_accumulate(nxt0, nxt1, nxtw);
                            v_prefetch(nxt0, nxt1, nxtw);
                            v_accumulate(alt0, alt1, altw);
                        } else {
                            v_prefetch(alt0, alt1, altw);
                            v_accumulate(nxt0, nxt1, nxtw);
                            nxt0 = alt0;
                            nxt1 = alt1;
#pragma unroll
                            for (int jj = 0; jj < KSV / 2; ++jj) nxtw[jj] = altw[jj];
                        }
                    }
                    const bool live = ov >= oy0;
                    if (live) {
                        float* trow = reinterpret_cast<float*>(tmp + size_t(ov - g0) * geom.tmp_px);
                        if (C == 4) {
                            if (v_active0)
                                *reinterpret_cast<float4*>(trow + emit_off[0]) =
                                    make_float4(vacc[c][0].x, vacc[c][0].y, vacc[c][1].x, vacc[c][1].y);
                            if (v_active1)
                                *reinterpret_cast<float4*>(trow + emit_off[4]) =
                                    make_float4(vacc[c][2].x, vacc[c][2].y, vacc[c][3].x, vacc[c][3].y);
                        } else {
                            if (v_active0) {
                                trow[emit_off[0]] = vacc[c][0].x;
                                trow[emit_off[1]] = vacc[c][0].y;
                                trow[emit_off[2]] = vacc[c][1].x;
                                trow[emit_off[3]] = vacc[c][1].y;
                            }
                            if (v_active1) {
                                trow[emit_off[4]] = vacc[c][2].x;
                                trow[emit_off[5]] = vacc[c][2].y;
                                trow[emit_off[6]] = vacc[c][3].x;
                                trow[emit_off[7]] = vacc[c][3].y;
                            }
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) vacc[c][q] = make_float2(0.0f, 0.0f);
                    ++ov;
                    if (live && --remaining == 0) {
                        c_start = (c + 1) % KV;
                        goto vertical_done;
                    }
                }
            }
            c_start = 0;
        }
    vertical_done:
        compute_barrier();

        // ============================ horizontal phase
        int n_heads = 0;
        int head0 = 0;
        if (h_active) {
            float2 hacc[KH][2];
#pragma unroll
            for (int j = 0; j < KH; ++j) hacc[j][0] = hacc[j][1] = make_float2(0.0f, 0.0f);
            int x = seg_lo;
            const float4* px = my_row + seg_lo;
            const float4* wh = hring + size_t(seg_lo) * (KSH / 2);
            float4 nxtp = make_float4(0.f, 0.f, 0.f, 0.f);
            float4 nxth[KSH / 2];
            if (x < seg_hi) {
                nxtp = *px;
#pragma unroll
                for (int jj = 0; jj < KSH / 2; ++jj) nxth[jj] = __ldg(wh + jj);
            }
            auto h_prefetch = [&](float4& p, float4 (&w)[KSH / 2]) {
                ++x;
                ++px;
                wh += KSH / 2;
                if (x < seg_hi) {
                    p = *px;
#pragma unroll
                    for (int jj = 0; jj < KSH / 2; ++jj) w[jj] = __ldg(wh + jj);
                }
            };
            auto h_accumulate = [&](const float4& p, const float4 (&w)[KSH / 2]) {
                const float2 plo = make_float2(p.x, p.y), phi = make_float2(p.z, p.w);
#pragma unroll
                for (int j = 0; j < KH; ++j) {
                    const float4 ww = w[j >> 1];
                    const float2 wj = (j & 1) ? make_float2(ww.z, ww.w) : make_float2(ww.x, ww.y);
                    hacc[j][0] = __ffma2_rn(wj, plo, hacc[j][0]);
                    hacc[j][1] = __ffma2_rn(wj, phi, hacc[j][1]);
                }
            };
            bool seg_done = false;
            for (int oh0 = o_ring0; oh0 < ox1 && !seg_done; oh0 += KH) {
#pragma unroll
                for (int c = 0; c < KH; ++c) {
                    const int oh = oh0 + c;
                    if (oh >= ox1) { seg_done = true; break; }
                    const int r_end = __ldg(hright + oh);
                    const int xend = min(r_end, seg_hi);
                    while (x < xend) {
                        float4 altp;
                        float4 alth[KSH / 2];
                        if (x + 1 < xend) {
                            h_prefetch(altp, alth);
                            h_accumulate(nxtp, nxth);
                            h_prefetch(nxtp, nxth);
                            h_accumulate(altp, alth);
                        } else {
                            h_prefetch(altp, alth);
                            h_accumulate(nxtp, nxth);
                            nxtp = altp;
#pragma unroll
                            for (int jj = 0; jj < KSH / 2; ++jj) nxth[jj] = alth[jj];
                        }
                    }
                    if (r_end > seg_hi) { seg_done = true; break; }  // window continues in the next segment
                    if (r_end > seg_lo && oh >= ox0) {
                        const float4 v = make_float4(hacc[c][0].x, hacc[c][0].y, hacc[c][1].x, hacc[c][1].y);
                        if (__ldg(hleft + oh) >= seg_lo) {
                            store_pixel<C>(my_out, oh - ox0, v);  // complete inside this segment
                        } else {
                            // head: the window started in an earlier segment; park the partial sum in a
                            // tmp column this thread has already consumed
                            if (n_heads == 0) head0 = oh;
                            my_row[seg_lo + n_heads] = v;
                            ++n_heads;
                        }
                    }
                    hacc[c][0] = hacc[c][1] = make_float2(0.0f, 0.0f);
                }
            }
            // tails: partial sums of the windows still open at seg_hi (slot = output index mod KH)
            if (sidx + 1 < n_seg) {
#pragma unroll
                for (int j = 0; j < KH; ++j)
                    my_row[seg_lo + KH + j] = make_float4(hacc[j][0].x, hacc[j][0].y, hacc[j][1].x, hacc[j][1].y);
            }
        }
        compute_barrier();

        // ============================ fix-up: heads + tails of earlier segments -> finished outputs
        for (int i = 0; i < n_heads; ++i) {
            const int oh = head0 + i;
            float4 v = my_row[seg_lo + i];
            const int first = __ldg(hleft + oh);
            const int slot = oh % KH;
            for (int sg = sidx - 1; sg >= 0; --sg) {
                const int sg_lo = xl + sg * seg_len;
                if (sg_lo + seg_len <= first) break;  // the window starts after that segment
                const float4 t = my_row[sg_lo + KH + slot];
                v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
            }
            store_pixel<C>(my_out, oh - ox0, v);
        }
        compute_barrier();

        // ============================ store the finished rows (coalesced, any alignment)
        for (int row = warp; row < n_rows; row += kComputeWarps) {
            const uint32_t* srow = out_stage + size_t(row) * geom.out_pitch_w;
            const uint8_t* sbytes = reinterpret_cast<const uint8_t*>(srow);
            uint8_t* g = J->dst + size_t(g0 + row) * J->dst_pitch + size_t(ox0) * C;
            const int head = min(out_bytes, int((4 - (reinterpret_cast<uintptr_t>(g) & 3)) & 3));
            const int nwords = (out_bytes - head) >> 2;
            if (lane < head) g[lane] = sbytes[lane];
            uint32_t* gw = reinterpret_cast<uint32_t*>(g + head);
            const int sh8 = head * 8;
            for (int k = lane; k < nwords; k += 32) gw[k] = __funnelshift_r(srow[k], srow[k + 1], sh8);
            const int done = head + nwords * 4;
            if (lane < out_bytes - done) g[done + lane] = sbytes[done + lane];
        }
        // No barrier needed here: the next vertical phase only writes tmp (all reads of tmp finished
        // before the barrier above) and ends in a barrier before out_stage is written again.
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

size_t fused_smem_bytes(int channels, const FusedGeom& g) {
    return size_t(kHeaderBytes) + size_t(ring_rows_for(channels)) * kSrcRowBytes +
           size_t(kTmpRows) * g.tmp_px * sizeof(float4) + size_t(kTmpRows) * g.out_pitch_w * 4;
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
int fused_max_segments() { return kMaxSegs; }

bool fused_supported(int channels, int kv, int kh) {
    return (channels == 3 || channels == 4) && kv >= 6 && kv <= 7 && kh >= 6 && kh <= 7;
}

template <int C, int KV, int KH>
static cudaError_t launch_one(const DevJob* jobs, const WorkItem* items, const FusedGeom& geom,
                              cudaStream_t stream) {
    const size_t smem = fused_smem_bytes(C, geom);
    // Opt in to > 48 KB dynamic shared memory (per device; cheap, so done on every launch).
    cudaError_t e = cudaFuncSetAttribute(fused_ring_kernel<C, KV, KH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         int(smem));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(fused_ring_kernel<C, KV, KH>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    fused_ring_kernel<C, KV, KH><<<geom.n_items, kThreads, smem, stream>>>(jobs, items, geom);
    return cudaGetLastError();
}

cudaError_t launch_fused(int channels, int kv, int kh, const DevJob* jobs, const WorkItem* items,
                         const FusedGeom& geom, cudaStream_t stream) {
#define IKC_CASE(C_, KV_, KH_) \
    if (channels == C_ && kv == KV_ && kh == KH_) return launch_one<C_, KV_, KH_>(jobs, items, geom, stream);
    IKC_CASE(4, 6, 6) IKC_CASE(4, 6, 7) IKC_CASE(4, 7, 6) IKC_CASE(4, 7, 7)
    IKC_CASE(3, 6, 6) IKC_CASE(3, 6, 7) IKC_CASE(3, 7, 6) IKC_CASE(3, 7, 7)
#undef IKC_CASE
    return cudaErrorInvalidValue;
}

}  // namespace ikc
