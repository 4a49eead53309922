// banded.cu -- fused single-launch resize kernel for 8-bit downscales whose VERTICAL pass runs on the
// 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM) -- sm_100a only.
//
// Why: resize_image's arithmetic (/root/reference/src/transform.rs:85-89 -> image 0.25.8 imageops::resize =
// vertical_sample then horizontal_sample) costs 6*C FMAs per source pixel in the vertical pass of a Lanczos3
// downscale whatever the ratio, which puts the pass above B200's FP32 ridge: on the CUDA cores the FMA pipe,
// not HBM, is the ceiling (fused.cu reached 24 % of the HBM roofline).  The vertical pass is a banded matrix
// product   tmp[dh x (sw*C)] = Wv[dh x sh] . src[sh x (sw*C)]   and the tensor cores run it for free:
//
//   A (M x K) = the source bytes, M = 128 byte columns, K = 16 source rows.  A byte is exact in f16; dropped
//               into an f16 word as it is, it is the denormal b * 2^-24 (one PRMT per two bytes, no arithmetic).
//               Layout: MN-major, no swizzle (8x8 core matrices), written by the converter warps.
//   B (N x K) = the weights of the N = 32 or 48 output rows the 16 source rows can touch (host-built tiles,
//               PassPlan::band_tiles): w * 2^14 split into f16 hi + lo (22 bits of every weight), two MMAs
//               into the same f32 accumulator.
//   D (M x N) = f32 accumulators in TMEM: lane = byte column, column = output row.  Per 128-byte block a
//               ring of 4 groups x 16 output rows (64 columns); a finished group is read with tcgen05.ld,
//               written to the shared-memory intermediate tile (so: transposed to row-major), zeroed and
//               handed back.  Every MMA accumulates (no per-instruction "first write" bookkeeping).
//
// The HORIZONTAL pass stays on the CUDA cores (its input is the unclamped f32 intermediate, its cost is
// 6*C/ratio FMAs per source pixel): output-stationary, lane = intermediate row, half warp = x segment,
// window pixels held in a rotating register file for the uniform interior of a 2:1 resize, general tap loop
// otherwise.  clamp + round-half-away + u8 pack at its end, as in the reference.
//
// Warp roles of a CTA (256 threads, 2 CTAs per SM, 256 TMEM columns each):
//   warp 0    producer: 1-D TMA bulk copies of source rows (16 rows per stage) and of the weight tiles
//   warp 1    MMA issuer (one elected lane), owns the TMEM allocation
//   warps 2-3 converters: staged u8 rows -> f16 operand tiles (LDS.128, 8 PRMT, 2 STS.128 per 16 bytes)
//   warps 4-7 epilogue + horizontal pass (TMEM lane quarter = warp % 4)
// All hand-offs are mbarriers; the only block-wide barriers are in the prologue and the teardown.
#include <cuda_runtime.h>

#include <cstdint>

#include "device_types.hpp"
#include "launch.hpp"
#include "plan.hpp"

#ifndef IKC_BANDED_CONV
#define IKC_BANDED_CONV 0
#endif

namespace ikc {
namespace {

constexpr int kBlocks = 4;                         // 128-byte column blocks per strip
constexpr int kStripBytes = kBlocks * 128;         // staged source bytes per row
constexpr int kURowPitch = kStripBytes + 16;       // u8 staging row pitch: spreads 8 rows over all banks (LDS.128)
constexpr int kChunk = kBandChunk;                 // source rows per stage = K of one MMA
constexpr int kUStages = 3, kFStages = 2, kBStages = 4;
constexpr int kRing = 4;                           // accumulator groups per block held in TMEM
constexpr int kGroup = kBandGroup;                 // output rows per group
constexpr int kTmemCols = kBlocks * kRing * kGroup;  // 256
constexpr int kFBlockBytes = 128 * kChunk * 2;     // one f16 operand tile: 4096 B
constexpr int kFStageBytes = kBlocks * kFBlockBytes;
constexpr int kThreads = 256;
constexpr int kConvThreads = 64;
constexpr int kEpiThreads = 128;
constexpr int kSegs = 8;                           // horizontal segments: one per half warp of the epilogue warps
constexpr int kHeaderBytes = 256;
constexpr size_t kBandedMaxSmem = 113 * 1024;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(
            smem_addr(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_wait_parked(uint64_t* bar, uint32_t parity) {  // suspend-time hint: no busy polling
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(
            smem_addr(bar)),
        "r"(parity), "r"(20000u)
        : "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }

// ---- tcgen05 wrappers
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// Shared-memory operand descriptor, no swizzle: start address, leading / stride byte offsets (16-byte units).
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return uint64_t((addr >> 4) & 0x3fffu) | (uint64_t((lbo >> 4) & 0x3fffu) << 16) | (uint64_t((sbo >> 4) & 0x3fffu) << 32) |
           (uint64_t(1) << 46);
}
// Instruction descriptor: D = f32, A = B = f16, A MN-major, B K-major, M = 128, N = n.
__device__ __forceinline__ uint32_t instr_desc(uint32_t n) {
    return (1u << 4) | (1u << 15) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void mma_f16_acc(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(1u)
        : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(addr));
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_zero16(uint32_t addr) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(addr), "r"(0u)
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- pixel helpers (same quantisation as fused.cu: clamp, round half away as trunc(v + 0.5), saturating pack)
__device__ __forceinline__ uint32_t pack_pixel(float4 v_plus_half) {
    const int r = __float2int_rz(v_plus_half.x), g = __float2int_rz(v_plus_half.y);
    const int b = __float2int_rz(v_plus_half.z), a = __float2int_rz(v_plus_half.w);
    uint32_t hi, px;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(a), "r"(b), "r"(0));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(px) : "r"(g), "r"(r), "r"(hi));
    return px;
}
constexpr float kRoundBias = 0.5f;

template <int C>
__device__ __forceinline__ void store_pixel(uint8_t* dst_px, float4 v_plus_half, int co) {
    uint32_t w = pack_pixel(v_plus_half);
    if (C <= 2 && co >= 3) {  // grey (+ alpha) -> r, g, b (, a)
        const uint32_t grey = w & 0xffu;
        const uint32_t alpha = C == 2 ? (w >> 8) & 0xffu : 0xffu;
        w = grey * 0x010101u | (alpha << 24);
    } else if (C == 3) {
        w |= 0xff000000u;     // rgb -> rgba: opaque
    }
    if (co == 4) {
        *reinterpret_cast<uint32_t*>(dst_px) = w;
    } else {
        dst_px[0] = uint8_t(w);
        if (co >= 2) dst_px[1] = uint8_t(w >> 8);
        if (co >= 3) dst_px[2] = uint8_t(w >> 16);
    }
}

// One intermediate pixel (C floats at `p`) as a float4; missing channels read as zero.
template <int C>
__device__ __forceinline__ float4 load_px(const float* p) {
    if (C == 4) return *reinterpret_cast<const float4*>(p);
    if (C == 3) return make_float4(p[0], p[1], p[2], 0.0f);
    if (C == 2) { const float2 v = *reinterpret_cast<const float2*>(p); return make_float4(v.x, v.y, 0.0f, 0.0f); }
    return make_float4(p[0], 0.0f, 0.0f, 0.0f);
}

// Row pitch of the intermediate tile, in floats: an odd number of pixel-sized (16 / 8 / 4 byte) units, so the 16 rows
// a half warp reads with one LDS.128 / .64 / .32 fall into different banks.
__host__ __device__ constexpr int tmp_pitch_floats(int channels) { return channels == 4 ? kStripBytes + 4 : channels == 2 ? kStripBytes + 2 : kStripBytes + 1; }

constexpr bool kConv = IKC_BANDED_CONV != 0;

}  // namespace

// Shared memory: [mbarriers | u8 staging ring (3 x 16 rows x 528 B) | f16 operand ring (2 x 4 blocks x 4 KB) |
//                weight-tile ring (4 x band_n * 64 B) | horizontal weights of the strip (float2 per output and tap)
//                | (left, right) of the strip's outputs | intermediate tile: 16 rows x tmp_pitch floats]
template <int C, bool CONV>
__global__ void __launch_bounds__(kThreads, 2)
banded_kernel(const DevJob* __restrict__ jobs, const WorkItem* __restrict__ items, const BandGeom geom) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int kTmpPitch = tmp_pitch_floats(C);

    uint64_t* const bars = reinterpret_cast<uint64_t*>(smem);
    uint64_t* const u_full = bars;                 // [kUStages] source rows landed (tx bytes)
    uint64_t* const u_empty = u_full + kUStages;   // [kUStages] every converter thread has read the stage
    uint64_t* const f_full = u_empty + kUStages;   // [kFStages] every converter thread has written the stage
    uint64_t* const f_empty = f_full + kFStages;   // [kFStages] the MMAs that read the stage have completed
    uint64_t* const b_full = f_empty + kFStages;   // [kBStages] weight tile landed
    uint64_t* const b_empty = b_full + kBStages;   // [kBStages] the MMAs that read it have completed
    uint64_t* const t_full = b_empty + kBStages;   // [kRing]    every MMA into the group has completed
    uint64_t* const t_empty = t_full + kRing;      // [kRing]    the 4 epilogue warps have drained and zeroed the group
    uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(smem + 240);
    uint8_t* const ustage = smem + kHeaderBytes;
    uint8_t* const fstage = ustage + kUStages * kChunk * kURowPitch;
    uint8_t* const bstage = fstage + kFStages * kFStageBytes;
    const uint32_t btile = uint32_t(geom.band_n) * 64u;  // bytes of one chunk's hi + lo tiles
    float2* const hw = reinterpret_cast<float2*>(bstage + kBStages * btile);
    int2* const hlr = reinterpret_cast<int2*>(hw + ((geom.hw_pairs + 1) & ~1));
    float* const tmp = reinterpret_cast<float*>(hlr + ((geom.max_out + 1) & ~1));

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;

    const WorkItem it = items[blockIdx.x];
    const DevJob* __restrict__ J = jobs + it.job;
    const int ox0 = it.ox0, ox1 = it.ox1, oy0 = it.oy0, oy1 = it.oy1;
    const int32_t* __restrict__ hleft = J->h.left;
    const int32_t* __restrict__ hright = J->h.right;
    const int32_t* __restrict__ gbase = J->v.band_gbase;

    // Strip geometry along x: source bytes [b0, b0 + nb), 16-byte aligned at both ends.
    const int xl = __ldg(hleft + ox0);
    const int xr = __ldg(hright + ox1 - 1);
    const int row_bytes = int(J->sw) * C;
    const int b0 = (xl * C) & ~15;
    const int b1 = min((xr * C + 15) & ~15, (row_bytes + 15) & ~15);
    const int nb = b1 - b0;
    const int nblk = (nb + 127) >> 7;
    // Chunk geometry along y: source rows [y_first, y_last) -> chunks [k0, k1] of the pass's global chunk grid.
    const int y_first = __ldg(J->v.left + oy0);
    const int y_last = __ldg(J->v.right + oy1 - 1);
    const int k0 = y_first / kChunk, k1 = (y_last - 1) / kChunk;
    const int nchunks = k1 - k0 + 1;
    const int NG = geom.band_n / kGroup;           // groups one chunk's window spans
    const int g0 = __ldg(gbase + k0);              // first group any MMA of this item touches
    const int g_end = __ldg(gbase + k1) + NG;      // one past the last

    if (tid == 0) {
        for (int s = 0; s < kUStages; ++s) { mbar_init(u_full + s, 1); mbar_init(u_empty + s, kConvThreads); }
        for (int s = 0; s < kFStages; ++s) { mbar_init(f_full + s, kConvThreads); mbar_init(f_empty + s, 1); }
        for (int s = 0; s < kBStages; ++s) { mbar_init(b_full + s, 1); mbar_init(b_empty + s, 1); }
        for (int s = 0; s < kRing; ++s) { mbar_init(t_full + s, 1); mbar_init(t_empty + s, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: 256 columns of this SM's 512 (two CTAs are resident)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(tmem_slot)), "n"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // Horizontal tables of the strip: (left, right) and the weights (x 2^10, duplicated for FFMA2) of every output.
    const int n_out = ox1 - ox0;
    const int hstride = J->h.stride;
    for (int i = tid; i < n_out; i += kThreads) hlr[i] = make_int2(__ldg(hleft + ox0 + i), __ldg(hright + ox0 + i));
    {
        const float* __restrict__ wsrc = J->h.w + size_t(ox0) * hstride;
        for (int i = tid; i < n_out * hstride; i += kThreads) {
            const float w = __ldg(wsrc + i) * kBandScaleH;
            hw[i] = make_float2(w, w);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------------------ producer
        if (lane == 0) {
            const uint8_t* const gsrc = J->src + b0;
            const size_t src_pitch = J->src_pitch;
            const uint8_t* const tiles = reinterpret_cast<const uint8_t*>(J->v.band_tiles);
            for (int i = 0; i < nchunks; ++i) {
                const int k = k0 + i;
                const int su = i % kUStages, sb = i % kBStages;
                mbar_wait_parked(u_empty + su, ((i / kUStages) & 1) ^ 1);
                const int r0 = max(k * kChunk, y_first), r1 = min((k + 1) * kChunk, y_last);  // rows outside carry zero weight for every live output
                mbar_expect_tx(u_full + su, uint32_t(r1 - r0) * uint32_t(nb));
                for (int r = r0; r < r1; ++r)
                    bulk_load(ustage + (su * kChunk + (r - k * kChunk)) * kURowPitch, gsrc + size_t(r) * src_pitch, uint32_t(nb), u_full + su);
                mbar_wait_parked(b_empty + sb, ((i / kBStages) & 1) ^ 1);
                mbar_expect_tx(b_full + sb, btile);
                bulk_load(bstage + sb * btile, tiles + size_t(k) * btile, btile, b_full + sb);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------------------------ MMA issuer
        int acquired = g0;   // groups [g0, acquired) belong to the MMAs (zeroed by the epilogue warps)
        int completed = g0;  // groups [g0, completed) have been committed to the epilogue
        const uint32_t f_addr = smem_addr(fstage), b_addr = smem_addr(bstage);
        int gb_next = g0;
        for (int i = 0; i < nchunks; ++i) {
            const int k = k0 + i;
            const int gb = gb_next;
            gb_next = (i + 1 < nchunks) ? __ldg(gbase + k + 1) : 0;
            while (acquired < gb + NG) {
                const int rel = acquired - g0;
                mbar_wait(t_empty + (rel & (kRing - 1)), (rel / kRing) & 1);
                ++acquired;
            }
            const int sf = i % kFStages, sb = i % kBStages;
            mbar_wait(f_full + sf, (i / kFStages) & 1);
            mbar_wait(b_full + sb, (i / kBStages) & 1);
            tc_fence_after();
            if (lane == 0) {
                const int s0 = (gb - g0) & (kRing - 1);
                const uint32_t n1 = uint32_t(min(NG, kRing - s0) * kGroup), n2 = uint32_t(geom.band_n) - n1;
                const uint32_t id1 = instr_desc(n1), id2 = instr_desc(n2);
                const uint32_t b_lbo = uint32_t(geom.band_n) * 16u;  // between the two halves of the 16 source rows
#pragma unroll 1
                for (int b = 0; b < nblk; ++b) {
                    const uint64_t a_desc = smem_desc(f_addr + uint32_t(sf * kFStageBytes + b * kFBlockBytes), 2048u, 128u);
                    const uint32_t d_col = tmem + uint32_t(b * kRing * kGroup);
#pragma unroll
                    for (int part = 0; part < 2; ++part) {  // hi weights, then what f16 rounding lost of them
                        const uint32_t bt = b_addr + uint32_t(sb) * btile + uint32_t(part) * (btile >> 1);
                        mma_f16_acc(d_col + uint32_t(s0 * kGroup), a_desc, smem_desc(bt, b_lbo, 128u), id1);
                        if (n2) mma_f16_acc(d_col, a_desc, smem_desc(bt + (n1 >> 3) * 128u, b_lbo, 128u), id2);  // the window wraps around the ring
                    }
                }
                tc_commit(f_empty + sf);
                tc_commit(b_empty + sb);
                const int final_below = (i + 1 < nchunks) ? gb_next : acquired;  // groups below it get no more contributions
                for (; completed < final_below; ++completed) tc_commit(t_full + ((completed - g0) & (kRing - 1)));
            }
            __syncwarp();
        }
    } else if (warp < 4) {
        // ------------------------------------------------------------------------------ converters
        const int ct = tid - 64;
        const int npieces = nb >> 4;
        for (int i = 0; i < nchunks; ++i) {
            const int su = i % kUStages, sf = i % kFStages;
            mbar_wait(u_full + su, (i / kUStages) & 1);
            mbar_wait(f_empty + sf, ((i / kFStages) & 1) ^ 1);
            const uint8_t* const ubase = ustage + su * kChunk * kURowPitch;
            uint8_t* const fbase = fstage + sf * kFStageBytes;
#pragma unroll
            for (int t = 0; t < (kChunk * (kStripBytes / 16)) / kConvThreads; ++t) {
                const int task = t * kConvThreads + ct;
                const int r8 = task & 7, piece = (task >> 3) & 31, rg = task >> 8;  // 8 lanes: 8 rows of one 16-byte piece
                if (piece < npieces) {
                    const uint4 v = *reinterpret_cast<const uint4*>(ubase + (rg * 8 + r8) * kURowPitch + piece * 16);
                    uint4 lo, hi;  // 16 bytes -> 16 f16 denormals: bytes 0-7, bytes 8-15
                    lo.x = __byte_perm(v.x, 0u, 0x4140); lo.y = __byte_perm(v.x, 0u, 0x4342);
                    lo.z = __byte_perm(v.y, 0u, 0x4140); lo.w = __byte_perm(v.y, 0u, 0x4342);
                    hi.x = __byte_perm(v.z, 0u, 0x4140); hi.y = __byte_perm(v.z, 0u, 0x4342);
                    hi.z = __byte_perm(v.w, 0u, 0x4140); hi.w = __byte_perm(v.w, 0u, 0x4342);
                    // element (m, k) of a block lives at (k / 8) * 2048 + (m / 8) * 128 + (k % 8) * 16 + (m % 8) * 2
                    uint8_t* const d = fbase + (piece >> 3) * kFBlockBytes + rg * 2048 + (piece & 7) * 256 + r8 * 16;
                    *reinterpret_cast<uint4*>(d) = lo;
                    *reinterpret_cast<uint4*>(d + 128) = hi;
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
            mbar_arrive(f_full + sf);
            mbar_arrive(u_empty + su);
        }
    } else {
        // ------------------------------------------------------------------------------ epilogue + horizontal pass
        const int q = warp & 3;                                  // TMEM lane quarter this warp may touch
        const uint32_t tlane = tmem + (uint32_t(q * 32) << 16);
        // hand all ring slots to the MMA warp, zeroed
        for (int c = 0; c < kTmemCols; c += 16) tmem_zero16(tlane + uint32_t(c));
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0)
            for (int s = 0; s < kRing; ++s) mbar_arrive(t_empty + s);

        const int hrow = lane & 15;
        const int seg = 2 * q + (lane >> 4);
        const int per = (n_out + kSegs - 1) / kSegs;
        const int os = ox0 + seg * per;
        const int oe = min(os + per, ox1);
        const int CO = CONV ? J->out_channels : C;
        uint8_t* const dst_base = J->dst;
        const size_t dst_pitch = J->dst_pitch;
        const float* const my_row = tmp + hrow * kTmpPitch - b0;   // indexed by source byte column x * C + c
        // uniform interior of a 2:1 pass: every output of the segment has the same 12 weights and starts 2 pixels
        // after its predecessor
        const bool uni2 = os < oe && J->h.uni_step == 2 && hstride == 12 && os >= J->h.uni_lo && oe <= J->h.uni_hi;

        for (int g = g0; g < g_end; ++g) {
            const int rel = g - g0;
            const int slot = rel & (kRing - 1);
            mbar_wait(t_full + slot, (rel / kRing) & 1);
            tc_fence_after();
            const int row0 = g * kGroup;
            const bool live = row0 < oy1 && row0 + kGroup > oy0;
            if (live) {  // TMEM -> registers -> intermediate tile (lane = byte column, so consecutive floats)
                float* const tcol = tmp + q * 32 + lane;
#pragma unroll 1
                for (int b = 0; b < nblk; ++b) {
                    float v[16];
                    tmem_ld16(tlane + uint32_t(b * kRing * kGroup + slot * kGroup), v);
                    tmem_ld_wait();
#pragma unroll
                    for (int r = 0; r < 16; ++r) tcol[r * kTmpPitch + b * 128] = v[r];
                }
            }
            for (int b = 0; b < nblk; ++b) tmem_zero16(tlane + uint32_t(b * kRing * kGroup + slot * kGroup));
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(t_empty + slot);
            if (!live) continue;
            epi_barrier();  // the whole tile is in shared memory

            const int orow = row0 + hrow;
            const bool row_live = orow >= oy0 && orow < oy1;
            uint8_t* const my_dst = dst_base + size_t(orow) * dst_pitch;
            if (os < oe) {
                if (uni2) {
                    float2 w[12];
#pragma unroll
                    for (int t = 0; t < 12; ++t) w[t] = hw[(os - ox0) * 12 + t];
                    const float* px = my_row + hlr[os - ox0].x * C;  // first pixel of the first window
                    float4 p[12];
#pragma unroll
                    for (int t = 0; t < 10; ++t) p[t] = load_px<C>(px + t * C);
                    px += 10 * C;
                    for (int o = os; o < oe; o += 6) {
#pragma unroll
                        for (int u = 0; u < 6; ++u) {
                            if (o + u < oe) {
                                p[(2 * u + 10) % 12] = load_px<C>(px);
                                p[(2 * u + 11) % 12] = load_px<C>(px + C);
                                px += 2 * C;
                                float2 a01 = make_float2(kRoundBias, kRoundBias), a23 = a01;
                                float2 c01 = make_float2(0.0f, 0.0f), c23 = c01;
#pragma unroll
                                for (int t = 0; t < 12; t += 2) {
                                    const float4 e = p[(2 * u + t) % 12], f = p[(2 * u + t + 1) % 12];
                                    a01 = __ffma2_rn(w[t], make_float2(e.x, e.y), a01);
                                    a23 = __ffma2_rn(w[t], make_float2(e.z, e.w), a23);
                                    c01 = __ffma2_rn(w[t + 1], make_float2(f.x, f.y), c01);
                                    c23 = __ffma2_rn(w[t + 1], make_float2(f.z, f.w), c23);
                                }
                                if (row_live) store_pixel<C>(my_dst + size_t(o + u) * CO, make_float4(a01.x + c01.x, a01.y + c01.y, a23.x + c23.x, a23.y + c23.y), CO);
                            }
                        }
                    }
                } else {
                    for (int o = os; o < oe; ++o) {
                        const int2 lr = hlr[o - ox0];
                        const float2* wrow = hw + (o - ox0) * hstride;
                        const float* px = my_row + lr.x * C;
                        const int n = lr.y - lr.x;
                        float2 a01 = make_float2(kRoundBias, kRoundBias), a23 = a01;
                        float2 c01 = make_float2(0.0f, 0.0f), c23 = c01;
                        int t = 0;
                        for (; t + 1 < n; t += 2) {
                            const float4 e = load_px<C>(px + t * C), f = load_px<C>(px + (t + 1) * C);
                            const float2 we = wrow[t], wf = wrow[t + 1];
                            a01 = __ffma2_rn(we, make_float2(e.x, e.y), a01);
                            a23 = __ffma2_rn(we, make_float2(e.z, e.w), a23);
                            c01 = __ffma2_rn(wf, make_float2(f.x, f.y), c01);
                            c23 = __ffma2_rn(wf, make_float2(f.z, f.w), c23);
                        }
                        if (t < n) {
                            const float4 e = load_px<C>(px + t * C);
                            const float2 we = wrow[t];
                            a01 = __ffma2_rn(we, make_float2(e.x, e.y), a01);
                            a23 = __ffma2_rn(we, make_float2(e.z, e.w), a23);
                        }
                        if (row_live) store_pixel<C>(my_dst + size_t(o) * CO, make_float4(a01.x + c01.x, a01.y + c01.y, a23.x + c23.x, a23.y + c23.y), CO);
                    }
                }
            }
            __syncwarp();
            epi_barrier();  // the tile may be overwritten
        }
    }

    // ---------------------------------------------------------------------------------- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols));
    }
}

// ---- launcher ---------------------------------------------------------------------------------

#if !IKC_BANDED_CONV
size_t banded_smem_bytes(int channels, const BandGeom& g) {
    const size_t tmp_pitch = size_t(tmp_pitch_floats(channels));
    return size_t(kHeaderBytes) + size_t(kUStages) * kChunk * kURowPitch + size_t(kFStages) * kFStageBytes +
           size_t(kBStages) * size_t(g.band_n) * 64 + size_t((g.hw_pairs + 1) & ~1) * sizeof(float2) +
           size_t((g.max_out + 1) & ~1) * sizeof(int2) + size_t(kGroup) * tmp_pitch * sizeof(float);
}
size_t banded_max_smem() { return kBandedMaxSmem; }
int banded_max_src_bytes() { return kStripBytes; }
int banded_group_rows() { return kGroup; }
bool banded_supported(int channels, int band_n) { return channels >= 1 && channels <= 4 && (band_n == 32 || band_n == 48); }
#endif

template <int C>
static cudaError_t launch_one(const DevJob* jobs, const WorkItem* items, const BandGeom& geom, cudaStream_t stream) {
    const size_t smem = banded_smem_bytes(C, geom);
    if (smem > kBandedMaxSmem) return cudaErrorInvalidValue;
    // always the planner-wide maximum: the attribute is shared by every thread launching on this device
    cudaError_t e = cudaFuncSetAttribute(banded_kernel<C, kConv>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kBandedMaxSmem));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(banded_kernel<C, kConv>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    banded_kernel<C, kConv><<<geom.n_items, kThreads, smem, stream>>>(jobs, items, geom);
    return cudaGetLastError();
}

#if IKC_BANDED_CONV
cudaError_t launch_banded_conv(int channels, const DevJob* jobs, const WorkItem* items, const BandGeom& geom, cudaStream_t stream) {
#else
cudaError_t launch_banded_conv(int channels, const DevJob* jobs, const WorkItem* items, const BandGeom& geom, cudaStream_t stream);  // banded_conv.cu

cudaError_t launch_banded(int channels, bool convert, const DevJob* jobs, const WorkItem* items, const BandGeom& geom,
                          cudaStream_t stream) {
    if (convert) return launch_banded_conv(channels, jobs, items, geom, stream);
#endif
    switch (channels) {
        case 1: return launch_one<1>(jobs, items, geom, stream);
        case 2: return launch_one<2>(jobs, items, geom, stream);
        case 3: return launch_one<3>(jobs, items, geom, stream);
        case 4: return launch_one<4>(jobs, items, geom, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace ikc
