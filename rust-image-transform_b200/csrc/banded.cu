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
//   warp 0    producer: one 2-D TMA box (16 source rows x 512 bytes) and one bulk copy of the weight tiles per stage
//   warp 1    MMA issuer (one elected lane), owns the TMEM allocation
//   warps 2-3 converters: staged u8 rows -> f16 operand tiles (LDS.128, 8 PRMT, 2 STS.128 per 16 bytes)
//   warps 4-7 epilogue + horizontal pass (TMEM lane quarter = warp % 4)
// All hand-offs are mbarriers; the only block-wide barriers are in the prologue and the teardown.
#include <cuda_runtime.h>

#include <cstdint>

#include "banded_common.cuh"
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
constexpr int kURowPitch = kStripBytes;            // u8 staging rows are dense: one 2-D TMA box (16 rows x 512 bytes) per stage
constexpr int kChunk = kBandChunk;                 // source rows per stage = K of one MMA
constexpr int kUStages = 3, kFStages = 2, kBStages = 4;
constexpr int kRing = 4;                           // accumulator groups per block held in TMEM
constexpr int kGroup = kBandGroup;                 // output rows per group
constexpr int kTmemCols = kBlocks * kRing * kGroup;  // 256
constexpr int kFBlockBytes = 128 * kChunk * 2;     // one f16 operand tile: 4096 B
constexpr int kFStageBytes = kBlocks * kFBlockBytes;
constexpr int kThreads = 256;
constexpr int kConvThreads = 64;
constexpr int kSegs = 8;                           // horizontal segments: one per half warp of the epilogue warps
constexpr int kHeaderBytes = 256;
constexpr size_t kBandedMaxSmem = 113 * 1024;
constexpr int kRegsIo = 72, kRegsEpi = 184;        // 128 x 72 + 128 x 184 = 256 x 128
constexpr int kTmpPad = 64;                        // floats after the intermediate tile: the look-ahead loads of the last row end here

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

    // Register budget by role (2 CTAs x 256 threads x 128 registers fill the SM's file): the producer / MMA /
    // converter warpgroup gives registers up, the epilogue warpgroup -- whose horizontal pass keeps a 12-pixel
    // window, its look-ahead and 12 weight pairs in registers -- takes them.
    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsIo));
    if (warp == 0) {
        // ------------------------------------------------------------------------------ producer
        if (lane == 0) {
            const void* const src_map = J->src_map;
            asm volatile("prefetch.tensormap [%0];" ::"l"(src_map) : "memory");
            const uint8_t* const tiles = reinterpret_cast<const uint8_t*>(J->v.band_tiles);
            for (int i = 0; i < nchunks; ++i) {
                const int k = k0 + i;
                const int su = i % kUStages, sb = i % kBStages;
                mbar_wait_parked(u_empty + su, ((i / kUStages) & 1) ^ 1);
                // one box = the chunk's 16 rows x 512 bytes of the strip; words past the raster's pitch or rows read as zero
                mbar_expect_tx(u_full + su, uint32_t(kChunk * kStripBytes));
                tma_load_2d(ustage + su * kChunk * kURowPitch, src_map, b0 >> 2, k * kChunk, u_full + su);
                mbar_wait_parked(b_empty + sb, ((i / kBStages) & 1) ^ 1);
                mbar_expect_tx(b_full + sb, btile);
                bulk_load(bstage + sb * btile, tiles + size_t(k) * btile, btile, b_full + sb);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------------------------ MMA issuer
        // One lane runs the whole role (waits included): tcgen05.mma / commit are single-thread instructions, and a
        // converged 32-lane loop around them costs an election per instruction.  Everything an MMA needs is a 32-bit
        // add away from values computed once: the descriptors' high words are constant (stride offset 128, version 1)
        // and only the 14-bit address field of the low words moves.
        if (lane == 0) {
            int acquired = g0;   // groups [g0, acquired) belong to the MMAs (zeroed by the epilogue warps)
            int completed = g0;  // groups [g0, completed) have been committed to the epilogue
            constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);
            const uint32_t a_lo0 = ((smem_addr(fstage) >> 4) & 0x3fffu) | ((2048u >> 4) << 16);
            const uint32_t b_lo0 = ((smem_addr(bstage) >> 4) & 0x3fffu) | (((uint32_t(geom.band_n) * 16u) >> 4) << 16);
            const uint32_t band_n = uint32_t(geom.band_n);
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
                const int s0 = (gb - g0) & (kRing - 1);
                const uint32_t n1 = uint32_t(min(NG, kRing - s0) * kGroup), n2 = band_n - n1;  // n2 > 0: the window wraps around the ring
                const uint32_t id1 = instr_desc(n1), id2 = instr_desc(n2);
                const uint32_t a_lo = a_lo0 + uint32_t(sf) * (kFStageBytes >> 4);
                const uint32_t bh_lo = b_lo0 + uint32_t(sb) * (btile >> 4);   // hi weights
                const uint32_t bl_lo = bh_lo + (btile >> 5);                  // what f16 rounding lost of them
                const uint32_t wrap = (n1 >> 3) * (128u >> 4);               // the wrapped part's first output row in the tile
                const uint32_t d1 = tmem + uint32_t(s0 * kGroup);
#pragma unroll
                for (int b = 0; b < kBlocks; ++b) {
                    if (b < nblk) {
                        const uint64_t a_desc = make_u64(a_lo + uint32_t(b) * (kFBlockBytes >> 4), kDescHi);
                        const uint32_t dc = uint32_t(b * kRing * kGroup);
                        mma_f16_acc(d1 + dc, a_desc, make_u64(bh_lo, kDescHi), id1);
                        mma_f16_acc(d1 + dc, a_desc, make_u64(bl_lo, kDescHi), id1);
                        if (n2) {
                            mma_f16_acc(tmem + dc, a_desc, make_u64(bh_lo + wrap, kDescHi), id2);
                            mma_f16_acc(tmem + dc, a_desc, make_u64(bl_lo + wrap, kDescHi), id2);
                        }
                    }
                }
                tc_commit(f_empty + sf);
                tc_commit(b_empty + sb);
                const int final_below = (i + 1 < nchunks) ? gb_next : acquired;  // groups below it get no more contributions
                for (; completed < final_below; ++completed) tc_commit(t_full + ((completed - g0) & (kRing - 1)));
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------------------ converters
        // Thread task t (0..7) of a stage: row (t / 4) * 8 + r8 of block t % 4, 16-byte piece (pc + r8) % 8 of the block's
        // 128 bytes.  The skew makes both sides conflict-free: the 8 lanes of a quarter warp read 8 rows (dense 512-byte
        // pitch) at 8 different 16-byte offsets, and write row r8 (16 bytes) of 8 different core matrices.
        const int ct = tid - 64;
        const int r8 = ct & 7, pc = ((ct >> 3) + r8) & 7;
        const int u_off = r8 * kURowPitch + pc * 16;
        const int f_off = pc * 256 + r8 * 16;  // element (m, k) of a block: (k / 8) * 2048 + (m / 8) * 128 + (k % 8) * 16 + (m % 8) * 2
        for (int i = 0; i < nchunks; ++i) {
            const int su = i % kUStages, sf = i % kFStages;
            mbar_wait(u_full + su, (i / kUStages) & 1);
            const uint8_t* const ubase = ustage + su * kChunk * kURowPitch + u_off;
            uint4 v[8];
#pragma unroll
            for (int t = 0; t < 8; ++t)
                if ((t & 3) < nblk) v[t] = *reinterpret_cast<const uint4*>(ubase + (t >> 2) * 8 * kURowPitch + (t & 3) * 128);
            mbar_wait(f_empty + sf, ((i / kFStages) & 1) ^ 1);
            uint8_t* const fbase = fstage + sf * kFStageBytes + f_off;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                if ((t & 3) < nblk) {
                    uint4 lo, hi;  // 16 bytes -> 16 f16 denormals: bytes 0-7, bytes 8-15
                    lo.x = __byte_perm(v[t].x, 0u, 0x4140); lo.y = __byte_perm(v[t].x, 0u, 0x4342);
                    lo.z = __byte_perm(v[t].y, 0u, 0x4140); lo.w = __byte_perm(v[t].y, 0u, 0x4342);
                    hi.x = __byte_perm(v[t].z, 0u, 0x4140); hi.y = __byte_perm(v[t].z, 0u, 0x4342);
                    hi.z = __byte_perm(v[t].w, 0u, 0x4140); hi.w = __byte_perm(v[t].w, 0u, 0x4342);
                    uint8_t* const d = fbase + (t & 3) * kFBlockBytes + (t >> 2) * 2048;
                    *reinterpret_cast<uint4*>(d) = lo;
                    *reinterpret_cast<uint4*>(d + 128) = hi;
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
            mbar_arrive(f_full + sf);
            mbar_arrive(u_empty + su);
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsEpi));
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
        float2 uw[12];  // the stretch's 12 tap weights (duplicated pairs)
#pragma unroll
        for (int t = 0; t < 12; ++t) uw[t] = uni2 ? hw[(os - ox0) * 12 + t] : make_float2(0.0f, 0.0f);

        for (int g = g0; g < g_end; ++g) {
            const int rel = g - g0;
            const int slot = rel & (kRing - 1);
            mbar_wait(t_full + slot, (rel / kRing) & 1);
            tc_fence_after();
            const int row0 = g * kGroup;
            const bool live = row0 < oy1 && row0 + kGroup > oy0;
            if (live) {  // TMEM -> registers -> intermediate tile (lane = byte column, so consecutive floats)
                float* const tcol = tmp + q * 32 + lane;
                float v[kBlocks][16];
#pragma unroll
                for (int b = 0; b < kBlocks; ++b)
                    if (b < nblk) tmem_ld16(tlane + uint32_t(b * kRing * kGroup + slot * kGroup), v[b]);
                tmem_ld_wait();
#pragma unroll
                for (int b = 0; b < kBlocks; ++b) {
                    if (b < nblk) {
#pragma unroll
                        for (int r = 0; r < 16; ++r) tcol[r * kTmpPitch + b * 128] = v[b][r];
                    }
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
                    // Blocks of 8 outputs as straight-line code: the 26 pixels their windows span are loaded at once
                    // (output j of the block reads pixels 2j .. 2j + 11), then 8 x 24 independent-enough FFMA2 chains.
                    // Reads past the segment's last window stay inside the tile's padding and are never stored.
                    for (int o = os; o < oe; o += 8) {
                        const float* px = my_row + hlr[o - ox0].x * C;
                        float4 p[26];
#pragma unroll
                        for (int t = 0; t < 26; ++t) p[t] = load_px<C>(px + t * C);
                        uint32_t word[8];
#pragma unroll
                        for (int j = 0; j < 8; j += 2) {  // two outputs at a time: eight accumulation chains in flight
                            float2 a01 = make_float2(kRoundBias, kRoundBias), a23 = a01, c01 = make_float2(0.0f, 0.0f), c23 = c01;
                            float2 d01 = a01, d23 = a01, e01 = c01, e23 = c01;
#pragma unroll
                            for (int t = 0; t < 12; t += 2) {
                                const float4 x0 = p[2 * j + t], x1 = p[2 * j + t + 1], y0 = p[2 * j + 2 + t], y1 = p[2 * j + 3 + t];
                                a01 = __ffma2_rn(uw[t], make_float2(x0.x, x0.y), a01);
                                a23 = __ffma2_rn(uw[t], make_float2(x0.z, x0.w), a23);
                                d01 = __ffma2_rn(uw[t], make_float2(y0.x, y0.y), d01);
                                d23 = __ffma2_rn(uw[t], make_float2(y0.z, y0.w), d23);
                                c01 = __ffma2_rn(uw[t + 1], make_float2(x1.x, x1.y), c01);
                                c23 = __ffma2_rn(uw[t + 1], make_float2(x1.z, x1.w), c23);
                                e01 = __ffma2_rn(uw[t + 1], make_float2(y1.x, y1.y), e01);
                                e23 = __ffma2_rn(uw[t + 1], make_float2(y1.z, y1.w), e23);
                            }
                            a01 = __fadd2_rn(a01, c01); a23 = __fadd2_rn(a23, c23);
                            d01 = __fadd2_rn(d01, e01); d23 = __fadd2_rn(d23, e23);
                            word[j] = pack_pixel(make_float4(a01.x, a01.y, a23.x, a23.y));
                            word[j + 1] = pack_pixel(make_float4(d01.x, d01.y, d23.x, d23.y));
                        }
                        if (row_live) {
                            uint8_t* const d = my_dst + size_t(o) * CO;
                            if (C == 4 && !CONV && o + 8 <= oe && (reinterpret_cast<uintptr_t>(d) & 15) == 0) {
                                *reinterpret_cast<uint4*>(d) = make_uint4(word[0], word[1], word[2], word[3]);
                                *reinterpret_cast<uint4*>(d + 16) = make_uint4(word[4], word[5], word[6], word[7]);
                            } else {
#pragma unroll
                                for (int j = 0; j < 8; ++j)
                                    if (o + j < oe) store_word<C>(d + j * CO, word[j], CO);
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
           size_t((g.max_out + 1) & ~1) * sizeof(int2) + (size_t(kGroup) * tmp_pitch + kTmpPad) * sizeof(float);
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
