// banded8u.cu -- exact 2x upscales of Rgb8 / Rgba8 rasters (BASELINE config 4: 1920x1080 -> 3840x2160 CatmullRom) with the
// vertical pass on the tensor cores and the horizontal pass from registers: sm_100a.
//
// The layout of banded8t.cu (accumulator lane = OUTPUT row) applied to an upscale:
//   A (M x K) = the vertical weights of a band of 128 output rows (64 + taps - 1 source rows: 3 chunks of 32), two signed
//               base-256 digits, host-built K-major s8 tiles (plan.cpp, the Band8T form of a 2x upscale pass);
//   B (N x K) = the raw source bytes of 32 pixels (N = 96 byte columns for Rgb8, 128 for Rgba8), K = 32 source rows, as
//               one 2-D TMA box lands (32 rows x 128 bytes, 128-byte swizzle);
//   D (M x N) = s32 accumulators in TMEM, lane = output row, column = byte column (hi digits at 0, lo digits at 128).
// An epilogue thread owns one output row.  It reads its intermediate pixels from TMEM eight at a time, keeps the last
// `T` of them (the frame of the pass: both outputs of source pixel k read source pixels k + off .. k + off + T - 1) in
// registers and, for every new pixel, finishes one PAIR of outputs with C x T packed FMAs (weight pair (even, odd) x
// broadcast sample).  Eight source pixels make 16 output pixels = 48 (Rgb8) or 64 (Rgba8) bytes: whole 16-byte stores.
// The block's boxes are fetched `adv` pixels to the right of the pairs it completes (adv = off + T - 1: the last pixel a
// pair needs), so the output stream of a block starts on a 16-byte boundary for any frame.
// Interior pixels take the pass's one set of weight pairs from registers; pixels near the left / right border (clamped
// windows, framed with zero weights by the planner) load theirs.
//
// One CTA per SM, 384 threads: warp 0 producer (TMA), warp 1 MMA issuer, warps 4-11 two epilogue quads that stream
// through the two halves of the item's columns, each with its own accumulator tile (as in banded8t.cu).
#include <cuda_runtime.h>

#include <cstdint>

#include "banded_common.cuh"
#include "device_types.hpp"
#include "launch.hpp"
#include "plan.hpp"

namespace ikc {
namespace {

constexpr int kURows = kBand8TRows;                      // output rows per band = TMEM lanes = M
constexpr int kUChunk = kBand8Chunk;                     // source rows per MMA = K
constexpr int kUBoxBytes = 128 * kUChunk;                // one box: 32 rows x 128 bytes, swizzled
constexpr int kUMaxChunks = 4;                           // weight tiles per band and digit the kernel has room for
constexpr int kUBlockStage = kUMaxChunks * kUBoxBytes;   // the boxes of one block
constexpr int kUStages = 4;                              // ring of source blocks
constexpr int kUWTile = kURows * kUChunk;                // one weight tile (band, chunk, digit)
constexpr int kUStreams = 2;
constexpr int kUTileCols = 256;                          // TMEM columns of a tile: hi digits at 0, lo digits at 128
constexpr int kUTmemCols = kUStreams * kUTileCols;
constexpr int kUThreads = 128 + kUStreams * 128;
constexpr int kUHeaderBytes = 1024;
constexpr int kURegsIo = 40, kURegsEpi = 232;
constexpr int kUBlockPx = 32;                            // source pixels per block
constexpr int kUUnitPx = 8;                              // source pixels per straight-line unit

__device__ __forceinline__ uint32_t srcU_desc_hi() { return (1024u >> 4) | (1u << 14) | (2u << 29); }
// D = s32, A = s8 (K-major), B = u8 (MN-major), M = 128, N = n.
__device__ __forceinline__ uint32_t instr_desc_i8u(uint32_t n) {
    return (2u << 4) | (1u << 7) | (0u << 10) | (1u << 16) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void mma_i8u(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// N consecutive TMEM columns of this thread's lane.
template <int N>
__device__ __forceinline__ void tmem_ldu(uint32_t addr, int* v);
template <>
__device__ __forceinline__ void tmem_ldu<8>(uint32_t addr, int* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(addr));
}
template <>
__device__ __forceinline__ void tmem_ldu<16>(uint32_t addr, int* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(addr));
}
// the 8 * C columns of one unit (hi or lo digits)
template <int C>
__device__ __forceinline__ void tmem_ld_unit(uint32_t addr, int (&v)[8 * C]) {
    if (C == 4) {
        tmem_ldu<16>(addr, v);
        tmem_ldu<16>(addr + 16, v + 16);
    } else {
        tmem_ldu<16>(addr, v);
        tmem_ldu<8>(addr + 16, v + 16);
    }
}

// One stream: pairs (source pixels k: outputs 2k, 2k + 1) [k0, k1), and the blocks [b0, b1] of 32 NEW source pixels it walks.
// Block b brings pixels 32 b .. 32 b + 31 and completes the pairs `adv` behind them; the first block only fills the
// window, the last one also flushes the words a 16-byte group still waits for.
struct URange {
    int k0, k1, b0, b1;
    __device__ __forceinline__ int blocks() const { return k1 > k0 ? b1 - b0 + 1 : 0; }
};
__device__ __forceinline__ int floor_div32(int v) { return (v + (1 << 20)) / 32 - (1 << 15); }
__device__ __forceinline__ URange u_range(int ox0, int ox1, int s, int off, int adv) {
    const int ka = ox0 >> 1, kb = (ox1 + 1) >> 1;                              // pairs of the item
    const int nblk = (kb - ka + kUBlockPx - 1) / kUBlockPx;
    const int mid = min(kb, ka + ((nblk + 1) / 2) * kUBlockPx);                 // first stream: whole blocks of pairs
    URange r;
    r.k0 = s == 0 ? ka : mid;
    r.k1 = s == 0 ? mid : kb;
    r.b0 = floor_div32(r.k0 + off);                                             // first pixel the first pair reads (may be < 0)
    r.b1 = floor_div32(r.k1 - 1 + adv + kUUnitPx);                              // last pixel of the last pair, plus one unit of flush
    return r;
}

// One 32-byte store (a whole L2 sector) per lane: STG.E.ENL2.256.
__device__ __forceinline__ void st_global_256(uint8_t* p, const uint32_t* w) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]),
                 "r"(w[7])
                 : "memory");
}
// Words by which the row's output stream lags a 32-byte boundary at a unit starting with new pixel x0 = 8 * (2 m + parity).
template <int C, int ADV>
__host__ __device__ constexpr int carry_words(int parity) { return (((2 * C * (8 * parity - ADV)) % 32 + 32) % 32) / 4; }

__device__ __forceinline__ uint32_t pack4(float a, float b, float c, float d) {  // four (value + 0.5) -> four saturated bytes
    return pack_pixel(make_float4(a, b, c, d));
}

// Eight NEW source pixels x0 .. x0 + 7 through the filter: they complete the pairs x0 - ADV .. x0 - ADV + 7, i.e. 16 output
// pixels of this thread's row = 4 C words, which leave as 32-byte groups (one whole sector per lane and store: with lane =
// row a warp's store touches 32 different rows, so anything smaller than a sector is a partial write).  The row's output
// stream lags 32-byte alignment by CW words here: `carry` holds the words of the group in flight.  win: the last T pixels.
// TRIM: the pass's first frame tap feeds only the even output and its last one only the odd output (every symmetric kernel
// at exactly 2x), so those two taps are scalar FMAs instead of packed ones.
template <int C, int T, int ADV, int CW, bool EDGE, bool TRIM>
__device__ __forceinline__ void push_unit(float (&win)[T][C], uint32_t (&carry)[7], const int (&hi)[8 * C], const int (&lo)[8 * C],
                                          const float2 (&up)[T], int x0, const float2* __restrict__ pairs, int n_in, float unscale,
                                          uint8_t* __restrict__ dst_row, int lo_b, int hi_b, bool row_live) {
    static_assert(C == 3 || C == 4, "Rgb8 / Rgba8");
    uint32_t word[CW + 4 * C];
#pragma unroll
    for (int j = 0; j < CW; ++j) word[j] = carry[j];
    uint32_t* const w = word + CW;   // the unit's own words
    // the T - 1 pixels kept from before and the eight new ones, as one run: pair i of the unit reads s[i .. i + T - 1]
    float smp[T - 1 + kUUnitPx][C];
#pragma unroll
    for (int t = 0; t + 1 < T; ++t)
#pragma unroll
        for (int c = 0; c < C; ++c) smp[t][c] = win[t + 1][c];
#pragma unroll
    for (int i = 0; i < kUUnitPx; ++i)
#pragma unroll
        for (int c = 0; c < C; ++c) smp[T - 1 + i][c] = __int2float_rn(hi[i * C + c] * kBand8Base + lo[i * C + c]);
#pragma unroll
    for (int t = 0; t < T; ++t)
#pragma unroll
        for (int c = 0; c < C; ++c) win[t][c] = smp[kUUnitPx - 1 + t][c];   // (win[0] is only a placeholder: T - 1 pixels carry over)
#pragma unroll
    for (int h = 0; h < kUUnitPx; h += 4) {   // four pairs at a time: 4 C independent accumulation chains, taps outermost
        float2 acc[4][C];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int c = 0; c < C; ++c) acc[i][c] = make_float2(kRoundBias, kRoundBias);
#pragma unroll
        for (int t = 0; t < T; ++t) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float2 wt = up[t];
                if (EDGE) {
                    const float2 g = __ldg(pairs + size_t(min(max(x0 - ADV + h + i, 0), n_in - 1)) * T + t);
                    wt = make_float2(g.x * unscale, g.y * unscale);
                }
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const float v = smp[h + i + t][c];
                    if (TRIM && t == 0) acc[i][c].x = fmaf(wt.x, v, acc[i][c].x);
                    else if (TRIM && t == T - 1) acc[i][c].y = fmaf(wt.y, v, acc[i][c].y);
                    else acc[i][c] = __ffma2_rn(wt, make_float2(v, v), acc[i][c]);
                }
            }
        }
        // bytes: per pair the even output (C channels), then the odd one
#pragma unroll
        for (int i = 0; i < 4; i += 2) {
            const float2(&a)[C] = acc[i];
            const float2(&b)[C] = acc[i + 1];
            if (C == 4) {
                w[2 * (h + i)] = pack4(a[0].x, a[1].x, a[2].x, a[3].x);
                w[2 * (h + i) + 1] = pack4(a[0].y, a[1].y, a[2].y, a[3].y);
                w[2 * (h + i) + 2] = pack4(b[0].x, b[1].x, b[2].x, b[3].x);
                w[2 * (h + i) + 3] = pack4(b[0].y, b[1].y, b[2].y, b[3].y);
            } else {
                w[3 * ((h + i) >> 1)] = pack4(a[0].x, a[1].x, a[2].x, a[0].y);
                w[3 * ((h + i) >> 1) + 1] = pack4(a[1].y, a[2].y, b[0].x, b[1].x);
                w[3 * ((h + i) >> 1) + 2] = pack4(b[2].x, b[0].y, b[1].y, b[2].y);
            }
        }
    }
    constexpr int kGroups = (CW + 4 * C) / 8, kLeft = (CW + 4 * C) % 8;
#pragma unroll
    for (int j = 0; j < kLeft; ++j) carry[j] = word[8 * kGroups + j];
    if (!row_live) return;
    const int gbyte0 = 2 * (x0 - ADV) * C - 4 * CW;                             // first byte of the first whole group (a multiple of 32)
#pragma unroll
    for (int g = 0; g < kGroups; ++g) {
        const int gb = gbyte0 + 32 * g;
        if (gb >= lo_b && gb + 32 <= hi_b) {
            st_global_256(dst_row + gb, word + 8 * g);
        } else if (gb + 32 > lo_b && gb < hi_b) {   // the ragged end of the row: byte by byte
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int bb = 0; bb < 4; ++bb) {
                    const int at = gb + 4 * j + bb;
                    if (at >= lo_b && at < hi_b) dst_row[at] = uint8_t(word[8 * g + j] >> (8 * bb));
                }
        }
    }
}

}  // namespace

// Shared memory: [mbarriers (1 KB) | source block ring: 4 x (4 boxes x 4 KB) | weight tiles of the band: chunks x 2 digits x 4 KB]
template <int C, int T>
__global__ void __launch_bounds__(kUThreads, 1)
banded8u_kernel(const DevJob* __restrict__ jobs, const WorkItem* __restrict__ items) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int kN = kUBlockPx * C;                              // byte columns per block = N of one MMA
    uint64_t* const bars = reinterpret_cast<uint64_t*>(smem);
    uint64_t* const u_full = bars;
    uint64_t* const u_empty = u_full + kUStages;
    uint64_t* const t_full = u_empty + kUStages;
    uint64_t* const t_empty = t_full + kUStreams;
    uint64_t* const w_full = t_empty + kUStreams;
    uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(smem + 512);
    uint8_t* const ustage = smem + kUHeaderBytes;
    uint8_t* const wtiles = ustage + kUStages * kUBlockStage;

    int tid;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
    const int warp = tid >> 5;
    const int lane = tid & 31;

    const WorkItem it = items[blockIdx.x];
    const DevJob* __restrict__ J = jobs + it.job;
    const int band = it.oy0 / kURows;
    const int nch = J->v.band8t_chunks;
    const int k_lo = __ldg(J->v.band8t_klo + band);
    constexpr int kAdv = (T - 1) / 2;                              // pair k is complete once source pixel k + kAdv has arrived (frame offset -(T-1)/2)
    constexpr int kCW0 = carry_words<C, kAdv>(0), kCW1 = carry_words<C, kAdv>(1);   // words the output stream lags a 32-byte boundary
    static_assert((2 * C * kAdv) % 4 == 0, "the output stream must stay word aligned");

    if (tid == 0) {
        if (smem_addr(smem) & 1023u) __trap();
        for (int s = 0; s < kUStages; ++s) { mbar_init(u_full + s, 1); mbar_init(u_empty + s, 1); }
        for (int s = 0; s < kUStreams; ++s) { mbar_init(t_full + s, 1); mbar_init(t_empty + s, 4); }
        mbar_init(w_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(tmem_slot)), "n"(kUTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const URange sr0 = u_range(it.ox0, it.ox1, 0, -kAdv, kAdv), sr1 = u_range(it.ox0, it.ox1, 1, -kAdv, kAdv);

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kURegsIo));
        const int nb0 = sr0.blocks(), nb1 = sr1.blocks();
        const int nbmax = max(nb0, nb1);
        if (warp == 0) {
            // -------------------------------------------------------------------------- producer
            const bool leader = elect_one();
            const void* const src_map = J->src_map8;
            if (leader) {
                asm volatile("prefetch.tensormap [%0];" ::"l"(src_map) : "memory");
                const uint32_t wbytes = uint32_t(nch) * 2u * kUWTile;
                mbar_expect_tx(w_full, wbytes);
                bulk_load(wtiles, J->v.band8t_tiles + size_t(band) * nch * 2 * kUWTile, wbytes, w_full);
            }
            int su = 0;
            uint32_t pu = 1;
            for (int t = 0; t < nbmax; ++t) {
                for (int s = 0; s < kUStreams; ++s) {
                    if (t >= (s == 0 ? nb0 : nb1)) continue;
                    const int blk = (s == 0 ? sr0.b0 : sr1.b0) + t;
                    mbar_wait_parked(u_empty + su, pu);
                    if (leader) {
                        mbar_expect_tx(u_full + su, uint32_t(nch) * kUBoxBytes);
                        for (int c = 0; c < nch; ++c)   // (x is a multiple of 16 bytes, as TMA wants; negative or past the row: zeros)
                            tma_load_2d(ustage + su * kUBlockStage + c * kUBoxBytes, src_map, C * blk * kUBlockPx, k_lo + c * kUChunk, u_full + su);
                    }
                    if (++su == kUStages) { su = 0; pu ^= 1; }
                }
            }
            __syncwarp();
        } else if (warp == 1) {
            // -------------------------------------------------------------------------- MMA issuer
            const bool leader = elect_one();
            const uint32_t bars_a = smem_addr(bars);
            const uint32_t a_lo0 = ((smem_addr(wtiles) >> 4) & 0x3fffu) | (((uint32_t(kURows) * 16u) >> 4) << 16);
            constexpr uint32_t kADescHi = (128u >> 4) | (1u << 14);
            const uint32_t b_lo0 = ((smem_addr(ustage) >> 4) & 0x3fffu) | ((1024u >> 4) << 16);
            const uint32_t idesc = instr_desc_i8u(kN);
            mbar_wait_at(bars_a + uint32_t(2 * kUStages + 2 * kUStreams) * 8, 0);   // weights
            int su = 0;
            uint32_t pu = 0, pt = 1;
            for (int t = 0; t < nbmax; ++t) {
#pragma unroll
                for (int s = 0; s < kUStreams; ++s) {
                    if (t >= (s == 0 ? nb0 : nb1)) continue;
                    mbar_wait_at(bars_a + uint32_t(2 * kUStages + kUStreams + s) * 8, pt);
                    mbar_wait_at(bars_a + uint32_t(su) * 8, pu);
                    tc_fence_after();
                    if (leader) {
                        const uint32_t d_hi = tmem + uint32_t(s * kUTileCols), d_lo = d_hi + 128;
                        const uint32_t b_lo = b_lo0 + uint32_t(su) * (kUBlockStage >> 4);
                        for (int c = 0; c < nch; ++c)   // one accumulator at a time: switching costs ~100 cycles per MMA
                            mma_i8u(d_hi, make_u64(a_lo0 + uint32_t(c) * (2u * kUWTile >> 4), kADescHi),
                                    make_u64(b_lo + uint32_t(c) * (kUBoxBytes >> 4), srcU_desc_hi()), idesc, c > 0);
                        for (int c = 0; c < nch; ++c)
                            mma_i8u(d_lo, make_u64(a_lo0 + uint32_t(c) * (2u * kUWTile >> 4) + (kUWTile >> 4), kADescHi),
                                    make_u64(b_lo + uint32_t(c) * (kUBoxBytes >> 4), srcU_desc_hi()), idesc, c > 0);
                        tc_commit(u_empty + su);
                        tc_commit(t_full + s);
                    }
                    if (++su == kUStages) { su = 0; pu ^= 1; }
                }
                pt ^= 1;
            }
            __syncwarp();
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kURegsEpi));
        // ------------------------------------------------------------------------------ epilogue: one output row per thread
        const int q = warp & 3;
        const int s = (warp >> 2) - 1;
        const URange sr = s == 0 ? sr0 : sr1;
        const int orow = it.oy0 + q * 32 + lane;
        const bool row_live = orow < it.oy1;
        uint8_t* const dst_row = J->dst + size_t(min(orow, it.oy1 - 1)) * J->dst_pitch;
        const int row_bytes = int(J->dw) * C;
        const uint32_t taddr = tmem + (uint32_t(q * 32) << 16) + uint32_t(s * kUTileCols);
        const uint32_t t_full_a = smem_addr(t_full + s), t_empty_a = smem_addr(t_empty + s);
        // the pairs carry 2^74 (they also feed the CUDA-core kernel's denormal operands); the vertical sums carry 2^shift
        const float unscale = __int_as_float((127 - 74 - J->v.band8_shift) << 23);
        const float2* __restrict__ pairs = J->h.up2_pairs_h;
        const int uni_lo = J->h.up2_uni_lo, uni_hi = J->h.up2_uni_hi, n_in = J->h.n_in;
        float2 up[T];
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const float2 g = __ldg(pairs + size_t(uni_lo) * T + t);
            up[t] = make_float2(g.x * unscale, g.y * unscale);
        }
        // the fast path drops the zero halves of the first and last frame tap; a pass whose pairs are not of that shape takes the
        // general path everywhere
        const bool trim_ok = up[0].y == 0.0f && up[T - 1].x == 0.0f;
        float win[T][C];
#pragma unroll
        for (int t = 0; t < T; ++t)
#pragma unroll
            for (int c = 0; c < C; ++c) win[t][c] = 0.0f;
        uint32_t carry[7] = {0u, 0u, 0u, 0u, 0u, 0u, 0u};
        const int lo_b = 2 * C * sr.k0, hi_b = min(2 * C * sr.k1, row_bytes);   // this stream's bytes of the row

        const int nb = sr.blocks();
        for (int n = 0; n < nb; ++n) {
            const int xb = (sr.b0 + n) * kUBlockPx;                    // first new pixel of the block
            mbar_wait_at(t_full_a, n & 1);
            tc_fence_after();
            int hiA[8 * C], loA[8 * C], hiB[8 * C], loB[8 * C];
            tmem_ld_unit<C>(taddr, hiA);
            tmem_ld_unit<C>(taddr + 128, loA);
            tmem_ld_unit<C>(taddr + 8 * C, hiB);
            tmem_ld_unit<C>(taddr + 128 + 8 * C, loB);
            tmem_ld_wait();
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                const int x0 = xb + 16 * h;
                const bool in0 = trim_ok && x0 - kAdv >= uni_lo && x0 - kAdv + 8 <= uni_hi;
                const bool in1 = trim_ok && x0 - kAdv + 8 >= uni_lo && x0 - kAdv + 16 <= uni_hi;
                if (in0) push_unit<C, T, kAdv, kCW0, false, true>(win, carry, hiA, loA, up, x0, pairs, n_in, unscale, dst_row, lo_b, hi_b, row_live);
                else push_unit<C, T, kAdv, kCW0, true, false>(win, carry, hiA, loA, up, x0, pairs, n_in, unscale, dst_row, lo_b, hi_b, row_live);
                if (h == 0) {
                    tmem_ld_unit<C>(taddr + 16 * C, hiA);
                    tmem_ld_unit<C>(taddr + 128 + 16 * C, loA);
                }
                if (in1) push_unit<C, T, kAdv, kCW1, false, true>(win, carry, hiB, loB, up, x0 + 8, pairs, n_in, unscale, dst_row, lo_b, hi_b, row_live);
                else push_unit<C, T, kAdv, kCW1, true, false>(win, carry, hiB, loB, up, x0 + 8, pairs, n_in, unscale, dst_row, lo_b, hi_b, row_live);
                if (h == 0) {
                    tmem_ld_unit<C>(taddr + 24 * C, hiB);
                    tmem_ld_unit<C>(taddr + 128 + 24 * C, loB);
                    tmem_ld_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_at(t_empty_a);          // the tile may be overwritten
                }
            }
        }
    }

    // ---------------------------------------------------------------------------------- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kUTmemCols));
    }
}

// ---- launcher ---------------------------------------------------------------------------------

bool banded8u_supported(int channels, int taps_h, int off_h, int chunks_v) {
    if (!(channels == 3 || channels == 4) || !(taps_h == 3 || taps_h == 5 || taps_h == 7)) return false;
    const int adv = (taps_h - 1) / 2;
    return off_h == -adv && (2 * channels * adv) % 4 == 0 && chunks_v >= 1 && chunks_v <= kUMaxChunks;
}
int banded8u_band_rows() { return kURows; }

template <int C, int T>
static cudaError_t launch_u(const DevJob* jobs, const WorkItem* items, int n_items, cudaStream_t stream) {
    const size_t smem = size_t(kUHeaderBytes) + size_t(kUStages) * kUBlockStage + size_t(kUMaxChunks) * 2 * kUWTile;
    cudaError_t e = cudaFuncSetAttribute(banded8u_kernel<C, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(banded8u_kernel<C, T>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    banded8u_kernel<C, T><<<n_items, kUThreads, smem, stream>>>(jobs, items);
    return cudaGetLastError();
}

cudaError_t launch_banded8u(int channels, int taps, const DevJob* jobs, const WorkItem* items, int n_items, cudaStream_t stream) {
    if (channels == 3 && taps == 5) return launch_u<3, 5>(jobs, items, n_items, stream);   // (taps 3 and 7: the Rgb8 stream is not word aligned)
    if (channels == 4 && taps == 3) return launch_u<4, 3>(jobs, items, n_items, stream);
    if (channels == 4 && taps == 5) return launch_u<4, 5>(jobs, items, n_items, stream);
    if (channels == 4 && taps == 7) return launch_u<4, 7>(jobs, items, n_items, stream);
    return cudaErrorInvalidValue;
}

}  // namespace ikc
