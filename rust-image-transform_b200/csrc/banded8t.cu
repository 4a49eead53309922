// banded8t.cu -- fused single-launch kernel for Rgba8 downscales that are exactly 2:1 horizontally (the 4K -> 1080p
// Lanczos3 resize of BASELINE config 2 and its relatives), with NO shared-memory intermediate: sm_100a.
//
// banded8.cu computes the vertical pass on the tensor cores with TMEM lane = source byte column, so the horizontal pass
// has to go through shared memory (every intermediate value is converted, stored, re-loaded ~6 times, two team barriers
// per 16 rows).  Here the operands swap roles:
//
//   A (M x K) = the vertical weights of one BAND of 128 output rows: round(w * 2^S) in two signed base-256 digits, one
//               host-built K-major s8 tile per chunk of 32 source rows and digit (plan.hpp: Band8T), resident in
//               shared memory for the CTA's whole life;
//   B (N x K) = u8 source bytes, N = 128 byte columns (32 Rgba pixels), K = 32 source rows, exactly as a 2-D TMA box
//               (32 rows x 128 bytes, 128-byte swizzle) lands: no thread touches the source;
//   D (M x N) = s32 accumulators in TMEM, LANE = OUTPUT ROW, column = byte column (hi digits | lo digits).
//
// A thread of the epilogue therefore owns one output row and reads a run of its intermediate pixels straight from TMEM
// into registers (tcgen05.ld, digits recombined with one IMAD + one I2FP per value) and walks the horizontal filter
// along x entirely in registers: eight open outputs in static accumulator slots, one "super-step" of 16 pixels = half an
// accumulator tile as straight-line code (each pixel feeds the six outputs whose 12-tap windows contain it; an output
// finishes every second pixel; four finished pixels leave as one 16-byte store).  Interior super-steps take the twelve
// tap weights of the uniform 2:1 stretch from registers; the few at the image borders look their weights up.
//
// One CTA per SM, 384 threads:
//   warp 0      producer: the band's weight tiles once, then per block of 128 byte columns one TMA box per chunk
//   warp 1      MMA issuer (one elected lane): per accumulator tile 2 x chunks MMAs (first chunk overwrites: no zeroing)
//   warps 4-11  two epilogue QUADS (TMEM lane quarter = warp % 4).  Each quad streams along x through its own half of the
//               item's output columns (its own accumulator tile; its MMAs run while the other quad computes), so the two quads never
//               synchronise with each other, and warps never synchronise at all (mbarriers with the MMA warp only).
#include <cuda_runtime.h>

#include <cstdint>

#include "banded_common.cuh"
#include "device_types.hpp"
#include "launch.hpp"
#include "plan.hpp"

namespace ikc {
namespace {

constexpr int kTRows = kBand8TRows;                      // output rows per band = TMEM lanes = M
constexpr int kTChunk = kBand8Chunk;                     // source rows per MMA = K
constexpr int kTBlockBytes = 128;                        // byte columns per source box
constexpr int kTStageBytes = kTBlockBytes * kTChunk;     // one box: 32 rows x 128 bytes, swizzled
constexpr int kTWTile = kTRows * kTChunk;                // one weight tile (band, chunk, digit)
constexpr int kTStages = 3;                              // ring of source BLOCKS (each: the band's chunks x one box)
constexpr int kTBlockStage = kBand8TMaxChunks * kTStageBytes;   // 40 KB
constexpr int kTStreams = 2;                             // epilogue quads, each with its own column range
constexpr int kTUnitCols = kTBlockBytes;                 // byte columns per accumulator tile = N = two super-steps of 16 pixels
constexpr int kTTileCols = 2 * kTUnitCols;               // TMEM columns of a tile: hi digits | lo digits
constexpr int kTTmemCols = kTStreams * kTTileCols;       // 512: one tile per stream
constexpr int kTThreads = 128 + kTStreams * 128;
constexpr int kTHeaderBytes = 1024;
constexpr int kTRegsIo = 40, kTRegsEpi = 232;            // 128 x 40 + 256 x 232 <= 64 K registers
constexpr int kTPx = 16;                                 // pixels per super-step
constexpr int kTSlots = 8;                               // accumulator slots (outputs o and o + 8 share one)
constexpr int kTTaps = 12, kTLead = 5;                   // output o reads pixels 2 o - 5 .. 2 o + 6

static_assert(kTTmemCols == 512, "the accumulator tiles fill TMEM");

// Shared-memory descriptor of the source tile (B operand): MN-major, 128-byte swizzle, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint32_t srcT_desc_hi() { return (1024u >> 4) | (1u << 14) | (2u << 29); }
// Instruction descriptor: D = s32, A = s8 (K-major), B = u8 (MN-major), M = 128, N = n.
__device__ __forceinline__ uint32_t instr_desc_i8t(uint32_t n) {
    return (2u << 4) | (1u << 7) | (0u << 10) | (1u << 16) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// One 32-byte store (a whole L2 sector) per lane: with lane = row a warp's store touches 32 different rows, so anything
// smaller than a sector is a partial write.
__device__ __forceinline__ void st_global_256(uint8_t* p, const uint32_t (&w)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]),
                 "r"(w[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t addr, int (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(addr));
}

// Column range of one stream: outputs [x0, x1) and the blocks (128 byte columns = 32 pixels) [b0, b1] its windows span.
struct StreamRange {
    int x0, x1, b0, b1;
    __device__ __forceinline__ int blocks() const { return x1 > x0 ? b1 - b0 + 1 : 0; }
};
__device__ __forceinline__ StreamRange stream_range(int ox0, int ox1, int s) {
    const int mid = min(ox1, ox0 + ((((ox1 - ox0) + 1) / 2 + 7) & ~7));   // first stream: a multiple of 8 outputs
    StreamRange r;
    r.x0 = s == 0 ? ox0 : mid;
    r.x1 = s == 0 ? mid : ox1;
    r.b0 = max(2 * r.x0 - kTLead, 0) / (2 * kTPx);
    // (outputs leave in aligned groups of eight, when the group's last one completes: run until that one has, even if it
    // lies beyond the stream's -- or the image's -- last output)
    r.b1 = (2 * ((r.x1 - 1) | 7) + kTTaps - kTLead - 1) / (2 * kTPx);
    return r;
}

// The horizontal filter state of one output row: eight accumulator slots (slot = output & 7), two float2 per slot
// (channels 0-1, 2-3), the packed pixels of the 16-byte group in flight.
struct RowState {
    float2 acc[kTSlots][2];
    uint32_t word[8];   // the packed pixels of the 32-byte group in flight (outputs 8 m .. 8 m + 7)
};

// Weight of source pixel x in output o, times `unscale` (border super-steps; the table is the pass's [n_out][stride]).
__device__ __forceinline__ float edge_weight(const DevPass& h, int o, int x, float unscale) {
    if (o < 0 || o >= h.n_out) return 0.0f;
    const int l = __ldg(h.left + o), r = __ldg(h.right + o);
    return (x >= l && x < r) ? __ldg(h.w + size_t(o) * h.stride + (x - l)) * unscale : 0.0f;
}

// Eight pixels (half a super-step: pixels 8 * HALF .. 8 * HALF + 7 of super-step Q) pushed through the filter.
//   hi / lo : the digits of the 32 intermediate values, as tcgen05.ld delivered them
//   uw      : the uniform stretch's twelve tap weights (x 2^-shift, duplicated pairs)      (EDGE == false)
//   EDGE    : weights are looked up per (output, pixel) instead
// Output 8 Q + j is complete after pixel 2 j + 6 of the super-step; it is stored when its group of four is, if in range
// (FULL: the caller knows every group the super-step completes is).
template <int HALF, bool EDGE, bool FULL>
__device__ __forceinline__ void push_half(RowState& st, const int (&hi)[32], const int (&lo)[32], const float2 (&uw)[kTTaps], int Q,
                                          const DevPass& h, float unscale, uint8_t* __restrict__ dst_row, int x0, int x1, bool row_live) {
#pragma unroll
    for (int ii = 0; ii < 8; ++ii) {
        constexpr int kBase = 8 * HALF;
        const int i = kBase + ii;                       // pixel of the super-step
        float v[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) v[c] = __int2float_rn(hi[4 * ii + c] * kBand8Base + lo[4 * ii + c]);
        const float2 v01 = make_float2(v[0], v[1]), v23 = make_float2(v[2], v[3]);
        // outputs j (relative to 8 Q) whose windows [2 j - 5, 2 j + 6] contain pixel i
#pragma unroll
        for (int j = -3; j <= 10; ++j) {
            const int t = i - (2 * j - kTLead);         // tap index
            if (t < 0 || t >= kTTaps) continue;
            const int slot = (j + kTSlots) & (kTSlots - 1);
            float2 w = uw[t];
            if (EDGE) {                                 // outputs outside the uniform stretch: look the weight up
                const int o = 8 * Q + j;
                if (o < h.uni_lo || o >= h.uni_hi) {
                    const float we = edge_weight(h, o, kTPx * Q + i, unscale);
                    w = make_float2(we, we);
                }
            }
            st.acc[slot][0] = __ffma2_rn(w, v01, st.acc[slot][0]);
            st.acc[slot][1] = __ffma2_rn(w, v23, st.acc[slot][1]);
        }
        if ((i & 1) == 0) {                             // output j = (i - 6) / 2 is complete
            const int j = (i - 6) / 2;                  // -3 .. 4
            const int slot = (j + kTSlots) & (kTSlots - 1);
            const int o = 8 * Q + j;
            st.word[(j + 8) & 7] = pack_pixel(make_float4(st.acc[slot][0].x, st.acc[slot][0].y, st.acc[slot][1].x, st.acc[slot][1].y));
            st.acc[slot][0] = st.acc[slot][1] = make_float2(kRoundBias, kRoundBias);
            if (((j + 8) & 7) == 7) {                   // outputs o - 7 .. o: one aligned group of eight = one 32-byte sector
                const int og = o - 7;
                if (FULL) {                             // the whole super-step lies inside the stream's columns: no range tests
                    if (row_live) st_global_256(dst_row + size_t(og) * 4, st.word);
                } else if (row_live) {
                    if (og >= x0 && o < x1) {
                        st_global_256(dst_row + size_t(og) * 4, st.word);
                    } else if (o >= x0 && og < x1) {
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            if (og + k >= x0 && og + k < x1) *reinterpret_cast<uint32_t*>(dst_row + size_t(og + k) * 4) = st.word[k];
                    }
                }
            }
        }
    }
}

}  // namespace

// Shared memory: [mbarriers (1 KB) | source block ring: 3 x (10 boxes x 4 KB) | weight tiles of the band: chunks x 2 digits x 4 KB]
__global__ void __launch_bounds__(kTThreads, 1)
banded8t_kernel(const DevJob* __restrict__ jobs, const WorkItem* __restrict__ items, const Band8TGeom geom) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* const bars = reinterpret_cast<uint64_t*>(smem);
    uint64_t* const u_full = bars;                                 // [kTStages] the block's boxes landed
    uint64_t* const u_empty = u_full + kTStages;                   // [kTStages] the MMAs that read them have completed
    uint64_t* const t_full = u_empty + kTStages;                   // [streams] every MMA of the tile has completed
    uint64_t* const t_empty = t_full + kTStreams;                  // [streams] the quad's four warps have read it
    uint64_t* const w_full = t_empty + kTStreams;                  // the weight tiles landed
    uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(smem + 512);
    uint8_t* const ustage = smem + kTHeaderBytes;
    uint8_t* const wtiles = ustage + kTStages * kTBlockStage;
    (void)geom;

    int tid;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
    const int warp = tid >> 5;
    const int lane = tid & 31;

    const WorkItem it = items[blockIdx.x];
    const DevJob* __restrict__ J = jobs + it.job;
    const int band = it.oy0 / J->v.band8t_rows;                    // (bands are band8t_rows <= 128 outputs high)
    const int nch = J->v.band8t_chunks;                            // tiles per band and digit (geom.chunks sizes shared memory)
    const int k_lo = __ldg(J->v.band8t_klo + band);

    if (tid == 0) {
        if (smem_addr(smem) & 1023u) __trap();
        for (int s = 0; s < kTStages; ++s) { mbar_init(u_full + s, 1); mbar_init(u_empty + s, 1); }
        for (int s = 0; s < kTStreams; ++s) { mbar_init(t_full + s, 1); mbar_init(t_empty + s, 4); }
        mbar_init(w_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(tmem_slot)), "n"(kTTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const StreamRange sr0 = stream_range(it.ox0, it.ox1, 0), sr1 = stream_range(it.ox0, it.ox1, 1);

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kTRegsIo));
        // Blocks of the two streams, interleaved: the order both the producer and the MMA warp walk.
        const int nb0 = sr0.blocks(), nb1 = sr1.blocks();
        const int nbmax = max(nb0, nb1);
        if (warp == 0) {
            // -------------------------------------------------------------------------- producer
            const bool leader = elect_one();
            const void* const src_map = J->src_map8;
            if (leader) {
                asm volatile("prefetch.tensormap [%0];" ::"l"(src_map) : "memory");
                const uint32_t wbytes = uint32_t(nch) * 2u * kTWTile;
                mbar_expect_tx(w_full, wbytes);
                bulk_load(wtiles, J->v.band8t_tiles + size_t(band) * nch * 2 * kTWTile, wbytes, w_full);
            }
            int su = 0;
            uint32_t pu = 1;
            for (int t = 0; t < nbmax; ++t) {
                for (int s = 0; s < kTStreams; ++s) {
                    if (t >= (s == 0 ? nb0 : nb1)) continue;
                    const int blk = (s == 0 ? sr0.b0 : sr1.b0) + t;
                    mbar_wait_parked(u_empty + su, pu);
                    if (leader) {
                        mbar_expect_tx(u_full + su, uint32_t(nch) * kTStageBytes);
                        for (int c = 0; c < nch; ++c)
                            tma_load_2d(ustage + su * kTBlockStage + c * kTStageBytes, src_map, blk * kTBlockBytes, k_lo + c * kTChunk, u_full + su);
                    }
                    if (++su == kTStages) { su = 0; pu ^= 1; }
                }
            }
            __syncwarp();
        } else if (warp == 1) {
            // -------------------------------------------------------------------------- MMA issuer
            // All of a tile's MMAs into one accumulator run back to back (hi digits over every chunk, then lo digits):
            // consecutive tcgen05.mma.kind::i8 into DIFFERENT accumulators cost 171 cycles each instead of 75
            // (tools/tc_probe_i8.cu).
            const bool leader = elect_one();
            const uint32_t bars_a = smem_addr(bars);
            const uint32_t a_lo0 = ((smem_addr(wtiles) >> 4) & 0x3fffu) | (((uint32_t(kTRows) * 16u) >> 4) << 16);   // LBO: the two halves of K
            constexpr uint32_t kADescHi = (128u >> 4) | (1u << 14);                                               // SBO: 8-row groups
            const uint32_t b_lo0 = ((smem_addr(ustage) >> 4) & 0x3fffu) | ((1024u >> 4) << 16);
            const uint32_t idesc = instr_desc_i8t(kTUnitCols);
            mbar_wait_at(bars_a + uint32_t(2 * kTStages + 2 * kTStreams) * 8, 0);   // weights
            int su = 0;
            uint32_t pu = 0;
            uint32_t pt = 1;                                        // phase of the streams' t_empty barriers (first pass: free)
            for (int t = 0; t < nbmax; ++t) {
#pragma unroll
                for (int s = 0; s < kTStreams; ++s) {
                    if (t >= (s == 0 ? nb0 : nb1)) continue;
                    mbar_wait_at(bars_a + uint32_t(2 * kTStages + kTStreams + s) * 8, pt);
                    mbar_wait_at(bars_a + uint32_t(su) * 8, pu);
                    tc_fence_after();
                    if (leader) {
                        const uint32_t d_hi = tmem + uint32_t(s * kTTileCols), d_lo = d_hi + kTUnitCols;
                        const uint32_t b_lo = b_lo0 + uint32_t(su) * (kTBlockStage >> 4);
                        for (int c = 0; c < nch; ++c)
                            mma_i8(d_hi, make_u64(a_lo0 + uint32_t(c) * (2u * kTWTile >> 4), kADescHi),
                                   make_u64(b_lo + uint32_t(c) * (kTStageBytes >> 4), srcT_desc_hi()), idesc, c > 0);
                        for (int c = 0; c < nch; ++c)
                            mma_i8(d_lo, make_u64(a_lo0 + uint32_t(c) * (2u * kTWTile >> 4) + (kTWTile >> 4), kADescHi),
                                   make_u64(b_lo + uint32_t(c) * (kTStageBytes >> 4), srcT_desc_hi()), idesc, c > 0);
                        tc_commit(u_empty + su);
                        tc_commit(t_full + s);
                    }
                    if (++su == kTStages) { su = 0; pu ^= 1; }
                }
                pt ^= 1;
            }
            __syncwarp();
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kTRegsEpi));
        // ------------------------------------------------------------------------------ epilogue: one output row per thread
        const int q = warp & 3;                                    // TMEM lane quarter
        const int s = (warp >> 2) - 1;                             // stream
        const StreamRange sr = s == 0 ? sr0 : sr1;
        const int orow = it.oy0 + q * 32 + lane;
        const bool row_live = orow < it.oy1;
        uint8_t* const dst_row = J->dst + size_t(min(orow, it.oy1 - 1)) * J->dst_pitch;
        const uint32_t taddr = tmem + (uint32_t(q * 32) << 16) + uint32_t(s * kTTileCols);
        const uint32_t t_full_a = smem_addr(t_full + s), t_empty_a = smem_addr(t_empty + s);
        const float unscale = __int_as_float((127 - J->v.band8_shift) << 23);   // 2^-shift: the vertical sums are integers x 2^shift
        const DevPass& hp = J->h;
        const int uni_lo = hp.uni_lo, uni_hi = hp.uni_hi;
        float2 uw[kTTaps];
#pragma unroll
        for (int t = 0; t < kTTaps; ++t) {
            const float w = __ldg(hp.w + size_t(uni_lo) * hp.stride + t) * unscale;
            uw[t] = make_float2(w, w);
        }
        RowState st;
#pragma unroll
        for (int k = 0; k < kTSlots; ++k) st.acc[k][0] = st.acc[k][1] = make_float2(kRoundBias, kRoundBias);
#pragma unroll
        for (int k = 0; k < 8; ++k) st.word[k] = 0;

        // One block = two super-steps = four quarters of 8 pixels in two register buffers.  Quarters 2 and 3 are fetched
        // (tcgen05.ld) while quarters 0 and 1 are pushed through the filter, and the tile goes back to the MMA warp as soon
        // as they have landed: the next block's MMAs then have half a block of this quad's arithmetic to hide behind.  The
        // super-step body exists once (a two-trip loop): the straight-line code of a whole block would not stay in the
        // instruction cache.
        const int nb = sr.blocks();
        for (int n = 0; n < nb; ++n) {
            const int Q = 2 * (sr.b0 + n);                         // the block's first super-step
            mbar_wait_at(t_full_a, n & 1);
            tc_fence_after();
            int hiA[32], loA[32], hiB[32], loB[32];
            tmem_ld32(taddr, hiA);
            tmem_ld32(taddr + kTUnitCols, loA);
            tmem_ld32(taddr + 32, hiB);
            tmem_ld32(taddr + kTUnitCols + 32, loB);
            tmem_ld_wait();
#pragma unroll 1
            for (int ss = 0; ss < 2; ++ss) {
                // every output the super-step touches lies in the uniform stretch (tap weights from registers) and every
                // group of four it completes inside the stream's columns (no range tests): the fast path
                const int qs = Q + ss;
                const bool interior = 8 * qs - 3 >= uni_lo && 8 * qs + 10 < uni_hi;
                const bool fast = interior && 8 * qs - 8 >= sr.x0 && 8 * qs - 1 < sr.x1;   // the group this super-step completes
                if (fast) push_half<0, false, true>(st, hiA, loA, uw, qs, hp, unscale, dst_row, sr.x0, sr.x1, row_live);
                else if (interior) push_half<0, false, false>(st, hiA, loA, uw, qs, hp, unscale, dst_row, sr.x0, sr.x1, row_live);
                else push_half<0, true, false>(st, hiA, loA, uw, qs, hp, unscale, dst_row, sr.x0, sr.x1, row_live);
                if (ss == 0) {
                    tmem_ld32(taddr + 64, hiA);
                    tmem_ld32(taddr + kTUnitCols + 64, loA);
                }
                if (fast) push_half<1, false, true>(st, hiB, loB, uw, qs, hp, unscale, dst_row, sr.x0, sr.x1, row_live);
                else if (interior) push_half<1, false, false>(st, hiB, loB, uw, qs, hp, unscale, dst_row, sr.x0, sr.x1, row_live);
                else push_half<1, true, false>(st, hiB, loB, uw, qs, hp, unscale, dst_row, sr.x0, sr.x1, row_live);
                if (ss == 0) {
                    tmem_ld32(taddr + 96, hiB);
                    tmem_ld32(taddr + kTUnitCols + 96, loB);
                    tmem_ld_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_at(t_empty_a);      // the tile may be overwritten
                }
            }
        }
    }

    // ---------------------------------------------------------------------------------- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTTmemCols));
    }
}

// ---- launcher ---------------------------------------------------------------------------------

size_t banded8t_smem_bytes(const Band8TGeom& g) {
    return size_t(kTHeaderBytes) + size_t(kTStages) * kTBlockStage + size_t(g.chunks) * 2 * kTWTile;
}

cudaError_t launch_banded8t(const DevJob* jobs, const WorkItem* items, const Band8TGeom& geom, cudaStream_t stream) {
    if (geom.chunks < 1 || geom.chunks > kBand8TMaxChunks) return cudaErrorInvalidValue;
    const size_t max_smem = size_t(kTHeaderBytes) + size_t(kTStages) * kTBlockStage + size_t(kBand8TMaxChunks) * 2 * kTWTile;
    // always the planner-wide maximum: the attribute is shared by every thread launching on this device
    cudaError_t e = cudaFuncSetAttribute(banded8t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(max_smem));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(banded8t_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    banded8t_kernel<<<geom.n_items, kTThreads, banded8t_smem_bytes(geom), stream>>>(jobs, items, geom);
    return cudaGetLastError();
}

}  // namespace ikc
