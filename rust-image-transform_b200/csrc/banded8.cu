// banded8.cu -- fused single-launch resize kernel for 8-bit downscales whose VERTICAL pass is an INTEGER banded
// matrix product on the 5th-generation tensor cores (tcgen05.mma kind::i8, s32 accumulators in TMEM) -- sm_100a.
//
// Why a second tensor-core kernel: banded.cu (f16 operands) is bound by shared-memory bandwidth -- the converter
// warps read every staged byte and write it back as two (f16), and the MMA reads those twice (hi + lo weights).
// Here the source bytes are the A operand AS THEY ARE: a 2-D TMA box (32 rows x 128 bytes, 128-byte swizzle) lands
// in shared memory in exactly the MN-major layout kind::i8 reads, so no thread touches the source at all.
//
//   A (M x K) = u8 source bytes, M = 128 byte columns, K = 32 source rows, straight from TMA.
//   B (N x K) = the weights of the <= 32 output rows those source rows can touch, as integers
//               W = round(w * 2^S) split into L signed base-256 digits (host-built tiles, plan.hpp: Band8);
//               N = L * 32: one MMA per block and chunk computes every digit's partial sums.
//   D (M x N) = s32 accumulators in TMEM: lane = byte column, column = (output row mod 32) * L + digit.  The 32 rows are
//               a ring of 4 groups of 8; a finished group is read with tcgen05.ld, its digits recombined in f32
//               ((d1 * 256 + d0) -- exact products, one rounding), written to the shared-memory intermediate tile,
//               zeroed and handed back.  The integer sums are exact, so the vertical pass is deterministic
//               fixed-point arithmetic with 2^-S weight resolution (each output's weights sum to exactly one).
//
// The HORIZONTAL pass is banded.cu's: CUDA cores, output-stationary, lane = intermediate row, half warp = x segment.
//
// One PERSISTENT CTA per SM (512 threads, all 512 TMEM columns, ~205 KB of shared memory) walks the work items
// blockIdx.x, blockIdx.x + gridDim.x, ...; every role runs that sequence on its own, and ring slots / barrier phases follow
// a running group index, so loads and MMAs of the next item start while the epilogue still works on the current one:
//   warp 0      producer (TMA): 5-stage ring of source boxes + ring of weight tiles
//   warps 1-2   MMA issuers (one lane each, two blocks each)
//   warp 3      table warp: the next item's horizontal tables into the buffer the epilogue is not using
//   warps 4-15  three epilogue TEAMS of four warps (TMEM lane quarter = warp % 4).  A team owns every third
//               intermediate tile (16 output rows = 2 ring groups): it drains the tile's groups from TMEM as they become
//               final (recombines the digits, writes its own shared-memory tile, zeroes and returns the ring slots) and
//               then runs the tile's horizontal pass, while the other teams do the same for the next tiles and the MMAs
//               run up to two tiles ahead: the accumulator ring holds 8 groups (64 output rows), twice a chunk's window.
#include <cuda_runtime.h>

#include <cstdint>

#include "banded_common.cuh"
#include "device_types.hpp"
#include "launch.hpp"
#include "plan.hpp"

#ifndef IKC_BANDED8_CONV
#define IKC_BANDED8_CONV 0
#endif

namespace ikc {
namespace {

constexpr int k8Blocks = 4;                          // 128-byte column blocks per strip
constexpr int k8StripBytes = k8Blocks * 128;
constexpr int k8Chunk = kBand8Chunk;                 // source rows per stage = K of one i8 MMA
constexpr int k8BlockBytes = 128 * k8Chunk;          // one operand tile: 32 rows x 128 bytes
constexpr int k8UStageBytes = k8Blocks * k8BlockBytes;
constexpr int k8UStages = 5, k8BStages = 4;
constexpr int k8Ring = 8;                            // accumulator groups in TMEM: twice the chunk window
constexpr int k8WinGroups = kBand8Window / kBand8Group;  // groups one chunk's MMAs touch
constexpr int k8Teams = 3;                           // epilogue teams (4 warps each)
constexpr int k8MmaWarps = 2;                        // MMA-issuing warps (warps 1 ..): each owns k8Blocks / k8MmaWarps blocks
constexpr int k8Group = kBand8Group;                 // output rows per group
constexpr int k8Window = kBand8Window;               // output rows in the ring
constexpr int k8TileRows = 16;                       // intermediate rows per horizontal phase (two groups)
constexpr int k8Threads = 128 + k8Teams * 128;
constexpr int k8Segs = 8;
constexpr int k8HeaderBytes = 1024;                  // mbarriers; the operand stages behind it stay 1024-byte aligned (swizzle atoms)
constexpr size_t k8MaxSmem = 226 * 1024;
constexpr int k8RegsIo = 32, k8RegsEpi = 160;        // 128 x 32 + 384 x 160 = 512 x 128
constexpr int k8TmpPad = 64;

__host__ __device__ constexpr int tmp8_pitch_floats(int channels) {
    return channels == 4 ? k8StripBytes + 4 : channels == 2 ? k8StripBytes + 2 : k8StripBytes + 1;
}

// Shared-memory descriptor of the A tile: MN-major, 128-byte swizzle, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint32_t a8_desc_hi() { return (1024u >> 4) | (1u << 14) | (2u << 29); }  // SBO, version, SWIZZLE_128B
// Instruction descriptor: D = s32, A = u8 (MN-major), B = s8 (K-major), M = 128, N = n.
__device__ __forceinline__ uint32_t instr_desc_i8(uint32_t n) {
    return (2u << 4) | (0u << 7) | (1u << 10) | (1u << 15) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void mma_i8_acc(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(1u)
        : "memory");
}
// N consecutive TMEM columns of this thread's lane in one instruction.
template <int N>
__device__ __forceinline__ void tmem_ld_n(uint32_t addr, int (&v)[N]) {
    static_assert(N == 16, "one group of one block: 8 outputs x 2 digits");
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(addr));
}
template <int N>
__device__ __forceinline__ void tmem_zero_n(uint32_t addr) {
    static_assert(N == 16, "one group of one block (L = 2)");
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(addr), "r"(0u)
                 : "memory");
}
__device__ __forceinline__ void tmem_zero8(uint32_t addr) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(addr), "r"(0u) : "memory");
}

// Barrier over one epilogue team (named barriers 1 .. k8Teams).
__device__ __forceinline__ void team_barrier(int team) { asm volatile("bar.sync %0, 128;" ::"r"(team + 1) : "memory"); }


// One accumulator group (8 output rows x the strip's blocks) of this thread's TMEM lane: TMEM -> registers, the ring slot
// is zeroed and handed back right away (the stores retire while the digits are recombined), digits -> f32 -> the
// intermediate tile's column of this lane.  FULL: the strip has all k8Blocks blocks (no per-block predicates).
template <int L, int PITCH, bool FULL>
__device__ __forceinline__ void drain_group8(uint32_t taddr, uint32_t trow, int nblk, bool live, uint32_t empty_bar, int lane) {
    constexpr int kBlockCols = L * k8Ring * k8Group;
    if (live) {
        int v[k8Blocks][8 * L];
#pragma unroll
        for (int b = 0; b < k8Blocks; ++b)
            if (FULL || b < nblk) tmem_ld_n<8 * L>(taddr + uint32_t(b * kBlockCols), v[b]);
        tmem_ld_wait();
#pragma unroll
        for (int b = 0; b < k8Blocks; ++b)
            if (FULL || b < nblk) tmem_zero_n<8 * L>(taddr + uint32_t(b * kBlockCols));
#pragma unroll
        for (int b = 0; b < k8Blocks; ++b) {
            if (FULL || b < nblk) {
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    int t = v[b][r * L];                                        // most significant digit first
#pragma unroll
                    for (int d = 1; d < L; ++d) t = t * kBand8Base + v[b][r * L + d];
                    sts_f32(trow + uint32_t(r * PITCH + b * 128) * 4, __int2float_rn(t));
                }
            }
        }
    } else {
#pragma unroll
        for (int b = 0; b < k8Blocks; ++b)
            if (FULL || b < nblk) tmem_zero_n<8 * L>(taddr + uint32_t(b * kBlockCols));
    }
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive_at(empty_bar);
}

// One work item's geometry, as every role of the persistent CTA needs it.
struct Item8 {
    const DevJob* J;
    const int32_t* gbase;
    int ox0, ox1, oy0, oy1;
    int b0, nblk;          // strip along x: source bytes [b0, b0 + nblk * 128), 16-byte aligned at both ends
    int k0, nchunks;       // chunks [k0, k0 + nchunks) of the pass's global chunk grid
    int g0, g_end;         // groups (of 8 outputs) the item's MMAs touch: [g0, g_end)
};
template <int C>
__device__ __forceinline__ Item8 load_item8(const DevJob* __restrict__ jobs, const WorkItem* __restrict__ items, int idx) {
    const WorkItem it = items[idx];
    Item8 r;
    r.J = jobs + it.job;
    r.ox0 = it.ox0; r.ox1 = it.ox1; r.oy0 = it.oy0; r.oy1 = it.oy1;
    r.gbase = r.J->v.band8_gbase;
    const int xl = __ldg(r.J->h.left + it.ox0);
    const int xr = __ldg(r.J->h.right + it.ox1 - 1);
    const int row_bytes = int(r.J->sw) * C;
    r.b0 = (xl * C) & ~15;
    const int b1 = min((xr * C + 15) & ~15, (row_bytes + 15) & ~15);
    r.nblk = (b1 - r.b0 + 127) >> 7;
    const int y_first = __ldg(r.J->v.left + it.oy0);
    const int y_last = __ldg(r.J->v.right + it.oy1 - 1);
    r.k0 = y_first / k8Chunk;
    const int k1 = (y_last - 1) / k8Chunk;
    r.nchunks = k1 - r.k0 + 1;
    r.g0 = __ldg(r.gbase + r.k0);
    r.g_end = __ldg(r.gbase + k1) + k8WinGroups;
    return r;
}

constexpr bool k8Conv = IKC_BANDED8_CONV != 0;

}  // namespace

// Shared memory: [mbarriers (1 KB) | operand ring: 5 stages x 4 blocks x (32 rows x 128 B, swizzled) | weight-tile ring
//                (4 x L * 1 KB) | 2 x horizontal weights of a strip | 2 x (left, right) of a strip's outputs |
//                one intermediate tile per team: 16 rows x pitch floats]
template <int C, int L, bool CONV>
__global__ void __launch_bounds__(k8Threads, 1)
banded8_kernel(const DevJob* __restrict__ jobs, const WorkItem* __restrict__ items, const Band8Geom geom) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int kTmpPitch = tmp8_pitch_floats(C);
    constexpr int kN = L * k8Window;                // MMA N: the columns of one chunk window
    constexpr int kBlockCols = L * k8Ring * k8Group; // TMEM columns per block: the ring
    constexpr int kTmemCols = k8Blocks * kBlockCols; // 512 for L = 2
    constexpr int kTileFloats = k8TileRows * kTmpPitch + k8TmpPad;
    constexpr uint32_t kBTile = uint32_t(kN) * k8Chunk;  // bytes of one chunk's weight tile

    uint64_t* const bars = reinterpret_cast<uint64_t*>(smem);
    uint64_t* const u_full = bars;                   // [k8UStages] source boxes landed (tx bytes)
    uint64_t* const u_empty = u_full + k8UStages;    // [k8UStages] the MMAs that read the stage have completed
    uint64_t* const b_full = u_empty + k8UStages;    // [k8BStages]
    uint64_t* const b_empty = b_full + k8BStages;    // [k8BStages]
    uint64_t* const t_full = b_empty + k8BStages;    // [k8Ring] every MMA into the group has completed
    uint64_t* const t_empty = t_full + k8Ring;       // [k8Ring] the 4 warps of a team have drained and zeroed the group
    uint64_t* const tab_full = t_empty + k8Ring;     // [2] the strip tables of an item are in their buffer (table warp)
    uint64_t* const tab_empty = tab_full + 2;        // [2] every epilogue warp has left the item that used the buffer
    uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(smem + 512);
    uint8_t* const ustage = smem + k8HeaderBytes;
    uint8_t* const bstage = ustage + k8UStages * k8UStageBytes;
    const int hw_cap = (geom.hw_pairs + 1) & ~1, hlr_cap = (geom.max_out + 1) & ~1;
    float2* const hw_all = reinterpret_cast<float2*>(bstage + k8BStages * kBTile);       // [2][hw_cap]: item k uses buffer k & 1
    int2* const hlr_all = reinterpret_cast<int2*>(hw_all + 2 * hw_cap);                  // [2][hlr_cap]
    float* const tmp_all = reinterpret_cast<float*>(hlr_all + 2 * hlr_cap);              // k8Teams tiles

    int tid;  // read once: the compiler otherwise re-reads %tid.x (a ~20-cycle S2R) inside the epilogue loops
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
    const int warp = tid >> 5;
    const int lane = tid & 31;

    // Persistent CTA: items blockIdx.x, blockIdx.x + gridDim.x, ... ; every role walks the same sequence.  Ring slots and
    // barrier phases follow a running group index (the item's group index + the groups of the CTA's earlier items), so the
    // producer and the MMA warps run ahead into the next item while the epilogue teams finish the current one.
    const int n_items = geom.n_items;

    if (tid == 0) {
        if (smem_addr(smem) & 1023u) __trap();  // the swizzled operand tiles need 1024-byte aligned shared memory
        for (int s = 0; s < k8UStages; ++s) { mbar_init(u_full + s, 1); mbar_init(u_empty + s, k8MmaWarps); }
        for (int s = 0; s < k8BStages; ++s) { mbar_init(b_full + s, 1); mbar_init(b_empty + s, k8MmaWarps); }
        for (int s = 0; s < k8Ring; ++s) { mbar_init(t_full + s, k8MmaWarps); mbar_init(t_empty + s, 4); }
        for (int s = 0; s < 2; ++s) { mbar_init(tab_full + s, 1); mbar_init(tab_empty + s, k8Teams * 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(tmem_slot)), "n"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(k8RegsIo));
    if (warp == 0) {
        // ------------------------------------------------------------------------------ producer
        // (the whole warp runs the loop and polls; one elected lane issues the copies)
        const bool leader = elect_one();
        int su = 0, sb = 0;
        uint32_t pu = 1, pb = 1;  // phase bits of the empty barriers (first pass: free)
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const Item8 I = load_item8<C>(jobs, items, item);
            const void* const src_map = I.J->src_map8;
            if (leader) asm volatile("prefetch.tensormap [%0];" ::"l"(src_map) : "memory");
            const uint8_t* const tiles = reinterpret_cast<const uint8_t*>(I.J->v.band8_tiles) + size_t(I.k0) * kBTile;
            for (int i = 0; i < I.nchunks; ++i) {
                mbar_wait_parked(u_empty + su, pu);
                if (leader) {
                    // one box per 128-byte block: 32 rows x 128 bytes, swizzled into the MMA's operand layout; bytes past the
                    // raster's pitch or rows read as zero
                    mbar_expect_tx(u_full + su, uint32_t(I.nblk) * k8BlockBytes);
                    for (int b = 0; b < I.nblk; ++b)
                        tma_load_2d(ustage + su * k8UStageBytes + b * k8BlockBytes, src_map, I.b0 + b * 128, (I.k0 + i) * k8Chunk, u_full + su);
                }
                mbar_wait_parked(b_empty + sb, pb);
                if (leader) {
                    mbar_expect_tx(b_full + sb, kBTile);
                    bulk_load(bstage + sb * kBTile, tiles + size_t(i) * kBTile, kBTile, b_full + sb);
                }
                if (++su == k8UStages) { su = 0; pu ^= 1; }
                if (++sb == k8BStages) { sb = 0; pb ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp <= k8MmaWarps) {
        // ------------------------------------------------------------------------------ MMA issuers
        // The issue loop is one warp's latency-bound instruction stream, so the strip's blocks are split over k8MmaWarps
        // issuing warps; every hand-off barrier they commit to counts one arrival per issuer.
        // The whole warp runs the loop (warp-uniform control flow: the descriptor arithmetic stays on the uniform datapath,
        // every lane polls the barriers) and one elected lane issues the MMAs and commits.
        const int b_lo_blk = (warp - 1) * (k8Blocks / k8MmaWarps);
        const bool leader = elect_one();
        const uint32_t a_lo0 = ((smem_addr(ustage) >> 4) & 0x3fffu) | ((1024u >> 4) << 16);
        const uint32_t b_lo0 = ((smem_addr(bstage) >> 4) & 0x3fffu) | (((uint32_t(kN) * 16u) >> 4) << 16);
        constexpr uint32_t kBDescHi = (128u >> 4) | (1u << 14);
        const uint32_t bars_a = smem_addr(bars);
        int su = 0, sb = 0;
        uint32_t pu = 0, pb = 0;  // phase bits of the two operand rings
        int vbase = 0;            // groups of the CTA's earlier items
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const Item8 I = load_item8<C>(jobs, items, item);
            const int32_t* __restrict__ gbase = I.gbase;
            const int k0 = I.k0, nchunks = I.nchunks, nblk = I.nblk, g0 = I.g0;
            const int voff = vbase - g0;   // group g of this item is running group g + voff: slot (g + voff) & 7, use ((g + voff) >> 3)
            int acquired = g0;   // groups [g0, acquired) belong to the MMAs (zeroed by the epilogue warps)
            int completed = g0;  // groups [g0, completed) have been committed to the epilogue
            int gb_next = g0;
            for (int i = 0; i < nchunks; ++i) {
                const int gb = gb_next;
                gb_next = (i + 1 < nchunks) ? __ldg(gbase + k0 + i + 1) : 0;
                while (acquired < gb + k8WinGroups) {
                    const int v = acquired + voff;
                    mbar_wait_at(bars_a + uint32_t(2 * k8UStages + 2 * k8BStages + k8Ring + (v & (k8Ring - 1))) * 8, (v >> 3) & 1);
                    ++acquired;
                }
                mbar_wait_at(bars_a + uint32_t(su) * 8, pu);
                mbar_wait_at(bars_a + uint32_t(2 * k8UStages + sb) * 8, pb);
                tc_fence_after();
                const uint32_t a_lo = a_lo0 + uint32_t(su) * (k8UStageBytes >> 4);
                const uint32_t b_lo = b_lo0 + uint32_t(sb) * (kBTile >> 4);
                const int s0 = (gb + voff) & (k8Ring - 1);
                const uint32_t n1 = uint32_t(min(k8WinGroups, k8Ring - s0) * k8Group * L), n2 = uint32_t(kN) - n1;  // n2 > 0: the window wraps
                const uint32_t id1 = instr_desc_i8(n1), id2 = instr_desc_i8(n2);
                const uint32_t c1 = tmem + uint32_t(s0 * k8Group * L);
                if (leader) {
#pragma unroll
                    for (int bb = 0; bb < k8Blocks / k8MmaWarps; ++bb) {
                        const int b = b_lo_blk + bb;
                        if (b < nblk) {
                            const uint64_t a_desc = make_u64(a_lo + uint32_t(b) * (k8BlockBytes >> 4), a8_desc_hi());
                            mma_i8_acc(c1 + uint32_t(b * kBlockCols), a_desc, make_u64(b_lo, kBDescHi), id1);
                            if (n2) mma_i8_acc(tmem + uint32_t(b * kBlockCols), a_desc, make_u64(b_lo + (n1 >> 3) * (128u >> 4), kBDescHi), id2);
                        }
                    }
                    tc_commit(u_empty + su);
                    tc_commit(b_empty + sb);
                }
                const int final_below = (i + 1 < nchunks) ? gb_next : acquired;  // groups below it get no more contributions
                for (; completed < final_below; ++completed)
                    if (leader) tc_commit(t_full + ((completed + voff) & (k8Ring - 1)));
                if (++su == k8UStages) { su = 0; pu ^= 1; }
                if (++sb == k8BStages) { sb = 0; pb ^= 1; }
            }
            vbase += I.g_end - g0;
        }
        __syncwarp();
    }
    else if (warp == 3) {
        // ------------------------------------------------------------------------------ table warp
        // Horizontal tables of the next item's strip -- (left, right) and the weights (x 2^-shift, duplicated for FFMA2) of
        // every output -- go into the buffer the epilogue teams are not using, while they work on the current item.
        int k = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++k) {
            const WorkItem it = items[item];
            const DevJob* __restrict__ J = jobs + it.job;
            const int buf = k & 1;
            mbar_wait_parked(tab_empty + buf, uint32_t(((k >> 1) & 1) ^ 1));   // (first use of a buffer: free)
            float2* const hw = hw_all + buf * hw_cap;
            int2* const hlr = hlr_all + buf * hlr_cap;
            const int n_out = it.ox1 - it.ox0;
            const int hstride = J->h.stride;
            const int32_t* __restrict__ hleft = J->h.left;
            const int32_t* __restrict__ hright = J->h.right;
            for (int i = lane; i < n_out; i += 32) hlr[i] = make_int2(__ldg(hleft + it.ox0 + i), __ldg(hright + it.ox0 + i));
            const float* __restrict__ wsrc = J->h.w + size_t(it.ox0) * hstride;
            const float unscale = __int_as_float((127 - J->v.band8_shift) << 23);  // 2^-shift: the vertical sums are integers x 2^shift
            const int n = n_out * hstride;
#pragma unroll 4
            for (int i = lane; i < n; i += 32) {
                const float w = __ldg(wsrc + i) * unscale;
                hw[i] = make_float2(w, w);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(tab_full + buf);
        }
    }
    // (warp 2, when there is one MMA warp only, has no role: it gave its registers up and waits for the teardown)
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(k8RegsEpi));
        // ------------------------------------------------------------------------------ epilogue + horizontal pass
        const int q = warp & 3;                                  // TMEM lane quarter this warp may touch
        const int team = (warp >> 2) - 1;                        // 0 .. k8Teams - 1
        const uint32_t tlane = tmem + (uint32_t(q * 32) << 16);
        if (team == 0) {  // hand every ring slot to the MMA warp, zeroed
            for (int c = 0; c < kTmemCols; c += 8) tmem_zero8(tlane + uint32_t(c));
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0)
                for (int s = 0; s < k8Ring; ++s) mbar_arrive(t_empty + s);
        }
        float* const tmp = tmp_all + team * kTileFloats;         // this team's intermediate tile
        // Loop invariants pinned in registers: opaque to the compiler, which otherwise re-derives them (S2R %tid.x, shifts,
        // multiplies) for every group and tile to save registers it has no other use for.
        const uint32_t t_full_a = pin_u32(smem_addr(t_full), lane), t_empty_a = pin_u32(smem_addr(t_empty), lane);
        const uint32_t tcol_a = pin_u32(smem_addr(tmp + q * 32 + lane), lane);   // this lane's column of the tile (drain)
        const uint32_t tlane_p = pin_u32(tlane, lane);
        const int lane_p = int(pin_u32(uint32_t(lane), lane));
        const int team_p = int(pin_u32(uint32_t(team), lane));

        const int hrow = lane & 15;
        const int seg = 2 * q + (lane >> 4);
        int vbase = 0;            // groups of the CTA's earlier items
        int kitem = 0;
        int tbase = 0;            // tiles of the CTA's earlier items, mod k8Teams
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++kitem) {
        const Item8 I = load_item8<C>(jobs, items, item);
        const DevJob* __restrict__ J = I.J;
        const int ox0 = I.ox0, ox1 = I.ox1, oy0 = I.oy0, oy1 = I.oy1, b0 = I.b0, nblk = I.nblk, g0 = I.g0, g_end = I.g_end;
        const int voff = vbase - g0;   // group g of this item is running group g + voff: slot (g + voff) & 7, use ((g + voff) >> 3)
        // the strip's tables: filled by the table warp (buffer = item parity)
        const int n_out = ox1 - ox0;
        const int hstride = J->h.stride;
        const int tbuf = kitem & 1;
        const float2* const hw = hw_all + tbuf * hw_cap;
        const int2* const hlr = hlr_all + tbuf * hlr_cap;
        mbar_wait_at(smem_addr(tab_full + tbuf), uint32_t((kitem >> 1) & 1));
        const int CO = CONV ? J->out_channels : C;
        uint8_t* const dst_base = J->dst;
        const size_t dst_pitch = J->dst_pitch;
        // Segments of the strip's outputs.  Rgba8 -> Rgba8 with a 16-byte aligned destination: segments start on absolute
        // multiples of four outputs (the first one may begin up to three outputs before the strip: computed, never stored),
        // so that every run of four finished pixels is one 16-byte store.
        const bool vec_store = C == 4 && !CONV && ((reinterpret_cast<uintptr_t>(dst_base) | dst_pitch) & 15) == 0;
        const int a0 = vec_store ? (ox0 & ~3) : ox0;
        int per = (ox1 - a0 + k8Segs - 1) / k8Segs;
        if (vec_store) per = (per + 3) & ~3;
        const int os_raw = a0 + seg * per;
        const int os = max(os_raw, ox0);
        const int oe = min(os_raw + per, ox1);
        const float* const my_row = tmp + hrow * kTmpPitch - b0;   // indexed by source byte column x * C + c
        const uint32_t my_row_a = pin_u32(smem_addr(tmp + hrow * kTmpPitch) - uint32_t(b0) * 4, lane);
        const uint32_t uw_a = pin_u32(smem_addr(hw) + uint32_t(max(os - ox0, 0) * 12) * 8, lane);   // the segment's first weight row
        const bool uni2 = os < oe && J->h.uni_step == 2 && hstride == 12 && os >= J->h.uni_lo && oe <= J->h.uni_hi;
        const bool full = nblk == k8Blocks;
        // uniform 2:1 stretch: first pixel of output os_raw (outputs before the strip included: the stretch is linear)
        const int px_first = os < oe ? hlr[os - ox0].x - 2 * (os - os_raw) : 0;

        // This team's tiles: every k8Teams-th pair (even group, its successor).  Groups are drained one by one, as soon as
        // they are final, so that their ring slots go back to the MMAs early.
        // (the assignment of tiles to teams continues across items -- running tile index mod k8Teams -- so that the odd tiles
        // of an item do not always fall to team 0)
        const int gt0 = g0 & ~1;
        const int first = (team - tbase + k8Teams) % k8Teams;
        tbase = (tbase + ((g_end - gt0 + 1) >> 1)) % k8Teams;
        for (int gt = gt0 + 2 * first; gt < g_end; gt += 2 * k8Teams) {
            const int tile_row0 = (gt >> 1) * k8TileRows;          // output row of the intermediate tile's first row
            const bool live = tile_row0 < oy1 && tile_row0 + k8TileRows > oy0;
            const int g_lo = max(gt, g0), g_hi = min(gt + 2, g_end);
            for (int g = g_lo; g < g_hi; ++g) {
                const int slot = (g + voff) & (k8Ring - 1);
                mbar_wait_at(t_full_a + slot * 8, ((g + voff) >> 3) & 1);
                tc_fence_after();
                const uint32_t taddr = tlane_p + uint32_t(slot * k8Group * L);   // the group's first column within block 0
                const uint32_t trow = tcol_a + uint32_t((g & 1) * k8Group * kTmpPitch) * 4;
                if (full) drain_group8<L, kTmpPitch, true>(taddr, trow, nblk, live, t_empty_a + slot * 8, lane_p);
                else drain_group8<L, kTmpPitch, false>(taddr, trow, nblk, live, t_empty_a + slot * 8, lane_p);
            }
            if (!live) continue;
            team_barrier(team_p);  // the whole tile is in shared memory

            const int orow = tile_row0 + hrow;
            const bool row_live = orow >= oy0 && orow < oy1;
            uint8_t* const my_dst = dst_base + size_t(orow) * dst_pitch;
            if (os < oe) {
                if (uni2) {
                    // Blocks of 8 outputs as straight-line code: output j of the block reads pixels 2j .. 2j + 11 of the 26
                    // the block spans.  Reads past the segment's last window stay inside the tile's padding and are never
                    // stored.  Taps are accumulated in ascending order (the reference's order), four chains in flight.
                    float2 uw[12];  // the stretch's 12 tap weights (duplicated pairs)
#pragma unroll
                    for (int t = 0; t < 12; ++t) uw[t] = lds_f32x2(uw_a + t * 8);
                    for (int o = os_raw; o < oe; o += 8) {
                        const float* px = my_row + (px_first + 2 * (o - os_raw)) * C;
                        const uint32_t px_a = my_row_a + uint32_t((px_first + 2 * (o - os_raw)) * C) * 4;
                        uint32_t word[8];
                        float4 p[18];   // two halves of 4 outputs: 18 pixels live at a time
#pragma unroll
                        for (int t = 0; t < 18; ++t) p[t] = C == 4 ? lds_f32x4(px_a + t * 16) : load_px<C>(px + t * C);
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            if (half == 1) {
#pragma unroll
                                for (int t = 0; t < 10; ++t) p[t] = p[t + 8];          // (register renaming, no moves once unrolled)
#pragma unroll
                                for (int t = 10; t < 18; ++t) p[t] = C == 4 ? lds_f32x4(px_a + (t + 8) * 16) : load_px<C>(px + (t + 8) * C);
                            }
                            float2 acc[4][2];  // four outputs at a time: eight accumulation chains in flight
#pragma unroll
                            for (int j = 0; j < 4; ++j) acc[j][0] = acc[j][1] = make_float2(kRoundBias, kRoundBias);
#pragma unroll
                            for (int t = 0; t < 12; ++t) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const float4 x = p[2 * j + t];
                                    acc[j][0] = __ffma2_rn(uw[t], make_float2(x.x, x.y), acc[j][0]);
                                    acc[j][1] = __ffma2_rn(uw[t], make_float2(x.z, x.w), acc[j][1]);
                                }
                            }
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                word[4 * half + j] = pack_pixel(make_float4(acc[j][0].x, acc[j][0].y, acc[j][1].x, acc[j][1].y));
                        }
                        if (row_live) {
                            uint8_t* const d = my_dst + ptrdiff_t(o) * CO;
#pragma unroll
                            for (int half = 0; half < 2; ++half) {
                                const int oh = o + 4 * half;
                                if (vec_store && oh >= ox0 && oh + 4 <= oe) {
                                    *reinterpret_cast<uint4*>(d + 16 * half) = make_uint4(word[4 * half], word[4 * half + 1], word[4 * half + 2], word[4 * half + 3]);
                                } else {
#pragma unroll
                                    for (int j = 0; j < 4; ++j)
                                        if (oh + j >= ox0 && oh + j < oe) store_word<C>(d + (4 * half + j) * CO, word[4 * half + j], CO);
                                }
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
            team_barrier(team_p);  // the tile may be overwritten
        }
        vbase += g_end - g0;
        __syncwarp();
        if (lane == 0) mbar_arrive(tab_empty + tbuf);   // this warp is done with the item's tables
        }   // items
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

#if !IKC_BANDED8_CONV
size_t banded8_smem_bytes(int channels, const Band8Geom& g) {
    const size_t tmp_pitch = size_t(tmp8_pitch_floats(channels));
    return size_t(k8HeaderBytes) + size_t(k8UStages) * k8UStageBytes + size_t(k8BStages) * size_t(g.limbs) * k8Window * k8Chunk +
           2 * (size_t((g.hw_pairs + 1) & ~1) * sizeof(float2) + size_t((g.max_out + 1) & ~1) * sizeof(int2)) +
           size_t(k8Teams) * (size_t(k8TileRows) * tmp_pitch + k8TmpPad) * sizeof(float);
}
size_t banded8_max_smem() { return k8MaxSmem; }
int banded8_max_src_bytes() { return k8StripBytes; }
int banded8_tile_rows() { return k8TileRows; }
bool banded8_supported(int channels, int limbs) { return channels >= 1 && channels <= 4 && limbs == 2; }
#endif

template <int C, int L>
static cudaError_t launch_one8(const DevJob* jobs, const WorkItem* items, const Band8Geom& geom, cudaStream_t stream) {
    const size_t smem = banded8_smem_bytes(C, geom);
    if (smem > k8MaxSmem) return cudaErrorInvalidValue;
    // always the planner-wide maximum: the attribute is shared by every thread launching on this device
    cudaError_t e = cudaFuncSetAttribute(banded8_kernel<C, L, k8Conv>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(k8MaxSmem));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(banded8_kernel<C, L, k8Conv>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    // persistent grid: one CTA per SM (the kernel's occupancy), each walking items blockIdx.x, + gridDim.x, ...
    int dev = 0, sms = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    const int grid = geom.n_items < sms ? geom.n_items : sms;
    if (grid <= 0) return cudaErrorInvalidConfiguration;
    banded8_kernel<C, L, k8Conv><<<grid, k8Threads, smem, stream>>>(jobs, items, geom);
    return cudaGetLastError();
}

#if IKC_BANDED8_CONV
cudaError_t launch_banded8_conv(int channels, const DevJob* jobs, const WorkItem* items, const Band8Geom& geom, cudaStream_t stream) {
#else
cudaError_t launch_banded8_conv(int channels, const DevJob* jobs, const WorkItem* items, const Band8Geom& geom, cudaStream_t stream);  // banded8_conv.cu

cudaError_t launch_banded8(int channels, bool convert, const DevJob* jobs, const WorkItem* items, const Band8Geom& geom,
                           cudaStream_t stream) {
    if (convert) return launch_banded8_conv(channels, jobs, items, geom, stream);
#endif
    if (geom.limbs != 2) return cudaErrorInvalidValue;
    switch (channels) {
        case 1: return launch_one8<1, 2>(jobs, items, geom, stream);
        case 2: return launch_one8<2, 2>(jobs, items, geom, stream);
        case 3: return launch_one8<3, 2>(jobs, items, geom, stream);
        case 4: return launch_one8<4, 2>(jobs, items, geom, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace ikc
