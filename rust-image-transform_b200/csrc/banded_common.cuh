// banded_common.cuh -- device helpers shared by the tensor-core resize kernels (banded.cu: f16 operands, banded8.cu:
// 8-bit operands): mbarrier / TMA / tcgen05 wrappers and the pixel quantise-and-store code.  sm_100a only.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace ikc {
namespace {

constexpr int kEpiThreads = 128;                   // the epilogue warpgroup (named barrier 1)

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
// The same on a precomputed 32-bit shared address (epilogue loops: keeps the address arithmetic out of the loop).
__device__ __forceinline__ void mbar_wait_at(uint32_t bar_addr, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(bar_addr),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_at(uint32_t bar_addr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_addr) : "memory");
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
// 2-D TMA tile load (SASS: UTMALDG): box at (x word, y row) of the tensor map -> shared memory, bytes counted on an mbarrier.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tensor_map, int x, int y, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_addr(smem_dst)),
                 "l"(tensor_map), "r"(x), "r"(y), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }

// One lane of the (converged) warp: the lane that issues tcgen05.mma / commit while the whole warp runs the loop.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}

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
__device__ __forceinline__ uint64_t make_u64(uint32_t lo, uint32_t hi) {
    uint64_t v;
    asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "r"(lo), "r"(hi));
    return v;
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

// Store one packed pixel (bytes: its C channels, then don't-care lanes) as `co` destination channels.
template <int C>
__device__ __forceinline__ void store_word(uint8_t* dst_px, uint32_t w, int co) {
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
template <int C>
__device__ __forceinline__ void store_pixel(uint8_t* dst_px, float4 v_plus_half, int co) {
    store_word<C>(dst_px, pack_pixel(v_plus_half), co);
}

// Pins a loop invariant in a register: the value comes back from a warp shuffle of the lane with itself, which ptxas
// neither folds nor re-executes, so it cannot re-derive the value (S2R %tid.x + shifts + multiplies) at every use.
__device__ __forceinline__ uint32_t pin_u32(uint32_t v, int lane) { return __shfl_sync(0xffffffffu, v, lane); }

// Shared-memory accesses on a 32-bit shared address held in one register (the epilogue loops pin their base addresses
// instead of letting the compiler re-derive 64-bit pointers).
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ float4 lds_f32x4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float2 lds_f32x2(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
    return v;
}

// One intermediate pixel (C floats at `p`) as a float4; missing channels read as zero.
template <int C>
__device__ __forceinline__ float4 load_px(const float* p) {
    if (C == 4) return *reinterpret_cast<const float4*>(p);
    if (C == 3) return make_float4(p[0], p[1], p[2], 0.0f);
    if (C == 2) { const float2 v = *reinterpret_cast<const float2*>(p); return make_float4(v.x, v.y, 0.0f, 0.0f); }
    return make_float4(p[0], 0.0f, 0.0f, 0.0f);
}

}  // namespace
}  // namespace ikc
