// Dev tool: tcgen05.mma kind::i8 issue/completion rates on sm_100a for the operand configurations of banded8.cu
// (A = u8 MN-major sw128, B = s8 K-major) and banded8t.cu (A = s8 K-major, B = u8 MN-major sw128), N = 64 .. 256.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tc_probe_i8 tools/tc_probe_i8.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t swz) {
    return uint64_t((addr >> 4) & 0x3fffu) | (uint64_t((lbo >> 4) & 0x3fffu) << 16) | (uint64_t((sbo >> 4) & 0x3fffu) << 32) | (uint64_t(1) << 46) |
           (uint64_t(swz) << 61);
}
// mode 0: A = s8 K-major (8 distinct tiles, 4 KB apart), B = u8 MN-major sw128 (one tile).   mode 1: A = u8 MN-major sw128, B = s8 K-major.
__global__ void __launch_bounds__(128, 1) rate_kernel(int mode, int N, int reps, int same_d, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (uint32_t i = tid; i < (64 * 1024) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        uint64_t da[8], db;
        uint32_t idesc;
        if (mode == 0) {
            for (int i = 0; i < 8; ++i) da[i] = make_desc(smem_u32(smem) + i * 4096, 2048, 128, 0);
            db = make_desc(smem_u32(smem) + 32768, 1024, 1024, 2);
            idesc = (2u << 4) | (1u << 7) | (0u << 10) | (1u << 16) | ((uint32_t(N) >> 3) << 17) | ((128u >> 4) << 24);
        } else {
            for (int i = 0; i < 8; ++i) da[i] = make_desc(smem_u32(smem) + i * 4096, 1024, 1024, 2);
            db = make_desc(smem_u32(smem) + 32768, uint32_t(N) * 16, 128, 0);
            idesc = (2u << 4) | (0u << 7) | (1u << 10) | (1u << 15) | ((uint32_t(N) >> 3) << 17) | ((128u >> 4) << 24);
        }
        const int nd = 512 / N;  // distinct accumulator tiles
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                asm volatile(
                    "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                    "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem + uint32_t(same_d ? 0 : (i % nd) * N)),
                    "l"(da[i]), "l"(db), "r"(idesc), "r"(1u));
            }
        }
        const long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
        asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
        const long long t2 = clock64();
        cycles[0] = t2 - t0;
        cycles[1] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}
// The banded8t pattern: groups of `group` MMAs (distinct A and B tiles, N = 128) into one accumulator, then the next
// accumulator (4 of 128 columns); optionally 4 other warps hammer tcgen05.ld on the tiles meanwhile.
__global__ void __launch_bounds__(256, 1) pattern_kernel(int group, int groups, int with_ld, long long* cycles, uint32_t* sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ volatile int done;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (uint32_t i = tid; i < (160 * 1024) / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
    if (tid == 0) {
        done = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = (2u << 4) | (1u << 7) | (0u << 10) | (1u << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
        const long long t0 = clock64();
        for (int g = 0; g < groups; ++g) {
            const uint32_t d = tmem + uint32_t((g & 3) * 128);
            for (int i = 0; i < group; ++i) {
                const uint64_t da = make_desc(smem_u32(smem) + ((g & 1) * 10 + i % 10) * 4096, 2048, 128, 0);        // 20 weight tiles
                const uint64_t db = make_desc(smem_u32(smem) + 81920 + (((g >> 1) & 1) * 10 + i % 10) * 4096, 1024, 1024, 2);  // 20 source tiles
                asm volatile(
                    "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                    "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(uint32_t(i > 0)));
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
        asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
        cycles[0] = clock64() - t0;
        done = 1;
    } else if (warp >= 4 && with_ld) {
        const uint32_t lane_base = tmem + (uint32_t((warp & 3) * 32) << 16);
        uint32_t acc = 0;
        int it = 0;
        while (!done) {
            uint32_t v[32];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) : "r"(lane_base + uint32_t((it++ & 15) * 32)));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) acc ^= v[j];
            if (with_ld > 1) __nanosleep(with_ld);
        }
        sink[tid] = acc;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}
// tcgen05.ld throughput, two x32 loads in flight per wait, 4 or 8 warps.
__global__ void __launch_bounds__(256, 1) ldtm_kernel(int iters, long long* cycles, uint32_t* sink) {
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    const uint32_t lane_base = tmem + (uint32_t((warp & 3) * 32) << 16) + uint32_t((warp >> 2) * 256);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint32_t v[32], w[32];
#define LD32(arr, addr) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
            : "=r"(arr[0]), "=r"(arr[1]), "=r"(arr[2]), "=r"(arr[3]), "=r"(arr[4]), "=r"(arr[5]), "=r"(arr[6]), "=r"(arr[7]), "=r"(arr[8]), "=r"(arr[9]), "=r"(arr[10]), "=r"(arr[11]), "=r"(arr[12]), "=r"(arr[13]), "=r"(arr[14]), "=r"(arr[15]), "=r"(arr[16]), "=r"(arr[17]), "=r"(arr[18]), "=r"(arr[19]), "=r"(arr[20]), "=r"(arr[21]), "=r"(arr[22]), "=r"(arr[23]), "=r"(arr[24]), "=r"(arr[25]), "=r"(arr[26]), "=r"(arr[27]), "=r"(arr[28]), "=r"(arr[29]), "=r"(arr[30]), "=r"(arr[31]) : "r"(addr))
        LD32(v, lane_base + uint32_t((it & 3) * 64));
        LD32(w, lane_base + uint32_t((it & 3) * 64 + 32));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= v[j] ^ w[j];
    }
    __syncthreads();
    const long long t1 = clock64();
    if (tid == 0) cycles[0] = t1 - t0;
    sink[tid] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}
int main() {
    CHECK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    long long* d_cyc; uint32_t* d_sink;
    CHECK(cudaMalloc(&d_cyc, 16)); CHECK(cudaMalloc(&d_sink, 256 * 4));
    for (int mode = 0; mode < 2; ++mode)
        for (int same_d = 0; same_d < 2; ++same_d)
            for (int N : {32, 64, 128, 256}) {
                const int reps = 128;
                rate_kernel<<<1, 128, 64 * 1024>>>(mode, N, reps, same_d, d_cyc);
                CHECK(cudaDeviceSynchronize());
                long long cyc[2];
                CHECK(cudaMemcpy(cyc, d_cyc, 16, cudaMemcpyDeviceToHost));
                printf("i8 mma %s, %s D, M=128 N=%d K=32: %.1f cycles/MMA to completion, %.1f to issue (%.0f MAC/clk)\n",
                       mode == 0 ? "A=s8 K-major, B=u8 MN-sw128" : "A=u8 MN-sw128, B=s8 K-major", same_d ? "same" : "rotating", N,
                       double(cyc[0]) / (8 * reps), double(cyc[1]) / (8 * reps), 128.0 * N * 32 * 8 * reps / double(cyc[0]));
            }
    CHECK(cudaFuncSetAttribute(pattern_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    for (int with_ld : {0, 1, 400})
        for (int group : {1, 2, 4, 9, 18, 36}) {
            const int groups = 1152 / group;
            pattern_kernel<<<1, 256, 160 * 1024>>>(group, groups, with_ld, d_cyc, d_sink);
            CHECK(cudaDeviceSynchronize());
            long long cyc = 0;
            CHECK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
            printf("i8 mma N=128, runs of %d into one accumulator (distinct A and B tiles), tcgen05.ld traffic %s: %.1f cycles/MMA\n", group,
                   with_ld == 0 ? "none" : with_ld == 1 ? "4 warps flat out" : "4 warps, paced", double(cyc) / (group * groups));
        }
    for (int threads : {128, 256}) {
        const int iters = 1024;
        ldtm_kernel<<<1, threads>>>(iters, d_cyc, d_sink);
        CHECK(cudaDeviceSynchronize());
        long long cyc = 0;
        CHECK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
        printf("tcgen05.ld 2 x 32x32b.x32 per wait, %d warps x %d rounds: %lld cycles -> %.1f B/clk/SM\n", threads / 32, iters, cyc,
               double(threads) * 64 * 4 * iters / double(cyc));
    }
    printf("tc_probe_i8 done\n");
    return 0;
}
