// tc_probe.cu -- dev tool: pins down the tcgen05 facts the banded-GEMM vertical pass depends on, on a
// real B200 (results go to profiles/).  Not product code.
//   1. shared-memory operand layouts: A = MN-major f16 (data: M = byte columns, K = source rows),
//      B = K-major f16 (weights: N = output rows), no-swizzle ("interleave") and 128-byte swizzle;
//      which of LBO / SBO is which stride.  Method: pack random small integers under a hypothesis, run one
//      MMA, compare D with the exact product on the host.
//   2. f16 denormal inputs (a byte dropped into an f16 is b * 2^-24), accumulation of two MMAs (hi + lo
//      weights) into one f32 accumulator, accumulator column offsets that are not multiples of N.
//   3. cycles per MMA for M = 128, K = 16, N in {16, 32, 48, 64} issued back to back (is a narrow MMA
//      bound by reading A from shared memory?), and tcgen05.ld throughput.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

struct Case {
    uint32_t a_bytes, b_bytes;      // bytes of the operand images copied to shared memory
    uint32_t a_lbo, a_sbo, a_type;  // descriptor fields (bytes; type: 0 none, 2 = 128B swizzle)
    uint32_t b_lbo, b_sbo, b_type;
    uint32_t idesc;
    uint32_t n_mma;                 // MMAs issued; operand start addresses advance by a_step / b_step bytes
    uint32_t a_step, b_step;
    uint32_t d_col;                 // accumulator column offset inside the 128 allocated columns
    uint32_t reps;                  // timing mode: the whole n_mma sequence is repeated this many times
};

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t type) {
    uint64_t d = 0;
    d |= uint64_t((addr >> 4) & 0x3fff);
    d |= uint64_t((lbo >> 4) & 0x3fff) << 16;
    d |= uint64_t((sbo >> 4) & 0x3fff) << 32;
    d |= uint64_t(1) << 46;  // descriptor version (Blackwell)
    d |= uint64_t(type & 7) << 61;
    return d;
}

__global__ void __launch_bounds__(128, 1) probe_kernel(const uint8_t* a_img, const uint8_t* b_img, Case c, float* d_out,
                                                       long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t* sa = smem;
    uint8_t* sb = smem + ((c.a_bytes + 1023) & ~1023u);
    const int tid = threadIdx.x, warp = tid >> 5;

    for (uint32_t i = tid; i < c.a_bytes / 4; i += 128) reinterpret_cast<uint32_t*>(sa)[i] = reinterpret_cast<const uint32_t*>(a_img)[i];
    for (uint32_t i = tid; i < c.b_bytes / 4; i += 128) reinterpret_cast<uint32_t*>(sb)[i] = reinterpret_cast<const uint32_t*>(b_img)[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    const uint32_t lane_base = tmem + (uint32_t(warp * 32) << 16);

    // zero the 128 accumulator columns
    for (int cb = 0; cb < 128; cb += 16) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(lane_base + cb), "r"(0u));
    }
    asm volatile("tcgen05.wait::st.sync.aligned;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");

    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        t0 = clock64();
        for (uint32_t r = 0; r < c.reps; ++r) {
            for (uint32_t i = 0; i < c.n_mma; ++i) {
                const uint64_t da = make_desc(smem_u32(sa) + i * c.a_step, c.a_lbo, c.a_sbo, c.a_type);
                const uint64_t db = make_desc(smem_u32(sb) + i * c.b_step, c.b_lbo, c.b_sbo, c.b_type);
                asm volatile(
                    "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem + c.d_col),
                    "l"(da), "l"(db), "r"(c.idesc), "r"(1u));
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
    }
    {   // everyone waits for the MMAs
        asm volatile(
            "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(
                smem_u32(&bar))
            : "memory");
    }
    if (tid == 0) { t1 = clock64(); cycles[0] = t1 - t0; }
    asm volatile("tcgen05.fence::after_thread_sync;");

    // read all 128 columns back: thread = lane
    long long t2 = clock64();
    for (int cb = 0; cb < 128; cb += 32) {
        uint32_t v[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
              "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(lane_base + cb));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) d_out[size_t(tid) * 128 + cb + j] = __uint_as_float(v[j]);
    }
    long long t3 = clock64();
    if (tid == 0) cycles[1] = t3 - t2;

    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem));
}

// MMA issue rate with everything precomputed: 8 A tiles (4 KB apart), one B tile, fully unrolled.
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(uint32_t idesc, uint32_t a_lbo, uint32_t a_sbo, uint32_t a_type, uint32_t b_lbo,
                                                          int reps, int same_a, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (uint32_t i = tid; i < (40 * 1024) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        uint64_t da[8];
        for (int i = 0; i < 8; ++i) da[i] = make_desc(smem_u32(smem) + (same_a ? (i >> 1) : i) * 4096, a_lbo, a_sbo, a_type);
        const uint64_t db = make_desc(smem_u32(smem) + 32768, b_lbo, 128, 0);
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                asm volatile(
                    "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem + uint32_t((i & 3) * 64)),
                    "l"(da[i]), "l"(db), "r"(idesc), "r"(1u));
            }
        }
        const long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
        asm volatile(
            "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(
                smem_u32(&bar))
            : "memory");
        const long long t2 = clock64();
        cycles[0] = t2 - t0;
        cycles[1] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

// tcgen05.ld throughput: 4 warps x `iters` loads of 32 columns each.
__global__ void __launch_bounds__(128, 1) ldtm_kernel(int iters, long long* cycles, uint32_t* sink) {
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    const uint32_t lane_base = tmem + (uint32_t(warp * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint32_t v[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
              "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(lane_base + uint32_t((it & 3) * 32)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= v[j];
    }
    __syncthreads();
    const long long t1 = clock64();
    if (tid == 0) cycles[0] = t1 - t0;
    sink[tid] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem));
}

// ---------------------------------------------------------------------------------------------- host
static uint16_t f2h(float f) { return __half_as_ushort(__float2half_rn(f)); }

static uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
    uint32_t d = 0;
    d |= 1u << 4;                    // D format f32
    d |= 0u << 7;                    // A f16
    d |= 0u << 10;                   // B f16
    d |= uint32_t(a_mn_major) << 15;
    d |= uint32_t(b_mn_major) << 16;
    d |= uint32_t(N >> 3) << 17;
    d |= uint32_t(M >> 4) << 24;
    return d;
}

// Offsets (bytes) of element (mn, k) under the hypothesised canonical layouts.
static uint32_t off_mn_none(int mn, int k, uint32_t mn_stride, uint32_t k_stride) {  // MN-major, no swizzle
    return uint32_t(mn % 8) * 2 + uint32_t(mn / 8) * mn_stride + uint32_t(k % 8) * 16 + uint32_t(k / 8) * k_stride;
}
static uint32_t off_mn_sw128(int mn, int k, uint32_t mn_stride, uint32_t k_stride) {  // MN-major, 128B swizzle
    const uint32_t chunk = uint32_t((mn % 64) / 8) ^ uint32_t(k % 8);
    return chunk * 16 + uint32_t(mn % 8) * 2 + uint32_t(k % 8) * 128 + uint32_t(mn / 64) * mn_stride + uint32_t(k / 8) * k_stride;
}
static uint32_t off_k_none(int mn, int k, uint32_t mn_stride, uint32_t k_stride) {  // K-major, no swizzle
    return uint32_t(mn % 8) * 16 + uint32_t(mn / 8) * mn_stride + uint32_t(k % 8) * 2 + uint32_t(k / 8) * k_stride;
}

struct Runner {
    uint8_t *d_a, *d_b;
    float* d_d;
    long long* d_cyc;
    std::vector<float> D;
    long long cyc[2];
    Runner() : D(128 * 128) {
        CHECK(cudaMalloc(&d_a, 1 << 16));
        CHECK(cudaMalloc(&d_b, 1 << 16));
        CHECK(cudaMalloc(&d_d, 128 * 128 * 4));
        CHECK(cudaMalloc(&d_cyc, 16));
        CHECK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    }
    bool run(const std::vector<uint8_t>& a, const std::vector<uint8_t>& b, Case c) {
        c.a_bytes = uint32_t(a.size());
        c.b_bytes = uint32_t(b.size());
        CHECK(cudaMemcpy(d_a, a.data(), a.size(), cudaMemcpyHostToDevice));
        CHECK(cudaMemcpy(d_b, b.data(), b.size(), cudaMemcpyHostToDevice));
        CHECK(cudaMemset(d_d, 0xff, 128 * 128 * 4));
        probe_kernel<<<1, 128, 100 * 1024>>>(d_a, d_b, c, d_d, d_cyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("   kernel failed: %s\n", cudaGetErrorString(e)); return false; }
        CHECK(cudaMemcpy(D.data(), d_d, D.size() * 4, cudaMemcpyDeviceToHost));
        CHECK(cudaMemcpy(cyc, d_cyc, 16, cudaMemcpyDeviceToHost));
        return true;
    }
};

int main() {
    cudaDeviceProp prop{};
    CHECK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s, %d SMs\n", prop.name, prop.multiProcessorCount);
    Runner R;
    srand(1234);
    const int M = 128, K = 16;

    // ---- 1. layout hypotheses: A MN-major (none / sw128), B K-major none; LBO/SBO roles
    for (int a_sw = 0; a_sw < 2; ++a_sw) {
        for (int a_swap = 0; a_swap < 2; ++a_swap) {
            for (int b_swap = 0; b_swap < 2; ++b_swap) {
                const int N = 32;
                std::vector<int> A(M * K), B(N * K);
                for (auto& v : A) v = rand() % 17 - 8;
                for (auto& v : B) v = rand() % 17 - 8;
                // A image
                uint32_t a_mn_stride, a_k_stride;
                std::vector<uint8_t> a_img, b_img;
                if (a_sw == 0) { a_mn_stride = 128; a_k_stride = 16 * 128; a_img.assign(4096, 0); }   // [k/8][m/8][k%8][m%8]
                else { a_mn_stride = 1024; a_k_stride = 2048; a_img.assign(4096, 0); }               // [k/8][m/64][k%8][swizzled 128 B]
                for (int m = 0; m < M; ++m)
                    for (int k = 0; k < K; ++k) {
                        const uint32_t o = a_sw ? off_mn_sw128(m, k, a_mn_stride, a_k_stride) : off_mn_none(m, k, a_mn_stride, a_k_stride);
                        const uint16_t h = f2h(float(A[m * K + k]));
                        std::memcpy(&a_img[o], &h, 2);
                    }
                // B image [k/8][n/8][n%8][k%8]
                const uint32_t b_n_stride = 128, b_k_stride = uint32_t(N / 8) * 128;
                b_img.assign(size_t(N) * K * 2, 0);
                for (int n = 0; n < N; ++n)
                    for (int k = 0; k < K; ++k) {
                        const uint16_t h = f2h(float(B[n * K + k]));
                        std::memcpy(&b_img[off_k_none(n, k, b_n_stride, b_k_stride)], &h, 2);
                    }
                Case c{};
                // hypothesis "normal": MN-major none: SBO = mn-group stride, LBO = k-group stride;
                //                      MN-major sw128: LBO = mn (64-element) stride, SBO = k-group stride;
                //                      K-major none: SBO = n-group stride, LBO = k-group stride
                uint32_t a_lbo = a_sw ? a_mn_stride : a_k_stride, a_sbo = a_sw ? a_k_stride : a_mn_stride;
                if (a_swap) std::swap(a_lbo, a_sbo);
                uint32_t b_lbo = b_k_stride, b_sbo = b_n_stride;
                if (b_swap) std::swap(b_lbo, b_sbo);
                c.a_lbo = a_lbo; c.a_sbo = a_sbo; c.a_type = a_sw ? 2 : 0;
                c.b_lbo = b_lbo; c.b_sbo = b_sbo; c.b_type = 0;
                c.idesc = make_idesc(M, N, 1, 0);
                c.n_mma = 1; c.reps = 1; c.d_col = 0;
                if (!R.run(a_img, b_img, c)) continue;
                int bad = 0;
                for (int m = 0; m < M; ++m)
                    for (int n = 0; n < N; ++n) {
                        int ref = 0;
                        for (int k = 0; k < K; ++k) ref += A[m * K + k] * B[n * K + k];
                        if (R.D[m * 128 + n] != float(ref)) ++bad;
                    }
                printf("layout A=%s a_swap=%d b_swap=%d : %s (%d / %d mismatches)\n", a_sw ? "MN-sw128" : "MN-none", a_swap,
                       b_swap, bad ? "FAIL" : "PASS", bad, M * N);
            }
        }
    }

    // Everything below uses the "normal" no-swizzle hypothesis.
    auto pack_a = [&](const std::vector<float>& A, std::vector<uint8_t>* img, bool raw_bits, const std::vector<uint16_t>* bits) {
        img->assign(4096, 0);
        for (int m = 0; m < M; ++m)
            for (int k = 0; k < K; ++k) {
                const uint16_t h = raw_bits ? (*bits)[m * K + k] : f2h(A[m * K + k]);
                std::memcpy(&(*img)[off_mn_none(m, k, 128, 2048)], &h, 2);
            }
    };
    auto pack_b = [&](const std::vector<float>& B, int N, std::vector<uint8_t>* img) {
        img->assign(size_t(N) * K * 2, 0);
        for (int n = 0; n < N; ++n)
            for (int k = 0; k < K; ++k) {
                const uint16_t h = f2h(B[n * K + k]);
                std::memcpy(&(*img)[off_k_none(n, k, 128, uint32_t(N / 8) * 128)], &h, 2);
            }
    };

    // ---- 2. denormal data, hi + lo weights accumulated into one D, column offsets, N in {16, 32, 48}
    for (int N : {16, 32, 48}) {
        for (uint32_t d_col : {0u, 4u, 8u, 16u, 40u}) {
            std::vector<uint16_t> bits(M * K);
            std::vector<float> Af(M * K);
            for (int i = 0; i < M * K; ++i) { bits[i] = uint16_t(rand() & 255); Af[i] = std::ldexp(float(bits[i]), -24); }
            // weights: w * 2^14 split into f16 hi + lo
            std::vector<float> W(N * K), Whi(N * K), Wlo(N * K);
            for (int i = 0; i < N * K; ++i) {
                W[i] = (float(rand()) / RAND_MAX - 0.3f) * 0.7f;
                const float s = W[i] * 16384.0f;
                Whi[i] = __half2float(__float2half_rn(s));
                Wlo[i] = __half2float(__float2half_rn(s - Whi[i]));
            }
            std::vector<uint8_t> a_img, bh, bl, b_img;
            pack_a(Af, &a_img, true, &bits);
            pack_b(Whi, N, &bh);
            pack_b(Wlo, N, &bl);
            b_img = bh;
            b_img.insert(b_img.end(), bl.begin(), bl.end());
            Case c{};
            c.a_lbo = 2048; c.a_sbo = 128; c.a_type = 0;
            c.b_lbo = uint32_t(N / 8) * 128; c.b_sbo = 128; c.b_type = 0;
            c.idesc = make_idesc(M, N, 1, 0);
            c.n_mma = 2; c.a_step = 0; c.b_step = uint32_t(bh.size()); c.reps = 1; c.d_col = d_col;
            if (!R.run(a_img, b_img, c)) continue;
            double max_err = 0, max_ref = 0;
            int outside = 0;
            for (int m = 0; m < M; ++m)
                for (int n = 0; n < 128; ++n) {
                    const int nn = n - int(d_col);
                    const float got = R.D[m * 128 + n];
                    if (nn < 0 || nn >= N) { if (got != 0.0f) ++outside; continue; }
                    double ref = 0;  // exact value of sum_k b * w (w in f32), scaled as the MMA sees it
                    for (int k = 0; k < K; ++k) ref += double(Af[m * K + k]) * (double(Whi[nn * K + k]) + double(Wlo[nn * K + k]));
                    max_err = std::fmax(max_err, std::fabs(double(got) - ref));
                    max_ref = std::fmax(max_ref, std::fabs(ref));
                }
            // in pixel units: the intermediate is ref * 2^10
            printf("denormal+hi/lo N=%d d_col=%u : max |err| = %.3g (in u8 units: %.3g), max |ref| = %.3g, nonzero outside window: %d\n", N,
                   d_col, max_err, max_err * 1024.0, max_ref, outside);
        }
    }

    // ---- 3. timing: back-to-back MMAs on one accumulator
    for (int N : {16, 32, 48, 64, 128}) {
        std::vector<float> Af(M * K, 1.0f), Bf(size_t(N) * K, 0.0f);
        std::vector<uint8_t> a_img, b_img;
        pack_a(Af, &a_img, false, nullptr);
        pack_b(Bf, N, &b_img);
        std::vector<uint8_t> a_big;
        for (int r = 0; r < 8; ++r) a_big.insert(a_big.end(), a_img.begin(), a_img.end());  // 8 distinct A tiles
        Case c{};
        c.a_lbo = 2048; c.a_sbo = 128; c.a_type = 0;
        c.b_lbo = uint32_t(N / 8) * 128; c.b_sbo = 128; c.b_type = 0;
        c.idesc = make_idesc(M, N, 1, 0);
        c.n_mma = 8; c.a_step = 4096; c.b_step = 0; c.reps = 64; c.d_col = 0;
        if (!R.run(a_big, b_img, c)) continue;
        printf("timing M=128 N=%d K=16: %lld cycles for %d MMAs = %.1f cycles/MMA (A bytes/MMA 4096, B bytes %d)\n", N, R.cyc[0],
               c.n_mma * c.reps, double(R.cyc[0]) / (c.n_mma * c.reps), N * K * 2);
    }
    CHECK(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024));
    for (int a_sw = 0; a_sw < 2; ++a_sw)
        for (int same_a = 0; same_a < 2; ++same_a)
            for (int N : {16, 32, 64, 128, 256}) {
                long long* d_cyc;
                CHECK(cudaMalloc(&d_cyc, 16));
                const int reps = 128;
                mma_rate_kernel<<<1, 128, 48 * 1024>>>(make_idesc(128, N, 1, 0), a_sw ? 1024 : 2048, a_sw ? 2048 : 128, a_sw ? 2 : 0,
                                                      uint32_t(N / 8) * 128, reps, same_a, d_cyc);
                CHECK(cudaDeviceSynchronize());
                long long cyc[2];
                CHECK(cudaMemcpy(cyc, d_cyc, 16, cudaMemcpyDeviceToHost));
                printf("mma rate (precomputed descriptors) A=%s %s M=128 N=%d K=16: %.1f cycles/MMA to completion, %.1f to issue\n",
                       a_sw ? "MN-sw128" : "MN-none", same_a ? "pairs share A" : "distinct A", N, double(cyc[0]) / (8 * reps),
                       double(cyc[1]) / (8 * reps));
                CHECK(cudaFree(d_cyc));
            }
    {
        long long* d_cyc; uint32_t* d_sink;
        CHECK(cudaMalloc(&d_cyc, 8));
        CHECK(cudaMalloc(&d_sink, 128 * 4));
        const int iters = 1024;
        ldtm_kernel<<<1, 128>>>(iters, d_cyc, d_sink);
        CHECK(cudaDeviceSynchronize());
        long long cyc = 0;
        CHECK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
        printf("tcgen05.ld 32x32b.x32, 4 warps x %d loads (+wait each): %lld cycles = %.1f cycles per round of 4 x 4 KB -> %.1f B/clk/SM\n",
               iters, cyc, double(cyc) / iters, 4.0 * 4096.0 * iters / double(cyc));
    }
    printf("tc_probe done\n");
    return 0;
}
