// Instruction-throughput microbenchmarks for sm_100a (dev tool; results go to profiles/).
// Each kernel runs ITER iterations of an unrolled body with 8-16 independent chains per thread;
// reported as lane-operations per clock per SM at the measured SM clock.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>
#include <vector>

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

constexpr int ITER = 4096;

__global__ void k_ffma(float* out, float a, float b) {
    float r[16];
    for (int i = 0; i < 16; ++i) r[i] = threadIdx.x + i;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = fmaf(r[i], a, b);
    }
    float s = 0; for (int i = 0; i < 16; ++i) s += r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma3(float* out, float a, float b) {  // 3 distinct register operands, no reuse of a/b
    float r[16], w[16];
    for (int i = 0; i < 16; ++i) { r[i] = threadIdx.x + i; w[i] = a + i; }
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = fmaf(w[i], w[(i + 5) & 15], r[i]);
    }
    float s = 0; for (int i = 0; i < 16; ++i) s += r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + b;
}
__global__ void k_ffma2(float* out, float a, float b) {
    float2 r[16], w[8];
    for (int i = 0; i < 16; ++i) r[i] = make_float2(threadIdx.x + i, i);
    for (int i = 0; i < 8; ++i) w[i] = make_float2(a + i, a - i);
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = __ffma2_rn(w[i & 7], w[(i + 3) & 7], r[i]);
    }
    float s = 0; for (int i = 0; i < 16; ++i) s += r[i].x + r[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + b;
}
__global__ void k_hfma2(float* out, float a, float b) {
    __half2 r[16];
    const __half2 ha = __float2half2_rn(a), hb = __float2half2_rn(b);
    for (int i = 0; i < 16; ++i) r[i] = __float2half2_rn(float(threadIdx.x + i));
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = __hfma2(r[i], ha, hb);
    }
    float s = 0; for (int i = 0; i < 16; ++i) s += __low2float(r[i]) + __high2float(r[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_prmt_fadd(float* out, uint32_t seed) {  // byte -> float via PRMT + FADD(imm)
    uint32_t w[4];
    float acc[16];
    for (int i = 0; i < 4; ++i) w[i] = seed * (threadIdx.x + i + 1);
    for (int i = 0; i < 16; ++i) acc[i] = 0;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float f = __uint_as_float(__byte_perm(w[i >> 2] + it, 0x4B000000u, 0x7440u + (i & 3))) - 8388608.0f;
            acc[i] = f;  // last value only; keeps PRMT+FADD alive via the sum below
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) w[i] ^= __float_as_uint(acc[i * 4]) & 0xff;
    }
    float s = 0; for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_i2f(float* out, uint32_t seed) {  // (float)(x & 0xff)
    uint32_t w[16];
    float acc[16];
    for (int i = 0; i < 16; ++i) { w[i] = seed * (threadIdx.x + i + 1); acc[i] = 0; }
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) { acc[i] = float((w[i] >> 8) & 0xffu); w[i] += __float_as_uint(acc[i]); }
    }
    float s = 0; for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dp4a(float* out, uint32_t seed) {
    int r[16]; uint32_t a = seed * (threadIdx.x + 1), b = seed ^ 0x01020304u;
    for (int i = 0; i < 16; ++i) r[i] = i;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = __dp4a(int(a + i), int(b), r[i]);
    }
    int s = 0; for (int i = 0; i < 16; ++i) s += r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = float(s);
}
__global__ void k_dp2a(float* out, uint32_t seed) {
    int r[16]; uint32_t a = seed * (threadIdx.x + 1), b = seed ^ 0x01020304u;
    for (int i = 0; i < 16; ++i) r[i] = i;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = __dp2a_lo(int(a + i), int(b), r[i]);
    }
    int s = 0; for (int i = 0; i < 16; ++i) s += r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = float(s);
}
// the resize inner-loop mix: per 8 bytes: 8 PRMT + 8 FADD + 24 FFMA2 (K=6)
__global__ void k_mix(float* out, uint32_t seed, float a) {
    float2 acc[6][4];
    float2 w[6];
    for (int j = 0; j < 6; ++j) { w[j] = make_float2(a + j, a + j); for (int q = 0; q < 4; ++q) acc[j][q] = make_float2(0.f, 0.f); }
    uint32_t d0 = seed * (threadIdx.x + 1), d1 = d0 ^ 0x9e3779b9u;
    for (int it = 0; it < ITER; ++it) {
        float f[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            f[i] = __uint_as_float(__byte_perm(d0, 0x4B000000u, 0x7440u + i)) - 8388608.0f;
            f[4 + i] = __uint_as_float(__byte_perm(d1, 0x4B000000u, 0x7440u + i)) - 8388608.0f;
        }
#pragma unroll
        for (int j = 0; j < 6; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[j][q] = __ffma2_rn(w[j], make_float2(f[2 * q], f[2 * q + 1]), acc[j][q]);
        d0 += 0x01010101u; d1 += 0x03010201u;
    }
    float s = 0; for (int j = 0; j < 6; ++j) for (int q = 0; q < 4; ++q) s += acc[j][q].x + acc[j][q].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double time_ms(F launch) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    launch(); launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CHECK(cudaGetDeviceProperties(&p, 0));
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount, threads = 256, blocks = sms * 8;
    float* out; CHECK(cudaMalloc(&out, sizeof(float) * blocks * threads));
    printf("device %s, %d SMs, max clock %.0f MHz; lane-ops per clk per SM assume the max clock\n", p.name, sms, clk_khz / 1e3);
    auto report = [&](const char* name, double ms, double ops_per_thread_iter) {
        const double total = double(blocks) * threads * ITER * ops_per_thread_iter;
        const double per_clk_sm = total / (ms * 1e-3) / (clk_khz * 1e3) / sms;
        printf("%-28s %8.3f ms  %7.1f lane-ops/clk/SM  (%.2f Tops/s)\n", name, ms, per_clk_sm, total / (ms * 1e-3) / 1e12);
    };
    report("FFMA (r=r*a+b)", time_ms([&] { k_ffma<<<blocks, threads>>>(out, 1.0001f, 0.5f); }), 16);
    report("FFMA (3 reg operands)", time_ms([&] { k_ffma3<<<blocks, threads>>>(out, 1.0001f, 0.5f); }), 16);
    report("FFMA2 (counted as 2 FMA)", time_ms([&] { k_ffma2<<<blocks, threads>>>(out, 1.0001f, 0.5f); }), 32);
    report("HFMA2 (counted as 2 FMA)", time_ms([&] { k_hfma2<<<blocks, threads>>>(out, 1.0001f, 0.5f); }), 32);
    report("PRMT+FADD byte->f32", time_ms([&] { k_prmt_fadd<<<blocks, threads>>>(out, 12345u); }), 16);
    report("I2F (float)((x>>8)&255)", time_ms([&] { k_i2f<<<blocks, threads>>>(out, 12345u); }), 16);
    report("DP4A (counted as 4 MAC)", time_ms([&] { k_dp4a<<<blocks, threads>>>(out, 12345u); }), 64);
    report("DP2A (counted as 2 MAC)", time_ms([&] { k_dp2a<<<blocks, threads>>>(out, 12345u); }), 32);
    report("resize mix (FMA only counted)", time_ms([&] { k_mix<<<blocks, threads>>>(out, 12345u, 0.25f); }), 48);
    CHECK(cudaDeviceSynchronize());
    return 0;
}
