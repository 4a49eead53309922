// Dev tool: FFMA2 (fma.rn.f32x2) issue cost on sm_100a for a few operand patterns, against scalar FFMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_ffma2 tools/microbench_ffma2.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#define ITER 4096
// 1: every FFMA2 multiplies the same two pairs (operand reuse cache can serve both)
__global__ void k_same(float* out, float a) {
    float2 r[24]; float2 w0 = make_float2(a, a + 1), w1 = make_float2(a + 2, a + 3);
    for (int i = 0; i < 24; ++i) r[i] = make_float2(threadIdx.x + i, i);
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 24; ++i) r[i] = __ffma2_rn(w0, w1, r[i]);
    }
    float s = 0; for (int i = 0; i < 24; ++i) s += r[i].x + r[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 2: the ring kernel's pattern: acc[j][q] += w[j] * s[q], j-major
__global__ void k_ring_jq(float* out, float a) {
    float2 acc[6][4], w[6], s[4];
    for (int j = 0; j < 6; ++j) { w[j] = make_float2(a + j, a + j); for (int q = 0; q < 4; ++q) acc[j][q] = make_float2(j, q); }
    for (int q = 0; q < 4; ++q) s[q] = make_float2(threadIdx.x + q, a - q);
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int j = 0; j < 6; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[j][q] = __ffma2_rn(w[j], s[q], acc[j][q]);
    }
    float t = 0; for (int j = 0; j < 6; ++j) for (int q = 0; q < 4; ++q) t += acc[j][q].x + acc[j][q].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
// 3: same, q-major
__global__ void k_ring_qj(float* out, float a) {
    float2 acc[6][4], w[6], s[4];
    for (int j = 0; j < 6; ++j) { w[j] = make_float2(a + j, a + j); for (int q = 0; q < 4; ++q) acc[j][q] = make_float2(j, q); }
    for (int q = 0; q < 4; ++q) s[q] = make_float2(threadIdx.x + q, a - q);
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < 6; ++j) acc[j][q] = __ffma2_rn(w[j], s[q], acc[j][q]);
    }
    float t = 0; for (int j = 0; j < 6; ++j) for (int q = 0; q < 4; ++q) t += acc[j][q].x + acc[j][q].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
// 4: scalar FFMA, same arithmetic (48 FMAs per trip)
__global__ void k_ring_scalar(float* out, float a) {
    float acc[6][8], w[6], s[8];
    for (int j = 0; j < 6; ++j) { w[j] = a + j; for (int q = 0; q < 8; ++q) acc[j][q] = j + q; }
    for (int q = 0; q < 8; ++q) s[q] = threadIdx.x + q * a;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int j = 0; j < 6; ++j)
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[j][q] = fmaf(w[j], s[q], acc[j][q]);
    }
    float t = 0; for (int j = 0; j < 6; ++j) for (int q = 0; q < 8; ++q) t += acc[j][q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
// 5: FFMA2 where the multiplier pair is built from one scalar weight by the instruction's operand
//    selection is not expressible in CUDA C; instead: weights as scalars duplicated by MOV each trip
__global__ void k_ring_scalar_half(float* out, float a) {  // 24 FFMA2 + 24 FFMA interleaved
    float2 acc2[6][2], w2[6], s2[2]; float acc[6][4], w[6], s[4];
    for (int j = 0; j < 6; ++j) { w[j] = a + j; w2[j] = make_float2(a + j, a + j);
        for (int q = 0; q < 4; ++q) acc[j][q] = j + q; for (int q = 0; q < 2; ++q) acc2[j][q] = make_float2(j, q); }
    for (int q = 0; q < 4; ++q) s[q] = threadIdx.x + q * a;
    for (int q = 0; q < 2; ++q) s2[q] = make_float2(threadIdx.x + q, a - q);
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int j = 0; j < 6; ++j) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                acc2[j][q] = __ffma2_rn(w2[j], s2[q], acc2[j][q]);
                acc[j][2 * q] = fmaf(w[j], s[2 * q], acc[j][2 * q]);
                acc[j][2 * q + 1] = fmaf(w[j], s[2 * q + 1], acc[j][2 * q + 1]);
            }
        }
    }
    float t = 0; for (int j = 0; j < 6; ++j) { for (int q = 0; q < 4; ++q) t += acc[j][q]; for (int q = 0; q < 2; ++q) t += acc2[j][q].x + acc2[j][q].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
template <typename F> static double time_ms(F launch) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    launch(); launch(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) { cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    for (int threads : {128, 256, 512}) {
        const int sms = p.multiProcessorCount, blocks = sms * (1024 / threads) * 2;
        float* out; cudaMalloc(&out, sizeof(float) * blocks * threads);
        auto report = [&](const char* name, double ms, double fma_per_thread_iter) {
            const double ops = double(blocks) * threads * ITER * fma_per_thread_iter;
            printf("threads %3d  %-34s %8.3f ms  %7.1f FMA lanes/clk/SM\n", threads, name, ms, ops / (ms * 1e-3) / (clk_khz * 1e3) / sms);
        };
        report("FFMA2 same operands", time_ms([&] { k_same<<<blocks, threads>>>(out, 1.0001f); }), 48);
        report("FFMA2 ring pattern j-major", time_ms([&] { k_ring_jq<<<blocks, threads>>>(out, 1.0001f); }), 48);
        report("FFMA2 ring pattern q-major", time_ms([&] { k_ring_qj<<<blocks, threads>>>(out, 1.0001f); }), 48);
        report("FFMA scalar ring pattern", time_ms([&] { k_ring_scalar<<<blocks, threads>>>(out, 1.0001f); }), 48);
        report("FFMA2 + FFMA half and half", time_ms([&] { k_ring_scalar_half<<<blocks, threads>>>(out, 1.0001f); }), 48);
        cudaFree(out);
    }
    return 0;
}
