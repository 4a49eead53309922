// tile.cu -- fused single-launch tile kernel: the path for every 8-bit resize that neither the ring kernel
// (fused.cu: downscales with ring size 6-7) nor the 2x upscale kernel (up2.cu) takes: other upscales,
// mixed up/down ratios, Nearest/Triangle/CatmullRom at mild ratios, 1-2 channel upscales.
//
// One CTA computes a TW x TH tile of one image of the batch.  The source footprint of the tile is
// staged into shared memory once as f32 (a zero-extended byte is the denormal b * 2^-149; the staged weights
// carry the compensating powers of two); the vertical pass (image 0.25.8
// vertical_sample) writes an f32 tmp tile [TH][footprint columns] to shared memory; the horizontal pass
// (horizontal_sample) reads it, clamps, rounds half away from zero and stores u8.  Same order of
// passes and same unclamped f32 intermediate as the reference; sums use FMA, hence |delta| <= 1.
// Output-stationary loops over the window taps: right for few taps (upscales: 2-7), acceptable for
// moderate downscales; large-ratio downscales belong to the ring kernel.
#include <cuda_runtime.h>

#include <cstdint>

#include "device_types.hpp"
#include "launch.hpp"

namespace ikc {
namespace {

constexpr int kTileThreads = 256;

// clamp + round half away from zero as trunc(v + 0.5) with saturation (see fused.cu: equals f32::round on
// [0, 255] except one float); the + 0.5 is the accumulators' initial value.  cvt to u8 saturates both ends.
template <typename T>
__device__ __forceinline__ uint32_t quantize_tile(float v_plus_half) {
    uint32_t q;
    if (sizeof(T) == 1) asm("cvt.rzi.u8.f32 %0, %1;" : "=r"(q) : "f"(v_plus_half));
    else asm("cvt.rzi.u16.f32 %0, %1;" : "=r"(q) : "f"(v_plus_half));
    return q;
}

}  // namespace

// geom.pitch_f: floats per staged row (footprint columns * channels, rounded up to 4)
// geom.max_src_rows / max_tile_rows / max_tile_cols: extents the shared memory was sized for
// geom.vstride / hstride: taps per output in the staged weight tables (max over the batch's jobs)
//
// Shared memory: [src f32: max_src_rows x pitch][tmp f32: max_tile_rows x pitch]
//                [v (left,count): max_tile_rows][h (left,count): max_tile_cols]
//                [v weights: max_tile_rows x vstride][h weights: max_tile_cols x hstride][out bytes: rows x out_pitch]
// T = uint8_t or uint16_t samples (Luma16 / LumaA16 / Rgb16 / Rgba16 rasters from the PNG decoder: same
// passes, clamp to 65535; a 16-bit sample is still a denormal float, so the same weight scaling applies).
template <typename T>
__global__ void __launch_bounds__(kTileThreads, 4)
tile_kernel(const DevJob* __restrict__ jobs, const WorkItem* __restrict__ items, const TileGeom geom) {
    extern __shared__ __align__(16) float tile_smem[];
    const WorkItem it = items[blockIdx.x];
    const DevJob& J = jobs[it.job];
    const int C = J.channels;
    const int CO = J.out_channels;  // == C, or 3 / 4: DynamicImage::to_rgb8() / to_rgba8() applied by the store
    const int ox0 = it.ox0, ox1 = it.ox1, oy0 = it.oy0, oy1 = it.oy1;
    const int tw = ox1 - ox0, th = oy1 - oy0;
    const int tid = threadIdx.x;

    const int32_t* __restrict__ vleft = J.v.left;
    const int32_t* __restrict__ vright = J.v.right;
    const int32_t* __restrict__ hleft = J.h.left;
    const int32_t* __restrict__ hright = J.h.right;

    const int sx0 = __ldg(hleft + ox0), sx1 = __ldg(hright + ox1 - 1);   // source columns [sx0, sx1)
    const int sy0 = __ldg(vleft + oy0), sy1 = __ldg(vright + oy1 - 1);   // source rows    [sy0, sy1)
    const int ncol = (sx1 - sx0) * C;                                    // staged values per row
    const int nrow = sy1 - sy0;
    const int pitch = geom.pitch_f;

    float* src_f = tile_smem;                                       // [nrow][pitch] source footprint as f32
    float* tmp_f = src_f + size_t(geom.max_src_rows) * pitch;       // [th][pitch]   vertical pass output
    int2* vwin = reinterpret_cast<int2*>(tmp_f + size_t(geom.max_tile_rows) * pitch);  // (first - sy0, taps)
    int2* hwin = vwin + geom.max_tile_rows;                                             // ((first - sx0) * C, taps)
    float* vw_s = reinterpret_cast<float*>(hwin + geom.max_tile_cols);                  // [th][vstride]
    float* hw_s = vw_s + size_t(geom.max_tile_rows) * geom.vstride;                     // [tw][hstride]
    T* out_s = reinterpret_cast<T*>(hw_s + size_t(geom.max_tile_cols) * geom.hstride);  // [th][out_pitch bytes]
    const int out_pitch = geom.out_pitch_b / int(sizeof(T));  // in samples

    // ---- stage windows + weights of the tile's output rows / columns
    for (int i = tid; i < th; i += kTileThreads) {
        const int first = __ldg(vleft + oy0 + i);
        vwin[i] = make_int2(first - sy0, __ldg(vright + oy0 + i) - first);
    }
    for (int i = tid; i < tw; i += kTileThreads) {
        const int first = __ldg(hleft + ox0 + i);
        hwin[i] = make_int2((first - sx0) * C, __ldg(hright + ox0 + i) - first);
    }
    {
        // The weight rows of consecutive outputs are contiguous in HBM: flat copies.  (The staged row
        // strides equal the job's table strides; the planner only merges jobs with equal strides.)
        const float* __restrict__ gv = J.v.w + size_t(oy0) * geom.vstride;
        const float* __restrict__ gh = J.h.w + size_t(ox0) * geom.hstride;
        // The staged source bytes are used as denormal floats (b * 2^-149, no int -> float conversion); the
        // weights carry the compensating exact powers of two (device_types.hpp).
        for (int i = tid; i < th * geom.vstride; i += kTileThreads) vw_s[i] = __ldg(gv + i) * kRingScaleV;
        for (int i = tid; i < tw * geom.hstride; i += kTileThreads) hw_s[i] = __ldg(gh + i) * kRingScaleH;
    }
    // ---- stage the footprint: coalesced byte loads, zero-extended into float words
    {
        const uint8_t* base = J.src + size_t(sy0) * J.src_pitch + size_t(sx0) * C * sizeof(T);
        for (int row = tid / 32; row < nrow; row += kTileThreads / 32) {
            const T* g = reinterpret_cast<const T*>(base + size_t(row) * J.src_pitch);
            float* s = src_f + row * pitch;
            for (int c = tid % 32; c < ncol; c += 32) s[c] = __uint_as_float(uint32_t(__ldg(g + c)));
        }
    }
    __syncthreads();

    // ---- vertical pass: tmp[oy][col] = sum_i w_v[oy][i] * src[left_v[oy] + i][col]; 4 columns per thread
    {
        const int groups = (ncol + 3) >> 2;
        for (int oyl = tid >> 5; oyl < th; oyl += kTileThreads / 32)
        for (int q = tid & 31; q < groups; q += 32) {
            const int2 win = vwin[oyl];
            const float* __restrict__ w = vw_s + oyl * geom.vstride;
            const float4* s = reinterpret_cast<const float4*>(src_f + win.x * pitch) + q;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
            for (int i = 0; i < win.y; ++i) {
                const float wi = w[i];
                const float4 v = *s;
                acc.x = fmaf(v.x, wi, acc.x);
                acc.y = fmaf(v.y, wi, acc.y);
                acc.z = fmaf(v.z, wi, acc.z);
                acc.w = fmaf(v.w, wi, acc.w);
                s += pitch / 4;
            }
            reinterpret_cast<float4*>(tmp_f + oyl * pitch)[q] = acc;
        }
    }
    __syncthreads();

    // ---- horizontal pass: one output pixel (all channels) per thread, quantised into the out tile
    for (int oxl = tid & 63; oxl < tw; oxl += 64)
    for (int oyl = tid >> 6; oyl < th; oyl += kTileThreads / 64) {
        const int2 win = hwin[oxl];
        const float* __restrict__ w = hw_s + oxl * geom.hstride;
        const float* t = tmp_f + oyl * pitch + win.x;
        float a0 = 0.5f, a1 = 0.5f, a2 = 0.5f, a3 = 0.5f;  // + 0.5: round half away from zero at the end
        if (C == 4) {
#pragma unroll 4
            for (int i = 0; i < win.y; ++i) {
                const float wi = w[i];
                const float4 v = *reinterpret_cast<const float4*>(t + 4 * i);
                a0 = fmaf(v.x, wi, a0); a1 = fmaf(v.y, wi, a1); a2 = fmaf(v.z, wi, a2); a3 = fmaf(v.w, wi, a3);
            }
        } else if (C == 3) {
#pragma unroll 4
            for (int i = 0; i < win.y; ++i) {
                const float wi = w[i];
                a0 = fmaf(t[3 * i], wi, a0); a1 = fmaf(t[3 * i + 1], wi, a1); a2 = fmaf(t[3 * i + 2], wi, a2);
            }
        } else if (C == 2) {
            for (int i = 0; i < win.y; ++i) {
                const float wi = w[i];
                const float2 v = *reinterpret_cast<const float2*>(t + 2 * i);
                a0 = fmaf(v.x, wi, a0); a1 = fmaf(v.y, wi, a1);
            }
        } else {
            for (int i = 0; i < win.y; ++i) a0 = fmaf(t[i], w[i], a0);
        }
        T* d = out_s + oyl * out_pitch + oxl * CO;
        const T q0 = T(quantize_tile<T>(a0)), q1 = T(quantize_tile<T>(a1));
        const T q2 = T(quantize_tile<T>(a2)), q3 = T(quantize_tile<T>(a3));
        if (CO == C) {
            d[0] = q0;
            if (C > 1) d[1] = q1;
            if (C > 2) d[2] = q2;
            if (C > 3) d[3] = q3;
        } else {  // Luma(A) replicates its grey into r, g, b; alpha is the source's or 255
            d[0] = q0;
            d[1] = C >= 3 ? q1 : q0;
            d[2] = C >= 3 ? q2 : q0;
            if (CO == 4) d[3] = C == 2 ? q1 : T(255);  // (conversions exist for 8-bit rasters only)
        }
    }
    __syncthreads();

    // ---- store the tile: whole 32-bit words where the destination allows, bytes at the ragged ends
    {
        const int row_bytes = tw * CO * int(sizeof(T));
        const int warp = tid >> 5, lane = tid & 31;
        for (int row = warp; row < th; row += kTileThreads / 32) {
            const uint8_t* sb = reinterpret_cast<const uint8_t*>(out_s + row * out_pitch);
            const uint32_t* sw_ = reinterpret_cast<const uint32_t*>(sb);
            uint8_t* g = J.dst + size_t(oy0 + row) * J.dst_pitch + size_t(ox0) * CO * sizeof(T);
            const int head = min(row_bytes, int((4 - (reinterpret_cast<uintptr_t>(g) & 3)) & 3));
            const int nwords = (row_bytes - head) >> 2;
            if (lane < head) g[lane] = sb[lane];
            uint32_t* gw = reinterpret_cast<uint32_t*>(g + head);
            const int sh8 = head * 8;
            for (int k = lane; k < nwords; k += 32) gw[k] = __funnelshift_r(sw_[k], sw_[k + 1], sh8);
            const int done = head + nwords * 4;
            if (lane < row_bytes - done) g[done + lane] = sb[done + lane];
        }
    }
}

size_t tile_smem_bytes(const TileGeom& g) {
    return (size_t(g.max_src_rows) + size_t(g.max_tile_rows)) * g.pitch_f * sizeof(float) +
           (size_t(g.max_tile_rows) * g.vstride + size_t(g.max_tile_cols) * g.hstride) * sizeof(float) +
           (size_t(g.max_tile_rows) + size_t(g.max_tile_cols)) * sizeof(int2) + size_t(g.max_tile_rows) * g.out_pitch_b;
}

cudaError_t launch_tile(int bytes_per_sample, const DevJob* jobs, const WorkItem* items, const TileGeom& geom,
                        cudaStream_t stream) {
    const size_t smem = tile_smem_bytes(geom);
    const auto kernel = bytes_per_sample == 2 ? tile_kernel<uint16_t> : tile_kernel<uint8_t>;
    constexpr size_t kTileMaxSmem = 110 * 1024;  // the planner's budget (context.cpp: plan_tiles)
    if (smem > kTileMaxSmem) return cudaErrorInvalidValue;
    // always the same value: the attribute is shared by every thread launching on this device
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kTileMaxSmem));
    if (e != cudaSuccess) return e;
    kernel<<<geom.n_items, kTileThreads, smem, stream>>>(jobs, items, geom);
    return cudaGetLastError();
}

}  // namespace ikc
