// generic.cu -- generic two-launch resize path (sm_100a):
//     vertical_generic -> f32 intermediate in HBM -> horizontal_generic
// Handles every filter, ratio, channel count (1..4) and sample width (u8/u16).  In EXACT mode it
// multiplies and adds separately in ascending tap order, i.e. it is the reference's algorithm
// (image 0.25.8 imageops/sample.rs vertical_sample / horizontal_sample, reached from
// /root/reference/src/transform.rs:85-89) operation for operation: delta == 0 against the CPU
// restatement.  It is the verification path and the fallback for shapes the fused ring kernel
// (fused.cu) does not take; its intermediate costs HBM traffic, so it is not the headline path.
#include <cuda_runtime.h>

#include <cstdint>

#include "device_types.hpp"
#include "launch.hpp"

namespace ikc {

// Round half away from zero for v >= 0 (f32::round), exact for every float.
__device__ __forceinline__ float round_half_away_nonneg(float v) {
    const float t = truncf(v);
    return (v - t >= 0.5f) ? t + 1.0f : t;
}

template <typename T, bool EXACT>
__global__ void __launch_bounds__(256) vertical_generic(const DevJob job) {
    const uint32_t ncol = job.sw * uint32_t(job.channels);
    const uint32_t col = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t oy = blockIdx.y;
    if (col >= ncol) return;
    const int first = job.v.left[oy];
    const int taps = job.v.right[oy] - first;
    const float* __restrict__ w = job.v.w + size_t(oy) * job.v.stride;
    const uint8_t* p = job.src + size_t(first) * job.src_pitch + size_t(col) * sizeof(T);
    float acc = 0.0f;
    for (int i = 0; i < taps; ++i) {
        const float s = float(*reinterpret_cast<const T*>(p));
        acc = EXACT ? __fadd_rn(acc, __fmul_rn(s, w[i])) : fmaf(s, w[i], acc);
        p += job.src_pitch;
    }
    job.tmp[size_t(oy) * ncol + col] = acc;
}

// One thread per destination sample.  With a channel conversion (out_channels != channels, 8-bit only:
// DynamicImage::to_rgb8() / to_rgba8()) destination channel oc reads source channel c: grey replicated
// into r, g, b; alpha from the source when it has one, else the constant 255.
template <typename T, bool EXACT>
__global__ void __launch_bounds__(256) horizontal_generic(const DevJob job) {
    const uint32_t ch = uint32_t(job.channels), och = uint32_t(job.out_channels);
    const uint32_t nout = job.dw * och;
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t oy = blockIdx.y;
    if (idx >= nout) return;
    const uint32_t ox = idx / och;
    const uint32_t oc = idx - ox * och;
    uint32_t c = oc;
    if (och != ch) {
        if (oc == 3) {
            if (ch == 1 || ch == 3) {
                *reinterpret_cast<T*>(job.dst + size_t(oy) * job.dst_pitch + size_t(idx) * sizeof(T)) = T(255);
                return;
            }
            c = ch - 1;
        } else if (ch < 3) {
            c = 0;
        }
    }
    const int first = job.h.left[ox];
    const int taps = job.h.right[ox] - first;
    const float* __restrict__ w = job.h.w + size_t(ox) * job.h.stride;
    const float* __restrict__ t = job.tmp + size_t(oy) * (size_t(job.sw) * ch) + size_t(first) * ch + c;
    float acc = 0.0f;
    for (int i = 0; i < taps; ++i) {
        const float s = t[size_t(i) * ch];
        acc = EXACT ? __fadd_rn(acc, __fmul_rn(s, w[i])) : fmaf(s, w[i], acc);
    }
    const float hi = sizeof(T) == 1 ? 255.0f : 65535.0f;
    acc = acc < 0.0f ? 0.0f : (acc > hi ? hi : acc);
    const float r = round_half_away_nonneg(acc);
    *reinterpret_cast<T*>(job.dst + size_t(oy) * job.dst_pitch + size_t(idx) * sizeof(T)) = T(r);
}

cudaError_t launch_generic(const DevJob& job, bool exact, cudaStream_t stream) {
    const uint32_t ncol = job.sw * uint32_t(job.channels);
    const uint32_t nout = job.dw * uint32_t(job.out_channels);
    const dim3 block(256);
    const dim3 gv((ncol + 255) / 256, job.dh);
    const dim3 gh((nout + 255) / 256, job.dh);
    if (job.bps == 1) {
        if (exact) {
            vertical_generic<uint8_t, true><<<gv, block, 0, stream>>>(job);
            horizontal_generic<uint8_t, true><<<gh, block, 0, stream>>>(job);
        } else {
            vertical_generic<uint8_t, false><<<gv, block, 0, stream>>>(job);
            horizontal_generic<uint8_t, false><<<gh, block, 0, stream>>>(job);
        }
    } else {
        if (exact) {
            vertical_generic<uint16_t, true><<<gv, block, 0, stream>>>(job);
            horizontal_generic<uint16_t, true><<<gh, block, 0, stream>>>(job);
        } else {
            vertical_generic<uint16_t, false><<<gv, block, 0, stream>>>(job);
            horizontal_generic<uint16_t, false><<<gh, block, 0, stream>>>(job);
        }
    }
    return cudaGetLastError();
}

}  // namespace ikc
