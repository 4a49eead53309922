// up2.cu -- fused single-launch kernel for exact 2x upscales of 8-bit Rgb8/Rgba8 rasters (sm_100a):
// BASELINE config 4 (1920x1080 -> 3840x2160 CatmullRom), the large-output, write-bound path of
// imageops::resize (image 0.25.8 vertical_sample then horizontal_sample; reached from
// /root/reference/src/transform.rs:85-89).
//
// At exactly 2x the two outputs 2k and 2k+1 of a pass read the same T source samples
// (k + off .. k + off + T - 1; windows clamped at the image border are framed with zero weights,
// PassPlan::up2_pairs).  One packed FMA (fma.rn.f32x2) with the weight pair (w[2k][t], w[2k+1][t]) and
// the duplicated sample therefore advances both outputs, and every staged sample is reused from
// registers for all the outputs it feeds:
//   * a CTA owns 128 x 32 outputs = 64 x 16 source pixels (+ T-1 halo), staged once as bytes
//     (RGBA: 112 x 32 outputs = 56 x 16 source pixels, so that a staged row still has <= 256 byte columns);
//   * vertical pass: thread = one byte column of the staged rows; its 16+T-1 samples are used as they are
//     (a zero-extended byte is the denormal float b * 2^-149; the weight pairs carry the compensating
//     powers of two, device_types.hpp) and produce the 32 rows of the unclamped f32 intermediate
//     (times 2^-74) in shared memory;
//   * horizontal pass: warp = 8 (RGBA: 7) source pixels, lane = intermediate row; the pixels of the row
//     are loaded once with LDS.128 and produce 16 (14) finished pixels, packed with saturating F2IP
//     conversions (round-half-away: the accumulators start at 0.5) and written as 16- / 8-byte stores.
// Same pass order, f32 unclamped intermediate and final clamp + round as the reference; sums use FMA,
// hence max |delta| <= 1 LSB (EXACT mode runs generic.cu instead).
#include <cuda_runtime.h>

#include <cstdint>
#include <type_traits>

#include "device_types.hpp"
#include "launch.hpp"

namespace ikc {
namespace {

constexpr int kUpThreads = 128;
constexpr int kUpKY = 16;                // source rows per tile        ->  32 output rows
constexpr int kUpOutH = 2 * kUpKY;

template <int C, int T>
struct UpGeom {
    static constexpr int kSegPx = C == 4 ? 6 : 8;             // source pixels per warp in the horizontal pass
    static constexpr int kKX = kSegPx * (kUpThreads / 32);    // source pixels per tile row
    static constexpr int kOutW = 2 * kKX;                     // output columns per tile
    static constexpr int kCols = kKX + T - 1;                 // staged source pixels per row
    static constexpr int kRows = kUpKY + T - 1;               // staged source rows
    static constexpr int kColBytes = kCols * C;
    static constexpr int kRowBytes = (kColBytes + 15 + 15) & ~15;  // staged bytes per source row (+ chunk misalignment)
    // floats per intermediate row: a multiple of 4 with an odd number of 16-byte units, so that the 32
    // rows read by a warp's LDS.128 fall into different bank groups; >= 3 floats of slack for the
    // last segment's rounded-up vector loads
    static constexpr int kPitch4 = ((kColBytes + 3 + 3) / 4) | 1;
    static constexpr int kPitchF = kPitch4 * 4;
    static constexpr int kSegFloats = (kSegPx + T - 1) * C;    // intermediate values a lane needs
    static constexpr int kWords = 2 * kSegPx * C / 4;          // packed output words per lane
    static constexpr int kSegVec = (kSegFloats + 3) / 4;
    static constexpr size_t kSrcBytes = size_t(kRows) * kRowBytes;
    static constexpr size_t kTmpBytes = size_t(kUpOutH) * kPitchF * sizeof(float);
    static constexpr size_t kPairBytes = size_t(kUpKY + kKX) * T * sizeof(float2);
    static_assert(kColBytes <= kUpThreads, "one thread per staged byte column in the vertical pass");
    static constexpr size_t kSmem = kTmpBytes + 2 * (kSrcBytes + kPairBytes);  // source and pairs are double buffered
};

__device__ __forceinline__ float2 dup(float v) { return make_float2(v, v); }

// (b0, b1, b2, b3) -> one little-endian word of saturated bytes; inputs already hold value + 0.5.
// float -> s32 (rz) followed by the saturating pack compiles to two F2IP.U8.F32.TRUNC.
__device__ __forceinline__ uint32_t pack4(float b0, float b1, float b2, float b3) {
    uint32_t hi, w;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(__float2int_rz(b3)), "r"(__float2int_rz(b2)), "r"(0));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(w) : "r"(__float2int_rz(b1)), "r"(__float2int_rz(b0)), "r"(hi));
    return w;
}

// cp.async (LDGSTS): 16-byte global -> shared copy that bypasses registers; `valid` < 16 zero-fills the tail.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, int valid) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(uint32_t(__cvta_generic_to_shared(smem_dst))),
                 "l"(gmem_src), "r"(valid)
                 : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(uint32_t(__cvta_generic_to_shared(smem_dst))), "l"(gmem_src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Persistent CTAs: each walks the tile list with a grid stride and prefetches the next tile's source
// footprint and weight pairs (cp.async, double buffered) while it computes the current one.
template <int C, int T>
__global__ void __launch_bounds__(kUpThreads, 4)
up2_kernel(const DevJob* __restrict__ jobs, const WorkItem* __restrict__ items, int n_items) {
    using G = UpGeom<C, T>;
    extern __shared__ __align__(16) uint8_t up_smem[];
    float* tmp_s = reinterpret_cast<float*>(up_smem);                             // [32][kPitchF]
    uint8_t* src_buf = up_smem + G::kTmpBytes;                                    // 2 x [kRows][kRowBytes]
    float2* pair_buf = reinterpret_cast<float2*>(src_buf + 2 * G::kSrcBytes);     // 2 x ([kUpKY][T] + [kKX][T])
    constexpr int kSegPx = G::kSegPx, kOutW = G::kOutW;
    constexpr int kPairs = (kUpKY + G::kKX) * T;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // Geometry of a tile's staged footprint.  Rows are staged as whole 16-byte chunks when the source
    // allows it (16-byte aligned base and pitch, no columns left of the image); `mis` is then the offset
    // of the first wanted byte inside the first chunk.
    // The fields of a job the kernel needs, kept in shared memory (two slots: the job of the current tile
    // and the job of the next one; job indices never decrease along the tile list) so that no tile waits
    // for the DevJob in global memory.
    struct JobLite {
        const uint8_t* src;
        uint8_t* dst;
        size_t src_pitch, dst_pitch;
        const float2* vpairs;
        const float2* hpairs;
        int sw, sh, voff, hoff;
        int vlo, vhi, hlo, hhi;  // source rows / columns whose weight pairs are all the same (the interior)
    };
    __shared__ JobLite job_s[2];
    __shared__ int job_id_s[2];
    auto cache_job = [&](int job) {  // thread 0 only; readers come after the next barrier
        if (job_id_s[job & 1] == job) return;
        const DevJob& D = jobs[job];
        JobLite l;
        l.src = D.src; l.dst = D.dst; l.src_pitch = D.src_pitch; l.dst_pitch = D.dst_pitch;
        l.vpairs = D.v.up2_pairs_v; l.hpairs = D.h.up2_pairs_h;
        l.sw = int(D.sw); l.sh = int(D.sh); l.voff = D.v.up2_off; l.hoff = D.h.up2_off;
        l.vlo = D.v.up2_uni_lo; l.vhi = D.v.up2_uni_hi; l.hlo = D.h.up2_uni_lo; l.hhi = D.h.up2_uni_hi;
        job_s[job & 1] = l;
        job_id_s[job & 1] = job;
    };
    struct Foot {
        const JobLite* J;
        int kx0, ky0, sx_first, sy_first, mis;
        bool vec;
    };
    auto footprint = [&](const WorkItem& it) {
        Foot f;
        f.J = job_s + (it.job & 1);
        f.kx0 = it.ox0 >> 1;  // tile origin in source pixels (tiles start on even outputs)
        f.ky0 = it.oy0 >> 1;
        f.sx_first = f.kx0 + f.J->hoff;  // first staged column / row (may be < 0)
        f.sy_first = f.ky0 + f.J->voff;
        f.vec = f.sx_first >= 0 && ((reinterpret_cast<uintptr_t>(f.J->src) | f.J->src_pitch) & 15) == 0;
        f.mis = f.vec ? (f.sx_first * C) & 15 : 0;
        return f;
    };
    // Asynchronous part of staging tile `idx` into buffer `b`: weight pairs and (vector path) source rows.
    // Bytes that are not written keep whatever the buffer held: any byte converts to a finite float, and
    // everything outside the image meets a zero weight.
    auto prefetch = [&](const WorkItem& it, int b) {
        const Foot f = footprint(it);
        const int sw = int(f.J->sw), sh = int(f.J->sh);
        float2* vp = pair_buf + b * kPairs;
        float2* hp = vp + kUpKY * T;
        for (int i = tid; i < kUpKY * T; i += kUpThreads) {
            const int k = f.ky0 + i / T;
            if (k < sh) cp_async8(vp + i, f.J->vpairs + size_t(k) * T + i % T);
            else vp[i] = make_float2(0.0f, 0.0f);
        }
        for (int i = tid; i < G::kKX * T; i += kUpThreads) {
            const int k = f.kx0 + i / T;
            if (k < sw) cp_async8(hp + i, f.J->hpairs + size_t(k) * T + i % T);
            else hp[i] = make_float2(0.0f, 0.0f);
        }
        if (f.vec) {
            uint8_t* dst = src_buf + b * G::kSrcBytes;
            const int a0 = (f.sx_first * C) & ~15;             // byte offset of the first chunk in a source row
            const int nchunk = (f.mis + G::kColBytes + 15) >> 4;
            const int row_bytes = sw * C;
            for (int i = tid; i < G::kRows * nchunk; i += kUpThreads) {
                const int r = i / nchunk, j = i - r * nchunk;
                const int sy = f.sy_first + r, off = a0 + 16 * j;
                const int valid = min(16, row_bytes - off);
                if (sy >= 0 && sy < sh && valid > 0)
                    cp_async16(dst + r * G::kRowBytes + 16 * j, f.J->src + size_t(sy) * f.J->src_pitch + off, valid);
            }
        }
        cp_async_commit();
    };

    // Tile descriptors are read two tiles ahead, so that neither the prefetch nor the tile itself waits
    // for a global load of its own descriptor.
    const int stride = int(gridDim.x);
    auto item_at = [&](int idx) { return idx < n_items ? items[idx] : WorkItem{0, 0, 0, 0, 0}; };
    int buf = 0;
    WorkItem it = item_at(blockIdx.x), it_next = item_at(int(blockIdx.x) + stride);
    if (tid == 0) {
        job_id_s[0] = job_id_s[1] = -1;
        cache_job(it.job);
    }
    __syncthreads();
    if (int(blockIdx.x) < n_items) prefetch(it, 0);
    for (int idx = blockIdx.x; idx < n_items; idx += stride, buf ^= 1) {
        const WorkItem it_next2 = item_at(idx + 2 * stride);
        if (tid == 0 && idx + stride < n_items) cache_job(it_next.job);  // read by the prefetch after the barrier below
        cp_async_wait<0>();  // this tile's copies (issued one tile ago) have landed
        const Foot f = footprint(it);
        const JobLite& J = *f.J;
        const int sw = J.sw, sh = J.sh;
        uint8_t* src_s = src_buf + buf * G::kSrcBytes;
        const float2* vp_s = pair_buf + buf * kPairs;
        const float2* hp_s = vp_s + kUpKY * T;
        if (!f.vec) {  // unaligned source or a tile on the image's left edge: byte loads, nothing left of / above the image
            for (int r = warp; r < G::kRows; r += kUpThreads / 32) {
                const int sy = f.sy_first + r;
                const bool row_ok = sy >= 0 && sy < sh;
                const uint8_t* g = J.src + ptrdiff_t(sy) * ptrdiff_t(J.src_pitch) + ptrdiff_t(f.sx_first) * C;
                for (int b = lane; b < G::kColBytes; b += 32) {
                    const int sx = f.sx_first + b / C;
                    if (row_ok && sx >= 0 && sx < sw) src_s[r * G::kRowBytes + b] = __ldg(g + b);
                }
            }
        }
        __syncthreads();  // the tile's footprint and pairs are in shared memory; the previous tile is finished,
                          // so its buffers may be refilled while this one is computed
        if (idx + stride < n_items) prefetch(it_next, buf ^ 1);

        // ---- vertical pass: tmp[2k + p][col] = sum_t pair[k][t].p * src[k + t][col]
        if (tid < G::kColBytes) {
            float s[G::kRows];
#pragma unroll
            for (int r = 0; r < G::kRows; ++r)  // the zero-extended byte *is* the float b * 2^-149 (a denormal); the weight
                s[r] = __uint_as_float(uint32_t(src_s[r * G::kRowBytes + f.mis + tid]));  // pairs carry 2^75 (V) and 2^74 (H)
            if (f.ky0 >= J.vlo && f.ky0 + kUpKY <= J.vhi) {  // interior rows: one set of pairs, held in registers
                float2 w[T];
#pragma unroll
                for (int t = 0; t < T; ++t) w[t] = vp_s[t];
#pragma unroll
                for (int k = 0; k < kUpKY; ++k) {
                    float2 acc = make_float2(0.0f, 0.0f);
#pragma unroll
                    for (int t = 0; t < T; ++t) acc = __ffma2_rn(w[t], dup(s[k + t]), acc);
                    tmp_s[(2 * k) * G::kPitchF + tid] = acc.x;
                    tmp_s[(2 * k + 1) * G::kPitchF + tid] = acc.y;
                }
            } else {
#pragma unroll
                for (int k = 0; k < kUpKY; ++k) {
                    float2 acc = make_float2(0.0f, 0.0f);
#pragma unroll
                    for (int t = 0; t < T; ++t) acc = __ffma2_rn(vp_s[k * T + t], dup(s[k + t]), acc);
                    tmp_s[(2 * k) * G::kPitchF + tid] = acc.x;
                    tmp_s[(2 * k + 1) * G::kPitchF + tid] = acc.y;
                }
            }
        }
        __syncthreads();

        // ---- horizontal pass: warp = kSegPx source pixels, lane = intermediate row
        uint32_t words[G::kWords];  // the 2 * kSegPx finished pixels of this lane's row, packed
        {
            float v[G::kSegVec * 4];
            const float4* trow = reinterpret_cast<const float4*>(tmp_s + lane * G::kPitchF + warp * (kSegPx * C));
#pragma unroll
            for (int i = 0; i < G::kSegVec; ++i) {
                const float4 q = trow[i];
                v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
            }
            float o[2 * kSegPx * C];  // o[px * C + c], value + 0.5
            // interior columns share one set of pairs, held in registers; border tiles read them per pixel
            auto hpass = [&](auto uniform) {
                float2 wu[T];
                if (decltype(uniform)::value) {
#pragma unroll
                    for (int t = 0; t < T; ++t) wu[t] = hp_s[t];
                }
#pragma unroll
                for (int k = 0; k < kSegPx; ++k) {
                    float2 acc[C];
#pragma unroll
                    for (int c = 0; c < C; ++c) acc[c] = make_float2(0.5f, 0.5f);
#pragma unroll
                    for (int t = 0; t < T; ++t) {
                        const float2 w = decltype(uniform)::value ? wu[t] : hp_s[(warp * kSegPx + k) * T + t];
#pragma unroll
                        for (int c = 0; c < C; ++c) acc[c] = __ffma2_rn(w, dup(v[(k + t) * C + c]), acc[c]);
                    }
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        o[(2 * k) * C + c] = acc[c].x;
                        o[(2 * k + 1) * C + c] = acc[c].y;
                    }
                }
            };
            if (f.kx0 >= J.hlo && f.kx0 + G::kKX <= J.hhi) hpass(std::true_type{});
            else hpass(std::false_type{});
#pragma unroll
            for (int j = 0; j < G::kWords; ++j) words[j] = pack4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
        }

        // ---- store: 2 * kSegPx pixels = kWords words per lane
        const int tw = it.ox1 - it.ox0, th = it.oy1 - it.oy0;
        const size_t dst_pitch = J.dst_pitch;
        uint8_t* const tile_dst = J.dst + size_t(it.oy0) * dst_pitch + size_t(it.ox0) * C;
        const bool vec_ok = tw == kOutW && ((reinterpret_cast<uintptr_t>(J.dst) | dst_pitch) & 15) == 0;
        if (vec_ok) {  // whole tile columns, 16-byte aligned rows (tiles start at multiples of kOutW pixels)
            if (lane < th) {
                uint8_t* d = tile_dst + size_t(lane) * dst_pitch + warp * (G::kWords * 4);
                if (G::kWords % 4 == 0) {
#pragma unroll
                    for (int j = 0; j < G::kWords / 4; ++j)
                        reinterpret_cast<uint4*>(d)[j] = make_uint4(words[4 * j], words[4 * j + 1], words[4 * j + 2], words[4 * j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < G::kWords / 2; ++j) reinterpret_cast<uint2*>(d)[j] = make_uint2(words[2 * j], words[2 * j + 1]);
                }
            }
        } else {       // ragged right edge or unaligned destination: through shared memory, byte-exact
            __syncthreads();  // every lane has read its intermediate row
            uint32_t* stage = reinterpret_cast<uint32_t*>(tmp_s);  // [32][kOutW * C / 4] words
            constexpr int kStageWords = kOutW * C / 4;
#pragma unroll
            for (int j = 0; j < G::kWords; ++j) stage[lane * kStageWords + warp * G::kWords + j] = words[j];
            __syncthreads();
            const uint8_t* sb = reinterpret_cast<const uint8_t*>(stage);
            const int row_bytes = tw * C;
            for (int row = warp; row < th; row += kUpThreads / 32) {
                uint8_t* g = tile_dst + size_t(row) * dst_pitch;
                for (int b = lane; b < row_bytes; b += 32) g[b] = sb[row * (kStageWords * 4) + b];
            }
        }
        it = it_next;
        it_next = it_next2;
    }
}

template <int C, int T>
cudaError_t launch_up2_one(const DevJob* jobs, const WorkItem* items, int n_items, cudaStream_t stream) {
    using G = UpGeom<C, T>;
    static_assert(G::kTmpBytes >= size_t(kUpOutH) * G::kOutW * C, "the store staging reuses the intermediate tile");
    cudaError_t e = cudaFuncSetAttribute(up2_kernel<C, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(G::kSmem));
    if (e != cudaSuccess) return e;
    // persistent grid: as many CTAs as are resident at once on this device
    int dev = 0, sms = 0, per_sm = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, up2_kernel<C, T>, kUpThreads, G::kSmem)) != cudaSuccess) return e;
    const int grid = n_items < sms * per_sm ? n_items : sms * per_sm;
    if (grid <= 0) return cudaErrorInvalidConfiguration;
    up2_kernel<C, T><<<grid, kUpThreads, G::kSmem, stream>>>(jobs, items, n_items);
    return cudaGetLastError();
}

}  // namespace

int up2_tile_w(int channels) { return channels == 4 ? UpGeom<4, 1>::kOutW : UpGeom<3, 1>::kOutW; }
int up2_tile_h() { return kUpOutH; }

bool up2_supported(int channels, int taps_v, int taps_h) {
    return (channels == 3 || channels == 4) && taps_v == taps_h && (taps_v == 1 || taps_v == 3 || taps_v == 5 || taps_v == 7);
}

cudaError_t launch_up2(int channels, int taps, const DevJob* jobs, const WorkItem* items, int n_items, cudaStream_t stream) {
#define IKC_UP_CASE(C_, T_) \
    if (channels == C_ && taps == T_) return launch_up2_one<C_, T_>(jobs, items, n_items, stream);
    IKC_UP_CASE(3, 1) IKC_UP_CASE(3, 3) IKC_UP_CASE(3, 5) IKC_UP_CASE(3, 7)
    IKC_UP_CASE(4, 1) IKC_UP_CASE(4, 3) IKC_UP_CASE(4, 5) IKC_UP_CASE(4, 7)
#undef IKC_UP_CASE
    return cudaErrorInvalidValue;
}

}  // namespace ikc
