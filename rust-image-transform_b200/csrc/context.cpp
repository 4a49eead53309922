// context.cpp -- see context.hpp.
#include "context.hpp"

#include <chrono>
#include <deque>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "launch.hpp"

#if defined(__x86_64__)
#include <emmintrin.h>
#endif
#include <cuda.h>  // CUtensorMap + cuTensorMapEncodeTiled's signature; the entry point is fetched at run time (no libcuda link)

namespace ikc {

// ---- TMA tensor maps ------------------------------------------------------------------------------

namespace {
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q{};
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}
}  // namespace

// Source raster as [rows][pitch / 4] 32-bit words; box = box_rows x box_bytes.  False if the driver refuses.
bool encode_src_map(uint8_t (&out)[128], const void* base, uint32_t rows, size_t pitch, uint32_t box_bytes, uint32_t box_rows) {
    static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap is 128 bytes");
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    alignas(64) CUtensorMap map;
    const cuuint64_t gdim[2] = {cuuint64_t(pitch / 4), cuuint64_t(rows)};
    const cuuint64_t gstride[1] = {cuuint64_t(pitch)};
    const cuuint32_t box[2] = {box_bytes / 4, box_rows};
    const cuuint32_t estride[2] = {1, 1};
    const CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void*>(base), gdim, gstride, box, estride,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
    std::memcpy(out, &map, 128);
    return true;
}

// Source raster as [rows][pitch] bytes; box = 32 rows x 128 bytes with the 128-byte swizzle (banded8's operand tiles).
bool encode_src_map8(uint8_t (&out)[128], const void* base, uint32_t rows, size_t pitch) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    alignas(64) CUtensorMap map;
    const cuuint64_t gdim[2] = {cuuint64_t(pitch), cuuint64_t(rows)};
    const cuuint64_t gstride[1] = {cuuint64_t(pitch)};
    const cuuint32_t box[2] = {128, uint32_t(kBand8Chunk)};
    const cuuint32_t estride[2] = {1, 1};
    const CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), gdim, gstride, box, estride,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
    std::memcpy(out, &map, 128);
    return true;
}

// ---- errors ---------------------------------------------------------------------------------------

static thread_local std::string g_last_error;
void set_last_error(const std::string& s) { g_last_error = s; }
const char* last_error() { return g_last_error.c_str(); }

void fail(Status s, const std::string& what) { throw Error{s, what}; }
void check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return;
    const Status s = (e == cudaErrorMemoryAllocation) ? kOom : kCudaError;
    fail(s, std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
}

// ---- buffers --------------------------------------------------------------------------------------

void Buffer::release() {
    if (!p) return;
    if (pinned_host) cudaFreeHost(p); else cudaFree(p);
    p = nullptr;
    cap = 0;
}
void Buffer::reserve(size_t bytes) {
    if (bytes <= cap) return;
    size_t want = std::max<size_t>(bytes, size_t(4) << 20);  // (growing means cudaFree + cudaMalloc: a device-wide stall; start big enough for thumbnails)
    want = std::max(want, cap + cap / 2);
    want = (want + 255) & ~size_t(255);
    release();
    if (pinned_host) check_cuda(cudaMallocHost(&p, want), "cudaMallocHost(staging)");
    else check_cuda(cudaMalloc(&p, want), "cudaMalloc(device buffer)");
    cap = want;
}

int Lane::trim(size_t keep) {
    int n = 0;
    for (Buffer* b : {&h_in, &h_out, &d_in, &d_out, &d_scratch})
        if (b->cap > keep) {
            b->release();
            ++n;
        }
    return n;
}

DevTables::~DevTables() {
    if (!base && !ready) return;
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(device);
    if (base) cudaFreeAsync(base, stream);   // stream-ordered: no device-wide synchronisation on eviction
    if (ready) cudaEventDestroy(ready);
    cudaSetDevice(cur);
}

void DevTables::wait_ready(cudaStream_t s) {
    if (settled.load(std::memory_order_acquire) || !ready) return;
    if (cudaEventQuery(ready) == cudaSuccess) {
        settled.store(true, std::memory_order_release);
        return;
    }
    check_cuda(cudaStreamWaitEvent(s, ready, 0), "cudaStreamWaitEvent(weight tables)");
}

// memcpy INTO pinned staging that a DMA is about to read, with non-temporal stores: the lines go to memory instead of
// staying dirty in the copying cores' caches, where the device's reads would have to find them.  Measured on the B200
// box (PCIe Gen5, 16 vCPU): a 6.2 MB source staged by 5 threads with plain memcpy then took 650 us to DMA (9.5 GB/s),
// with streaming stores 120 us; the whole ikc_resize_u8 call went from 790 us to 215 us (160 us from pinned memory).
// IKC_COPY_NT=0 switches back to memcpy.
static const bool kStreamStores = [] { const char* v = std::getenv("IKC_COPY_NT"); return !(v && *v == '0'); }();
static void stage_copy(uint8_t* dst, const uint8_t* src, size_t n) {
#if defined(__x86_64__)
    if (kStreamStores && n >= 256) {
        const size_t head = (16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15;
        std::memcpy(dst, src, head);
        dst += head; src += head; n -= head;
        size_t i = 0;
        for (; i + 64 <= n; i += 64) {
            const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i));
            const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 16));
            const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 32));
            const __m128i e = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 48));
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), a);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 16), b);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 32), c);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 48), e);
        }
        std::memcpy(dst + i, src + i, n - i);
        _mm_sfence();
        return;
    }
#endif
    std::memcpy(dst, src, n);
}

// ---- copy pool ------------------------------------------------------------------------------------

CopyPool::CopyPool(int helpers) {
    for (int i = 0; i < helpers; ++i) threads_.emplace_back([this] { worker(); });
}
CopyPool::~CopyPool() {
    {
        std::lock_guard<std::mutex> lk(mu_);
        stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : threads_) t.join();
}
constexpr size_t kWakeMorePieces = 8;   // a joining helper wakes the next one only if at least this many pieces are unclaimed
static inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#endif
}
void CopyPool::worker() {
    uint64_t seen = 0;
    for (;;) {
        const std::function<void(size_t)>* fn;
        size_t n;
        bool pass_on = false;
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return stop_ || epoch_ != seen; });
            if (stop_) return;
            seen = epoch_;
            if (!open_ || wanted_ <= 0) continue;   // the job is over already, or has all the helpers it asked for
            --wanted_;
            ++inside_;
            fn = fn_;
            n = n_;
            pass_on = wanted_ > 0;
        }
        // Waking a thread costs the waker tens of microseconds (more on a virtual machine), so the caller wakes one helper
        // only and each helper that joins wakes the next while plenty of pieces are left.
        if (pass_on && next_.load(std::memory_order_relaxed) + kWakeMorePieces <= n) cv_.notify_one();
        for (;;) {
            const size_t i = next_.fetch_add(1, std::memory_order_relaxed);
            if (i >= n) break;
            try { (*fn)(i); } catch (...) {}
            done_.fetch_add(1, std::memory_order_release);
        }
        std::lock_guard<std::mutex> lk(mu_);
        --inside_;
    }
}
void CopyPool::parallel_for(size_t n, const std::function<void(size_t)>& fn, const std::function<void()>* tick, int wake) {
    // `tick` (optional) runs on the calling thread only: after each piece it has done itself and while it waits for the
    // helpers' last pieces -- the staged upload uses it to queue the DMA of every chunk whose pieces have all landed.
    const int helpers = wake < 0 ? int(threads_.size()) : std::min(wake, int(threads_.size()));
    if (n <= 1 || helpers <= 0 || !run_mu_.try_lock()) {  // nothing to share, or the helpers are busy with another caller
        for (size_t i = 0; i < n; ++i) {
            fn(i);
            if (tick) (*tick)();
        }
        return;
    }
    std::lock_guard<std::mutex> run_lock(run_mu_, std::adopt_lock);
    {
        std::lock_guard<std::mutex> lk(mu_);
        fn_ = &fn;
        n_ = n;
        next_.store(0, std::memory_order_relaxed);
        done_.store(0, std::memory_order_relaxed);
        wanted_ = helpers;
        open_ = true;
        ++epoch_;
    }
    cv_.notify_one();
    for (;;) {  // the caller works too: a helper that wakes late finds less left, nobody waits for it to wake
        const size_t i = next_.fetch_add(1, std::memory_order_relaxed);
        if (i >= n) break;
        fn(i);
        done_.fetch_add(1, std::memory_order_release);
        if (tick) (*tick)();
    }
    while (done_.load(std::memory_order_acquire) < n) {  // the helpers' last pieces: tens of microseconds
        if (tick) (*tick)();
        cpu_relax();
    }
    for (;;) {  // nobody may still hold `fn` when this returns
        {
            std::lock_guard<std::mutex> lk(mu_);
            open_ = false;
            if (inside_ == 0) break;
        }
        cpu_relax();
    }
    fn_ = nullptr;
    n_ = 0;
}

// ---- validation -----------------------------------------------------------------------------------

static constexpr uint32_t kMaxDim = 65535u;
static constexpr uint64_t kMaxPixels = 1ull << 28;

void validate_job(const JobDesc& d) {
    if (d.channels < 1) fail(kInvalidArg, "channels must be >= 1");
    if (d.channels > 4) fail(kUnsupported, "more than 4 interleaved channels is not supported");
    if (d.bps != 1 && d.bps != 2) fail(kUnsupported, "only 8- and 16-bit samples are supported");
    if (d.filter < 0 || d.filter > 4) fail(kInvalidArg, "unknown filter");
    if (d.oc() != d.channels) {  // to_rgb8() / to_rgba8() fused into the store
        if (d.oc() != 3 && d.oc() != 4) fail(kInvalidArg, "destination channels must equal the source's, or be 3 or 4");
        if (d.bps != 1) fail(kUnsupported, "channel conversion is only available for 8-bit rasters");
    }
    if (d.sw > kMaxDim || d.sh > kMaxDim || d.dw > kMaxDim || d.dh > kMaxDim)
        fail(kTooLarge, "image dimension exceeds IKC_MAX_DIM");
    if (uint64_t(d.sw) * d.sh > kMaxPixels || uint64_t(d.dw) * d.dh > kMaxPixels)
        fail(kTooLarge, "image area exceeds IKC_MAX_PIXELS");
    if (d.dw != 0 && d.dh != 0) {
        if (!d.dst) fail(kInvalidArg, "dst is null");
        if (d.dst_pitch < size_t(d.dw) * d.oc() * d.bps) fail(kInvalidArg, "dst_pitch smaller than a row");
        if (d.bps == 2 && (d.dst_pitch & 1)) fail(kInvalidArg, "dst_pitch must be even for 16-bit samples");
    }
    if (d.sw != 0 && d.sh != 0) {
        if (!d.src) fail(kInvalidArg, "src is null");
        if (d.src_pitch < size_t(d.sw) * d.channels * d.bps) fail(kInvalidArg, "src_pitch smaller than a row");
        if (d.bps == 2 && (d.src_pitch & 1)) fail(kInvalidArg, "src_pitch must be even for 16-bit samples");
    }
}

// ---- device ---------------------------------------------------------------------------------------

static constexpr int kLanesPerDevice = 4;       // what a batch worker pipelines over
static constexpr int kMaxLanesPerDevice = 16;   // tickets (begin / end) may grow the pool up to this
static constexpr size_t kMaxCachedTables = 512;

Device::Device(Context* ctx, int ordinal, int index) : ctx_(ctx), ordinal_(ordinal), index_(index) {
    check_cuda(cudaSetDevice(ordinal), "cudaSetDevice");
    cudaDeviceProp prop{};
    check_cuda(cudaGetDeviceProperties(&prop, ordinal), "cudaGetDeviceProperties");
    if (prop.major < 10)
        fail(kCudaError, std::string("device ") + prop.name + " is not sm_100-class; this library is built for sm_100a only");
    sm_count_ = prop.multiProcessorCount;
    check_cuda(cudaStreamCreateWithFlags(&table_stream_, cudaStreamNonBlocking), "cudaStreamCreate(table stream)");
    {   // freed table memory stays in the pool instead of going back to the driver (which would synchronise)
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, ordinal) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    for (int i = 0; i < kLanesPerDevice; ++i) {
        auto l = std::make_unique<Lane>();
        check_cuda(cudaStreamCreateWithFlags(&l->stream, cudaStreamNonBlocking), "cudaStreamCreate");
        free_.push_back(l.get());
        lanes_.push_back(std::move(l));
    }
}

Device::~Device() {
    cudaSetDevice(ordinal_);
    for (auto& l : lanes_)
        if (l->stream) cudaStreamSynchronize(l->stream);
    tabs_.clear();
    if (table_stream_) cudaStreamSynchronize(table_stream_);
    for (auto& b : stage_) {
        if (b.done) cudaEventDestroy(b.done);
        if (b.p) cudaFreeHost(b.p);
    }
    // (the table stream itself is left to the context's teardown: tables still referenced by a caller's prepared batch
    // free themselves on it)
    for (auto& l : lanes_) {
        if (l->stream) {
            cudaStreamSynchronize(l->stream);
            cudaStreamDestroy(l->stream);
        }
        for (cudaEvent_t e : l->out_events) cudaEventDestroy(e);
        l->h_in.release(); l->h_out.release(); l->h_desc.release();
        l->d_in.release(); l->d_out.release(); l->d_scratch.release(); l->d_desc.release();
    }
}

Lane* Device::acquire_lane(bool may_grow) {
    // First come, first served: with a plain condition variable a thread that has just released a lane wins it back
    // against the waiters again and again, and the waiters' latency grows a tail of tens of milliseconds.
    std::unique_lock<std::mutex> lk(mu_);
    if (may_grow && free_.empty() && lanes_.size() < size_t(kMaxLanesPerDevice)) {
        // every lane is held by a ticket (ikc_resize_begin_u8): one more lane instead of a wait that could be a deadlock
        auto l = std::make_unique<Lane>();
        if (cudaSetDevice(ordinal_) == cudaSuccess && cudaStreamCreateWithFlags(&l->stream, cudaStreamNonBlocking) == cudaSuccess) {
            Lane* p = l.get();
            lanes_.push_back(std::move(l));
            return p;
        }
    }
    const uint64_t ticket = next_ticket_++;
    cv_.wait(lk, [&] { return ticket == serving_ && !free_.empty(); });
    ++serving_;
    Lane* l = free_.back();
    free_.pop_back();
    cv_.notify_all();  // the next ticket may find a free lane too
    return l;
}
void Device::release_lane(Lane* l) {
    // One oversized request (a 268 MP raster is more than a gigabyte) must not keep that much page-locked and device
    // memory on the lane for the life of the process: buffers above the keep size go back now (the lane's stream is
    // idle at every release), ordinary ones stay so that steady traffic never allocates.
    if (l->trim(kLaneKeepBytes)) ctx_->stats.staging_trims.fetch_add(1, std::memory_order_relaxed);
    {
        std::lock_guard<std::mutex> lk(mu_);
        free_.push_back(l);
    }
    cv_.notify_all();
}

std::shared_ptr<DevTables> Device::tables(int filter, uint32_t n_in, uint32_t n_out, bool vertical) {
    const auto key = std::make_tuple(filter, n_in, n_out, vertical);
    {
        std::lock_guard<std::mutex> lk(mu_);
        auto it = tabs_.find(key);
        if (it != tabs_.end()) {
            ctx_->stats.table_hits.fetch_add(1, std::memory_order_relaxed);
            return it->second;
        }
    }
    ctx_->stats.table_misses.fetch_add(1, std::memory_order_relaxed);
    auto host = ctx_->pass(filter, n_in, n_out, vertical);
    if (!host) fail(kInvalidArg, "cannot plan pass");
    // One stream-ordered allocation (left | right | w | ring forms | 2x-upscale pairs | band forms, each 256-byte aligned),
    // filled from one pinned staging block by one asynchronous copy on the device's table stream.  Nothing here waits
    // for the device: consumers order their streams behind `ready` (DevTables::wait_ready).
    auto up = [](size_t b) { return (b + 255) & ~size_t(255); };
    std::vector<float> up2_v(host->up2_pairs.size()), up2_h(host->up2_pairs.size());
    for (size_t i = 0; i < up2_v.size(); ++i) {
        up2_v[i] = host->up2_pairs[i] * kRingScaleV;  // exact: powers of two
        up2_h[i] = host->up2_pairs[i] * kRingScaleH;
    }
    struct Part { const void* src; size_t bytes; size_t off; };
    std::vector<Part> parts;
    size_t total = 0;
    auto add = [&](const void* src, size_t bytes) {
        parts.push_back(Part{src, bytes, total});
        total += up(bytes);
        return parts.size() - 1;
    };
    const size_t i_left = add(host->left.data(), sizeof(int32_t) * n_out);
    const size_t i_right = add(host->right.data(), sizeof(int32_t) * n_out);
    const size_t i_w = add(host->w.data(), sizeof(float) * host->w.size());
    const size_t i_ring_v = add(host->ring_v.data(), sizeof(float) * host->ring_v.size());
    const size_t i_ring_h = add(host->ring_h.data(), sizeof(float) * host->ring_h.size());
    const size_t i_up2_v = add(up2_v.data(), sizeof(float) * up2_v.size());
    const size_t i_up2_h = add(up2_h.data(), sizeof(float) * up2_h.size());
    const size_t i_band = add(host->band_tiles.data(), sizeof(uint16_t) * host->band_tiles.size());
    const size_t i_gbase = add(host->band_gbase.data(), sizeof(int32_t) * host->band_gbase.size());
    const size_t i_band8 = add(host->band8.tiles.data(), host->band8.tiles.size());
    const size_t i_gbase8 = add(host->band8.gbase.data(), sizeof(int32_t) * host->band8.gbase.size());
    const size_t i_band8t = add(host->band8t.tiles.data(), host->band8t.tiles.size());
    const size_t i_klo8t = add(host->band8t.k_lo.data(), sizeof(int32_t) * host->band8t.k_lo.size());
    total = std::max<size_t>(total, 256);

    auto t = std::make_shared<DevTables>();
    t->device = ordinal_;
    t->host = host;
    t->stream = table_stream_;
    check_cuda(cudaSetDevice(ordinal_), "cudaSetDevice");
    check_cuda(cudaEventCreateWithFlags(&t->ready, cudaEventDisableTiming), "cudaEventCreate(weight tables)");
    check_cuda(cudaMallocAsync(&t->base, total, table_stream_), "cudaMallocAsync(weight tables)");
    {
        // a free staging block (its previous copy has completed, or is waited for by this thread only)
        size_t k;
        {
            std::lock_guard<std::mutex> lk(stage_mu_);
            for (k = 0; k < stage_.size(); ++k)
                if (!stage_[k].busy) break;
            if (k == stage_.size()) stage_.push_back(StageBlock{});
            stage_[k].busy = true;
        }
        StageBlock blk;
        {
            std::lock_guard<std::mutex> lk(stage_mu_);
            blk = stage_[k];
        }
        try {
            if (blk.done) check_cuda(cudaEventSynchronize(blk.done), "cudaEventSynchronize(table staging)");
            else check_cuda(cudaEventCreateWithFlags(&blk.done, cudaEventDisableTiming), "cudaEventCreate(table staging)");
            if (blk.cap < total) {
                if (blk.p) cudaFreeHost(blk.p);
                blk.p = nullptr;
                blk.cap = 0;
                const size_t want = std::max<size_t>(total + total / 2, size_t(1) << 20);
                check_cuda(cudaMallocHost(&blk.p, want), "cudaMallocHost(table staging)");
                blk.cap = want;
            }
            uint8_t* hp = static_cast<uint8_t*>(blk.p);
            for (const Part& pt : parts)
                if (pt.bytes) stage_copy(hp + pt.off, static_cast<const uint8_t*>(pt.src), pt.bytes);
            check_cuda(cudaMemcpyAsync(t->base, hp, total, cudaMemcpyHostToDevice, table_stream_), "cudaMemcpyAsync(weight tables)");
            check_cuda(cudaEventRecord(blk.done, table_stream_), "cudaEventRecord(table staging)");
            check_cuda(cudaEventRecord(t->ready, table_stream_), "cudaEventRecord(weight tables)");
        } catch (...) {
            std::lock_guard<std::mutex> lk(stage_mu_);
            blk.busy = false;
            stage_[k] = blk;
            throw;
        }
        std::lock_guard<std::mutex> lk(stage_mu_);
        blk.busy = false;
        stage_[k] = blk;
    }
    const uint8_t* base = static_cast<const uint8_t*>(t->base);
    auto at = [&](size_t i) { return base + parts[i].off; };
    t->pass.left = reinterpret_cast<const int32_t*>(at(i_left));
    t->pass.right = reinterpret_cast<const int32_t*>(at(i_right));
    t->pass.w = reinterpret_cast<const float*>(at(i_w));
    t->pass.ring_v = host->ring_v.empty() ? nullptr : reinterpret_cast<const float*>(at(i_ring_v));
    t->pass.ring_h = host->ring_h.empty() ? nullptr : reinterpret_cast<const float*>(at(i_ring_h));
    t->pass.up2_pairs_v = up2_v.empty() ? nullptr : reinterpret_cast<const float2*>(at(i_up2_v));
    t->pass.up2_pairs_h = up2_h.empty() ? nullptr : reinterpret_cast<const float2*>(at(i_up2_h));
    t->pass.band_tiles = host->band_n ? reinterpret_cast<const uint16_t*>(at(i_band)) : nullptr;
    t->pass.band_gbase = host->band_n ? reinterpret_cast<const int32_t*>(at(i_gbase)) : nullptr;
    t->pass.band_n = host->band_n;
    t->pass.band8_tiles = host->band8.limbs ? reinterpret_cast<const int8_t*>(at(i_band8)) : nullptr;
    t->pass.band8_gbase = host->band8.limbs ? reinterpret_cast<const int32_t*>(at(i_gbase8)) : nullptr;
    t->pass.band8_limbs = host->band8.limbs;
    t->pass.band8_shift = host->band8.shift;
    t->pass.band8t_tiles = host->band8t.chunks ? reinterpret_cast<const int8_t*>(at(i_band8t)) : nullptr;
    t->pass.band8t_klo = host->band8t.chunks ? reinterpret_cast<const int32_t*>(at(i_klo8t)) : nullptr;
    t->pass.band8t_chunks = host->band8t.chunks;
    t->pass.band8t_rows = host->band8t.rows;
    t->pass.up2_off = host->up2_off;
    t->pass.up2_taps = host->up2_taps;
    t->pass.up2_uni_lo = host->up2_uni_lo;
    t->pass.up2_uni_hi = host->up2_uni_hi;
    t->pass.stride = int32_t(host->stride);
    t->pass.ring_k = host->ring_k;
    t->pass.ring_stride = host->ring_stride;
    t->pass.n_in = int32_t(n_in);
    t->pass.n_out = int32_t(n_out);
    t->pass.max_count = int32_t(host->max_count);
    t->pass.uni_step = host->uni_step;
    t->pass.uni_lo = host->uni_lo;
    t->pass.uni_hi = host->uni_hi;

    std::lock_guard<std::mutex> lk(mu_);
    auto it = tabs_.find(key);
    if (it != tabs_.end()) return it->second;  // another thread won the race; ours is freed on return
    if (tabs_.size() >= kMaxCachedTables) {    // drop the oldest entry (holders keep theirs alive)
        tabs_.erase(tab_order_.front());
        tab_order_.erase(tab_order_.begin());
    }
    tabs_[key] = t;
    tab_order_.push_back(key);
    return t;
}

// ---- context --------------------------------------------------------------------------------------

static int copy_helpers() {
    if (const char* v = std::getenv("IKC_COPY_HELPERS")) return std::max(0, std::atoi(v));   // (tuning knob)
    const unsigned hw = std::thread::hardware_concurrency();
    return int(std::min(4u, hw > 4 ? hw / 4 : 0u));
}

Context::Context(const int* ids, int n) : copy_pool(copy_helpers()) {
    int visible = 0;
    cudaError_t e = cudaGetDeviceCount(&visible);
    if (e != cudaSuccess || visible <= 0)
        fail(kCudaError, std::string("no usable CUDA device (there is no CPU fallback): ") +
                             (e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e)));
    std::vector<int> want;
    if (ids && n > 0) want.assign(ids, ids + n);
    else for (int i = 0; i < visible; ++i) want.push_back(i);
    for (size_t i = 0; i < want.size(); ++i) {
        if (want[i] < 0 || want[i] >= visible) fail(kInvalidArg, "device id out of range");
        devs_.push_back(std::make_unique<Device>(this, want[i], int(i)));
    }
}

Context::~Context() {
    submit_.clear();  // joins the dispatcher threads before the devices go
    devs_.clear();
}

std::shared_ptr<const PassPlan> Context::pass(int filter, uint32_t n_in, uint32_t n_out, bool vertical) {
    const auto key = std::make_tuple(filter, n_in, n_out, vertical);
    {
        std::lock_guard<std::mutex> lk(pass_mu_);
        auto it = passes_.find(key);
        if (it != passes_.end()) return it->second;
    }
    auto p = build_pass(filter, n_in, n_out, vertical);
    std::lock_guard<std::mutex> lk(pass_mu_);
    if (passes_.size() >= kMaxCachedTables) {
        passes_.erase(pass_order_.front());
        pass_order_.erase(pass_order_.begin());
    }
    if (passes_.emplace(key, p).second) pass_order_.push_back(key);
    return p;
}

// ---- launch planning ------------------------------------------------------------------------------

namespace {

int up16(int v) { return (v + 15) & ~15; }

// Source bytes a strip [a, b) of output columns needs per row (16-byte aligned at both ends).
int strip_bytes(const PassPlan& h, int a, int b, int ch, int sw) {
    const int xl = h.left[a], xr = h.right[b - 1];
    const int b0 = (xl * ch) & ~15;
    const int b1 = std::min(up16(xr * ch), up16(sw * ch));
    return b1 - b0;
}

// Cut [0, dw) into column strips whose source footprint fits the kernel's staging row.
// `align` > 1: interior strip boundaries fall on multiples of it where a strip is wide enough (the banded kernel
// then stores whole 16-byte groups of pixels).
bool cut_strips(const PassPlan& h, int ch, int sw, int max_src, int max_out, std::vector<std::pair<int, int>>* out,
                int align = 1) {
    const int dw = int(h.n_out);
    auto greedy = [&](int cap, std::vector<std::pair<int, int>>* res) {
        res->clear();
        int a = 0;
        while (a < dw) {
            if (strip_bytes(h, a, a + 1, ch, sw) > max_src) return false;
            int b = a + 1;
            while (b < dw && b - a < cap && strip_bytes(h, a, b + 1, ch, sw) <= max_src) ++b;
            if (align > 1 && b < dw && b / align * align > a) b = b / align * align;
            res->emplace_back(a, b);
            a = b;
        }
        return true;
    };
    if (!greedy(max_out, out)) return false;
    const int n = int(out->size());
    std::vector<std::pair<int, int>> even;
    if (n > 1 && greedy((dw + n - 1) / n, &even) && int(even.size()) == n) *out = even;
    return true;
}


// Cut a job into output tiles for the tile kernel and grow `geom` to cover their source footprints.
// Returns false (and leaves the outputs untouched) when no tile shape fits the shared-memory budget.
bool plan_tiles(const PassPlan& v, const PassPlan& h, int ch, int och, int bps, int job, std::vector<WorkItem>* items,
                TileGeom* geom) {
    const int dw = int(h.n_out), dh = int(v.n_out);
    auto footprint = [](const PassPlan& p, int n_out, int t) {  // widest source span of any tile of t outputs
        int worst = 0;
        for (int a = 0; a < n_out; a += t) {
            const int b = std::min(n_out, a + t);
            worst = std::max(worst, p.right[b - 1] - p.left[a]);
        }
        return worst;
    };
    const int ths[] = {32, 16, 8, 4, 2, 1}, tws[] = {64, 32, 16, 8};
    int best_tw = 0, best_th = 0, best_rows = 0, best_pitch = 0;
    for (size_t limit : {size_t(40) << 10, size_t(110) << 10}) {
        for (int t_h : ths) {
            for (int t_w : tws) {
                const int rows = footprint(v, dh, t_h);
                const int pitch = (footprint(h, dw, t_w) * ch + 3) & ~3;
                TileGeom probe{};
                probe.pitch_f = pitch; probe.max_src_rows = rows; probe.max_tile_rows = t_h; probe.max_tile_cols = t_w;
                probe.vstride = int(v.stride); probe.hstride = int(h.stride);
                probe.out_pitch_b = ((t_w * och * bps + 3) & ~3) + 4;
                const size_t smem = tile_smem_bytes(probe);
                if (smem <= limit) { best_tw = t_w; best_th = t_h; best_rows = rows; best_pitch = pitch; break; }
            }
            if (best_tw) break;
        }
        if (best_tw) break;
    }
    if (!best_tw) return false;
    // the launch uses one geometry for every tile of every job: check the merged one still fits
    TileGeom merged = *geom;
    merged.pitch_f = std::max(merged.pitch_f, best_pitch);
    merged.max_src_rows = std::max(merged.max_src_rows, best_rows);
    merged.max_tile_rows = std::max(merged.max_tile_rows, best_th);
    merged.max_tile_cols = std::max(merged.max_tile_cols, best_tw);
    // one launch shares one geometry: jobs whose weight-table strides differ go to another path
    if (geom->pitch_f != 0 && (geom->vstride != int(v.stride) || geom->hstride != int(h.stride))) return false;
    merged.vstride = int(v.stride);
    merged.hstride = int(h.stride);
    merged.out_pitch_b = std::max(merged.out_pitch_b, ((best_tw * och * bps + 3) & ~3) + 4);
    if (tile_smem_bytes(merged) > (size_t(110) << 10)) return false;
    *geom = merged;
    for (int oy = 0; oy < dh; oy += best_th)
        for (int ox = 0; ox < dw; ox += best_tw)
            items->push_back(WorkItem{job, ox, std::min(dw, ox + best_tw), oy, std::min(dh, oy + best_th)});
    return true;
}

}  // namespace

LaunchPlan Context::plan(Device& dev, const JobDesc* descs, size_t n, int* status, bool exact) {
    LaunchPlan lp;
    struct Cand {
        int job;
        std::vector<std::pair<int, int>> strips;
        int ch, kv, kh;
        int sv, sh;  // uniform steps the ring kernel has a specialised loop for, else 0
        bool convert;
        int band_n;  // > 0: goes to the banded (tensor-core) kernel instead of the ring kernel
        int band8;   // > 0: goes to the banded8 (integer tensor-core) kernel; digits per weight
    };
    std::vector<Cand> cands;
    std::vector<WorkItem> tile_items[2];  // [bytes per sample - 1]: one tile-kernel launch per sample type
    std::vector<int> b8t_jobs;            // jobs of the banded8t launch (cut into items once the batch is known)
    std::vector<int> b8u_jobs;            // jobs of the banded8u launches
    TileGeom tile_geom[2]{};
    lp.jobs.reserve(n);
    for (size_t i = 0; i < n; ++i) {
        status[i] = kOk;
        try {
            const JobDesc& d = descs[i];
            validate_job(d);
            if (d.sw == 0 || d.sh == 0 || d.dw == 0 || d.dh == 0 || (d.sw == d.dw && d.sh == d.dh))
                fail(kInvalidArg, "degenerate resize (empty or same-size) must be handled by the caller");
            auto tv = dev.tables(d.filter, d.sh, d.dh, true);
            auto th = dev.tables(d.filter, d.sw, d.dw, false);
            DevJob j{};
            j.src = static_cast<const uint8_t*>(d.src);
            j.dst = static_cast<uint8_t*>(d.dst);
            j.tmp = nullptr;
            j.src_pitch = d.src_pitch;
            j.dst_pitch = d.dst_pitch;
            j.sw = d.sw; j.sh = d.sh; j.dw = d.dw; j.dh = d.dh;
            j.channels = d.channels;
            j.out_channels = d.oc();
            j.bps = d.bps;
            j.v = tv->pass;
            j.h = th->pass;
            const int idx = int(lp.jobs.size());
            lp.jobs.push_back(j);
            lp.keepalive.push_back(tv);
            lp.keepalive.push_back(th);

            const bool tma_ok = (reinterpret_cast<uintptr_t>(d.src) & 15) == 0 && (d.src_pitch & 15) == 0 &&
                                (d.oc() != 4 || ((reinterpret_cast<uintptr_t>(d.dst) | d.dst_pitch) & 3) == 0);
            bool fused = !exact && d.bps == 1 && d.sh >= d.dh && d.sw >= d.dw &&
                         fused_supported(d.channels, tv->pass.ring_k, th->pass.ring_k) && tma_ok;
            Cand c{idx, {}, d.channels, tv->pass.ring_k, th->pass.ring_k, 0, 0, d.oc() != d.channels, 0, 0};
            // Rgba8 downscales that are exactly 2:1 horizontally (and at most ~2:1 vertically): the row-band kernel, whose
            // horizontal pass runs from registers (banded8t.cu).  Work items: bands of 128 output rows x column segments.
            if (!exact && mode.load() == 0 && d.bps == 1 && d.channels == 4 && d.oc() == 4 && tma_ok &&
                ((reinterpret_cast<uintptr_t>(d.dst) | d.dst_pitch) & 31) == 0 && tv->pass.band8t_tiles && tv->pass.band8_limbs == 2 &&
                th->host->h2_12 && encode_src_map8(lp.jobs[size_t(idx)].src_map8, d.src, d.sh, d.src_pitch)) {
                FusedGroup* g = nullptr;
                for (auto& gg : lp.groups)
                    if (gg.band8t) g = &gg;
                if (!g) {
                    lp.groups.push_back(FusedGroup{4, 0, 0, {}, {}, {}});
                    g = &lp.groups.back();
                    g->band8t = true;
                }
                g->b8tgeom.chunks = std::max(g->b8tgeom.chunks, tv->pass.band8t_chunks);
                b8t_jobs.push_back(idx);
                continue;
            }
            // First choice for the other downscales: the banded8 kernel (vertical pass as an integer product on the tensor
            // cores, source bytes used as they are).
            bool banded8 = !exact && mode.load() == 0 && d.bps == 1 && d.sh >= d.dh && d.sw >= d.dw && tma_ok &&
                           tv->pass.band8_tiles && banded8_supported(d.channels, tv->pass.band8_limbs);
            if (banded8) {
                Band8Geom probe{tv->pass.band8_limbs, 2, 2, 0};
                const size_t fixed = banded8_smem_bytes(d.channels, probe);
                const size_t room = banded8_max_smem() > fixed ? banded8_max_smem() - fixed : 0;
                // (two table buffers per CTA: the next item's tables are loaded while the current item is processed)
                const int max_out = int(std::min<size_t>(room / (16 * (size_t(th->pass.stride) + 1)), 512));
                banded8 = max_out >= 1 &&
                          cut_strips(*th->host, d.channels, int(d.sw), banded8_max_src_bytes(), max_out, &c.strips) &&
                          encode_src_map8(lp.jobs[size_t(idx)].src_map8, d.src, d.sh, d.src_pitch);
                if (banded8) {
                    c.band8 = tv->pass.band8_limbs;
                    cands.push_back(std::move(c));
                    continue;
                }
                c.strips.clear();
            }
            // First choice for downscales: the banded kernel (vertical pass on the tensor cores).  The strip's
            // horizontal tables must fit what its other shared-memory tenants leave.
            bool banded = !exact && mode.load() != 2 && d.bps == 1 && d.sh >= d.dh && d.sw >= d.dw && tma_ok && tv->pass.band_tiles &&
                          banded_supported(d.channels, tv->pass.band_n);
            if (banded) {
                BandGeom probe{tv->pass.band_n, 2, 2, 0};
                const size_t fixed = banded_smem_bytes(d.channels, probe);  // (with the rounding slack of both tables)
                const size_t room = banded_max_smem() > fixed ? banded_max_smem() - fixed : 0;
                // per output: its weights (8 bytes per tap) and its (left, right)
                const int max_out = int(std::min<size_t>(room / (8 * (size_t(th->pass.stride) + 1)), 512));
                banded = max_out >= 1 &&
                         cut_strips(*th->host, d.channels, int(d.sw), banded_max_src_bytes(), max_out, &c.strips);
                if (banded) banded = encode_src_map(lp.jobs[size_t(idx)].src_map, d.src, d.sh, d.src_pitch,
                                                    uint32_t(banded_max_src_bytes()), uint32_t(kBandChunk));
                if (banded) {
                    c.band_n = tv->pass.band_n;
                    cands.push_back(std::move(c));
                    continue;
                }
                c.strips.clear();
            }
            if (fused_has_uniform(d.channels, c.kv, c.kh, tv->pass.uni_step, th->pass.uni_step)) {
                c.sv = tv->pass.uni_step;
                c.sh = th->pass.uni_step;
            }
            if (fused) {
                const int max_out = 256;  // outputs per strip (bounds the kernel's left/right table)
                fused = cut_strips(*th->host, d.channels, int(d.sw), fused_max_src_bytes(d.channels), max_out,
                                   &c.strips);
            }
            const bool up2 = !fused && !exact && d.bps == 1 && d.oc() == d.channels && d.dw == 2 * d.sw && d.dh == 2 * d.sh &&
                             tv->pass.up2_pairs_v && th->pass.up2_pairs_h &&
                             up2_supported(d.channels, tv->pass.up2_taps, th->pass.up2_taps);
            // exact 2x upscales of Rgb8 / Rgba8: the tensor-core kernel first (vertical pass as an integer product, lane =
            // output row; horizontal pass from registers), the CUDA-core tile kernel of up2.cu otherwise
            const bool up2_tc = up2 && mode.load() == 0 && tma_ok && ((reinterpret_cast<uintptr_t>(d.dst) | d.dst_pitch) & 31) == 0 &&
                                tv->pass.band8t_tiles && tv->pass.band8_limbs == 0 && tv->pass.up2_taps == th->pass.up2_taps &&
                                banded8u_supported(d.channels, th->pass.up2_taps, th->pass.up2_off, tv->pass.band8t_chunks) &&
                                encode_src_map8(lp.jobs[size_t(idx)].src_map8, d.src, d.sh, d.src_pitch);
            if (fused) cands.push_back(std::move(c));
            else if (up2_tc) {
                FusedGroup* g = nullptr;
                for (auto& gg : lp.groups)
                    if (gg.band8u_taps == th->pass.up2_taps && gg.channels == d.channels) g = &gg;
                if (!g) {
                    lp.groups.push_back(FusedGroup{d.channels, 0, 0, {}, {}, {}});
                    g = &lp.groups.back();
                    g->band8u_taps = th->pass.up2_taps;
                }
                b8u_jobs.push_back(idx);
            } else if (up2) {
                FusedGroup* g = nullptr;
                for (auto& gg : lp.groups)
                    if (gg.up_taps == tv->pass.up2_taps && gg.channels == d.channels) g = &gg;
                if (!g) {
                    lp.groups.push_back(FusedGroup{d.channels, 0, 0, {}, {}, {}, 0, 0, false, tv->pass.up2_taps});
                    g = &lp.groups.back();
                }
                const int tile_w = up2_tile_w(d.channels), tile_h = up2_tile_h();
                for (int oy = 0; oy < int(d.dh); oy += tile_h)
                    for (int ox = 0; ox < int(d.dw); ox += tile_w)
                        g->items.push_back(WorkItem{idx, ox, std::min(int(d.dw), ox + tile_w), oy, std::min(int(d.dh), oy + tile_h)});
            } else if (!exact && plan_tiles(*tv->host, *th->host, d.channels, d.oc(), d.bps, idx, &tile_items[d.bps - 1],
                                            &tile_geom[d.bps - 1])) {
                // taken by the tile kernel
            } else {
                lp.generic_jobs.push_back(idx);
                lp.scratch_floats = std::max(lp.scratch_floats, size_t(d.sw) * d.channels * d.dh);
            }
        } catch (const Error& e) {
            status[i] = e.status;
            set_last_error(e.what);
        }
    }
    for (int b = 0; b < 2; ++b) {
        if (tile_items[b].empty()) continue;
        FusedGroup g{0, 0, 0, std::move(tile_items[b]), {}, tile_geom[b]};
        g.tgeom.n_items = int(g.items.size());
        g.bps = b + 1;
        lp.groups.push_back(std::move(g));
    }
    if (!b8t_jobs.empty()) {
        // Items: band x column segment (boundaries on multiples of eight outputs, no segment narrower than 128 outputs:
        // each pays one block of pre-roll).  One CTA per SM: the segment count minimises (waves) x (work per item).
        FusedGroup* g = nullptr;
        for (auto& gg : lp.groups)
            if (gg.band8t) g = &gg;
        size_t bands = 0;
        for (int idx : b8t_jobs) {
            const DevJob& j = lp.jobs[size_t(idx)];
            bands += (j.dh + uint32_t(j.v.band8t_rows) - 1) / uint32_t(j.v.band8t_rows);
        }
        int want = 1;
        double best = 1e30;
        for (int sgs = 1; sgs <= 64; ++sgs) {
            const double waves = std::ceil(double(bands) * sgs / double(dev.sm_count()));
            const double cost = waves * (1.0 / sgs + 0.03);
            if (cost < best - 1e-12) { best = cost; want = sgs; }
        }
        for (int idx : b8t_jobs) {
            const DevJob& j = lp.jobs[size_t(idx)];
            const int dw = int(j.dw), dh = int(j.dh), rows = j.v.band8t_rows;
            const int segs = std::max(1, std::min(want, dw / 128));
            for (int oy = 0; oy < dh; oy += rows)
                for (int k = 0; k < segs; ++k) {
                    const int x0 = k == 0 ? 0 : int(int64_t(dw) * k / segs) & ~7;
                    const int x1 = k == segs - 1 ? dw : int(int64_t(dw) * (k + 1) / segs) & ~7;
                    if (x1 > x0) g->items.push_back(WorkItem{idx, x0, x1, oy, std::min(dh, oy + rows)});
                }
        }
        g->b8tgeom.n_items = int(g->items.size());
    }
    if (!b8u_jobs.empty()) {
        // Items: band of 128 output rows x column segment in whole blocks of 64 output pixels; the segment count minimises
        // (waves) x (work per item), each segment paying one block of pre-roll per stream.
        const int rows = banded8u_band_rows();
        size_t bands = 0;
        for (int idx : b8u_jobs) bands += (lp.jobs[size_t(idx)].dh + rows - 1) / rows;
        int want = 1;
        double best = 1e30;
        for (int sgs = 1; sgs <= 64; ++sgs) {
            const double waves = std::ceil(double(bands) * sgs / double(dev.sm_count()));
            const double cost = waves * (1.0 / sgs + 0.04);
            if (cost < best - 1e-12) { best = cost; want = sgs; }
        }
        for (int idx : b8u_jobs) {
            const DevJob& j = lp.jobs[size_t(idx)];
            FusedGroup* g = nullptr;
            for (auto& gg : lp.groups)
                if (gg.band8u_taps == j.h.up2_taps && gg.channels == j.channels) g = &gg;
            const int dw = int(j.dw), dh = int(j.dh);
            const int blocks = (dw + 63) / 64;
            const int segs = std::max(1, std::min(want, blocks / 4));
            for (int oy = 0; oy < dh; oy += rows)
                for (int k = 0; k < segs; ++k) {
                    const int x0 = int(int64_t(blocks) * k / segs) * 64;
                    const int x1 = k == segs - 1 ? dw : int(int64_t(blocks) * (k + 1) / segs) * 64;
                    if (x1 > x0) g->items.push_back(WorkItem{idx, x0, x1, oy, std::min(dh, oy + rows)});
                }
        }
    }
    if (cands.empty()) return lp;

    // Row chunks: enough CTAs to fill every SM twice over (two CTAs are resident per SM), but
    // no chunk so short that the vertical halo (taps - ratio rows per chunk) dominates.
    size_t total_strips = 0;
    for (auto& c : cands) total_strips += c.strips.size();
    for (auto& c : cands) {
        const size_t slots = size_t(dev.sm_count()) * (c.band8 ? 1 : 2);  // the banded8 kernel runs one CTA per SM
        const DevJob& j = lp.jobs[c.job];
        const int group_rows = c.band8 ? banded8_tile_rows() : c.band_n ? banded_group_rows() : fused_group_rows();
        // Pick the chunk count that minimises (tail-wave waste) x (vertical halo recompute), assuming
        // the other jobs of the batch are cut the same way.
        const PassPlan& vp = *lp.keepalive[size_t(c.job) * 2]->host;
        const double ratio_v = double(j.sh) / double(j.dh);
        // the banded kernel reads whole 16-row chunks and pays a fixed prologue (TMEM allocation, tables) per item
        const double halo_rows = std::max(0.0, double(vp.max_count) - ratio_v) + (c.band_n ? 36.0 : c.band8 ? 52.0 : 0.0);
        const int max_chunks = std::max(1, std::min(64, int(j.dh) / (2 * group_rows)));
        int n_chunks = 1;
        double best = 1e30;
        for (int nc = 1; nc <= max_chunks; ++nc) {
            const double items = double(total_strips) * nc;
            const double waves = items / double(slots);
            const double tail = std::ceil(waves) / waves;
            const double halo = 1.0 + halo_rows / (double(j.sh) / nc);
            const double cost = tail * halo;
            if (cost < best - 1e-9) { best = cost; n_chunks = nc; }
        }
        FusedGroup* g = nullptr;
        for (auto& gg : lp.groups)
            if (gg.channels == c.ch && gg.convert == c.convert && gg.band8_limbs == c.band8 &&
                (c.band8 ? true
                 : c.band_n ? gg.band_n == c.band_n
                            : (gg.band_n == 0 && gg.kv == c.kv && gg.kh == c.kh && gg.sv == c.sv && gg.sh == c.sh)))
                g = &gg;
        if (!g) {
            const bool tc = c.band_n || c.band8;
            lp.groups.push_back(FusedGroup{c.ch, tc ? -1 : c.kv, tc ? -1 : c.kh, {}, {}, {}, c.sv, c.sh, c.convert});
            g = &lp.groups.back();
            g->band_n = c.band_n;
            g->bgeom.band_n = c.band_n;
            g->band8_limbs = c.band8;
            g->b8geom.limbs = c.band8;
        }
        const PassPlan& hp = *lp.keepalive[size_t(c.job) * 2 + 1]->host;
        for (int k = 0; k < n_chunks; ++k) {
            // banded kernel: chunk boundaries on multiples of the 16-row group, so no group straddles two items
            int oy0 = int(int64_t(j.dh) * k / n_chunks);
            int oy1 = int(int64_t(j.dh) * (k + 1) / n_chunks);
            if (c.band_n || c.band8) {
                oy0 = k == 0 ? 0 : (oy0 + group_rows / 2) / group_rows * group_rows;
                oy1 = k == n_chunks - 1 ? int(j.dh) : (oy1 + group_rows / 2) / group_rows * group_rows;
            }
            if (oy1 <= oy0) continue;
            for (auto& s : c.strips) {
                g->items.push_back(WorkItem{c.job, s.first, s.second, oy0, oy1});
                if (c.band8) {
                    g->b8geom.max_out = std::max(g->b8geom.max_out, s.second - s.first);
                    g->b8geom.hw_pairs = std::max(g->b8geom.hw_pairs, (s.second - s.first) * int(hp.stride));
                    continue;
                }
                if (c.band_n) {
                    g->bgeom.max_out = std::max(g->bgeom.max_out, s.second - s.first);
                    g->bgeom.hw_pairs = std::max(g->bgeom.hw_pairs, (s.second - s.first) * int(hp.stride));
                    continue;
                }
                // tmp row capacity: every staged source byte of the strip lands in some pixel slot
                const int xl = hp.left[s.first];
                const int b0 = (xl * c.ch) & ~15;
                const int nb = strip_bytes(hp, s.first, s.second, c.ch, int(j.sw));
                const int pxb = b0 / c.ch;
                g->geom.tmp_px = std::max(g->geom.tmp_px, (b0 + nb - 1) / c.ch - pxb + 2);
            }
        }
    }
    for (auto& g : lp.groups) {
        if (g.band8_limbs) {
            g.b8geom.n_items = int(g.items.size());
            if (banded8_smem_bytes(g.channels, g.b8geom) > banded8_max_smem())
                fail(kUnsupported, "internal: banded8 kernel shared-memory budget exceeded");
            continue;
        }
        if (g.band_n) {
            g.bgeom.n_items = int(g.items.size());
            if (banded_smem_bytes(g.channels, g.bgeom) > banded_max_smem())
                fail(kUnsupported, "internal: banded kernel shared-memory budget exceeded");
            continue;
        }
        if (g.kv == 0) continue;   // tile-kernel / 2x-upscale launch: geometry already final
        g.geom.tmp_px |= 1;        // odd pixel pitch: conflict-free float4 column walks
        g.geom.n_items = int(g.items.size());
        if (fused_smem_bytes(g.channels, g.kv, g.kh, g.geom) > 113 * 1024)
            fail(kUnsupported, "internal: fused kernel shared-memory budget exceeded");
    }
    return lp;
}

size_t Context::desc_bytes(const LaunchPlan& lp) const {
    size_t b = (sizeof(DevJob) * lp.jobs.size() + 15) & ~size_t(15);
    for (auto& g : lp.groups) b += (sizeof(WorkItem) * g.items.size() + 15) & ~size_t(15);
    return std::max<size_t>(b, 16);
}

void Context::fill_desc(const LaunchPlan& lp, uint8_t* host, float* scratch) const {
    // (byte copies: DevJob is 64-byte aligned for its tensor map, `host` need not be)
    for (size_t i = 0; i < lp.jobs.size(); ++i) {
        DevJob j = lp.jobs[i];
        if (std::find(lp.generic_jobs.begin(), lp.generic_jobs.end(), int(i)) != lp.generic_jobs.end()) j.tmp = scratch;
        std::memcpy(host + i * sizeof(DevJob), &j, sizeof(DevJob));
    }
    size_t off = (sizeof(DevJob) * lp.jobs.size() + 15) & ~size_t(15);
    for (auto& g : lp.groups) {
        std::memcpy(host + off, g.items.data(), sizeof(WorkItem) * g.items.size());
        off += (sizeof(WorkItem) * g.items.size() + 15) & ~size_t(15);
    }
}

void Context::launch_resident(const LaunchPlan& lp, const uint8_t* d_desc_base, cudaStream_t stream, bool exact) {
    for (auto& t : lp.keepalive) t->wait_ready(stream);   // freshly uploaded weight tables: order behind their copy
    const DevJob* d_jobs = reinterpret_cast<const DevJob*>(d_desc_base);
    size_t off = (sizeof(DevJob) * lp.jobs.size() + 15) & ~size_t(15);
    for (auto& g : lp.groups) {
        const WorkItem* d_items = reinterpret_cast<const WorkItem*>(d_desc_base + off);
        if (g.band8u_taps) check_cuda(launch_banded8u(g.channels, g.band8u_taps, d_jobs, d_items, int(g.items.size()), stream), "launch banded8u_kernel");
        else if (g.band8t) check_cuda(launch_banded8t(d_jobs, d_items, g.b8tgeom, stream), "launch banded8t_kernel");
        else if (g.band8_limbs) check_cuda(launch_banded8(g.channels, g.convert, d_jobs, d_items, g.b8geom, stream), "launch banded8_kernel");
        else if (g.band_n) check_cuda(launch_banded(g.channels, g.convert, d_jobs, d_items, g.bgeom, stream), "launch banded_kernel");
        else if (g.up_taps) check_cuda(launch_up2(g.channels, g.up_taps, d_jobs, d_items, int(g.items.size()), stream), "launch up2_kernel");
        else if (g.kv == 0) check_cuda(launch_tile(g.bps, d_jobs, d_items, g.tgeom, stream), "launch tile_kernel");
        else check_cuda(launch_fused(g.channels, g.kv, g.kh, g.sv, g.sh, g.convert, d_jobs, d_items, g.geom, stream), "launch fused_ring_kernel");
        off += (sizeof(WorkItem) * g.items.size() + 15) & ~size_t(15);
        launches.fetch_add(1, std::memory_order_relaxed);
        const int family = g.band8u_taps ? 7 : g.band8t ? 0 : g.band8_limbs ? 1 : g.band_n ? 2 : g.up_taps ? 4 : g.kv == 0 ? 5 : 3;
        stats.launches_by_family[family].fetch_add(1, std::memory_order_relaxed);
    }
    for (int idx : lp.generic_jobs) {
        check_cuda(launch_generic(lp.jobs[idx], exact, stream), "launch generic kernels");
        launches.fetch_add(2, std::memory_order_relaxed);
        stats.launches_by_family[6].fetch_add(2, std::memory_order_relaxed);
    }
}

void Context::enqueue(Device& dev, LaunchPlan& lp, Buffer& h_desc, Buffer& d_desc, Buffer& d_scratch,
                      cudaStream_t stream, bool exact) {
    (void)dev;
    if (lp.jobs.empty()) return;
    if (lp.scratch_floats) d_scratch.reserve(lp.scratch_floats * sizeof(float));
    for (int idx : lp.generic_jobs) lp.jobs[idx].tmp = static_cast<float*>(d_scratch.p);
    const size_t bytes = desc_bytes(lp);
    h_desc.reserve(bytes);
    d_desc.reserve(bytes);
    fill_desc(lp, static_cast<uint8_t*>(h_desc.p), static_cast<float*>(d_scratch.p));
    check_cuda(cudaMemcpyAsync(d_desc.p, h_desc.p, bytes, cudaMemcpyHostToDevice, stream), "upload descriptors");
    launch_resident(lp, static_cast<const uint8_t*>(d_desc.p), stream, exact);
}

// ---- host-buffer path -----------------------------------------------------------------------------

namespace {

bool is_pinned(const void* p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

size_t device_pitch(size_t row_bytes) { return (row_bytes + 255) & ~size_t(255); }

struct HostJobState {  // one in-flight host job on a lane
    JobDesc d;
    size_t in_pitch = 0, out_pitch = 0;
    bool out_staged = false;
    uint32_t out_chunk_rows = 0;  // staged output: rows per D2H chunk (one event each)
    LaunchPlan lp;                // kept until the job has finished: it holds the references to the weight tables
};

static size_t env_kb(const char* name, size_t dflt_bytes) {
    const char* v = std::getenv(name);
    if (!v || !*v) return dflt_bytes;
    const long kb = std::atol(v);
    return kb > 0 ? size_t(kb) << 10 : dflt_bytes;
}
// (tuning knobs, read once: IKC_STAGE_CHUNK_KB / IKC_COPY_PIECE_KB override the defaults for experiments)
const size_t kStageChunkBytes = env_kb("IKC_STAGE_CHUNK_KB", size_t(1) << 20);   // DMA granule of the staged upload
const size_t kCopyPieceBytes = env_kb("IKC_COPY_PIECE_KB", size_t(256) << 10);   // memcpy granule handed to one copy-pool thread
const size_t kOutChunkBytes = env_kb("IKC_OUT_CHUNK_KB", size_t(4) << 20);       // D2H granule of the staged download (one event each)
const size_t kPoolMinBytes = size_t(3) << 19;   // copies below 1.5 MB are done by the calling thread: waking the helpers costs more


// rows [y0, y0 + rows) of a pitched raster <-> tight rows, split over the copy pool
void pooled_copy_rows(CopyPool& pool, uint8_t* dst, size_t dst_pitch, const uint8_t* src, size_t src_pitch, size_t row_bytes,
                      uint32_t rows, bool to_staging) {
    if (rows == 0 || row_bytes == 0) return;
    // to_staging: `dst` is pinned memory a DMA reads next (streaming stores); otherwise it is the caller's result buffer,
    // which the caller reads next (ordinary stores: it should stay in cache)
    auto copy = [to_staging](uint8_t* d, const uint8_t* s, size_t n) {
        if (to_staging) stage_copy(d, s, n);
        else std::memcpy(d, s, n);
    };
    auto copy_rows = [&](uint32_t a, uint32_t b) {
        if (dst_pitch == row_bytes && src_pitch == row_bytes) copy(dst + size_t(a) * row_bytes, src + size_t(a) * row_bytes, size_t(b - a) * row_bytes);
        else for (uint32_t y = a; y < b; ++y) copy(dst + size_t(y) * dst_pitch, src + size_t(y) * src_pitch, row_bytes);
    };
    if (size_t(rows) * row_bytes < kPoolMinBytes) {
        copy_rows(0, rows);
        return;
    }
    const uint32_t piece_rows = uint32_t(std::max<size_t>(1, kCopyPieceBytes / row_bytes));
    const size_t pieces = (rows + piece_rows - 1) / piece_rows;
    pool.parallel_for(pieces, [&](size_t i) {
        const uint32_t a = uint32_t(i) * piece_rows;
        copy_rows(a, std::min(rows, a + piece_rows));
    });
}

// Stage + H2D + kernels + D2H for one job on one lane; everything asynchronous on the lane stream
// except the pageable staging memcpy.  finish_host_job() completes it.
void start_host_job(Context& ctx, Device& dev, Lane& l, const JobDesc& d, HostJobState* st, bool exact) {
    const size_t in_row = size_t(d.sw) * d.channels * d.bps;
    const size_t out_row = size_t(d.dw) * d.oc() * d.bps;
    st->d = d;
    st->in_pitch = device_pitch(in_row);
    st->out_pitch = device_pitch(out_row);
    l.d_in.reserve(st->in_pitch * d.sh);
    l.d_out.reserve(st->out_pitch * d.dh);
    if (!is_pinned(d.src)) {
        // Pageable source: rows are packed into the lane's pinned staging a few MB at a time (the copy pool shares
        // each chunk's memcpy) and every chunk's DMA is queued as soon as it is staged, so the copy of chunk
        // k + 1 overlaps the DMA of chunk k.
        l.h_in.reserve(in_row * d.sh);
        uint8_t* hp = static_cast<uint8_t*>(l.h_in.p);
        // One queue of memcpy pieces over the whole raster, no barrier between chunks: the calling thread copies pieces
        // itself and, between pieces, queues the DMA of every chunk (a run of pieces) that is complete, in order.  Helper
        // threads join whenever they wake -- a per-chunk barrier made every chunk wait for the slowest wake-up
        // (measured: 790 us for a 6 MB source against 160 us from pinned memory).
        const uint32_t piece_rows = uint32_t(std::max<size_t>(1, kCopyPieceBytes / std::max<size_t>(in_row, 1)));
        const uint32_t chunk_pieces = uint32_t(std::max<size_t>(1, kStageChunkBytes / (size_t(piece_rows) * std::max<size_t>(in_row, 1))));
        const uint32_t chunk_rows = piece_rows * chunk_pieces;
        const size_t n_pieces = (size_t(d.sh) + piece_rows - 1) / piece_rows;
        const size_t n_chunks = (n_pieces + chunk_pieces - 1) / chunk_pieces;
        std::vector<std::atomic<uint32_t>> left(n_chunks);
        for (size_t c = 0; c < n_chunks; ++c)
            left[c].store(uint32_t(std::min<size_t>(chunk_pieces, n_pieces - c * chunk_pieces)), std::memory_order_relaxed);
        size_t issued = 0;
        cudaError_t dma_err = cudaSuccess;
        const std::function<void()> issue_ready = [&] {   // calling thread only; never throws (helpers hold references)
            while (issued < n_chunks && left[issued].load(std::memory_order_acquire) == 0) {
                const uint32_t y0 = uint32_t(issued) * chunk_rows, rows = std::min(chunk_rows, d.sh - y0);
                const cudaError_t e = cudaMemcpy2DAsync(static_cast<uint8_t*>(l.d_in.p) + size_t(y0) * st->in_pitch, st->in_pitch,
                                                        hp + size_t(y0) * in_row, in_row, in_row, rows, cudaMemcpyHostToDevice, l.stream);
                if (e != cudaSuccess && dma_err == cudaSuccess) dma_err = e;
                ++issued;
            }
        };
        const uint8_t* sp = static_cast<const uint8_t*>(d.src);
        const size_t src_pitch = d.src_pitch;
        ctx.copy_pool.parallel_for(n_pieces, [&](size_t i) {
            const uint32_t a = uint32_t(i) * piece_rows, b = std::min(d.sh, a + piece_rows);
            if (src_pitch == in_row) stage_copy(hp + size_t(a) * in_row, sp + size_t(a) * in_row, size_t(b - a) * in_row);
            else for (uint32_t y = a; y < b; ++y) stage_copy(hp + size_t(y) * in_row, sp + size_t(y) * src_pitch, in_row);
            left[i / chunk_pieces].fetch_sub(1, std::memory_order_release);
        }, &issue_ready);
        issue_ready();
        check_cuda(dma_err, "H2D copy (staged chunk)");
    } else {
        check_cuda(cudaMemcpy2DAsync(l.d_in.p, st->in_pitch, d.src, d.src_pitch, in_row, d.sh, cudaMemcpyHostToDevice, l.stream),
                   "H2D copy");
    }
    JobDesc dj = d;
    dj.src = l.d_in.p;
    dj.dst = l.d_out.p;
    dj.src_pitch = st->in_pitch;
    dj.dst_pitch = st->out_pitch;
    int status = kOk;
    st->lp = ctx.plan(dev, &dj, 1, &status, exact);
    if (status != kOk) fail(Status(status), last_error());
    ctx.enqueue(dev, st->lp, l.h_desc, l.d_desc, l.d_scratch, l.stream, exact);
    st->out_staged = !is_pinned(d.dst);
    if (st->out_staged) {
        // Pageable destination: the result comes back in chunks, each followed by an event, so that
        // finish_host_job can copy chunk k out of the pinned staging while chunk k + 1 is still in flight.
        l.h_out.reserve(out_row * d.dh);
        uint8_t* hp = static_cast<uint8_t*>(l.h_out.p);
        st->out_chunk_rows = uint32_t(std::max<size_t>(1, kOutChunkBytes / std::max<size_t>(out_row, 1)));
        size_t c = 0;
        for (uint32_t y0 = 0; y0 < d.dh; y0 += st->out_chunk_rows, ++c) {
            const uint32_t rows = std::min(st->out_chunk_rows, d.dh - y0);
            check_cuda(cudaMemcpy2DAsync(hp + size_t(y0) * out_row, out_row, static_cast<const uint8_t*>(l.d_out.p) + size_t(y0) * st->out_pitch,
                                         st->out_pitch, out_row, rows, cudaMemcpyDeviceToHost, l.stream),
                       "D2H copy (staged chunk)");
            if (c >= l.out_events.size()) {
                cudaEvent_t e;
                check_cuda(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
                l.out_events.push_back(e);
            }
            check_cuda(cudaEventRecord(l.out_events[c], l.stream), "cudaEventRecord");
        }
    } else {
        check_cuda(cudaMemcpy2DAsync(d.dst, d.dst_pitch, l.d_out.p, st->out_pitch, out_row, d.dh, cudaMemcpyDeviceToHost, l.stream),
                   "D2H copy");
    }
}

void finish_host_job(Context& ctx, Lane& l, HostJobState& st) {
    struct DropPlan {  // the table references go once the stream has drained (also when a copy below throws)
        HostJobState& s; cudaStream_t q;
        ~DropPlan() { cudaStreamSynchronize(q); s.lp = LaunchPlan{}; }
    } drop{st, l.stream};
    if (st.out_staged) {
        const JobDesc& d = st.d;
        const size_t out_row = size_t(d.dw) * d.oc() * d.bps;
        const uint8_t* hp = static_cast<const uint8_t*>(l.h_out.p);
        size_t c = 0;
        for (uint32_t y0 = 0; y0 < d.dh; y0 += st.out_chunk_rows, ++c) {
            const uint32_t rows = std::min(st.out_chunk_rows, d.dh - y0);
            check_cuda(cudaEventSynchronize(l.out_events[c]), "resize (chunk sync)");
            pooled_copy_rows(ctx.copy_pool, static_cast<uint8_t*>(d.dst) + size_t(y0) * d.dst_pitch, d.dst_pitch,
                             hp + size_t(y0) * out_row, out_row, out_row, rows, false);
        }
    }
    check_cuda(cudaStreamSynchronize(l.stream), "resize (stream sync)");
}

// Cases imageops::resize answers without resampling.  Returns true if handled.
bool trivial_resize(const JobDesc& d) {
    if (d.dw == 0 || d.dh == 0) return true;
    const size_t out_row = size_t(d.dw) * d.oc() * d.bps;
    if (d.sw == 0 || d.sh == 0) {  // nothing to sample from: zeroed ImageBuffer::new(nw, nh)
        const bool opaque = d.oc() == 4 && (d.channels == 1 || d.channels == 3);  // to_rgba8() of it: alpha = 255
        for (uint32_t y = 0; y < d.dh; ++y) {
            uint8_t* row = static_cast<uint8_t*>(d.dst) + size_t(y) * d.dst_pitch;
            std::memset(row, 0, out_row);
            if (opaque) for (uint32_t x = 0; x < d.dw; ++x) row[size_t(x) * 4 + 3] = 255;
        }
        return true;
    }
    if (d.sw == d.dw && d.sh == d.dh && d.oc() != d.channels) {  // same dimensions: the layout conversion alone
        const int c = d.channels, co = d.oc();
        for (uint32_t y = 0; y < d.dh; ++y) {
            const uint8_t* s = static_cast<const uint8_t*>(d.src) + size_t(y) * d.src_pitch;
            uint8_t* o = static_cast<uint8_t*>(d.dst) + size_t(y) * d.dst_pitch;
            for (uint32_t x = 0; x < d.dw; ++x, s += c, o += co) {
                o[0] = s[0];
                o[1] = c >= 3 ? s[1] : s[0];
                o[2] = c >= 3 ? s[2] : s[0];
                if (co == 4) o[3] = (c == 2 || c == 4) ? s[c - 1] : 255;
            }
        }
        return true;
    }
    if (d.sw == d.dw && d.sh == d.dh) {  // same dimensions: plain copy
        for (uint32_t y = 0; y < d.dh; ++y)
            std::memcpy(static_cast<uint8_t*>(d.dst) + size_t(y) * d.dst_pitch,
                        static_cast<const uint8_t*>(d.src) + size_t(y) * d.src_pitch, out_row);
        return true;
    }
    return false;
}

}  // namespace

namespace {
struct CallTimer {  // one host-buffer call: counted, timed, failures noticed through unwinding
    Stats& s;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    bool ok = false;
    explicit CallTimer(Stats& st, uint64_t images = 1) : s(st) { s.calls.fetch_add(images, std::memory_order_relaxed); }
    ~CallTimer() {
        s.busy_ns.fetch_add(uint64_t(std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count()),
                            std::memory_order_relaxed);
        if (!ok) s.failed.fetch_add(1, std::memory_order_relaxed);
    }
};
void count_bytes(Stats& s, const JobDesc& d) {
    s.src_bytes.fetch_add(uint64_t(d.sw) * d.sh * d.channels * d.bps, std::memory_order_relaxed);
    s.dst_bytes.fetch_add(uint64_t(d.dw) * d.dh * d.oc() * d.bps, std::memory_order_relaxed);
}
}  // namespace

void Context::resize_host(const JobDesc& d, int* device_index_out) {
    CallTimer timer(stats);
    Device& dev = device(next_device());
    if (device_index_out) *device_index_out = dev.index();
    resize_host_on(dev, d);
    timer.ok = true;
}

// One image on one lane of `dev`: staged / in-place H2D, kernels, D2H, all on the lane's stream.
void Context::resize_host_on(Device& dev, const JobDesc& d) {
    validate_job(d);
    if (trivial_resize(d)) {
        stats.trivial.fetch_add(1, std::memory_order_relaxed);
        return;
    }
    count_bytes(stats, d);
    check_cuda(cudaSetDevice(dev.ordinal()), "cudaSetDevice");
    Lane* l = dev.acquire_lane();
    HostJobState st;  // (outlives the handler below: its plan holds the weight tables the enqueued kernels read)
    try {
        start_host_job(*this, dev, *l, d, &st, mode.load() == 1);
        finish_host_job(*this, *l, st);
    } catch (...) {
        cudaStreamSynchronize(l->stream);
        dev.release_lane(l);
        throw;
    }
    dev.release_lane(l);
}

void Context::resize_batch_host(JobDesc* descs, size_t n, int* status, int* device_out) {
    CallTimer timer(stats, n);
    timer.ok = true;  // (per-job failures are counted below)
    const int G = device_count();
    const bool exact = mode.load() == 1;
    std::vector<std::string> errors{size_t(G), std::string()};
    Context& self = *this;
    auto worker = [&](int g) {
        Device& dev = device(g);
        if (cudaSetDevice(dev.ordinal()) != cudaSuccess) {
            for (size_t i = size_t(g); i < n; i += size_t(G)) status[i] = kCudaError;
            return;
        }
        // software pipeline over the device's lanes: while lane k waits for its DMA, the next jobs
        // are staged and enqueued on the other lanes
        std::lock_guard<std::mutex> batch_lock(dev.batch_mu);  // one batch at a time owns all lanes
        const int L = dev.lane_count();
        std::vector<HostJobState> st{static_cast<size_t>(L)};   // declared before the lanes: destroyed after they are drained
        std::vector<long> inflight(static_cast<size_t>(L), -1);
        struct LaneSet {  // every lane of the device, handed back (drained) however the worker ends
            Device& dev;
            std::vector<Lane*> lanes;
            ~LaneSet() {
                for (Lane* l : lanes) {
                    cudaStreamSynchronize(l->stream);
                    dev.release_lane(l);
                }
            }
        } held{dev, {}};
        size_t next = size_t(g);                    // first job this worker has not accounted for yet
        try {
            for (int k = 0; k < L; ++k) held.lanes.push_back(dev.acquire_lane());
            auto& lanes = held.lanes;
            auto drain = [&](int k) {
                if (inflight[size_t(k)] < 0) return;
                const size_t i = size_t(inflight[size_t(k)]);
                inflight[size_t(k)] = -1;
                try {
                    finish_host_job(self, *lanes[size_t(k)], st[size_t(k)]);
                } catch (const Error& e) {
                    status[i] = e.status;
                    errors[size_t(g)] = e.what;
                }
            };
            int k = 0;
            // Small images (thumbnail sources, icons) are not worth a plan, a descriptor upload, a launch and two copies each:
            // runs of them are staged together and share ONE upload, plan, launch per kernel variant and download.
            std::vector<size_t> small;
            size_t small_bytes = 0;
            auto flush_small = [&] {
                if (small.empty()) return;
                drain(k);
                std::vector<JobDesc> gd(small.size());
                std::vector<int> gst(small.size(), kOk);
                std::vector<std::string> gerr(small.size());
                for (size_t j = 0; j < small.size(); ++j) gd[j] = descs[small[j]];
                try {
                    resize_group_host(dev, gd.data(), gd.size(), gst.data(), gerr.data(), lanes[size_t(k)]);
                } catch (const Error& e) {
                    for (size_t j = 0; j < small.size(); ++j) { gst[j] = e.status; gerr[j] = e.what; }
                }
                for (size_t j = 0; j < small.size(); ++j) {
                    status[small[j]] = gst[j];
                    if (gst[j] != kOk) errors[size_t(g)] = gerr[j];
                }
                small.clear();
                small_bytes = 0;
                k = (k + 1) % L;
            };
            constexpr size_t kSmallJob = size_t(1) << 20, kSmallGroupJobs = 32, kSmallGroupBytes = size_t(16) << 20;
            for (size_t i = size_t(g); i < n; i += size_t(G)) {
                next = i + size_t(G);
                status[i] = kOk;
                device_out[i] = g;
                const size_t job_bytes = size_t(descs[i].sw) * descs[i].sh * size_t(descs[i].channels & 0xff) * size_t(std::max(descs[i].bps, 1)) +
                                         size_t(descs[i].dw) * descs[i].dh * 4u * size_t(std::max(descs[i].bps, 1));
                if (job_bytes <= kSmallJob) {   // (validated, counted and answered inside the group call)
                    small.push_back(i);
                    small_bytes += job_bytes;
                    if (small.size() >= kSmallGroupJobs || small_bytes >= kSmallGroupBytes) flush_small();
                    continue;
                }
                flush_small();
                try {
                    validate_job(descs[i]);
                    if (trivial_resize(descs[i])) {
                        stats.trivial.fetch_add(1, std::memory_order_relaxed);
                        continue;
                    }
                    count_bytes(stats, descs[i]);
                    drain(k);
                    start_host_job(self, dev, *lanes[size_t(k)], descs[i], &st[size_t(k)], exact);
                    inflight[size_t(k)] = long(i);
                    k = (k + 1) % L;
                } catch (const Error& e) {
                    status[i] = e.status;
                    errors[size_t(g)] = e.what;
                    cudaStreamSynchronize(lanes[size_t(k)]->stream);
                }
            }
            flush_small();
            for (int q = 0; q < L; ++q) drain(q);
        } catch (...) {
            // anything that is not an ikc::Error (std::bad_alloc from the planner's vectors, ...): nothing may cross the
            // C boundary or escape a worker thread.  Jobs in flight and the jobs not reached yet are reported as failed.
            const bool oom = [] { try { throw; } catch (const std::bad_alloc&) { return true; } catch (...) { return false; } }();
            for (long i : inflight)
                if (i >= 0) status[size_t(i)] = oom ? kOom : kCudaError;
            for (size_t i = next; i < n; i += size_t(G)) status[i] = oom ? kOom : kCudaError;
            errors[size_t(g)] = oom ? "out of host memory while planning a batch" : "unexpected exception in a batch worker";
        }
    };
    if (G == 1) worker(0);
    else {
        std::vector<std::thread> th;
        for (int g = 0; g < G; ++g) th.emplace_back(worker, g);
        for (auto& t : th) t.join();
    }
    for (auto& e : errors) if (!e.empty()) set_last_error(e);
    for (size_t i = 0; i < n; ++i)
        if (status[i] != kOk) stats.failed.fetch_add(1, std::memory_order_relaxed);
}

// ---- split host call: begin / end ----------------------------------------------------------------

namespace {
struct Ticket {
    Context* ctx;
    Device* dev;
    Lane* lane;
    HostJobState st;
};
}  // namespace

void* Context::begin_host(const JobDesc& d) {
    CallTimer timer(stats);
    validate_job(d);
    if (trivial_resize(d)) {
        stats.trivial.fetch_add(1, std::memory_order_relaxed);
        timer.ok = true;
        return nullptr;
    }
    count_bytes(stats, d);
    Device& dev = device(next_device());
    check_cuda(cudaSetDevice(dev.ordinal()), "cudaSetDevice");
    auto t = std::make_unique<Ticket>();
    t->ctx = this;
    t->dev = &dev;
    t->lane = dev.acquire_lane(true);
    try {
        start_host_job(*this, dev, *t->lane, d, &t->st, mode.load() == 1);
    } catch (...) {
        cudaStreamSynchronize(t->lane->stream);
        dev.release_lane(t->lane);
        throw;
    }
    timer.ok = true;
    return t.release();
}

void Context::end_host(void* ticket) {
    if (!ticket) return;
    std::unique_ptr<Ticket> t(static_cast<Ticket*>(ticket));
    const auto t0 = std::chrono::steady_clock::now();
    struct Release {
        Ticket& t;
        ~Release() { cudaStreamSynchronize(t.lane->stream); t.st.lp = LaunchPlan{}; t.dev->release_lane(t.lane); }
    } release{*t};
    cudaSetDevice(t->dev->ordinal());
    try {
        finish_host_job(*this, *t->lane, t->st);
    } catch (...) {
        stats.failed.fetch_add(1, std::memory_order_relaxed);
        throw;
    }
    stats.busy_ns.fetch_add(uint64_t(std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count()),
                            std::memory_order_relaxed);
}

// ---- coalescing submit queue ---------------------------------------------------------------------

void Context::resize_group_host(Device& dev, const JobDesc* descs, size_t n, int* status, std::string* errors, Lane* lane) {
    check_cuda(cudaSetDevice(dev.ordinal()), "cudaSetDevice");
    struct Slot { size_t in_off = 0, out_off = 0, in_pitch = 0, out_pitch = 0; bool live = false; };
    std::vector<Slot> slot(n);
    size_t total_in = 0, total_out = 0;
    for (size_t i = 0; i < n; ++i) {
        status[i] = kOk;
        try {
            const JobDesc& d = descs[i];
            validate_job(d);
            if (trivial_resize(d)) {
                stats.trivial.fetch_add(1, std::memory_order_relaxed);
                continue;
            }
            count_bytes(stats, d);
            Slot& s = slot[i];
            s.in_pitch = device_pitch(size_t(d.sw) * d.channels * d.bps);
            s.out_pitch = device_pitch(size_t(d.dw) * d.oc() * d.bps);
            s.in_off = total_in;
            s.out_off = total_out;
            total_in += s.in_pitch * d.sh;
            total_out += s.out_pitch * d.dh;
            s.live = true;
        } catch (const Error& e) {
            status[i] = e.status;
            errors[i] = e.what;
        }
    }
    if (total_in == 0) return;
    LaunchPlan lp;  // (declared before the lane guard: its table references go only after the guard has drained the stream)
    Lane* l = lane ? lane : dev.acquire_lane();
    struct Release {
        Device& d; Lane* l; bool own;
        ~Release() { cudaStreamSynchronize(l->stream); if (own) d.release_lane(l); }
    } release{dev, l, lane == nullptr};
    l->h_in.reserve(total_in);
    l->d_in.reserve(total_in);
    l->h_out.reserve(total_out);
    l->d_out.reserve(total_out);
    uint8_t* hin = static_cast<uint8_t*>(l->h_in.p);
    // inputs: every image's rows into one pinned block laid out like the device buffer, then ONE copy
    for (size_t i = 0; i < n; ++i) {
        if (!slot[i].live) continue;
        const JobDesc& d = descs[i];
        pooled_copy_rows(copy_pool, hin + slot[i].in_off, slot[i].in_pitch, static_cast<const uint8_t*>(d.src), d.src_pitch,
                         size_t(d.sw) * d.channels * d.bps, d.sh, true);
    }
    check_cuda(cudaMemcpyAsync(l->d_in.p, hin, total_in, cudaMemcpyHostToDevice, l->stream), "H2D copy (group)");
    std::vector<JobDesc> dj;
    std::vector<size_t> who;
    for (size_t i = 0; i < n; ++i) {
        if (!slot[i].live) continue;
        JobDesc j = descs[i];
        j.src = static_cast<uint8_t*>(l->d_in.p) + slot[i].in_off;
        j.dst = static_cast<uint8_t*>(l->d_out.p) + slot[i].out_off;
        j.src_pitch = slot[i].in_pitch;
        j.dst_pitch = slot[i].out_pitch;
        dj.push_back(j);
        who.push_back(i);
    }
    const bool exact = mode.load() == 1;
    std::vector<int> st(dj.size(), kOk);
    lp = plan(dev, dj.data(), dj.size(), st.data(), exact);
    for (size_t k = 0; k < dj.size(); ++k)
        if (st[k] != kOk) {
            status[who[k]] = st[k];
            errors[who[k]] = last_error();
            slot[who[k]].live = false;
        }
    enqueue(dev, lp, l->h_desc, l->d_desc, l->d_scratch, l->stream, exact);
    check_cuda(cudaMemcpyAsync(l->h_out.p, l->d_out.p, total_out, cudaMemcpyDeviceToHost, l->stream), "D2H copy (group)");
    check_cuda(cudaStreamSynchronize(l->stream), "resize (group sync)");
    const uint8_t* hout = static_cast<const uint8_t*>(l->h_out.p);
    for (size_t i = 0; i < n; ++i) {
        if (!slot[i].live) continue;
        const JobDesc& d = descs[i];
        pooled_copy_rows(copy_pool, static_cast<uint8_t*>(d.dst), d.dst_pitch, hout + slot[i].out_off, slot[i].out_pitch,
                         size_t(d.dw) * d.oc() * d.bps, d.dh, false);
    }
}

// Per device: callers enqueue a job and sleep; two dispatcher threads (so that one group's staging overlaps the other's
// GPU work) each take whatever has queued up -- one job when the service is idle, dozens under load -- and run it as one
// group.
class SubmitQueue {
public:
    SubmitQueue(Context* ctx, Device* dev) : ctx_(ctx), dev_(dev) {
        for (int i = 0; i < 2; ++i) threads_.emplace_back([this] { run(); });
    }
    ~SubmitQueue() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : threads_) t.join();
    }
    void submit(const JobDesc& d) {
        Pending p{d, kOk, {}, false};
        std::unique_lock<std::mutex> lk(mu_);
        static const bool inline_lone = [] { const char* v = std::getenv("IKC_SUBMIT_INLINE"); return !(v && *v == '0'); }();
        if (inline_lone && q_.empty() && busy_ == 0) {
            // Nothing queued, nothing in flight: there is nobody to share a launch with, so the caller runs its image
            // itself on the single-image path (chunked staging overlapped with the DMA, no hand-off to a dispatcher
            // thread and back: two thread wake-ups, ~100 us on a virtual machine).  Requests that arrive meanwhile find
            // busy_ != 0, queue up and are coalesced by the dispatchers as before.
            ++busy_;
            lk.unlock();
            struct Done {
                SubmitQueue& q;
                ~Done() { std::lock_guard<std::mutex> g(q.mu_); --q.busy_; }
            } done{*this};
            ctx_->resize_host_on(*dev_, d);
            ctx_->stats.submit_batches.fetch_add(1, std::memory_order_relaxed);
            ctx_->stats.submit_jobs.fetch_add(1, std::memory_order_relaxed);
            return;
        }
        q_.push_back(&p);
        cv_.notify_one();
        done_cv_.wait(lk, [&] { return p.done; });
        lk.unlock();
        if (p.status != kOk) fail(Status(p.status), p.err);
    }

private:
    struct Pending { JobDesc d; int status; std::string err; bool done; };
    static constexpr size_t kMaxJobs = 64, kMaxBytes = size_t(256) << 20;
    void run() {
        for (;;) {
            std::vector<Pending*> group;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || !q_.empty(); });
                if (q_.empty()) return;  // stopping
                size_t bytes = 0;
                while (!q_.empty() && group.size() < kMaxJobs && bytes < kMaxBytes) {
                    Pending* p = q_.front();
                    q_.pop_front();
                    bytes += size_t(p->d.sw) * p->d.sh * size_t(p->d.channels & 0xff) * p->d.bps;
                    group.push_back(p);
                }
                ++busy_;
            }
            const size_t n = group.size();
            std::vector<JobDesc> descs(n);
            std::vector<int> st(n, kOk);
            std::vector<std::string> errs(n);
            for (size_t i = 0; i < n; ++i) descs[i] = group[i]->d;
            try {
                ctx_->resize_group_host(*dev_, descs.data(), n, st.data(), errs.data());
            } catch (const Error& e) {
                for (size_t i = 0; i < n; ++i) { st[i] = e.status; errs[i] = e.what; }
            } catch (const std::bad_alloc&) {
                for (size_t i = 0; i < n; ++i) { st[i] = kOom; errs[i] = "out of host memory in the submit queue"; }
            } catch (...) {
                for (size_t i = 0; i < n; ++i) { st[i] = kCudaError; errs[i] = "unexpected exception in the submit queue"; }
            }
            ctx_->stats.submit_batches.fetch_add(1, std::memory_order_relaxed);
            ctx_->stats.submit_jobs.fetch_add(n, std::memory_order_relaxed);
            {
                std::lock_guard<std::mutex> lk(mu_);
                --busy_;
                for (size_t i = 0; i < n; ++i) {
                    group[i]->status = st[i];
                    group[i]->err = std::move(errs[i]);
                    group[i]->done = true;
                }
            }
            done_cv_.notify_all();
        }
    }
    Context* ctx_;
    Device* dev_;
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    std::deque<Pending*> q_;
    std::vector<std::thread> threads_;
    int busy_ = 0;   // groups in flight (dispatchers) + callers running their own image (mu_)
    bool stop_ = false;
};

void Context::submit_host(const JobDesc& d) {
    CallTimer timer(stats);
    SubmitQueue* q = nullptr;
    {
        std::lock_guard<std::mutex> lk(submit_mu_);
        if (submit_.empty()) submit_.resize(devs_.size());
        const size_t g = size_t(next_device());
        if (!submit_[g]) submit_[g] = std::make_unique<SubmitQueue>(this, devs_[g].get());
        q = submit_[g].get();
    }
    q->submit(d);
    timer.ok = true;
}

}  // namespace ikc
