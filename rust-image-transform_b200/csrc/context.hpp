// context.hpp -- process-wide context behind the C ABI: devices, lanes (stream + pinned staging +
// device buffers), weight-table caches, launch planning.  Product code.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include "device_types.hpp"
#include "plan.hpp"

namespace ikc {

// Status codes mirror enum ikc_status in include/imagekit_cuda.h.
enum Status : int { kOk = 0, kInvalidArg = 1, kUnsupported = 2, kTooLarge = 3, kCudaError = 4, kOom = 5 };

struct Error {
    Status status;
    std::string what;
};
[[noreturn]] void fail(Status s, const std::string& what);
void check_cuda(cudaError_t e, const char* what);

void set_last_error(const std::string& s);
const char* last_error();

// A host- or device-resident resize request (pointers are interpreted by the caller of plan).
struct JobDesc {
    const void* src;
    void* dst;
    uint32_t sw, sh, dw, dh;
    size_t src_pitch, dst_pitch;
    int channels;
    int bps;  // bytes per sample
    int filter;
    int out_channels = 0;  // destination samples per pixel; 0 = same as `channels`
    int oc() const { return out_channels ? out_channels : channels; }
};

// Device copy of one PassPlan; frees its memory when the last holder lets go.
struct DevTables {
    int device = -1;
    void* base = nullptr;  // single stream-ordered allocation holding left | right | w | ring | band forms
    cudaStream_t stream = nullptr;   // the device's table stream: the upload and, at destruction, the free are ordered on it
    cudaEvent_t ready = nullptr;     // recorded after the upload; consumers' streams wait on it until it has been seen complete
    std::atomic<bool> settled{false};
    DevPass pass{};
    std::shared_ptr<const PassPlan> host;
    // Orders `s` after the upload (no host-side wait).  Holders must keep their reference until the work they enqueued
    // has completed: the destructor returns the memory with cudaFreeAsync, which does not wait for other streams.
    void wait_ready(cudaStream_t s);
    ~DevTables();
};

struct FusedGroup {  // one kernel launch over a list of work items
    int channels, kv, kh;  // ring kernel variant; kv == 0 marks a tile-kernel launch
    std::vector<WorkItem> items;
    FusedGeom geom{};
    TileGeom tgeom{};
    int sv = 0, sh = 0;    // ring kernel: uniform vertical / horizontal step the launch is specialised for (0 = none)
    bool convert = false;  // ring kernel: the jobs store another channel count than they read
    int up_taps = 0;       // > 0 (with kv == 0): an exact-2x upscale launch (up2.cu) with this tap frame
    int bps = 1;           // tile-kernel launch: bytes per sample of its jobs (1 or 2)
    int band_n = 0;        // > 0: a banded (tensor-core) launch whose weight tiles span band_n output rows
    BandGeom bgeom{};
    int band8_limbs = 0;   // > 0: a banded8 (integer tensor-core) launch with this many digits per weight
    Band8Geom b8geom{};
    bool band8t = false;   // a banded8t (row-band integer tensor-core) launch
    int band8u_taps = 0;   // > 0: a banded8u (tensor-core 2x upscale) launch with this horizontal tap frame
    Band8TGeom b8tgeom{};
};

// Everything needed to enqueue a set of device-resident jobs.
struct LaunchPlan {
    std::vector<DevJob> jobs;               // tmp pointers are patched at enqueue time
    std::vector<FusedGroup> groups;         // fused launches (jobs referenced by index)
    std::vector<int> generic_jobs;          // indices that take the two-launch path
    std::vector<std::shared_ptr<DevTables>> keepalive;
    size_t scratch_floats = 0;              // max f32 intermediate any generic job needs
    int launches() const { return int(groups.size()) + 2 * int(generic_jobs.size()); }
};

struct Buffer {  // device or pinned-host buffer that grows on demand (Lane::trim gives oversized ones back)
    void* p = nullptr;
    size_t cap = 0;
    bool pinned_host = false;
    void reserve(size_t bytes);
    void release();
};

// A few helper threads that share the memcpy between pageable caller memory and the pinned staging buffers
// (one thread moves ~15 GB/s, PCIe Gen5 ~55 GB/s).  parallel_for never blocks behind another caller: if the
// helpers are busy the calling thread does the whole job itself.  Pieces are claimed with an atomic counter: the mutex
// and the condition variable are touched only to wake the helpers and when one of them goes back to sleep.
class CopyPool {
public:
    explicit CopyPool(int helpers);
    ~CopyPool();
    // `wake`: how many helpers to wake for this job at most (each wake-up costs the caller a futex call; < 0: all).
    void parallel_for(size_t n, const std::function<void(size_t)>& fn, const std::function<void()>* tick = nullptr, int wake = -1);
    int helpers() const { return int(threads_.size()); }

private:
    void worker();
    std::vector<std::thread> threads_;
    std::mutex mu_, run_mu_;
    std::condition_variable cv_;
    // the job in flight (published under mu_, pieces claimed lock-free)
    const std::function<void(size_t)>* fn_ = nullptr;
    size_t n_ = 0;
    std::atomic<size_t> next_{0}, done_{0};
    int inside_ = 0;          // helpers that have joined the job and not left it yet (mu_)
    int wanted_ = 0;          // helpers still allowed to join it (mu_)
    bool open_ = false;       // (mu_)
    uint64_t epoch_ = 0;      // (mu_)
    bool stop_ = false;       // (mu_)
};

struct Lane {
    cudaStream_t stream = nullptr;
    std::vector<cudaEvent_t> out_events;  // one per staged D2H chunk of the job in flight
    Buffer h_in{nullptr, 0, true}, h_out{nullptr, 0, true}, h_desc{nullptr, 0, true};
    Buffer d_in, d_out, d_scratch, d_desc;
    // Frees every buffer larger than `keep` bytes (the lane is idle); returns how many were freed.
    int trim(size_t keep);
};

class Context;

class Device {
public:
    Device(Context* ctx, int ordinal, int index);
    ~Device();
    int ordinal() const { return ordinal_; }
    int index() const { return index_; }
    int sm_count() const { return sm_count_; }

    // vertical: the pass is used vertically (its tensor-core operand tiles are built and uploaded too)
    std::shared_ptr<DevTables> tables(int filter, uint32_t n_in, uint32_t n_out, bool vertical);
    Lane* acquire_lane(bool may_grow = false);
    void release_lane(Lane* l);
    int lane_count() const { return kBatchLanes; }   // lanes a batch worker takes (tickets may have grown the pool beyond it)
    static constexpr int kBatchLanes = 4;
    static constexpr size_t kLaneKeepBytes = size_t(256) << 20;   // lane buffers above this are freed when the lane is released
    std::mutex batch_mu;  // serialises batch workers, which take every lane of the device

private:
    Context* ctx_;
    int ordinal_, index_, sm_count_ = 148;
    std::mutex mu_;
    std::condition_variable cv_;
    std::vector<std::unique_ptr<Lane>> lanes_;
    std::vector<Lane*> free_;
    uint64_t next_ticket_ = 0, serving_ = 0;   // lanes are handed out in arrival order
    std::map<std::tuple<int, uint32_t, uint32_t, bool>, std::shared_ptr<DevTables>> tabs_;
    std::vector<std::tuple<int, uint32_t, uint32_t, bool>> tab_order_;
    // Table uploads: one stream, a few reusable pinned staging blocks (each guarded by the event of its last copy).
    struct StageBlock { void* p = nullptr; size_t cap = 0; cudaEvent_t done = nullptr; bool busy = false; };
    cudaStream_t table_stream_ = nullptr;
    std::mutex stage_mu_;
    std::vector<StageBlock> stage_;
};

// Monotonic counters behind ikc_get_stats.
struct Stats {
    std::atomic<uint64_t> calls{0}, failed{0}, trivial{0};
    std::atomic<uint64_t> launches_by_family[8]{};   // banded8t, banded8, banded f16, ring, up2, tile, generic, banded8u
    std::atomic<uint64_t> src_bytes{0}, dst_bytes{0}, busy_ns{0};
    std::atomic<uint64_t> table_hits{0}, table_misses{0};
    std::atomic<uint64_t> submit_batches{0}, submit_jobs{0};
    std::atomic<uint64_t> staging_trims{0};
};

class SubmitQueue;

class Context {
public:
    Context(const int* ids, int n);
    ~Context();
    int device_count() const { return int(devs_.size()); }
    Device& device(int i) { return *devs_[i]; }
    int next_device() { return int(rr_.fetch_add(1, std::memory_order_relaxed) % devs_.size()); }

    std::shared_ptr<const PassPlan> pass(int filter, uint32_t n_in, uint32_t n_out, bool vertical);

    std::atomic<int> mode{0};
    std::atomic<uint64_t> launches{0};
    Stats stats;
    CopyPool copy_pool;

    // Plan device-resident jobs for `dev` (device pointers in descs).  Per-job failures are
    // reported through `status` (size n) and leave that job out of the plan.
    LaunchPlan plan(Device& dev, const JobDesc* descs, size_t n, int* status, bool exact);
    // Upload descriptors into `d_desc` (via pinned `h_desc`) and enqueue every launch on `stream`.
    void enqueue(Device& dev, LaunchPlan& lp, Buffer& h_desc, Buffer& d_desc, Buffer& d_scratch,
                 cudaStream_t stream, bool exact);
    // Enqueue only (descriptors already resident at `d_desc_base`).
    void launch_resident(const LaunchPlan& lp, const uint8_t* d_desc_base, cudaStream_t stream, bool exact);
    size_t desc_bytes(const LaunchPlan& lp) const;
    void fill_desc(const LaunchPlan& lp, uint8_t* host, float* scratch) const;

    // Host-buffer resize of one image through a lane of some device.
    void resize_host(const JobDesc& d, int* device_index_out);
    void resize_host_on(Device& dev, const JobDesc& d);
    void resize_batch_host(JobDesc* descs, size_t n, int* status, int* device_out);
    // Split host-buffer resize (ikc_resize_begin_u8 / ikc_resize_end): begin returns an opaque ticket (nullptr for a
    // raster answered without a kernel), end completes and frees it.  Both throw Error.
    void* begin_host(const JobDesc& d);
    void end_host(void* ticket);
    // Host-buffer resize through the coalescing submit queue (ikc_submit_u8); throws Error on this job's failure.
    void submit_host(const JobDesc& d);
    // One group of host jobs on one lane of `dev`: one staged upload, one plan, one launch per kernel variant, one
    // download.  Per-job results in status / errors (size n).
    // `lane`: a lane the caller already holds (else one is acquired for the call).
    void resize_group_host(Device& dev, const JobDesc* descs, size_t n, int* status, std::string* errors, Lane* lane = nullptr);

private:
    std::vector<std::unique_ptr<Device>> devs_;
    std::atomic<uint64_t> rr_{0};
    std::mutex pass_mu_;
    std::map<std::tuple<int, uint32_t, uint32_t, bool>, std::shared_ptr<const PassPlan>> passes_;
    std::vector<std::tuple<int, uint32_t, uint32_t, bool>> pass_order_;
    std::mutex submit_mu_;
    std::vector<std::unique_ptr<SubmitQueue>> submit_;   // one per device, created on first use
};

// Validation shared by every entry point; throws Error.
void validate_job(const JobDesc& d);

struct PreparedBatch {
    Context* ctx;
    int device_index;
    bool exact;  // arithmetic mode the batch was planned for
    LaunchPlan lp;
    Buffer d_desc, d_scratch;
};

}  // namespace ikc
