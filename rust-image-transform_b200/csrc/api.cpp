// api.cpp -- the extern "C" boundary declared in include/imagekit_cuda.h.  Nothing throws across it.
#include <cstdio>
#include <cstring>
#include <string>
#include <new>
#include <vector>

#include "context.hpp"
#include "imagekit_cuda.h"
#include "launch.hpp"

using namespace ikc;

struct ikc_ctx {
    Context impl;
    ikc_ctx(const int* ids, int n) : impl(ids, n) {}
};
struct ikc_batch {
    PreparedBatch impl;
};

namespace {

template <typename F>
int guarded(F&& f) {
    try {
        f();
        return IKC_OK;
    } catch (const Error& e) {
        set_last_error(e.what);
        return int(e.status);
    } catch (const std::bad_alloc&) {
        set_last_error("host allocation failed");
        return IKC_ERR_OOM;
    } catch (const std::exception& e) {
        set_last_error(std::string("internal error: ") + e.what());
        return IKC_ERR_CUDA;
    } catch (...) {
        set_last_error("internal error");
        return IKC_ERR_CUDA;
    }
}

// `ch` may carry the destination channel count in bits 8..15 (IKC_CHANNELS(src, dst)); 0 there = same as source.
JobDesc make_desc(const void* src, uint32_t sw, uint32_t sh, size_t sp, int ch, void* dst, uint32_t dw, uint32_t dh,
                  size_t dp, int filter, int bps) {
    const int out = (ch >> 8) & 0xff;
    if (ch < 0 || (ch >> 16) != 0) fail(kInvalidArg, "bad channels value");
    return JobDesc{src, dst, sw, sh, dw, dh, sp, dp, ch & 0xff, bps, filter, out};
}

}  // namespace

extern "C" {

int ikc_version(void) { return IKC_VERSION_MAJOR * 1000 + IKC_VERSION_MINOR; }
const char* ikc_last_error(void) { return last_error(); }

int ikc_create(const int* device_ids, int n, ikc_ctx** out) {
    if (!out) {
        set_last_error("out is null");
        return IKC_ERR_INVALID_ARG;
    }
    *out = nullptr;
    return guarded([&] { *out = new ikc_ctx(device_ids, n); });
}

void ikc_destroy(ikc_ctx* ctx) {
    try {
        delete ctx;
    } catch (...) {
    }
}

int ikc_device_count(const ikc_ctx* ctx) { return ctx ? ctx->impl.device_count() : 0; }

int ikc_set_mode(ikc_ctx* ctx, int mode) {
    if (!ctx || (mode != IKC_MODE_FAST && mode != IKC_MODE_EXACT && mode != IKC_MODE_FAST_FP32 && mode != IKC_MODE_FAST_F16)) {
        set_last_error("bad context or mode");
        return IKC_ERR_INVALID_ARG;
    }
    ctx->impl.mode.store(mode);
    return IKC_OK;
}
int ikc_get_mode(const ikc_ctx* ctx) { return ctx ? ctx->impl.mode.load() : -1; }
uint64_t ikc_kernel_launches(const ikc_ctx* ctx) { return ctx ? ctx->impl.launches.load() : 0; }

int ikc_target_dims(uint32_t ow, uint32_t oh, int has_w, uint32_t w, int has_h, uint32_t h, uint32_t* tw,
                    uint32_t* th) {
    if (!tw || !th) {
        set_last_error("tw/th is null");
        return -IKC_ERR_INVALID_ARG;
    }
    return target_dims(ow, oh, has_w != 0, w, has_h != 0, h, tw, th);
}

int ikc_check_dims(uint32_t sw, uint32_t sh, uint32_t dw, uint32_t dh) {
    return guarded([&] {
        if (sw > IKC_MAX_DIM || sh > IKC_MAX_DIM || dw > IKC_MAX_DIM || dh > IKC_MAX_DIM)
            fail(kTooLarge, "image dimension exceeds IKC_MAX_DIM");
        if (uint64_t(sw) * sh > IKC_MAX_PIXELS || uint64_t(dw) * dh > IKC_MAX_PIXELS)
            fail(kTooLarge, "image area exceeds IKC_MAX_PIXELS");
    });
}

uint32_t ikc_pass_table(int filter, uint32_t n_in, uint32_t n_out, uint32_t* left, uint32_t* count, float* weights,
                        uint32_t stride) {
    uint32_t need = 0;
    guarded([&] {
        auto p = build_pass(filter, n_in, n_out);
        if (!p) return;
        need = p->stride;
        if (!weights) return;
        if (!left || !count || stride < p->stride) {
            need = 0;
            return;
        }
        for (uint32_t o = 0; o < n_out; ++o) {
            left[o] = uint32_t(p->left[o]);
            count[o] = uint32_t(p->count[o]);
            for (uint32_t i = 0; i < stride; ++i)
                weights[size_t(o) * stride + i] = i < p->stride ? p->w[size_t(o) * p->stride + i] : 0.0f;
        }
        need = stride;
    });
    return need;
}

uint32_t ikc_pass_band(int filter, uint32_t n_in, uint32_t n_out, uint32_t* band_n, int32_t* gbase, uint16_t* tiles,
                       size_t tiles_cap) {
    uint32_t chunks = 0;
    guarded([&] {
        auto p = build_pass(filter, n_in, n_out);
        if (!p || p->band_n == 0) return;
        const uint32_t n = uint32_t(p->band_gbase.size()) - 1;
        if (band_n) *band_n = uint32_t(p->band_n);
        if (gbase || tiles) {
            if (!gbase || !tiles || tiles_cap < p->band_tiles.size()) return;
            std::memcpy(gbase, p->band_gbase.data(), sizeof(int32_t) * p->band_gbase.size());
            std::memcpy(tiles, p->band_tiles.data(), sizeof(uint16_t) * p->band_tiles.size());
        }
        chunks = n;
    });
    return chunks;
}

uint32_t ikc_pass_band8(int filter, uint32_t n_in, uint32_t n_out, uint32_t* limbs, uint32_t* shift, int32_t* gbase, int8_t* tiles,
                        size_t tiles_cap) {
    uint32_t chunks = 0;
    guarded([&] {
        auto p = build_pass(filter, n_in, n_out);
        if (!p || p->band8.limbs == 0) return;
        if (limbs) *limbs = uint32_t(p->band8.limbs);
        if (shift) *shift = uint32_t(p->band8.shift);
        if (gbase || tiles) {
            if (!gbase || !tiles || tiles_cap < p->band8.tiles.size()) return;
            std::memcpy(gbase, p->band8.gbase.data(), sizeof(int32_t) * p->band8.gbase.size());
            std::memcpy(tiles, p->band8.tiles.data(), p->band8.tiles.size());
        }
        chunks = uint32_t(p->band8.gbase.size()) - 1;
    });
    return chunks;
}

uint32_t ikc_pass_band8t(int filter, uint32_t n_in, uint32_t n_out, uint32_t* chunks, uint32_t* rows, int32_t* k_lo, int8_t* tiles,
                         size_t tiles_cap) {
    uint32_t bands = 0;
    guarded([&] {
        auto p = build_pass(filter, n_in, n_out);
        if (!p || p->band8t.chunks == 0) return;
        if (chunks) *chunks = uint32_t(p->band8t.chunks);
        if (rows) *rows = uint32_t(p->band8t.rows);
        if (k_lo || tiles) {
            if (!k_lo || !tiles || tiles_cap < p->band8t.tiles.size()) return;
            std::memcpy(k_lo, p->band8t.k_lo.data(), sizeof(int32_t) * p->band8t.k_lo.size());
            std::memcpy(tiles, p->band8t.tiles.data(), p->band8t.tiles.size());
        }
        bands = uint32_t(p->band8t.k_lo.size());
    });
    return bands;
}

int ikc_pass_info(int filter, uint32_t n_in, uint32_t n_out, ikc_pass_info_t* out) {
    if (!out) return IKC_ERR_INVALID_ARG;
    return guarded([&] {
        auto p = build_pass(filter, n_in, n_out);
        if (!p) fail(kInvalidArg, "cannot plan pass");
        out->stride = p->stride;
        out->max_count = p->max_count;
        out->ring_k = p->ring_k;
        out->uni_step = p->uni_step;
        out->uni_lo = p->uni_lo;
        out->uni_hi = p->uni_hi;
        out->up2_taps = p->up2_taps;
        out->up2_off = p->up2_off;
        out->up2_uni_lo = p->up2_uni_lo;
        out->up2_uni_hi = p->up2_uni_hi;
    });
}

int ikc_resize_u8(ikc_ctx* ctx, const uint8_t* src, uint32_t sw, uint32_t sh, size_t src_pitch, int channels,
                  uint8_t* dst, uint32_t dw, uint32_t dh, size_t dst_pitch, int filter) {
    if (!ctx) {
        set_last_error("ctx is null");
        return IKC_ERR_INVALID_ARG;
    }
    return guarded([&] {
        ctx->impl.resize_host(make_desc(src, sw, sh, src_pitch, channels, dst, dw, dh, dst_pitch, filter, 1), nullptr);
    });
}

int ikc_submit_u8(ikc_ctx* ctx, const uint8_t* src, uint32_t sw, uint32_t sh, size_t src_pitch, int channels, uint8_t* dst,
                  uint32_t dw, uint32_t dh, size_t dst_pitch, int filter) {
    if (!ctx) {
        set_last_error("ctx is null");
        return IKC_ERR_INVALID_ARG;
    }
    return guarded([&] { ctx->impl.submit_host(make_desc(src, sw, sh, src_pitch, channels, dst, dw, dh, dst_pitch, filter, 1)); });
}

struct ikc_ticket {
    ikc_ctx* ctx;
    void* impl;   // Context's ticket; nullptr: answered at begin (empty or same-size raster)
};

int ikc_resize_begin_u8(ikc_ctx* ctx, const uint8_t* src, uint32_t sw, uint32_t sh, size_t src_pitch, int channels, uint8_t* dst,
                        uint32_t dw, uint32_t dh, size_t dst_pitch, int filter, ikc_ticket** out) {
    if (out) *out = nullptr;
    if (!ctx || !out) {
        set_last_error("ctx or out is null");
        return IKC_ERR_INVALID_ARG;
    }
    return guarded([&] {
        auto t = std::make_unique<ikc_ticket>();
        t->ctx = ctx;
        t->impl = ctx->impl.begin_host(make_desc(src, sw, sh, src_pitch, channels, dst, dw, dh, dst_pitch, filter, 1));
        *out = t.release();
    });
}

int ikc_resize_end(ikc_ticket* ticket) {
    if (!ticket) {
        set_last_error("ticket is null");
        return IKC_ERR_INVALID_ARG;
    }
    std::unique_ptr<ikc_ticket> t(ticket);
    return guarded([&] { t->ctx->impl.end_host(t->impl); });
}

int ikc_get_stats(const ikc_ctx* ctx, ikc_stats_t* out) {
    if (!ctx || !out) {
        set_last_error("ctx or out is null");
        return IKC_ERR_INVALID_ARG;
    }
    const Stats& s = ctx->impl.stats;
    auto ld = [](const std::atomic<uint64_t>& a) { return a.load(std::memory_order_relaxed); };
    out->calls = ld(s.calls); out->failed = ld(s.failed); out->trivial = ld(s.trivial);
    out->launches = ctx->impl.launches.load(std::memory_order_relaxed);
    out->launches_banded8t = ld(s.launches_by_family[0]); out->launches_banded8 = ld(s.launches_by_family[1]);
    out->launches_banded_f16 = ld(s.launches_by_family[2]); out->launches_ring = ld(s.launches_by_family[3]);
    out->launches_up2 = ld(s.launches_by_family[4]); out->launches_tile = ld(s.launches_by_family[5]);
    out->launches_generic = ld(s.launches_by_family[6]);
    out->src_bytes = ld(s.src_bytes); out->dst_bytes = ld(s.dst_bytes); out->busy_ns = ld(s.busy_ns);
    out->table_hits = ld(s.table_hits); out->table_misses = ld(s.table_misses);
    out->submit_batches = ld(s.submit_batches); out->submit_jobs = ld(s.submit_jobs);
    out->launches_banded8u = ld(s.launches_by_family[7]);
    out->staging_trims = ld(s.staging_trims);
    return IKC_OK;
}

int ikc_resize_convert_u8(ikc_ctx* ctx, const uint8_t* src, uint32_t sw, uint32_t sh, size_t src_pitch, int src_channels,
                          uint8_t* dst, uint32_t dw, uint32_t dh, size_t dst_pitch, int dst_channels, int filter) {
    if (!ctx) {
        set_last_error("ctx is null");
        return IKC_ERR_INVALID_ARG;
    }
    return guarded([&] {
        if (src_channels < 1 || src_channels > 255 || dst_channels < 0 || dst_channels > 255)
            fail(kInvalidArg, "bad channel count");
        ctx->impl.resize_host(make_desc(src, sw, sh, src_pitch, IKC_CHANNELS(src_channels, dst_channels), dst, dw, dh,
                                        dst_pitch, filter, 1),
                              nullptr);
    });
}

int ikc_resize_u16(ikc_ctx* ctx, const uint16_t* src, uint32_t sw, uint32_t sh, size_t src_pitch, int channels,
                   uint16_t* dst, uint32_t dw, uint32_t dh, size_t dst_pitch, int filter) {
    if (!ctx) {
        set_last_error("ctx is null");
        return IKC_ERR_INVALID_ARG;
    }
    return guarded([&] {
        ctx->impl.resize_host(make_desc(src, sw, sh, src_pitch, channels, dst, dw, dh, dst_pitch, filter, 2), nullptr);
    });
}

int ikc_resize_image_u8(ikc_ctx* ctx, const uint8_t* src, uint32_t sw, uint32_t sh, int channels, int has_w,
                        uint32_t w, int has_h, uint32_t h, uint8_t* dst, size_t dst_capacity, uint32_t* tw,
                        uint32_t* th) {
    if (!ctx || !tw || !th) {
        set_last_error("null argument");
        return -IKC_ERR_INVALID_ARG;
    }
    int code = 0;
    const int rc = guarded([&] {
        if (channels < 1 || channels > 4) fail(kInvalidArg, "channels must be 1..4");
        code = target_dims(sw, sh, has_w != 0, w, has_h != 0, h, tw, th);
        if (*tw > IKC_MAX_DIM || *th > IKC_MAX_DIM || uint64_t(*tw) * *th > IKC_MAX_PIXELS)
            fail(kTooLarge, "requested size exceeds IKC_MAX_DIM / IKC_MAX_PIXELS");
        const size_t need = size_t(*tw) * *th * size_t(channels);
        if (need > dst_capacity) fail(kInvalidArg, "dst_capacity too small for the target size");
        if (need && !dst) fail(kInvalidArg, "dst is null");
        if (code != IKC_DIMS_RESAMPLE) {  // passthrough / clone / copy: bytes unchanged
            if (need) std::memcpy(dst, src, need);
            return;
        }
        ctx->impl.resize_host(make_desc(src, sw, sh, size_t(sw) * channels, channels, dst, *tw, *th,
                                        size_t(*tw) * channels, IKC_FILTER_LANCZOS3, 1),
                              nullptr);
    });
    return rc == IKC_OK ? code : -rc;
}

int ikc_resize_batch(ikc_ctx* ctx, ikc_job* jobs, size_t n) {
    if (!ctx || (!jobs && n)) {
        set_last_error("null argument");
        return IKC_ERR_INVALID_ARG;
    }
    int first_bad = IKC_OK;
    const int rc = guarded([&] {
        std::vector<JobDesc> d(n);
        std::vector<int> st(n, 0), dev(n, 0);
        for (size_t i = 0; i < n; ++i)
            d[i] = make_desc(jobs[i].src, jobs[i].sw, jobs[i].sh, jobs[i].src_pitch, jobs[i].channels, jobs[i].dst,
                             jobs[i].dw, jobs[i].dh, jobs[i].dst_pitch, jobs[i].filter, 1);
        ctx->impl.resize_batch_host(d.data(), n, st.data(), dev.data());
        for (size_t i = 0; i < n; ++i) {
            jobs[i].status = st[i];
            jobs[i].device = dev[i];
            if (st[i] != IKC_OK && first_bad == IKC_OK) first_bad = st[i];
        }
    });
    return rc != IKC_OK ? rc : first_bad;
}

int ikc_host_alloc(size_t bytes, void** out) {
    if (!out) return IKC_ERR_INVALID_ARG;
    *out = nullptr;
    return guarded([&] { check_cuda(cudaMallocHost(out, bytes ? bytes : 1), "cudaMallocHost"); });
}
void ikc_host_free(void* p) {
    if (p) cudaFreeHost(p);
}
int ikc_host_register(void* p, size_t bytes) {
    if (!p || !bytes) return IKC_ERR_INVALID_ARG;
    return guarded([&] { check_cuda(cudaHostRegister(p, bytes, cudaHostRegisterPortable), "cudaHostRegister"); });
}
int ikc_host_unregister(void* p) {
    if (!p) return IKC_ERR_INVALID_ARG;
    return guarded([&] { check_cuda(cudaHostUnregister(p), "cudaHostUnregister"); });
}

int ikc_resize_u8_device(ikc_ctx* ctx, int device_index, void* stream, const uint8_t* d_src, uint32_t sw,
                         uint32_t sh, size_t src_pitch, int channels, uint8_t* d_dst, uint32_t dw, uint32_t dh,
                         size_t dst_pitch, int filter) {
    if (!ctx || device_index < 0 || device_index >= ctx->impl.device_count()) {
        set_last_error("bad context or device index");
        return IKC_ERR_INVALID_ARG;
    }
    return guarded([&] {
        Context& c = ctx->impl;
        Device& dev = c.device(device_index);
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        check_cuda(cudaSetDevice(dev.ordinal()), "cudaSetDevice");
        JobDesc d = make_desc(d_src, sw, sh, src_pitch, channels, d_dst, dw, dh, dst_pitch, filter, 1);
        validate_job(d);
        const size_t row = size_t(dw) * d.oc();
        if (dw == 0 || dh == 0) return;
        if ((sw == 0 || sh == 0 || (sw == dw && sh == dh)) && d.oc() != d.channels)
            fail(kUnsupported, "device entry point: empty or same-size rasters cannot be combined with a channel conversion");
        if (sw == 0 || sh == 0) {
            check_cuda(cudaMemset2DAsync(d_dst, dst_pitch, 0, row, dh, s), "memset");
            check_cuda(cudaStreamSynchronize(s), "resize (stream sync)");
            return;
        }
        if (sw == dw && sh == dh) {
            check_cuda(cudaMemcpy2DAsync(d_dst, dst_pitch, d_src, src_pitch, row, dh, cudaMemcpyDeviceToDevice, s),
                       "copy");
            check_cuda(cudaStreamSynchronize(s), "resize (stream sync)");
            return;
        }
        const bool exact = c.mode.load() == 1;
        int status = kOk;
        LaunchPlan lp = c.plan(dev, &d, 1, &status, exact);
        if (status != kOk) fail(Status(status), last_error());
        // Descriptor staging comes from a lane; the lane is held until the caller's stream has
        // consumed the upload (the scratch of the generic path lives in the lane too).
        Lane* l = dev.acquire_lane();
        try {
            c.enqueue(dev, lp, l->h_desc, l->d_desc, l->d_scratch, s, exact);
            check_cuda(cudaStreamSynchronize(s), "resize (stream sync)");
        } catch (...) {
            dev.release_lane(l);
            throw;
        }
        dev.release_lane(l);
    });
}

int ikc_batch_prepare(ikc_ctx* ctx, int device_index, ikc_job* jobs, size_t n, ikc_batch** out) {
    if (!ctx || !out || (!jobs && n) || device_index < 0 || device_index >= ctx->impl.device_count()) {
        set_last_error("bad argument");
        return IKC_ERR_INVALID_ARG;
    }
    *out = nullptr;
    int first_bad = IKC_OK;
    const int rc = guarded([&] {
        Context& c = ctx->impl;
        Device& dev = c.device(device_index);
        check_cuda(cudaSetDevice(dev.ordinal()), "cudaSetDevice");
        std::vector<JobDesc> d(n);
        std::vector<int> st(n, 0);
        for (size_t i = 0; i < n; ++i)
            d[i] = make_desc(jobs[i].src, jobs[i].sw, jobs[i].sh, jobs[i].src_pitch, jobs[i].channels, jobs[i].dst,
                             jobs[i].dw, jobs[i].dh, jobs[i].dst_pitch, jobs[i].filter, 1);
        const bool exact = c.mode.load() == 1;
        auto b = std::unique_ptr<ikc_batch>(new ikc_batch{PreparedBatch{&c, device_index, exact, {}, {}, {}}});
        b->impl.lp = c.plan(dev, d.data(), n, st.data(), exact);
        for (size_t i = 0; i < n; ++i) {
            jobs[i].status = st[i];
            jobs[i].device = device_index;
            if (st[i] != IKC_OK && first_bad == IKC_OK) first_bad = st[i];
        }
        LaunchPlan& lp = b->impl.lp;
        if (lp.scratch_floats) b->impl.d_scratch.reserve(lp.scratch_floats * sizeof(float));
        for (int idx : lp.generic_jobs) lp.jobs[idx].tmp = static_cast<float*>(b->impl.d_scratch.p);
        const size_t bytes = c.desc_bytes(lp);
        std::vector<uint8_t> host(bytes);
        c.fill_desc(lp, host.data(), static_cast<float*>(b->impl.d_scratch.p));
        b->impl.d_desc.reserve(bytes);
        check_cuda(cudaMemcpy(b->impl.d_desc.p, host.data(), bytes, cudaMemcpyHostToDevice), "upload descriptors");
        check_cuda(cudaDeviceSynchronize(), "upload descriptors (sync)");  // pageable source: wait for the DMA
        *out = b.release();
    });
    return rc != IKC_OK ? rc : first_bad;
}

int ikc_batch_launch(ikc_batch* b, void* stream) {
    if (!b) {
        set_last_error("batch is null");
        return IKC_ERR_INVALID_ARG;
    }
    return guarded([&] {
        Context& c = *b->impl.ctx;
        check_cuda(cudaSetDevice(c.device(b->impl.device_index).ordinal()), "cudaSetDevice");
        c.launch_resident(b->impl.lp, static_cast<const uint8_t*>(b->impl.d_desc.p), static_cast<cudaStream_t>(stream),
                          b->impl.exact);
    });
}

int ikc_batch_launch_count(const ikc_batch* b) { return b ? b->impl.lp.launches() : 0; }

int ikc_batch_describe(const ikc_batch* b, char* out, size_t cap) {
    if (!b || !out || cap == 0) return IKC_ERR_INVALID_ARG;
    std::string s;
    for (auto& g : b->impl.lp.groups) {
        if (!s.empty()) s += "; ";
        if (g.band8u_taps) s += "banded8u_kernel<" + std::to_string(g.channels) + "," + std::to_string(g.band8u_taps) + ">";
        else if (g.band8t) s += "banded8t_kernel<4> (2 digits, " + std::to_string(g.b8tgeom.chunks) + " chunks per band)";
        else if (g.band8_limbs) s += "banded8_kernel<" + std::to_string(g.channels) + (g.convert ? ",conv" : "") + "> (" + std::to_string(g.band8_limbs) + " digits)";
        else if (g.band_n) s += "banded_kernel<" + std::to_string(g.channels) + (g.convert ? ",conv" : "") + "> (band_n " + std::to_string(g.band_n) + ")";
        else if (g.up_taps) s += "up2_kernel<" + std::to_string(g.channels) + "," + std::to_string(g.up_taps) + ">";
        else if (g.kv == 0) s += g.bps == 2 ? "tile_kernel<u16>" : "tile_kernel";
        else s += "fused_ring_kernel<" + std::to_string(g.channels) + "," + std::to_string(g.kv) + "," + std::to_string(g.kh) + "," + std::to_string(g.sv) + "," + std::to_string(g.sh) + ">";
        s += " x " + std::to_string(g.items.size()) +
             (g.up_taps ? " tiles (persistent CTAs)" : (g.band8_limbs && !g.band8t && !g.band8u_taps) ? " items (persistent CTAs, one per SM)" : " CTAs");
    }
    if (!b->impl.lp.generic_jobs.empty()) {
        if (!s.empty()) s += "; ";
        s += "vertical_generic+horizontal_generic x " + std::to_string(b->impl.lp.generic_jobs.size()) + " jobs";
    }
    std::snprintf(out, cap, "%s", s.c_str());
    return IKC_OK;
}

void ikc_batch_free(ikc_batch* b) {
    if (!b) return;
    try {
        cudaSetDevice(b->impl.ctx->device(b->impl.device_index).ordinal());
        cudaDeviceSynchronize();
        b->impl.d_desc.release();
        b->impl.d_scratch.release();
        delete b;
    } catch (...) {
    }
}

}  // extern "C"
