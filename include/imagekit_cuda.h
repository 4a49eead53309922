/*
 * imagekit_cuda.h -- C ABI of libimagekit_cuda.so, the B200 (sm_100a) drop-in for the one
 * data-parallel hot path of the imagekit service: the `resize_image` step.
 *
 * Reference interface this library replaces (citations into the reference repository):
 *   pub fn resize_image(img: DynamicImage, w: Option<u32>, h: Option<u32>)
 *       -> Result<DynamicImage, ImageKitError>                  src/transform.rs:62-66
 *   imported at src/lib.rs:32, called at src/lib.rs:180 (/img) and src/lib.rs:286 (/upload).
 * Its arithmetic is `DynamicImage::resize(tw, th, FilterType::Lanczos3)` (src/transform.rs:85-89)
 * of crate image 0.25.8 (Cargo.toml:20): dynimage.rs `resize`, math/utils.rs
 * `resize_dimensions`, imageops/sample.rs `resize`/`vertical_sample`/`horizontal_sample`.
 *
 * The reference has no plugin registry: the function itself is the seam.  A Rust crate
 * `imagekit-cuda` binds the entry points below through `extern "C"` (see INTEGRATION.md and
 * rust-image-transform_b200/crate/) and re-exports a `resize_image` with the reference's
 * exact signature.
 *
 * Rules of the boundary
 *   - plain pointers and sizes only; the caller owns every buffer it passes and the library
 *     never frees or retains caller memory past the call's return;
 *   - every function is thread-safe and re-entrant (the reference calls resize_image
 *     synchronously from N tokio worker threads: src/lib.rs:180,286);
 *   - nothing panics, aborts or throws across the boundary: status codes + ikc_last_error();
 *   - there is NO CPU fallback: without a usable CUDA device ikc_create fails with
 *     IKC_ERR_CUDA and nothing else can be called.
 */
#ifndef IMAGEKIT_CUDA_H
#define IMAGEKIT_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define IKC_API __attribute__((visibility("default")))
#else
#define IKC_API
#endif

#define IKC_VERSION_MAJOR 0
#define IKC_VERSION_MINOR 1

typedef struct ikc_ctx ikc_ctx;       /* process-wide context: devices, streams, staging, tables */
typedef struct ikc_batch ikc_batch;   /* a prepared, device-resident batch (replayable)          */

/* Status codes.  Rust maps any non-zero status to ImageKitError::TransformError(msg)
 * (src/lib.rs:38-39), which the handlers turn into HTTP 400 "Resize error: ..." (src/lib.rs:182,288). */
enum ikc_status {
    IKC_OK = 0,
    IKC_ERR_INVALID_ARG = 1,   /* null pointer, zero size, pitch < row bytes, bad enum          */
    IKC_ERR_UNSUPPORTED = 2,   /* layout the kernels do not handle (e.g. channels > 4)          */
    IKC_ERR_TOO_LARGE = 3,     /* request exceeds IKC_MAX_DIM / IKC_MAX_PIXELS (the reference has no
                                  bound on w/h -- src/lib.rs:61-63 -- and would try to allocate) */
    IKC_ERR_CUDA = 4,          /* CUDA runtime/driver error; text in ikc_last_error()           */
    IKC_ERR_OOM = 5            /* host or device allocation failed                              */
};

/* image::imageops::FilterType, same order as the crate's enum.  resize_image always uses
 * IKC_FILTER_LANCZOS3 (src/transform.rs:88); the others exist because imageops::resize has them. */
enum ikc_filter {
    IKC_FILTER_NEAREST = 0,
    IKC_FILTER_TRIANGLE = 1,
    IKC_FILTER_CATMULLROM = 2,
    IKC_FILTER_GAUSSIAN = 3,
    IKC_FILTER_LANCZOS3 = 4
};

/* What the dims rule decided (return value of ikc_target_dims when >= 0). */
enum ikc_dims_code {
    IKC_DIMS_RESAMPLE = 0,     /* resample to (*tw,*th)                                          */
    IKC_DIMS_PASSTHROUGH = 1,  /* w and h both None: input returned untouched (transform.rs:67-69) */
    IKC_DIMS_CLONE = 2,        /* requested size == current size: DynamicImage::resize clones    */
    IKC_DIMS_COPY = 3          /* aspect-fit size == current size: imageops::resize copies       */
};

/* Arithmetic mode of the resampling kernels. */
enum ikc_mode {
    IKC_MODE_FAST = 0,   /* fused single-launch kernels, FMA accumulation: max |delta| <= 1 LSB; downscales run
                            their vertical pass (and exact 2x upscales theirs) on the tensor cores as an exact integer
                            product (u8 source bytes x 16-bit fixed-point weights, s32 accumulation)          */
    IKC_MODE_EXACT = 1,  /* two-launch verification path: separate mul/add in ascending tap order,
                            vertical then horizontal, f32 intermediate in HBM: delta == 0 vs the
                            CPU restatement of image 0.25.8                                      */
    IKC_MODE_FAST_FP32 = 2, /* FAST without the tensor cores: downscales take the CUDA-core ring kernel
                            (the round-1 path; kept for A/B measurements and as the north-star's
                            "no tensor cores" variant); same +-1 LSB bound                        */
    IKC_MODE_FAST_F16 = 3   /* FAST with the f16 tensor-core kernel for downscales (f16 hi + lo weights = 22 bits,
                            f32 accumulation; a converter pass feeds it): finer weights, ~2x slower  */
};

#define IKC_MAX_DIM 65535u              /* per-axis bound on source and destination             */
#define IKC_MAX_PIXELS (1ull << 28)     /* per-image bound on width*height (268 MP)             */

/* Channel counts: 1=Luma, 2=LumaA, 3=Rgb, 4=Rgba.  Every `channels` argument may also carry a
 * destination channel count of 3 or 4 in bits 8..15 (0 there = same as the source): the store then
 * applies DynamicImage::to_rgb8() / to_rgba8() to the resized pixel (8-bit only), which
 * encode_image otherwise does on the CPU before every encode (src/transform.rs:123,131,140):
 * grey is replicated into r,g,b; alpha is the source's, or 255 if it has none; to_rgb8 drops it. */
#define IKC_CHANNELS(src_channels, dst_channels) ((int)(src_channels) | ((int)(dst_channels) << 8))

/* One resize of a batch.  Pointers are HOST pointers for ikc_resize_batch and DEVICE pointers
 * for ikc_batch_prepare.  Rows are `pitch` bytes apart, pixels are `channels` interleaved
 * samples (see IKC_CHANNELS; dst rows hold the destination channel count).  `status` and `device`
 * are written by the library. */
typedef struct ikc_job {
    const void* src;
    void* dst;
    uint32_t sw, sh;
    uint32_t dw, dh;
    size_t src_pitch;
    size_t dst_pitch;
    int32_t channels;
    int32_t filter;     /* enum ikc_filter */
    int32_t status;     /* out: enum ikc_status for this job */
    int32_t device;     /* out: index (into the ctx's device list) that ran it */
} ikc_job;

/* ---- context ------------------------------------------------------------------------------ */

/* Creates the process-wide context over `n` CUDA devices (device_ids == NULL or n <= 0: all
 * visible devices).  Per device: a pool of lanes (stream + pinned staging + device scratch) and
 * a weight-table cache.  Rust holds it in a OnceLock and shares it across all worker threads. */
IKC_API int ikc_create(const int* device_ids, int n, ikc_ctx** out);
IKC_API void ikc_destroy(ikc_ctx* ctx);
IKC_API int ikc_device_count(const ikc_ctx* ctx);
IKC_API int ikc_set_mode(ikc_ctx* ctx, int mode);          /* enum ikc_mode; default FAST */
IKC_API int ikc_get_mode(const ikc_ctx* ctx);
/* Number of resampling-kernel launches issued through this context so far. */
IKC_API uint64_t ikc_kernel_launches(const ikc_ctx* ctx);
/* Counters a /metrics handler can export for the resize step (the reference counts requests and latencies around its
 * handlers, src/lib.rs:318-338, :400-427; these are the ones only the library can see).  Monotonic since ikc_create;
 * ikc_get_stats fills a snapshot (fields are read one by one: a consistent total is not guaranteed while calls run). */
typedef struct ikc_stats_t {
    uint64_t calls;               /* images handed to a host-buffer entry point (ikc_resize_*, ikc_resize_batch, ikc_submit_u8) */
    uint64_t failed;              /* ... that returned an error */
    uint64_t trivial;             /* ... answered without a kernel (empty or same-size rasters) */
    uint64_t launches;            /* kernel launches, all entry points (== ikc_kernel_launches) */
    uint64_t launches_banded8t, launches_banded8, launches_banded_f16, launches_ring, launches_up2, launches_tile,
        launches_generic;         /* ... per kernel family */
    uint64_t src_bytes, dst_bytes;   /* raster bytes read / written by host-buffer calls */
    uint64_t busy_ns;             /* wall time inside host-buffer entry points, summed over calling threads */
    uint64_t table_hits, table_misses;   /* per-device weight-table cache */
    uint64_t submit_batches, submit_jobs;   /* ikc_submit_u8: launch groups formed / images they carried */
    uint64_t launches_banded8u;   /* the tensor-core 2x upscale kernel (appended: keeps the earlier fields in place) */
    uint64_t staging_trims;       /* times a lane gave back staging buffers above 256 MB after an oversized request */
} ikc_stats_t;
IKC_API int ikc_get_stats(const ikc_ctx* ctx, ikc_stats_t* out);
/* Thread-local text of the last failure on the calling thread ("" if none). */
IKC_API const char* ikc_last_error(void);
IKC_API int ikc_version(void);                               /* major*1000 + minor */

/* ---- dims rule (rows a1-a3 of the hot path) ------------------------------------------------- */

/* Replaces the size arithmetic of resize_image (src/transform.rs:67-87: f32 ratio, f32::round,
 * saturating cast, max(1)) followed by DynamicImage::resize's early-out and resize_dimensions
 * (image 0.25.8 math/utils.rs: f64 min-ratio fit-within, round, max(1), u32::MAX branch).
 * Returns an ikc_dims_code (>= 0) or -IKC_ERR_INVALID_ARG. Needs no context and no GPU. */
IKC_API int ikc_target_dims(uint32_t ow, uint32_t oh, int has_w, uint32_t w, int has_h, uint32_t h,
                            uint32_t* tw, uint32_t* th);

/* The library's size guard, callable before any buffer is allocated: IKC_OK, or IKC_ERR_TOO_LARGE when a
 * source or destination side exceeds IKC_MAX_DIM or an area exceeds IKC_MAX_PIXELS.  The reference has
 * no upper bound on w / h (src/lib.rs:61-63): the Rust wrapper calls this on the ikc_target_dims result
 * BEFORE sizing its output Vec, so an absurd request becomes a TransformError (HTTP 400, src/lib.rs:182)
 * instead of an allocation abort.  Needs no context and no GPU. */
IKC_API int ikc_check_dims(uint32_t sw, uint32_t sh, uint32_t dw, uint32_t dh);

/* Inspection of the window/weight table of one separable pass n_in -> n_out (the host-built table
 * the kernels consume; image 0.25.8 imageops/sample.rs window + weight loop).  With weights == NULL
 * returns the per-output stride needed; otherwise fills left[n_out], count[n_out] and
 * weights[n_out * stride] (zero padded) and returns the stride; 0 on bad arguments.  No GPU needed. */
IKC_API uint32_t ikc_pass_table(int filter, uint32_t n_in, uint32_t n_out, uint32_t* left, uint32_t* count,
                                float* weights, uint32_t stride);

/* What the planner derived for one separable pass (introspection for tests and tuning; no GPU
 * needed).  ring_k: most windows covering any source index (the fused downscale kernel needs 6 or 7).
 * uni_*: outputs [uni_lo, uni_hi) each end uni_step source indices after their predecessor with
 * full, bit-identical windows (0: none) -- the interior of an integer-ratio downscale.
 * up2_*: exact 2x upscale frame: up2_taps source indices from (o >> 1) + up2_off cover every
 * window (0: not a 2x upscale); source indices [up2_uni_lo, up2_uni_hi) share one set of weights. */
typedef struct ikc_pass_info_t {
    uint32_t stride, max_count;
    int32_t ring_k;
    int32_t uni_step, uni_lo, uni_hi;
    int32_t up2_taps, up2_off, up2_uni_lo, up2_uni_hi;
} ikc_pass_info_t;
IKC_API int ikc_pass_info(int filter, uint32_t n_in, uint32_t n_out, ikc_pass_info_t* out);

/* Band form of a downscale pass (n_in >= n_out), as the tensor-core vertical pass consumes it (inspection for
 * tests; no GPU needed).  The source indices are cut into chunks of 16; chunk k only touches the *band_n
 * (32 or 48) outputs starting at output 16 * gbase[k]; gbase[n_chunks] = number of 16-output groups.
 * tiles: per chunk two f16 operand tiles (hi, then lo) of band_n x 16 elements each, element (output n, index k)
 * at (k / 8) * band_n * 8 + (n / 8) * 64 + (n % 8) * 8 + (k % 8): the weight * 2^14 rounded to f16, and the f16 of the
 * rounding error.  Returns the number of chunks (0: the pass has no band form, or a buffer is too small);
 * with gbase == NULL and tiles == NULL only *band_n and the chunk count are returned. */
IKC_API uint32_t ikc_pass_band(int filter, uint32_t n_in, uint32_t n_out, uint32_t* band_n, int32_t* gbase,
                               uint16_t* tiles, size_t tiles_cap);

/* 8-bit band form of a downscale pass, as the integer tensor-core vertical pass consumes it (inspection for tests; no
 * GPU needed).  Chunks of 32 source indices; chunk k only touches the 32 outputs starting at output 8 * gbase[k]
 * (gbase[n_chunks] = number of 8-output groups).  Every weight is the integer round(w * 2^*shift), each output's
 * weights nudged to sum to exactly 2^*shift, split into *limbs signed base-256 digits (low digits in [-128, 127]).  tiles: per chunk one s8 operand
 * tile of (*limbs * 32) rows x 32 indices, row = (output - 8 * gbase[k]) * *limbs + digit with the most significant digit first,
 * element (row n, index k) at (k / 16) * (*limbs * 512) + (n / 8) * 128 + (n % 8) * 16 + (k % 16).  Returns the number of chunks
 * (0: the pass has no such form, or a buffer is too small); gbase == tiles == NULL: only *limbs, *shift and the count. */
IKC_API uint32_t ikc_pass_band8(int filter, uint32_t n_in, uint32_t n_out, uint32_t* limbs, uint32_t* shift, int32_t* gbase,
                                int8_t* tiles, size_t tiles_cap);

/* Row-band form of the same integer weights (the A operand of the kernel whose accumulator lanes are output rows;
 * inspection for tests; no GPU needed).  Band r = outputs [*rows r, *rows r + *rows), *rows <= 128 (the height that costs
 * least: 120 at exactly 2:1, where a band then spans 8 chunks instead of 9); its chunks of 32 source indices start at k_lo[r].  tiles: per band, chunk (*chunks of them) and digit (most significant first) one s8 operand tile of 128 rows
 * x 32 indices, element (row m, index k) at (k / 16) * 2048 + (m / 8) * 128 + (m % 8) * 16 + (k % 16).  Returns the
 * number of bands (0: the pass has no such form -- ratio well above 2, or no 8-bit band form at all -- or a buffer is too
 * small); k_lo == tiles == NULL: only *chunks, *rows and the count. */
IKC_API uint32_t ikc_pass_band8t(int filter, uint32_t n_in, uint32_t n_out, uint32_t* chunks, uint32_t* rows, int32_t* k_lo,
                                 int8_t* tiles, size_t tiles_cap);

/* ---- host-buffer entry points (the drop-in path; include H2D + D2H) -------------------------- */

/* Split form of ikc_resize_u8 for the upload-shaped pipeline (reference src/lib.rs:246-309: decode -> resize -> encode per
 * upload, duplicate decode in src/fetch.rs:104-121): ikc_resize_begin_u8 returns as soon as the copies and kernels of
 * this image are queued on a lane (pageable sources are staged before it returns, pinned ones -- ikc_host_alloc /
 * ikc_host_register -- must stay untouched until the end call); the worker decodes its next upload meanwhile, then
 * ikc_resize_end blocks until `dst` is complete and frees the ticket (also on failure).  A ticket holds one of the
 * device's lanes (4, growing to 16 under demand); when all are held, begin waits for an end -- so a thread must end the
 * ticket it holds before it begins another one, or many such threads can wait on each other for ever.  channels may be
 * IKC_CHANNELS(src, dst). */
typedef struct ikc_ticket ikc_ticket;
IKC_API int ikc_resize_begin_u8(ikc_ctx* ctx, const uint8_t* src, uint32_t sw, uint32_t sh, size_t src_pitch, int channels,
                                uint8_t* dst, uint32_t dw, uint32_t dh, size_t dst_pitch, int filter, ikc_ticket** out);
IKC_API int ikc_resize_end(ikc_ticket* ticket);

/* Same contract and result as ikc_resize_u8 / ikc_resize_convert_u8 (channels may be IKC_CHANNELS(src, dst)), for
 * handler threads that each resize one image at a time (src/lib.rs:180, :286): calls that arrive while the device is
 * busy are coalesced -- staged into one pinned block, uploaded with one copy, planned together and run as ONE launch per
 * kernel variant -- instead of one plan + descriptor upload + launch + two copies each.  Blocks until this caller's image
 * is done.  There is no waiting window: a call that finds nothing queued and nothing in flight runs at once on its own
 * thread, exactly as ikc_resize_u8 would; calls that arrive meanwhile form the next group. */
IKC_API int ikc_submit_u8(ikc_ctx* ctx, const uint8_t* src, uint32_t sw, uint32_t sh, size_t src_pitch, int channels,
                          uint8_t* dst, uint32_t dw, uint32_t dh, size_t dst_pitch, int filter);

/* Replaces imageops::resize(&buf, dw, dh, filter) for 8-bit rasters (image 0.25.8
 * imageops/sample.rs; reached from src/transform.rs:85-89 via resize_exact).  `src`/`dst` are
 * host buffers owned by the caller (Rust: `buf.as_raw().as_ptr()` and a pre-sized Vec<u8>).
 * Pinned buffers (ikc_host_alloc or cudaHostRegister'ed) are DMA'd directly; pageable ones go
 * through the lane's pinned staging.  Returns after the result is in `dst`.
 * Semantics kept from the crate: empty source -> zero-filled dst; same dims -> copy. */
IKC_API int ikc_resize_u8(ikc_ctx* ctx, const uint8_t* src, uint32_t sw, uint32_t sh, size_t src_pitch,
                          int channels, uint8_t* dst, uint32_t dw, uint32_t dh, size_t dst_pitch,
                          int filter);

/* ikc_resize_u8 followed by DynamicImage::to_rgb8() (dst_channels = 3) or to_rgba8() (dst_channels = 4)
 * of the result, fused into the kernels' store: replaces the resize at src/transform.rs:85-89 plus the
 * conversion at src/transform.rs:123/131 (jpeg, webp: rgb) or :140 (avif: rgba).  Same as passing
 * IKC_CHANNELS(src_channels, dst_channels) to ikc_resize_u8.  dst rows hold dw * dst_channels bytes. */
IKC_API int ikc_resize_convert_u8(ikc_ctx* ctx, const uint8_t* src, uint32_t sw, uint32_t sh, size_t src_pitch,
                                  int src_channels, uint8_t* dst, uint32_t dw, uint32_t dh, size_t dst_pitch,
                                  int dst_channels, int filter);

/* Same for 16-bit samples (Luma16/LumaA16/Rgb16/Rgba16 as produced by the PNG decoder);
 * pitches in bytes. */
IKC_API int ikc_resize_u16(ikc_ctx* ctx, const uint16_t* src, uint32_t sw, uint32_t sh, size_t src_pitch,
                           int channels, uint16_t* dst, uint32_t dw, uint32_t dh, size_t dst_pitch,
                           int filter);

/* Whole resize_image(img, w, h) for a tight 8-bit raster (src/transform.rs:62-90): dims rule +
 * Lanczos3.  `dst_capacity` bytes must hold tw*th*channels (query with ikc_target_dims first).
 * Writes the produced size to (*tw,*th) and returns an ikc_dims_code (>= 0; for codes 1..3 the
 * source bytes are copied to dst unchanged) or -(enum ikc_status) on failure. */
IKC_API int ikc_resize_image_u8(ikc_ctx* ctx, const uint8_t* src, uint32_t sw, uint32_t sh, int channels,
                                int has_w, uint32_t w, int has_h, uint32_t h, uint8_t* dst,
                                size_t dst_capacity, uint32_t* tw, uint32_t* th);

/* Batched resize of independent images (8-bit).  Jobs are sharded round-robin over the context's
 * devices (job i -> device i mod G); every device pipelines H2D / kernel / D2H over its lanes, and runs of small
 * images (<= 1 MB of pixels each) share one staged upload, one plan, one launch per kernel variant and one download.
 * No collective: images are independent.  Returns IKC_OK if every job succeeded, else the first
 * failing status; per-job results are in jobs[i].status. */
IKC_API int ikc_resize_batch(ikc_ctx* ctx, ikc_job* jobs, size_t n);

/* Pinned host memory for callers that want zero-copy staging (decode straight into it). */
IKC_API int ikc_host_alloc(size_t bytes, void** out);
IKC_API void ikc_host_free(void* p);
/* Page-lock memory the caller already owns (e.g. the Vec<u8> a decoder filled, src/transform.rs:27-43), so
 * that the entry points above DMA from / to it directly instead of staging through the lane's pinned
 * buffers.  Unregister before the memory is freed.  (cudaHostRegister / cudaHostUnregister.) */
IKC_API int ikc_host_register(void* p, size_t bytes);
IKC_API int ikc_host_unregister(void* p);

/* ---- device-resident entry points (no PCIe; what the roofline is measured on) --------------- */

/* One resize between device buffers on `device_index` (index into the ctx's device list),
 * executed on `stream` (a cudaStream_t passed as void*; NULL = the legacy default stream).
 * Returns once the stream has finished the resize (its descriptor staging is per call; the empty-source memset and
 * the same-size copy are synchronised the same way); use ikc_batch_prepare / ikc_batch_launch for fully asynchronous,
 * replayable launches.
 * Device buffers must span rows*pitch bytes; the fused kernel additionally wants a 16-byte
 * aligned base and pitch (otherwise the slower generic kernels are used). */
IKC_API int ikc_resize_u8_device(ikc_ctx* ctx, int device_index, void* stream, const uint8_t* d_src,
                                 uint32_t sw, uint32_t sh, size_t src_pitch, int channels, uint8_t* d_dst,
                                 uint32_t dw, uint32_t dh, size_t dst_pitch, int filter);

/* Prepare `n` device-resident jobs (8-bit) as one replayable batch: plans every job, uploads the
 * weight tables and work-item list once.  ikc_batch_launch enqueues the whole batch on `stream`
 * (one launch per kernel family present, normally one).  jobs[i].status is set at prepare time.  If some jobs are
 * rejected the call returns the first failing status and *out is STILL a valid batch over the accepted jobs (it must be
 * freed); *out is NULL only when nothing could be prepared (bad arguments, CUDA failure).  The batch references the
 * weight tables it uses: synchronise the streams it was launched on before ikc_batch_free. */
IKC_API int ikc_batch_prepare(ikc_ctx* ctx, int device_index, ikc_job* jobs, size_t n, ikc_batch** out);
IKC_API int ikc_batch_launch(ikc_batch* b, void* stream);
/* Kernel launches one ikc_batch_launch issues. */
IKC_API int ikc_batch_launch_count(const ikc_batch* b);
/* Human-readable list of the kernels (and CTA counts) one ikc_batch_launch issues. */
IKC_API int ikc_batch_describe(const ikc_batch* b, char* out, size_t cap);
IKC_API void ikc_batch_free(ikc_batch* b);

#ifdef __cplusplus
}
#endif
#endif /* IMAGEKIT_CUDA_H */
