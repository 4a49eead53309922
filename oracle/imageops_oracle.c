/*
 * imageops_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the arithmetic behind the reference's hot path
 *     resize_image(img, w, h)            /root/reference/src/transform.rs:62-90
 * whose pixel math lives in the third-party crate `image` = 0.25.8
 * (/root/reference/Cargo.toml:20, /root/reference/Cargo.lock:987-990), files
 * src/dynimage.rs (DynamicImage::resize), src/math/utils.rs (resize_dimensions)
 * and src/imageops/sample.rs (resize, vertical_sample, horizontal_sample and the
 * filter kernels).  That crate's source is NOT vendored under /root/reference and
 * there is no Rust toolchain in this image, so this file restates the published
 * algorithm operation by operation:
 *   - all sample arithmetic in IEEE binary32, multiply THEN add (Rust never
 *     contracts to FMA), taps accumulated in ascending index order;
 *   - vertical pass first into an unclamped, unrounded f32 intermediate, then the
 *     horizontal pass, clamp to [0,max] and round half away from zero at the end;
 *   - weights from libm sinf/expf exactly as Rust's f32::sin / f32::exp lower to
 *     on x86_64-unknown-linux-gnu;
 *   - truncated edge windows are renormalised, never clamp-replicated.
 *
 * PARITY STATUS: "parity unpinned" for pixel values.  The reference's own tests
 * (tests/transform.rs:11-96, 224-269) pin output DIMENSIONS only (every test image
 * is all-zero); those nine known answers are checked in tests/test_oracle_dims.py.
 * No golden pixel vector exists anywhere in the reference tree.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this file.  The product library never links it.
 *
 * Build: gcc -O3 -ffp-contract=off -fno-fast-math -shared -fPIC (see oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

enum { ORC_NEAREST = 0, ORC_TRIANGLE = 1, ORC_CATMULLROM = 2, ORC_GAUSSIAN = 3, ORC_LANCZOS3 = 4 };

/* ---- Rust cast / rounding semantics ------------------------------------ */

/* `x as u32` for f32: saturating, NaN -> 0. */
static uint32_t sat_f32_to_u32(float x) {
    if (!(x == x)) return 0u;
    if (x <= 0.0f) return 0u;
    if (x >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)x;
}
/* `x as u64` for f64: saturating, NaN -> 0. */
static uint64_t sat_f64_to_u64(double x) {
    if (!(x == x)) return 0ull;
    if (x <= 0.0) return 0ull;
    if (x >= 18446744073709551616.0) return 0xFFFFFFFFFFFFFFFFull;
    return (uint64_t)x;
}
static uint32_t sat_f64_to_u32(double x) {
    if (!(x == x)) return 0u;
    if (x <= 0.0) return 0u;
    if (x >= 4294967296.0) return 0xFFFFFFFFu;
    return (uint32_t)x;
}
/* `x as i64` for f32: saturating, NaN -> 0. */
static int64_t sat_f32_to_i64(float x) {
    if (!(x == x)) return 0;
    if (x <= -9223372036854775808.0f) return INT64_MIN;
    if (x >= 9223372036854775808.0f) return INT64_MAX;
    return (int64_t)x;
}
/* f32::round / f64::round: half away from zero == C roundf / round. */

/* ---- dims: transform.rs:62-90 -> DynamicImage::resize -> resize_dimensions --- */

/* image 0.25.8 src/math/utils.rs resize_dimensions(width,height,nwidth,nheight,fill) */
static void resize_dimensions(uint32_t width, uint32_t height, uint32_t nwidth, uint32_t nheight,
                              int fill, uint32_t* ow, uint32_t* oh) {
    double wratio = (double)nwidth / (double)width;
    double hratio = (double)nheight / (double)height;
    /* f64::max / f64::min: NaN-ignoring like fmax/fmin */
    double ratio = fill ? fmax(wratio, hratio) : fmin(wratio, hratio);
    uint64_t nw = sat_f64_to_u64(round((double)width * ratio));
    uint64_t nh = sat_f64_to_u64(round((double)height * ratio));
    if (nw < 1) nw = 1;
    if (nh < 1) nh = 1;
    if (nw > 0xFFFFFFFFull) {
        double r = 4294967295.0 / (double)width;
        uint32_t h2 = sat_f64_to_u32(round((double)height * r));
        *ow = 0xFFFFFFFFu;
        *oh = h2 < 1 ? 1 : h2;
    } else if (nh > 0xFFFFFFFFull) {
        double r = 4294967295.0 / (double)height;
        uint32_t w2 = sat_f64_to_u32(round((double)width * r));
        *ow = w2 < 1 ? 1 : w2;
        *oh = 0xFFFFFFFFu;
    } else {
        *ow = (uint32_t)nw;
        *oh = (uint32_t)nh;
    }
}

/*
 * Full dims rule of the hot path.  Returns:
 *   0  a resample to (*tw,*th) happens
 *   1  (None,None): input returned untouched          transform.rs:67-69
 *   2  requested == current dims: DynamicImage::resize clones        (dynimage.rs)
 *   3  fit-within result == current dims: imageops::resize copies    (sample.rs)
 */
ORC_API int orc_target_dims(uint32_t ow, uint32_t oh, int has_w, uint32_t w, int has_h, uint32_t h,
                            uint32_t* tw, uint32_t* th) {
    if (!has_w && !has_h) { *tw = ow; *th = oh; return 1; }
    uint32_t target_w, target_h;
    if (has_w) target_w = w;
    else {                                       /* transform.rs:74-77 */
        float ratio = (float)h / (float)oh;
        target_w = sat_f32_to_u32(roundf((float)ow * ratio));
    }
    if (has_h) target_h = h;
    else {                                       /* transform.rs:79-82 */
        float ratio = (float)w / (float)ow;
        target_h = sat_f32_to_u32(roundf((float)oh * ratio));
    }
    if (target_w < 1) target_w = 1;              /* transform.rs:86-87 */
    if (target_h < 1) target_h = 1;
    if (target_w == ow && target_h == oh) { *tw = ow; *th = oh; return 2; }
    resize_dimensions(ow, oh, target_w, target_h, 0, tw, th);
    if (*tw == ow && *th == oh) return 3;
    return 0;
}

/* ---- filter kernels: image 0.25.8 src/imageops/sample.rs ---------------- */

static const float PI_F32 = 3.14159274101257324f; /* core::f32::consts::PI */

static float sinc(float t) {
    float a = t * PI_F32;
    if (t == 0.0f) return 1.0f;
    return sinf(a) / a;
}
static float lanczos3_kernel(float x) {
    const float t = 3.0f;
    if (fabsf(x) < t) return sinc(x) * sinc(x / t);
    return 0.0f;
}
static float powi2(float a) { return a * a; }
static float powi3(float a) { float a2 = a * a; return a * a2; } /* __powisf2 order; commutative */
static float bc_cubic_spline(float x, float b, float c) {
    float a = fabsf(x);
    float k;
    if (a < 1.0f) {
        k = (12.0f - 9.0f * b - 6.0f * c) * powi3(a) + (-18.0f + 12.0f * b + 6.0f * c) * powi2(a) +
            (6.0f - 2.0f * b);
    } else if (a < 2.0f) {
        k = (-b - 6.0f * c) * powi3(a) + (6.0f * b + 30.0f * c) * powi2(a) +
            (-12.0f * b - 48.0f * c) * a + (8.0f * b + 24.0f * c);
    } else {
        k = 0.0f;
    }
    return k / 6.0f;
}
static float catmullrom_kernel(float x) { return bc_cubic_spline(x, 0.0f, 0.5f); }
static float gaussian(float x, float r) {
    return (1.0f / (sqrtf(2.0f * PI_F32) * r)) * expf(-powi2(x) / (2.0f * powi2(r)));
}
static float gaussian_kernel(float x) { return gaussian(x, 0.5f); }
static float triangle_kernel(float x) { return fabsf(x) < 1.0f ? 1.0f - fabsf(x) : 0.0f; }
static float box_kernel(float x) { (void)x; return 1.0f; }

typedef float (*kernel_fn)(float);
static int pick_filter(int filter, kernel_fn* k, float* support) {
    switch (filter) {
        case ORC_NEAREST:    *k = box_kernel;        *support = 0.0f; return 0;
        case ORC_TRIANGLE:   *k = triangle_kernel;   *support = 1.0f; return 0;
        case ORC_CATMULLROM: *k = catmullrom_kernel; *support = 2.0f; return 0;
        case ORC_GAUSSIAN:   *k = gaussian_kernel;   *support = 3.0f; return 0;
        case ORC_LANCZOS3:   *k = lanczos3_kernel;   *support = 3.0f; return 0;
    }
    return -1;
}

ORC_API float orc_kernel(int filter, float x) {
    kernel_fn k; float s;
    if (pick_filter(filter, &k, &s)) return NAN;
    return k(x);
}

/* window + weights for one output index; shared by both passes (identical code upstream) */
static uint32_t window(kernel_fn kern, float support, uint32_t n_in, float ratio, float sratio,
                       uint32_t o, uint32_t* left_out, float* ws /* >= right-left */) {
    float src_support = support * sratio;
    float inputx = ((float)o + 0.5f) * ratio;
    int64_t left = sat_f32_to_i64(floorf(inputx - src_support));
    if (left < 0) left = 0;
    if (left > (int64_t)n_in - 1) left = (int64_t)n_in - 1;
    int64_t right = sat_f32_to_i64(ceilf(inputx + src_support));
    if (right < left + 1) right = left + 1;
    if (right > (int64_t)n_in) right = (int64_t)n_in;
    inputx = inputx - 0.5f;
    float sum = 0.0f;
    uint32_t n = (uint32_t)(right - left);
    for (uint32_t i = 0; i < n; ++i) {
        float w = kern(((float)(uint32_t)(left + i) - inputx) / sratio);
        ws[i] = w;
        sum += w;
    }
    for (uint32_t i = 0; i < n; ++i) ws[i] /= sum;
    *left_out = (uint32_t)left;
    return n;
}

static uint32_t max_taps(float support, uint32_t n_in, uint32_t n_out) {
    float ratio = (float)n_in / (float)n_out;
    float sratio = ratio < 1.0f ? 1.0f : ratio;
    double m = 2.0 * (double)support * (double)sratio + 4.0;
    if (m > (double)n_in) m = (double)n_in;
    if (m < 1.0) m = 1.0;
    return (uint32_t)m + 1u;
}

/*
 * Weight table for one pass n_in -> n_out.  left[o], count[o] and weights[o*stride + i].
 * Pass stride = 0 and weights = NULL to query the needed stride (returned).
 */
ORC_API uint32_t orc_pass_table(int filter, uint32_t n_in, uint32_t n_out, uint32_t* left,
                                uint32_t* count, float* weights, uint32_t stride) {
    kernel_fn kern; float support;
    if (pick_filter(filter, &kern, &support) || n_in == 0 || n_out == 0) return 0;
    uint32_t cap = max_taps(support, n_in, n_out);
    if (!weights) return cap;
    float ratio = (float)n_in / (float)n_out;
    float sratio = ratio < 1.0f ? 1.0f : ratio;
    float* ws = (float*)malloc(sizeof(float) * (size_t)cap);
    for (uint32_t o = 0; o < n_out; ++o) {
        uint32_t l; uint32_t n = window(kern, support, n_in, ratio, sratio, o, &l, ws);
        left[o] = l; count[o] = n;
        for (uint32_t i = 0; i < stride; ++i) weights[(size_t)o * stride + i] = i < n ? ws[i] : 0.0f;
    }
    free(ws);
    return cap;
}

/* ---- the two passes ------------------------------------------------------ */

/* vertical_sample: src (u8 or u16, C interleaved) -> f32 tmp [nh][w][C] */
#define DEFINE_VERTICAL(NAME, T)                                                                    \
static void NAME(const T* src, uint32_t w, uint32_t h, size_t pitch_elems, int C, uint32_t nh,      \
                 kernel_fn kern, float support, float* tmp) {                                      \
    float ratio = (float)h / (float)nh;                                                            \
    float sratio = ratio < 1.0f ? 1.0f : ratio;                                                    \
    float* ws = (float*)malloc(sizeof(float) * (size_t)max_taps(support, h, nh));                  \
    for (uint32_t outy = 0; outy < nh; ++outy) {                                                   \
        uint32_t left; uint32_t n = window(kern, support, h, ratio, sratio, outy, &left, ws);      \
        float* trow = tmp + (size_t)outy * w * C;                                                  \
        for (uint32_t x = 0; x < w; ++x) {                                                         \
            float t[4] = {0.0f, 0.0f, 0.0f, 0.0f};                                                 \
            for (uint32_t i = 0; i < n; ++i) {                                                     \
                const T* p = src + (size_t)(left + i) * pitch_elems + (size_t)x * C;               \
                float wi = ws[i];                                                                  \
                for (int c = 0; c < C; ++c) t[c] = t[c] + (float)p[c] * wi;                        \
            }                                                                                      \
            for (int c = 0; c < C; ++c) trow[(size_t)x * C + c] = t[c];                            \
        }                                                                                          \
    }                                                                                              \
    free(ws);                                                                                      \
}
DEFINE_VERTICAL(vertical_sample_u8, uint8_t)
DEFINE_VERTICAL(vertical_sample_u16, uint16_t)

/* horizontal_sample: f32 tmp [h][w][C] -> dst (u8/u16), clamp + round half away */
#define DEFINE_HORIZONTAL(NAME, T, MAXV)                                                            \
static void NAME(const float* tmp, uint32_t w, uint32_t h, int C, uint32_t nw, kernel_fn kern,      \
                 float support, T* dst, size_t dst_pitch_elems) {                                  \
    const float maxv = MAXV, minv = 0.0f;                                                          \
    float ratio = (float)w / (float)nw;                                                            \
    float sratio = ratio < 1.0f ? 1.0f : ratio;                                                    \
    float* ws = (float*)malloc(sizeof(float) * (size_t)max_taps(support, w, nw));                  \
    for (uint32_t outx = 0; outx < nw; ++outx) {                                                   \
        uint32_t left; uint32_t n = window(kern, support, w, ratio, sratio, outx, &left, ws);      \
        for (uint32_t y = 0; y < h; ++y) {                                                         \
            float t[4] = {0.0f, 0.0f, 0.0f, 0.0f};                                                 \
            const float* row = tmp + (size_t)y * w * C;                                            \
            for (uint32_t i = 0; i < n; ++i) {                                                     \
                const float* p = row + (size_t)(left + i) * C;                                     \
                float wi = ws[i];                                                                  \
                for (int c = 0; c < C; ++c) t[c] = t[c] + p[c] * wi;                               \
            }                                                                                      \
            T* q = dst + (size_t)y * dst_pitch_elems + (size_t)outx * C;                           \
            for (int c = 0; c < C; ++c) {                                                          \
                float v = t[c];                                                                    \
                v = v < minv ? minv : (v > maxv ? maxv : v);                                       \
                q[c] = (T)roundf(v);                                                               \
            }                                                                                      \
        }                                                                                          \
    }                                                                                              \
    free(ws);                                                                                      \
}
DEFINE_HORIZONTAL(horizontal_sample_u8, uint8_t, 255.0f)
DEFINE_HORIZONTAL(horizontal_sample_u16, uint16_t, 65535.0f)

/*
 * imageops::resize(image, nwidth, nheight, filter) for 8-bit samples (resize_exact semantics:
 * the caller has already applied orc_target_dims).  Pitches are in BYTES.
 * Returns 0 ok, -1 bad argument, -2 out of memory.
 */
ORC_API int orc_resize_u8(const uint8_t* src, uint32_t sw, uint32_t sh, size_t src_pitch, int channels,
                          uint8_t* dst, uint32_t dw, uint32_t dh, size_t dst_pitch, int filter) {
    kernel_fn kern; float support;
    if (channels < 1 || channels > 4 || pick_filter(filter, &kern, &support)) return -1;
    if (dw == 0 || dh == 0) return 0;
    if (sw == 0 || sh == 0) {                     /* empty source -> zeroed ImageBuffer::new(nw,nh) */
        for (uint32_t y = 0; y < dh; ++y) memset(dst + (size_t)y * dst_pitch, 0, (size_t)dw * channels);
        return 0;
    }
    if (sw == dw && sh == dh) {                   /* same dims -> copy */
        for (uint32_t y = 0; y < dh; ++y)
            memcpy(dst + (size_t)y * dst_pitch, src + (size_t)y * src_pitch, (size_t)dw * channels);
        return 0;
    }
    float* tmp = (float*)malloc(sizeof(float) * (size_t)sw * dh * channels);
    if (!tmp) return -2;
    vertical_sample_u8(src, sw, sh, src_pitch, channels, dh, kern, support, tmp);
    horizontal_sample_u8(tmp, sw, dh, channels, dw, kern, support, dst, dst_pitch);
    free(tmp);
    return 0;
}

/* Same for 16-bit samples (Luma16/LumaA16/Rgb16/Rgba16 variants); pitches in BYTES. */
ORC_API int orc_resize_u16(const uint16_t* src, uint32_t sw, uint32_t sh, size_t src_pitch, int channels,
                           uint16_t* dst, uint32_t dw, uint32_t dh, size_t dst_pitch, int filter) {
    kernel_fn kern; float support;
    if (channels < 1 || channels > 4 || pick_filter(filter, &kern, &support)) return -1;
    if ((src_pitch | dst_pitch) & 1) return -1;
    if (dw == 0 || dh == 0) return 0;
    if (sw == 0 || sh == 0) {
        for (uint32_t y = 0; y < dh; ++y)
            memset((uint8_t*)dst + (size_t)y * dst_pitch, 0, (size_t)dw * channels * 2);
        return 0;
    }
    if (sw == dw && sh == dh) {
        for (uint32_t y = 0; y < dh; ++y)
            memcpy((uint8_t*)dst + (size_t)y * dst_pitch, (const uint8_t*)src + (size_t)y * src_pitch,
                   (size_t)dw * channels * 2);
        return 0;
    }
    float* tmp = (float*)malloc(sizeof(float) * (size_t)sw * dh * channels);
    if (!tmp) return -2;
    vertical_sample_u16(src, sw, sh, src_pitch / 2, channels, dh, kern, support, tmp);
    horizontal_sample_u16(tmp, sw, dh, channels, dw, kern, support, dst, dst_pitch / 2);
    free(tmp);
    return 0;
}

/* Expose the f32 intermediate of the vertical pass (tests compare it with the GPU's exact mode). */
ORC_API int orc_vertical_f32_u8(const uint8_t* src, uint32_t sw, uint32_t sh, size_t src_pitch, int channels,
                                float* tmp, uint32_t dh, int filter) {
    kernel_fn kern; float support;
    if (channels < 1 || channels > 4 || pick_filter(filter, &kern, &support)) return -1;
    if (sw == 0 || sh == 0 || dh == 0) return -1;
    vertical_sample_u8(src, sw, sh, src_pitch, channels, dh, kern, support, tmp);
    return 0;
}

/* resize_image(img, w, h) end to end for a tight 8-bit raster: dims rule + Lanczos3 (transform.rs:88).
 * dst must hold tw*th*channels bytes where (tw,th) come from orc_target_dims. Returns the dims code. */
ORC_API int orc_resize_image_u8(const uint8_t* src, uint32_t sw, uint32_t sh, int channels, int has_w,
                                uint32_t w, int has_h, uint32_t h, uint8_t* dst) {
    uint32_t tw, th;
    int code = orc_target_dims(sw, sh, has_w, w, has_h, h, &tw, &th);
    if (code != 0) {
        memcpy(dst, src, (size_t)sw * sh * channels);
        return code;
    }
    int rc = orc_resize_u8(src, sw, sh, (size_t)sw * channels, channels, dst, tw, th,
                           (size_t)tw * channels, ORC_LANCZOS3);
    return rc < 0 ? rc : 0;
}
