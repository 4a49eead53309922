// plan.cpp -- dims rule + window/weight tables (see plan.hpp).  Must be compiled with
// -ffp-contract=off and without fast-math: every float operation below is meant to round
// exactly once, as the Rust reference does.
#include "plan.hpp"

#include "device_types.hpp"

#include <algorithm>
#include <cstring>
#include <cmath>
#include <limits>

namespace ikc {
namespace {

// Rust `as` casts from float saturate and send NaN to zero.
template <typename Int, typename Flt>
Int saturating_cast(Flt v) {
    if (std::isnan(v)) return Int(0);
    const Flt lo = static_cast<Flt>(std::numeric_limits<Int>::min());
    // max() of a 32/64-bit integer is not representable; 2^bits (or 2^(bits-1)) is.
    const Flt hi_excl = std::ldexp(Flt(1), std::numeric_limits<Int>::digits);
    if (v <= lo) return std::numeric_limits<Int>::min();
    if (v >= hi_excl) return std::numeric_limits<Int>::max();
    return static_cast<Int>(v);
}

// image 0.25.8 math/utils.rs: resize_dimensions(width, height, nwidth, nheight, fill = false)
void fit_within(uint32_t width, uint32_t height, uint32_t nwidth, uint32_t nheight, uint32_t* rw,
                uint32_t* rh) {
    const double wratio = double(nwidth) / double(width);
    const double hratio = double(nheight) / double(height);
    const double ratio = std::fmin(wratio, hratio);
    const uint64_t nw = std::max<uint64_t>(saturating_cast<uint64_t>(std::round(double(width) * ratio)), 1);
    const uint64_t nh = std::max<uint64_t>(saturating_cast<uint64_t>(std::round(double(height) * ratio)), 1);
    constexpr uint64_t kU32Max = std::numeric_limits<uint32_t>::max();
    if (nw > kU32Max) {
        const double r = double(kU32Max) / double(width);
        *rw = uint32_t(kU32Max);
        *rh = std::max<uint32_t>(saturating_cast<uint32_t>(std::round(double(height) * r)), 1);
    } else if (nh > kU32Max) {
        const double r = double(kU32Max) / double(height);
        *rw = std::max<uint32_t>(saturating_cast<uint32_t>(std::round(double(width) * r)), 1);
        *rh = uint32_t(kU32Max);
    } else {
        *rw = uint32_t(nw);
        *rh = uint32_t(nh);
    }
}

constexpr float kPi = 3.14159274101257324f;  // f32::consts::PI

struct FilterDef {
    float support;
    float (*eval)(float);
};

float eval_box(float) { return 1.0f; }
float eval_triangle(float x) {
    const float a = std::fabs(x);
    return a < 1.0f ? 1.0f - a : 0.0f;
}
float eval_sinc(float t) {
    const float a = t * kPi;
    return t == 0.0f ? 1.0f : sinf(a) / a;
}
float eval_lanczos3(float x) {
    constexpr float t = 3.0f;
    return std::fabs(x) < t ? eval_sinc(x) * eval_sinc(x / t) : 0.0f;
}
// bc_cubic_spline with the b, c parameters kept symbolic so the constant sub-expressions round
// exactly as the crate's f32 code does; CatmullRom is (b, c) = (0, 1/2).
float eval_bc_spline(float x, float b, float c) {
    const float a = std::fabs(x);
    const float a2 = a * a;   // powi(2)
    const float a3 = a * a2;  // powi(3)
    float k = 0.0f;
    if (a < 1.0f) {
        k = (12.0f - 9.0f * b - 6.0f * c) * a3 + (-18.0f + 12.0f * b + 6.0f * c) * a2 + (6.0f - 2.0f * b);
    } else if (a < 2.0f) {
        k = (-b - 6.0f * c) * a3 + (6.0f * b + 30.0f * c) * a2 + (-12.0f * b - 48.0f * c) * a +
            (8.0f * b + 24.0f * c);
    }
    return k / 6.0f;
}
float eval_catmullrom(float x) { return eval_bc_spline(x, 0.0f, 0.5f); }
float eval_gaussian(float x) {
    constexpr float r = 0.5f;
    const float norm = 1.0f / (std::sqrt(2.0f * kPi) * r);
    return norm * expf(-(x * x) / (2.0f * (r * r)));
}

bool lookup_filter(int filter, FilterDef* out) {
    switch (filter) {
        case kNearest: *out = {0.0f, eval_box}; return true;
        case kTriangle: *out = {1.0f, eval_triangle}; return true;
        case kCatmullRom: *out = {2.0f, eval_catmullrom}; return true;
        case kGaussian: *out = {3.0f, eval_gaussian}; return true;
        case kLanczos3: *out = {3.0f, eval_lanczos3}; return true;
        default: return false;
    }
}

}  // namespace

// IEEE binary32 -> binary16, round to nearest even, with denormals and overflow to infinity.
uint16_t f32_to_f16_rn(float f) {
    uint32_t x;
    std::memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    const uint32_t a = x & 0x7fffffffu;
    if (a >= 0x7f800000u) return uint16_t(sign | 0x7c00u | ((a > 0x7f800000u) ? 0x200u : 0u));  // inf / nan
    if (a >= 0x477ff000u) return uint16_t(sign | 0x7c00u);                                      // rounds to inf
    if (a < 0x33000001u) return uint16_t(sign);                                                 // below half the smallest denormal
    const int e = int(a >> 23) - 127;
    uint32_t mant = (a & 0x7fffffu) | 0x800000u;  // 24 bits
    int shift;                                   // bits dropped from the 24-bit mantissa
    uint32_t base;
    if (e >= -14) { shift = 13; base = uint32_t(e + 15) << 10; mant &= 0x7fffffu; }
    else { shift = 13 + (-14 - e); base = 0; }
    const uint32_t kept = mant >> shift;
    const uint32_t rem = mant & ((1u << shift) - 1u);
    const uint32_t half = 1u << (shift - 1);
    uint32_t h = base + kept;
    if (rem > half || (rem == half && (kept & 1u))) ++h;  // carries ripple into the exponent correctly
    return uint16_t(sign | h);
}

float f16_to_f32(uint16_t h) {
    const uint32_t sign = uint32_t(h & 0x8000u) << 16;
    const uint32_t e = (h >> 10) & 0x1fu, m = h & 0x3ffu;
    float v;
    if (e == 0) v = std::ldexp(float(m), -24);
    else if (e == 31) v = m ? std::numeric_limits<float>::quiet_NaN() : std::numeric_limits<float>::infinity();
    else v = std::ldexp(float(m | 0x400u), int(e) - 25);
    uint32_t x;
    std::memcpy(&x, &v, 4);
    x |= sign;
    std::memcpy(&v, &x, 4);
    return v;
}

float filter_support(int filter) {
    FilterDef f;
    return lookup_filter(filter, &f) ? f.support : -1.0f;
}

int target_dims(uint32_t ow, uint32_t oh, bool has_w, uint32_t w, bool has_h, uint32_t h, uint32_t* tw,
                uint32_t* th) {
    if (!has_w && !has_h) {  // transform.rs:67-69
        *tw = ow;
        *th = oh;
        return 1;
    }
    // transform.rs:74-82: the missing side follows the other one's scale, in f32.
    uint32_t want_w = w, want_h = h;
    if (!has_w) want_w = saturating_cast<uint32_t>(std::round(float(ow) * (float(h) / float(oh))));
    if (!has_h) want_h = saturating_cast<uint32_t>(std::round(float(oh) * (float(w) / float(ow))));
    want_w = std::max<uint32_t>(want_w, 1);  // transform.rs:86-87
    want_h = std::max<uint32_t>(want_h, 1);
    if (want_w == ow && want_h == oh) {  // DynamicImage::resize: same size -> clone
        *tw = ow;
        *th = oh;
        return 2;
    }
    fit_within(ow, oh, want_w, want_h, tw, th);
    if (*tw == ow && *th == oh) return 3;  // imageops::resize: same size -> copy
    return 0;
}

std::shared_ptr<const PassPlan> build_pass(int filter, uint32_t n_in, uint32_t n_out, bool vertical_forms) {
    FilterDef f;
    if (!lookup_filter(filter, &f) || n_in == 0 || n_out == 0) return nullptr;
    auto plan = std::make_shared<PassPlan>();
    PassPlan& p = *plan;
    p.filter = filter;
    p.n_in = n_in;
    p.n_out = n_out;
    p.left.resize(n_out);
    p.count.resize(n_out);
    p.right.resize(n_out);

    const float ratio = float(n_in) / float(n_out);
    const float sratio = ratio < 1.0f ? 1.0f : ratio;
    const float reach = f.support * sratio;

    // Pass 1: windows + raw weights (ragged), exactly the crate's loop.
    std::vector<std::vector<float>> ragged(n_out);
    for (uint32_t o = 0; o < n_out; ++o) {
        float centre = (float(o) + 0.5f) * ratio;
        int64_t lo = saturating_cast<int64_t>(std::floor(centre - reach));
        lo = std::clamp<int64_t>(lo, 0, int64_t(n_in) - 1);
        int64_t hi = saturating_cast<int64_t>(std::ceil(centre + reach));
        hi = std::clamp<int64_t>(hi, lo + 1, int64_t(n_in));
        centre = centre - 0.5f;
        std::vector<float>& ws = ragged[o];
        ws.reserve(size_t(hi - lo));
        float total = 0.0f;
        for (int64_t i = lo; i < hi; ++i) {
            const float wi = f.eval((float(uint32_t(i)) - centre) / sratio);
            ws.push_back(wi);
            total += wi;
        }
        for (float& wi : ws) wi /= total;
        p.left[o] = int32_t(lo);
        p.count[o] = int32_t(hi - lo);
        p.right[o] = int32_t(hi);
        p.max_count = std::max<uint32_t>(p.max_count, uint32_t(hi - lo));
    }
    p.stride = p.max_count;
    p.w.assign(size_t(n_out) * p.stride, 0.0f);
    for (uint32_t o = 0; o < n_out; ++o)
        std::copy(ragged[o].begin(), ragged[o].end(), p.w.begin() + size_t(o) * p.stride);

    // Ring form: how many windows cover each source index (windows are monotone, so the
    // covering outputs are a contiguous run and get distinct residues mod ring_k).
    std::vector<int32_t> cover(size_t(n_in) + 1, 0);
    for (uint32_t o = 0; o < n_out; ++o) {
        cover[p.left[o]] += 1;
        cover[p.right[o]] -= 1;
    }
    int run = 0, k = 0;
    for (uint32_t y = 0; y < n_in; ++y) {
        run += cover[y];
        k = std::max(k, run);
    }
    p.ring_k = k;
    // Uniform stretch: the longest run of outputs that (a) end the same number of source indices after
    // their predecessor, (b) have full windows of ring_k * step taps that start where the window ring_k
    // outputs earlier ended, and (c) share one set of weights bit for bit -- the interior of an
    // integer-ratio downscale.  The fused kernel runs such a stretch from tap weights held in registers.
    if (k >= 1 && n_out >= 2) {
        auto uniform_with = [&](uint32_t o, uint32_t ref, int step) {
            if (p.right[o] - p.right[o - 1] != step || p.count[o] != k * step) return false;
            if (o >= uint32_t(k) && p.left[o] != p.right[o - uint32_t(k)]) return false;
            return std::equal(ragged[o].begin(), ragged[o].end(), ragged[ref].begin(),
                              [](float x, float y) { return std::memcmp(&x, &y, sizeof(float)) == 0; });
        };
        uint32_t best_lo = 0, best_len = 0;
        for (uint32_t lo = 1; lo < n_out;) {
            const int step = p.right[lo] - p.right[lo - 1];
            uint32_t hi = lo;
            if (step >= 1 && p.count[lo] == k * step)
                while (hi < n_out && uniform_with(hi, lo, step)) ++hi;
            if (hi - lo > best_len) { best_len = hi - lo; best_lo = lo; }
            lo = std::max(hi, lo + 1);
        }
        if (best_len >= 4u * uint32_t(k)) {
            p.uni_step = p.right[best_lo] - p.right[best_lo - 1];
            p.uni_lo = int(best_lo);
            p.uni_hi = int(best_lo + best_len);
        }
    }
    if (p.uni_step == 2 && p.count[size_t(p.uni_lo)] == 12 && p.left[size_t(p.uni_lo)] == 2 * p.uni_lo - 5) {
        bool ok = true;
        for (uint32_t o = 0; o < n_out && ok; ++o)
            ok = int64_t(p.left[o]) >= 2 * int64_t(o) - 5 && int64_t(p.right[o]) <= 2 * int64_t(o) + 7;
        p.h2_12 = ok;
    }
    // Frame form for exact 2x upscales.
    if (n_out == 2 * n_in) {
        int off = 0x7fffffff, end = -0x7fffffff;
        for (uint32_t o = 0; o < n_out; ++o) {
            off = std::min(off, p.left[o] - int(o >> 1));
            end = std::max(end, p.right[o] - int(o >> 1));
        }
        const int taps = end - off;
        if (taps >= 1 && taps <= 8) {
            p.up2_off = off;
            p.up2_taps = taps;
            p.up2_pairs.assign(size_t(n_in) * taps * 2, 0.0f);
            for (uint32_t o = 0; o < n_out; ++o) {
                const int base = int(o >> 1) + off;
                for (int32_t i = 0; i < p.count[o]; ++i)
                    p.up2_pairs[(size_t(o >> 1) * taps + size_t(p.left[o] + i - base)) * 2 + (o & 1)] = ragged[o][i];
            }
            // longest run of source indices with identical pairs
            const size_t row = size_t(taps) * 2;
            uint32_t best_lo = 0, best_len = 0;
            for (uint32_t lo = 0; lo < n_in;) {
                uint32_t hi = lo + 1;
                while (hi < n_in && std::memcmp(&p.up2_pairs[hi * row], &p.up2_pairs[lo * row], row * sizeof(float)) == 0) ++hi;
                if (hi - lo > best_len) { best_len = hi - lo; best_lo = lo; }
                lo = hi;
            }
            p.up2_uni_lo = int(best_lo);
            p.up2_uni_hi = int(best_lo + best_len);
            // Row-band integer form of the pass used vertically (banded8u.cu: the tensor-core 2x upscale kernel): the same
            // digits and tile layout as a downscale's Band8T; a band of 128 outputs spans 64 + taps - 1 source indices.
            float wmax = 0.0f;
            for (auto& ws : ragged)
                for (float wi : ws) wmax = std::max(wmax, std::fabs(wi));
            const uint32_t n_bands = (n_out + kBand8TRows - 1) / kBand8TRows;
            std::vector<int32_t> k_lo(n_bands, 0);
            int band_chunks = 0;
            for (uint32_t r = 0; r < n_bands; ++r) {
                const uint32_t o0 = r * kBand8TRows, o1 = std::min(n_out, o0 + kBand8TRows);
                k_lo[r] = p.left[o0];
                band_chunks = std::max(band_chunks, (p.right[o1 - 1] - k_lo[r] + kBand8Chunk - 1) / kBand8Chunk);
            }
            if (vertical_forms && wmax > 0.0f && band_chunks >= 1 && band_chunks <= kBand8TMaxChunks) {
                int shift = int(std::floor(std::log2(120.0 * kBand8Base / double(wmax))));
                shift = std::min(shift, 21);
                const double scale = std::ldexp(1.0, shift);
                constexpr size_t kTTile = size_t(kBand8TRows) * kBand8Chunk;
                p.band8.shift = shift;   // (band8.limbs stays 0: there is no chunk-window form of an upscale)
                p.band8t.chunks = band_chunks;
                p.band8t.k_lo = k_lo;
                p.band8t.tiles.assign(size_t(n_bands) * band_chunks * 2 * kTTile, 0);
                std::vector<int64_t> W;
                for (uint32_t o = 0; o < n_out; ++o) {
                    const auto& ws = ragged[o];
                    W.resize(ws.size());
                    int64_t sum = 0;
                    size_t big = 0;
                    for (size_t i = 0; i < ws.size(); ++i) {
                        W[i] = int64_t(std::llround(double(ws[i]) * scale));
                        sum += W[i];
                        if (std::fabs(ws[i]) > std::fabs(ws[big])) big = i;
                    }
                    W[big] += (int64_t(1) << shift) - sum;
                    const uint32_t r = o / kBand8TRows;
                    const int m = int(o % kBand8TRows);
                    for (size_t i = 0; i < ws.size(); ++i) {
                        const int32_t y = p.left[o] + int32_t(i);
                        const int ct = (y - k_lo[r]) / kBand8Chunk, kt = (y - k_lo[r]) % kBand8Chunk;
                        const int64_t lo = ((W[i] + kBand8Base / 2) & (kBand8Base - 1)) - kBand8Base / 2, hi = (W[i] - lo) / kBand8Base;
                        const size_t at = size_t(kt / 16) * (size_t(kBand8TRows) * 16) + size_t(m / 8) * 128 + size_t(m % 8) * 16 + size_t(kt % 16);
                        const size_t tt = (size_t(r) * band_chunks + size_t(ct)) * 2 * kTTile;
                        p.band8t.tiles[tt + at] = int8_t(hi);
                        p.band8t.tiles[tt + kTTile + at] = int8_t(lo);
                    }
                }
            }
        }
    }
    // Band form for the tensor-core vertical pass (downscales and 1:1 only).
    if (vertical_forms && n_in >= n_out && n_in >= uint32_t(kBandChunk)) {
        const uint32_t n_chunks = (n_in + kBandChunk - 1) / kBandChunk;
        const uint32_t n_groups = (n_out + kBandGroup - 1) / kBandGroup;
        std::vector<int32_t> gbase(n_chunks + 1, 0), ghi(n_chunks, 0);
        uint32_t o_lo = 0;  // first output whose window reaches index 16k or beyond
        int n_max = 0;
        for (uint32_t c = 0; c < n_chunks; ++c) {
            const int32_t y0 = int32_t(c) * kBandChunk, y1 = y0 + kBandChunk;
            while (o_lo < n_out && p.right[o_lo] <= y0) ++o_lo;
            uint32_t o_hi = o_lo;  // one past the last output whose window starts before y1
            while (o_hi < n_out && p.left[o_hi] < y1) ++o_hi;
            gbase[c] = int32_t(std::min(o_lo, n_out - 1) / kBandGroup);
            ghi[c] = o_hi > o_lo ? int32_t((o_hi - 1) / kBandGroup) : gbase[c];
            n_max = std::max(n_max, (ghi[c] - gbase[c] + 1) * kBandGroup);
        }
        gbase[n_chunks] = int32_t(n_groups);
        if (n_max <= kBandMaxN) {
            n_max = std::max(n_max, 2 * kBandGroup);
            p.band_n = n_max;
            p.band_gbase = gbase;
            const size_t tile = size_t(n_max) * kBandChunk;  // f16 elements of one operand tile
            p.band_tiles.assign(size_t(n_chunks) * 2 * tile, 0);
            const size_t kgroup_stride = size_t(n_max) * 8;  // elements between the two halves of the 16 indices
            for (uint32_t c = 0; c < n_chunks; ++c) {
                uint16_t* hi = p.band_tiles.data() + size_t(c) * 2 * tile;
                uint16_t* lo = hi + tile;
                for (int n = 0; n < n_max; ++n) {
                    const int64_t o = int64_t(gbase[c]) * kBandGroup + n;
                    if (o >= int64_t(n_out)) break;
                    for (int kk = 0; kk < kBandChunk; ++kk) {
                        const int32_t y = int32_t(c) * kBandChunk + kk;
                        if (y < p.left[o] || y >= p.right[o]) continue;
                        const float ws = ragged[size_t(o)][size_t(y - p.left[o])] * kBandScaleW;  // exact: a power of two
                        const uint16_t h = f32_to_f16_rn(ws);
                        const uint16_t l = f32_to_f16_rn(ws - f16_to_f32(h));                     // the difference is exact in f32
                        const size_t at = size_t(kk / 8) * kgroup_stride + size_t(n / 8) * 64 + size_t(n % 8) * 8 + size_t(kk % 8);
                        hi[at] = h;
                        lo[at] = l;
                    }
                }
            }
        }
    }
    // 8-bit band form: integer weights in base-256 digits.
    if (vertical_forms && n_in >= n_out && n_in >= uint32_t(kBand8Chunk) && p.max_count <= 240) {
        const int limbs = kBand8DefaultLimbs;
        const uint32_t n_chunks = (n_in + kBand8Chunk - 1) / kBand8Chunk;
        const uint32_t n_groups = (n_out + kBand8Group - 1) / kBand8Group;
        std::vector<int32_t> gbase(n_chunks + 1, 0);
        bool fits = true;
        uint32_t o_lo = 0;
        for (uint32_t c = 0; c < n_chunks && fits; ++c) {
            const int32_t y0 = int32_t(c) * kBand8Chunk, y1 = y0 + kBand8Chunk;
            while (o_lo < n_out && p.right[o_lo] <= y0) ++o_lo;
            uint32_t o_hi = o_lo;
            while (o_hi < n_out && p.left[o_hi] < y1) ++o_hi;
            gbase[c] = int32_t(std::min(o_lo, n_out - 1) / kBand8Group);
            if (o_hi > o_lo && int32_t(o_hi - 1) - gbase[c] * kBand8Group >= kBand8Window) fits = false;
        }
        gbase[n_chunks] = int32_t(n_groups);
        float wmax = 0.0f;
        for (auto& ws : ragged)
            for (float wi : ws) wmax = std::max(wmax, std::fabs(wi));
        if (fits && wmax > 0.0f) {
            // digits: the top one in [-127, 127], the others in [-128, 127]  =>  |W| <= 127 * 256^(limbs-1) + 127 * (...)
            const double top = 120.0 * std::pow(double(kBand8Base), limbs - 1);  // (headroom for the sum correction)
            int shift = int(std::floor(std::log2(top / double(wmax))));
            shift = std::min(shift, 21);  // 255 * 2^21 * sum|w| stays inside the s32 accumulators
            p.band8.limbs = limbs;
            p.band8.shift = shift;
            p.band8.gbase = gbase;
            const size_t tile = size_t(limbs) * kBand8Window * kBand8Chunk;
            p.band8.tiles.assign(size_t(n_chunks) * tile, 0);
            const double scale = std::ldexp(1.0, shift);
            std::vector<int64_t> W;
            // row-band form: applicable when no band spans more than kBand8TMaxChunks chunks.  The band height (<= 128
            // outputs, the MMA's M) is the one that costs least: bands x (chunks + the epilogue's share).
            int band_rows = kBand8TRows;
            uint32_t n_bands = (n_out + kBand8TRows - 1) / kBand8TRows;
            int band_chunks = 0;
            std::vector<int32_t> k_lo;
            if (limbs == 2) {
                uint64_t best = ~uint64_t(0);
                for (int rows = kBand8TRows; rows >= 96; rows -= 8) {
                    const uint32_t nb = (n_out + uint32_t(rows) - 1) / uint32_t(rows);
                    int chunks = 0;
                    for (uint32_t r = 0; r < nb; ++r) {
                        const uint32_t o0 = r * uint32_t(rows), o1 = std::min(n_out, o0 + uint32_t(rows));
                        chunks = std::max(chunks, (p.right[o1 - 1] - p.left[o0] + kBand8Chunk - 1) / kBand8Chunk);
                    }
                    // (per band and block of columns: 2 MMAs per chunk on the tensor pipe, and an epilogue worth about 15 chunks)
                    const uint64_t cost = uint64_t(nb) * uint64_t(chunks + 15);
                    if (chunks <= kBand8TMaxChunks && cost < best) {
                        best = cost;
                        band_rows = rows;
                        n_bands = nb;
                        band_chunks = chunks;
                    }
                }
                k_lo.assign(n_bands, 0);
                for (uint32_t r = 0; r < n_bands; ++r) k_lo[r] = p.left[r * uint32_t(band_rows)];  // chunks start at the band's first source index
            }
            constexpr size_t kTTile = size_t(kBand8TRows) * kBand8Chunk;
            if (band_chunks) {
                p.band8t.chunks = band_chunks;
                p.band8t.rows = band_rows;
                p.band8t.k_lo = k_lo;
                p.band8t.tiles.assign(size_t(n_bands) * band_chunks * 2 * kTTile, 0);
            }
            for (uint32_t o = 0; o < n_out; ++o) {
                const auto& ws = ragged[o];
                W.resize(ws.size());
                int64_t sum = 0;
                size_t big = 0;
                for (size_t i = 0; i < ws.size(); ++i) {
                    W[i] = int64_t(std::llround(double(ws[i]) * scale));
                    sum += W[i];
                    if (std::fabs(ws[i]) > std::fabs(ws[big])) big = i;
                }
                W[big] += (int64_t(1) << shift) - sum;  // the weights sum to one: a flat area stays exactly flat
                for (size_t i = 0; i < ws.size(); ++i) {
                    const int32_t y = p.left[o] + int32_t(i);
                    const uint32_t c = uint32_t(y) / kBand8Chunk;
                    const int kk = y % kBand8Chunk;
                    const int pos = int(o) - gbase[c] * kBand8Group;  // position in the chunk's window
                    int8_t* t = p.band8.tiles.data() + size_t(c) * tile;
                    int64_t v = W[i];
                    for (int d = limbs - 1; d >= 0; --d) {   // least significant digit first
                        int64_t digit;
                        if (d > 0) { digit = ((v + kBand8Base / 2) & (kBand8Base - 1)) - kBand8Base / 2; v = (v - digit) / kBand8Base; }
                        else digit = v;                       // the top digit takes what is left (|digit| <= 127 by the choice of shift)
                        const int n = pos * limbs + d;         // row of the tile: an output's digits are adjacent, most significant first
                        const size_t at = size_t(kk / 16) * (size_t(limbs) * kBand8Window * 16) + size_t(n / 8) * 128 + size_t(n % 8) * 16 + size_t(kk % 16);
                        t[at] = int8_t(digit);
                        if (band_chunks) {  // the same digit in the row-band tile (band, chunk, digit): row m, index kt
                            const uint32_t r = o / uint32_t(band_rows);
                            const int m = int(o % uint32_t(band_rows));
                            const int ct = (y - k_lo[r]) / kBand8Chunk, kt = (y - k_lo[r]) % kBand8Chunk;
                            const size_t tt = ((size_t(r) * band_chunks + size_t(ct)) * 2 + size_t(d)) * kTTile;
                            p.band8t.tiles[tt + size_t(kt / 16) * (size_t(kBand8TRows) * 16) + size_t(m / 8) * 128 + size_t(m % 8) * 16 + size_t(kt % 16)] = int8_t(digit);
                        }
                    }
                }
            }
        }
    }
    if (k >= 1 && k <= 8) {  // only the fused kernels use it; they handle ring_k <= 8
        p.ring_stride = (k + 1) & ~1;  // even: every ring row is a whole number of 16-byte loads
        p.ring_v.assign(size_t(n_in) * p.ring_stride * 2, 0.0f);
        p.ring_h.assign(size_t(n_in) * p.ring_stride * 2, 0.0f);
        for (uint32_t o = 0; o < n_out; ++o) {
            const int j = int(o % uint32_t(k));
            for (int32_t i = 0; i < p.count[o]; ++i) {
                const size_t at = (size_t(p.left[o] + i) * p.ring_stride + j) * 2;
                p.ring_v[at] = p.ring_v[at + 1] = ragged[o][i] * kRingScaleV;  // exact: powers of two
                p.ring_h[at] = p.ring_h[at + 1] = ragged[o][i] * kRingScaleH;
            }
        }
    }
    return plan;
}

}  // namespace ikc
