// ORACLE — TEST INFRASTRUCTURE ONLY (see fmath.hpp).
//
// Stage 4: per-tile-row painter. Restates forma/src/cpu/painter/mod.rs,
// forma/src/cpu/painter/styling.rs, forma/src/cpu/painter/layer_workbench/
// (mod.rs + passes/*.rs) and LinearLayout::write
// (forma/src/cpu/buffer/layout/mod.rs:265-295).
//
// The reference works on f32x8 vectors = 8 vertically adjacent pixels of one
// tile column. Here every lane is computed with scalar code in the same
// operation order; the only cross-lane semantics (the "all 8 coverages are
// zero -> skip" test, cpu/painter/mod.rs:317-319) are kept explicitly.
#pragma once

#include <immintrin.h>

#include <map>
#include <vector>

#include "raster.hpp"

namespace fo {

constexpr int kTile = 16;

// cpu/painter/mod.rs:169-215
struct Cover {
    int8_t c[kTile] = {0};
    bool is_empty(FillRule fr) const {
        for (int i = 0; i < kTile; ++i) {
            if (fr == kNonZero) {
                if (c[i] != 0) return false;
            } else {
                int8_t ab = (int8_t)(c[i] < 0 ? -c[i] : c[i]);  // wrapping abs
                if ((ab & 31) != 0) return false;
            }
        }
        return true;
    }
    bool is_full(FillRule fr) const {
        for (int i = 0; i < kTile; ++i) {
            int8_t ab = (int8_t)(c[i] < 0 ? -c[i] : c[i]);
            if (fr == kNonZero) {
                if (ab != 16) return false;
            } else {
                if ((ab & 31) != 16) return false;
            }
        }
        return true;
    }
};

struct CoverCarry {
    Cover cover;
    uint32_t layer_id;
};

inline float color_channel(const Color& c, Channel ch) {
    switch (ch) {
        case kRed: return c.r;
        case kGreen: return c.g;
        case kBlue: return c.b;
        case kAlpha: return c.a;
        case kZero: return 0.0f;
        default: return 1.0f;
    }
}

// cpu/painter/mod.rs:98-112 / :116-130 (same arithmetic for x8 and x4).
inline float linear_to_srgb(float l) {
    float s = std::sqrt(l);
    float s3 = l * s;
    float m = l * 12.92f;
    float n = std::fmaf(0.20101772f, s3, std::fmaf(-0.51280147f, l, std::fmaf(1.344401f, s, -0.030656587f)));
    return l <= 0.0031308f ? m : n;
}

// cpu/painter/mod.rs:135-154
inline uint8_t to_byte(float v) {
    float scaled = rclamp(v * 255.0f, 0.0f, 255.0f);
    float val = scaled + u2f(0x4B000000u);
    return (uint8_t)(f2u(val) & 0xFF);
}

// cpu/painter/mod.rs:156-162
inline void to_srgb_bytes(const float color[4], uint8_t out[4]) {
    out[0] = to_byte(linear_to_srgb(color[0]));
    out[1] = to_byte(linear_to_srgb(color[1]));
    out[2] = to_byte(linear_to_srgb(color[2]));
    out[3] = to_byte(color[3]);
}

// ---------------------------------------------------------------------------
// Scalar blend (cpu/painter/styling.rs:195-339) — used when folding solid tiles.
// ---------------------------------------------------------------------------
namespace scalar_blend {
inline float multiply(float d, float s) { return d * s; }
inline float screen(float d, float s) { return d + s - (d * s); }
inline float hard_light(float d, float s) {
    return s <= 0.5f ? multiply(d, 2.0f * s) : screen(d, 2.0f * s - 1.0f);
}
inline float lum(const Color& c) { return std::fmaf(c.r, 0.3f, std::fmaf(c.g, 0.59f, c.b * 0.11f)); }
inline float cmax(const Color& c) { return rmax(c.r, rmax(c.g, c.b)); }
inline float cmin(const Color& c) { return rmin(c.r, rmin(c.g, c.b)); }
inline float chan(const Color& c, int ch) { return ch == 0 ? c.r : ch == 1 ? c.g : c.b; }
inline float clip_color(int ch, const Color& color) {
    float l = lum(color), n = cmin(color), x = cmax(color);
    float c = chan(color, ch);
    if (n < 0.0f) {
        float l_n_recip_l = recip(l - n) * l;
        c = std::fmaf(l_n_recip_l, c - l, l);
    }
    if (x > 1.0f) {
        float l_1 = l - 1.0f;
        float x_l_recip = recip(x - l);
        c = std::fmaf(x_l_recip, std::fmaf(l, l_1 - c, c), l);
    }
    return c;
}
inline float set_lum(int ch, Color color, float l) {
    float d = l - lum(color);
    color.r += d;
    color.g += d;
    color.b += d;
    return clip_color(ch, color);
}
inline float sat(const Color& c) { return cmax(c) - cmin(c); }
inline Color set_sat(Color color, float s) {
    float* v[3] = {&color.r, &color.g, &color.b};
    bool a = color.r < color.g, b = color.r < color.b, c = color.g < color.b;
    int idx[3];
    // Color::sorted, cpu/painter/styling.rs:33-44 -> [min, mid, max]
    if (a && b && c) { idx[0] = 0; idx[1] = 1; idx[2] = 2; }
    else if (a && b && !c) { idx[0] = 0; idx[1] = 2; idx[2] = 1; }
    else if (a && !b) { idx[0] = 2; idx[1] = 0; idx[2] = 1; }
    else if (!a && b && c) { idx[0] = 1; idx[1] = 0; idx[2] = 2; }
    else if (!a && !c) { idx[0] = 2; idx[1] = 1; idx[2] = 0; }
    else { idx[0] = 1; idx[1] = 2; idx[2] = 0; }
    float *c_min = v[idx[0]], *c_mid = v[idx[1]], *c_max = v[idx[2]];
    if (*c_max > *c_min) {
        *c_mid = std::fmaf(s, *c_mid, -s * *c_min) / (*c_max - *c_min);
        *c_max = s;
    } else {
        *c_mid = 0.0f;
        *c_max = 0.0f;
    }
    *c_min = 0.0f;
    return color;
}
inline float soft_d(float d) { return d <= 0.25f ? ((16.0f * d - 12.0f) * d + 4.0f) * d : std::sqrt(d); }

inline float blend_channel(BlendMode mode, int ch, const Color& dst, const Color& src) {
    float d = chan(dst, ch), s = chan(src, ch);
    switch (mode) {
        case kOver: return s;
        case kMultiply: return multiply(d, s);
        case kScreen: return screen(d, s);
        case kOverlay: return hard_light(s, d);
        case kDarken: return rmin(d, s);
        case kLighten: return rmax(d, s);
        case kColorDodge: return d == 0.0f ? 0.0f : (s == 1.0f ? 1.0f : rmin(1.0f, d / (1.0f - s)));
        case kColorBurn: return d == 1.0f ? 1.0f : (s == 0.0f ? 0.0f : 1.0f - rmin(1.0f, (1.0f - d) / s));
        case kHardLight: return hard_light(d, s);
        case kSoftLight:
            return s <= 0.5f ? d - (1.0f - 2.0f * s) * d * (1.0f - d) : d + (2.0f * s - 1.0f) * (soft_d(d) - d);
        case kDifference: return std::fabs(d - s);
        case kExclusion: return d + s - 2.0f * d * s;
        case kColorMode: return set_lum(ch, src, lum(dst));
        case kLuminosity: return set_lum(ch, dst, lum(src));
        case kHue: return set_lum(ch, set_sat(src, sat(dst)), lum(dst));
        case kSaturation: return set_lum(ch, set_sat(dst, sat(src)), lum(dst));
    }
    return s;
}

// BlendMode::blend, cpu/painter/styling.rs:315-339
inline Color blend(BlendMode mode, const Color& dst, const Color& src) {
    float inv_dst_a = 1.0f - dst.a;
    float inv_dst_a_src_a = inv_dst_a * src.a;
    float inv_src_a = 1.0f - src.a;
    float dst_a_src_a = dst.a * src.a;
    float cr = std::fmaf(src.r, inv_dst_a_src_a, blend_channel(mode, 0, dst, src) * dst_a_src_a);
    float cg = std::fmaf(src.g, inv_dst_a_src_a, blend_channel(mode, 1, dst, src) * dst_a_src_a);
    float cb = std::fmaf(src.b, inv_dst_a_src_a, blend_channel(mode, 2, dst, src) * dst_a_src_a);
    Color o;
    o.r = std::fmaf(dst.r, inv_src_a, cr);
    o.g = std::fmaf(dst.g, inv_src_a, cg);
    o.b = std::fmaf(dst.b, inv_src_a, cb);
    o.a = std::fmaf(dst.a, inv_src_a, src.a);
    return o;
}
}  // namespace scalar_blend

// ---------------------------------------------------------------------------
// Per-lane form of the vector macro blend_function! (styling.rs:342-594).
// ---------------------------------------------------------------------------
namespace lane_blend {
// The x86 AVX2 shim of the reference evaluates f32x8::recip with the ~12-bit
// _mm256_rcp_ps (utils/simd/avx.rs:465-467); the portable shim uses 1/x
// (auto.rs:727-730). recip is only used by clip_color!. Mode 1 exists solely to
// pin the oracle against goldens that were rendered through the AVX2 shim.
inline int& recip_mode() {
    static int mode = 0;
    return mode;
}
// Mode 2: Arm FPRecipEstimate (vrecpeq_f32, utils/simd/aarch64.rs:520-530),
// restated from the Arm ARM pseudocode for normal inputs.
inline float arm_recpe(float v) {
    uint32_t u = f2u(v);
    uint32_t sign = u & 0x80000000u;
    int exp = (int)((u >> 23) & 0xFF);
    uint32_t frac = u & 0x7FFFFFu;
    if (exp == 0 || exp == 255) return recip(v);  // zero/denormal/inf/NaN: not needed for pinning
    uint32_t scaled = 256u | (frac >> 15);        // '1' : fraction<51:44>
    int result_exp = 253 - exp;
    uint32_t a = scaled * 2 + 1;
    uint32_t b = (1u << 19) / a;
    uint32_t est = (b + 1) / 2;                   // 256..511
    uint32_t fraction = (est & 0xFF) << 15;       // 8 bits at the top of 23
    if (result_exp == 0) {
        fraction = (1u << 22) | (fraction >> 1);
    } else if (result_exp == -1) {
        fraction = (1u << 21) | (fraction >> 2);
        result_exp = 0;
    } else if (result_exp < -1) {
        return u2f(sign);
    }
    return u2f(sign | ((uint32_t)result_exp << 23) | fraction);
}
inline float vrecip(float v) {
    if (recip_mode() == 1) return _mm_cvtss_f32(_mm_rcp_ss(_mm_set_ss(v)));
    if (recip_mode() == 2) return arm_recpe(v);
    return recip(v);
}
inline float lum(float r, float g, float b) { return std::fmaf(r, 0.3f, std::fmaf(g, 0.59f, b * 0.11f)); }
inline float sat(float r, float g, float b) { return rmax(r, rmax(g, b)) - rmin(r, rmin(g, b)); }
inline void clip_color(float r, float g, float b, float out[3]) {
    float l = lum(r, g, b);
    float n = rmin(r, rmin(g, b));
    float x = rmax(r, rmax(g, b));
    float l_1 = l - 1.0f;
    float x_l_recip = vrecip(x - l);
    float l_n_recip_l = vrecip(l - n) * l;
    float in[3] = {r, g, b};
    for (int i = 0; i < 3; ++i) {
        float c = in[i];
        float hi = std::fmaf(x_l_recip, std::fmaf(l, l_1 - c, c), l);
        float lo = n < 0.0f ? std::fmaf(l_n_recip_l, c - l, l) : c;
        out[i] = 1.0f < x ? hi : lo;
    }
}
inline void set_lum(float r, float g, float b, float l, float out[3]) {
    float d = l - lum(r, g, b);
    r += d;
    g += d;
    b += d;
    clip_color(r, g, b, out);
}
inline void set_sat(float sat_dst, float sr, float sg, float sb, float out[3]) {
    float src_min = rmin(sr, rmin(sg, sb));
    float src_max = rmax(sr, rmax(sg, sb));
    float src_mid = sr + sg + sb - src_min - src_max;
    bool min_lt_max = src_min < src_max;
    float sat_mid = min_lt_max ? (std::fmaf(sat_dst, -src_min, sat_dst * src_mid) / (src_max - src_min)) : 0.0f;
    float sat_max = min_lt_max ? sat_dst : 0.0f;
    float in[3] = {sr, sg, sb};
    for (int i = 0; i < 3; ++i) {
        float inner = in[i] == src_min ? 0.0f : sat_mid;
        out[i] = in[i] == src_max ? sat_max : inner;
    }
}
inline float hard(float d, float s, float sel) {
    return sel <= 0.5f ? d * s * 2.0f : 2.0f * (d + s - std::fmaf(d, s, 0.5f));
}
inline void blend(BlendMode mode, float dr, float dg, float db, float sr, float sg, float sb, float out[3]) {
    float d[3] = {dr, dg, db}, s[3] = {sr, sg, sb};
    switch (mode) {
        case kOver:
            for (int i = 0; i < 3; ++i) out[i] = s[i];
            return;
        case kMultiply:
            for (int i = 0; i < 3; ++i) out[i] = d[i] * s[i];
            return;
        case kScreen:
            for (int i = 0; i < 3; ++i) out[i] = std::fmaf(d[i], -s[i], d[i]) + s[i];
            return;
        case kOverlay:
            for (int i = 0; i < 3; ++i) out[i] = hard(d[i], s[i], d[i]);
            return;
        case kDarken:
            for (int i = 0; i < 3; ++i) out[i] = rmin(d[i], s[i]);
            return;
        case kLighten:
            for (int i = 0; i < 3; ++i) out[i] = rmax(d[i], s[i]);
            return;
        case kColorDodge:
            for (int i = 0; i < 3; ++i) out[i] = s[i] == 1.0f ? 1.0f : rmin(1.0f, d[i] / (1.0f - s[i]));
            return;
        case kColorBurn:
            for (int i = 0; i < 3; ++i) out[i] = s[i] == 0.0f ? 0.0f : 1.0f - rmin(1.0f, (1.0f - d[i]) / s[i]);
            return;
        case kHardLight:
            for (int i = 0; i < 3; ++i) out[i] = hard(d[i], s[i], s[i]);
            return;
        case kSoftLight:
            for (int i = 0; i < 3; ++i) {
                float dd = d[i] <= 0.25f ? std::fmaf(std::fmaf(16.0f, d[i], -12.0f), d[i], 4.0f) * d[i] : std::sqrt(d[i]);
                float k = std::fmaf(2.0f, s[i], -1.0f);
                out[i] = s[i] <= 0.5f ? std::fmaf(d[i] * (1.0f - d[i]), k, d[i]) : std::fmaf(dd - d[i], k, d[i]);
            }
            return;
        case kDifference:
            for (int i = 0; i < 3; ++i) out[i] = std::fabs(d[i] - s[i]);
            return;
        case kExclusion:
            for (int i = 0; i < 3; ++i) out[i] = std::fmaf(-2.0f * d[i], s[i], d[i]) + s[i];
            return;
        case kHue: {
            float t[3];
            set_sat(sat(dr, dg, db), sr, sg, sb, t);
            set_lum(t[0], t[1], t[2], lum(dr, dg, db), out);
            return;
        }
        case kSaturation: {
            float t[3];
            set_sat(sat(sr, sg, sb), dr, dg, db, t);
            set_lum(t[0], t[1], t[2], lum(dr, dg, db), out);
            return;
        }
        case kColorMode:
            set_lum(sr, sg, sb, lum(dr, dg, db), out);
            return;
        case kLuminosity:
            set_lum(dr, dg, db, lum(sr, sg, sb), out);
            return;
    }
}
}  // namespace lane_blend

// Gradient::color_at for one lane (cpu/painter/styling.rs:58-144). `x` is the
// pixel column, `y_base` the y of lane 0 of the f32x8, `lane` in 0..8.
inline void gradient_color_at(const Gradient& g, float x, float y_base, int lane, float out[4]) {
    float dx = g.end.x - g.start.x;
    float dy = g.end.y - g.start.y;
    float dot = dx * dx + dy * dy;
    float dot_recip = recip(dot);
    float t;
    if (g.type == kLinear) {
        float tx = (x - g.start.x) * dx * dot_recip;
        float ty = y_base - g.start.y;
        t = std::fmaf(((float)lane + ty) * dy, dot_recip, tx);
    } else {
        float px = x - g.start.x;
        float px2 = px * px;
        float py = (float)lane + (y_base - g.start.y);
        t = std::sqrt(std::fmaf(py, py, px2) * dot_recip);
    }
    uint32_t bits[4] = {0, 0, 0, 0};
    auto or_color = [&](const float c[4]) {
        for (int i = 0; i < 4; ++i) bits[i] |= f2u(c[i]);
    };
    bool acc = t <= g.stops[0].stop;
    if (acc) {
        const Color& s = g.stops[0].color;
        float c[4] = {s.r, s.g, s.b, s.a};
        or_color(c);
    }
    float start_stop = 0.0f;
    Color start_color = g.stops[0].color;
    for (size_t i = 1; i < g.stops.size(); ++i) {
        const Color& color = g.stops[i].color;
        float end_stop = g.stops[i].stop;
        bool mask = acc != (t < end_stop);
        if (mask) {
            float d = end_stop - start_stop;
            float local_t = (t - start_stop) * recip(d);
            float sc[4] = {start_color.r, start_color.g, start_color.b, start_color.a};
            float ec[4] = {color.r, color.g, color.b, color.a};
            float c[4];
            for (int k = 0; k < 4; ++k) c[k] = std::fmaf(local_t, ec[k], std::fmaf(-local_t, sc[k], sc[k]));
            or_color(c);
            acc = true;
        }
        start_stop = end_stop;
        start_color = color;
    }
    if (!acc) {
        const Color& s = g.stops.back().color;
        float c[4] = {s.r, s.g, s.b, s.a};
        or_color(c);
    }
    for (int i = 0; i < 4; ++i) out[i] = u2f(bits[i]);
}

// Texture::color_at for one lane (cpu/painter/styling.rs:146-193).
inline void texture_color_at(const Texture& tex, float x, float y_base, int lane, float out[4]) {
    float y = y_base + (float)lane;
    const Affine& t = tex.transform;
    float tx = std::fmaf(x, t.ux, std::fmaf(t.vx, y, t.tx));
    float ty = std::fmaf(x, t.uy, std::fmaf(t.vy, y, t.ty));
    uint32_t ix = sat_u32(rmin(tx, tex.image.max_x));
    uint32_t iy = sat_u32(rmin(ty, tex.image.max_y));
    uint32_t off = iy * tex.image.width + ix;
    const uint16_t* px = tex.image.data->data() + 4 * (size_t)off;
    for (int i = 0; i < 4; ++i) out[i] = f16_to(px[i]);
}

// cpu/painter/mod.rs:629-715
struct CachedTile {
    uint8_t tags = 0;
    uint32_t layer_count_v = 0;
    uint8_t solid[4] = {0, 0, 0, 0};
    bool has_layer_count() const { return tags & 2; }
    bool has_solid() const { return tags & 1; }
};

struct LayerCache {
    std::vector<CachedTile> tiles;
    bool has_clear = false;
    Color clear_color;
    bool has_size = false;
    size_t width = 0, height = 0;
    uint8_t id = 0;
    void clear() {  // BufferLayerCache::clear, cpu/buffer/mod.rs
        has_clear = false;
        for (auto& t : tiles) t = CachedTile();
    }
};

struct Rect {
    size_t hor0, hor1, vert0, vert1;  // tile ranges, cpu/renderer.rs:38-53
};

struct RenderTarget {
    uint8_t* buffer;
    size_t width, height, stride;
};

enum class WriteOp { None, Solid, ColorBuffer };

// cpu/painter/mod.rs:232-483 (Painter) + layer_workbench (per-tile driver).
struct Painter {
    int16_t areas[kTile * kTile];
    int8_t covers[(kTile + 1) * kTile];
    bool clip_active = false;
    float clip_mask[kTile * kTile];
    uint32_t clip_last = 0;
    float red[kTile * kTile], green[kTile * kTile], blue[kTile * kTile], alpha[kTile * kTile];
    uint8_t srgb[kTile * kTile * 4];

    void clear_cells() {
        std::memset(areas, 0, sizeof(areas));
        std::memset(covers, 0, sizeof(covers));
    }
    void acc_segment(uint64_t s) {
        int x = seg_local_x(s), y = seg_local_y(s);
        areas[x * kTile + y] = (int16_t)(areas[x * kTile + y] + seg_double_area(s));
        covers[(x + 1) * kTile + y] = (int8_t)(covers[(x + 1) * kTile + y] + seg_cover(s));
    }
    void acc_cover(const Cover& c) {
        for (int y = 0; y < kTile; ++y) covers[y] = (int8_t)(covers[y] + c.c[y]);
    }
    void clear(const Color& c) {
        for (int i = 0; i < kTile * kTile; ++i) {
            red[i] = c.r;
            green[i] = c.g;
            blue[i] = c.b;
            alpha[i] = c.a;
        }
    }

    static float coverage_of(int32_t da, FillRule fr) {
        if (fr == kNonZero) return rclamp(std::fabs((float)da * recip(512.0f)), 0.0f, 1.0f);
        int32_t v = 512 - std::abs((da & 1023) - 512);
        return (float)v * recip(512.0f);
    }

    // cpu/painter/mod.rs:290-347
    Cover paint_layer(size_t tile_x, size_t tile_y, uint32_t layer_id, const Props& props, bool apply_clip) {
        int8_t run[kTile] = {0};
        if (clip_active && clip_last < layer_id) clip_active = false;
        for (int x = 0; x <= kTile; ++x) {
            if (x != 0) {
                int px = x - 1;
                int32_t da[kTile];
                for (int y = 0; y < kTile; ++y) da[y] = 32 * (int32_t)run[y] + (int32_t)areas[px * kTile + y];
                for (int half = 0; half < 2; ++half) {
                    float cov[8];
                    bool all_zero = true;
                    if (props.fill_rule == kNonZero) {
                        for (int l = 0; l < 8; ++l) cov[l] = coverage_of(da[half * 8 + l], kNonZero);
                    } else {
                        for (int l = 0; l < 8; ++l) cov[l] = coverage_of(da[half * 8 + l], kEvenOdd);
                    }
                    for (int l = 0; l < 8; ++l)
                        if (!(cov[l] == 0.0f)) all_zero = false;
                    if (props.func == kDraw) {
                        if (all_zero) continue;
                        if (apply_clip && !clip_active) continue;
                        if (props.fill_type == kSolid && props.blend_mode == kOver) {
                            // The f32x8 of the reference (cpu/painter/mod.rs:406-447) for the common
                            // style, written so that the compiler turns it into 8-lane AVX2 code like
                            // the reference's SIMD shim: same operations as blend_at with blend = src.
                            const int base = px * kTile + half * 8;
                            const float sr = props.color.r, sg = props.color.g, sb = props.color.b, fa = props.color.a;
                            const bool clip = apply_clip && clip_active;
                            float* __restrict__ pr = red + base;
                            float* __restrict__ pg = green + base;
                            float* __restrict__ pb = blue + base;
                            float* __restrict__ pa = alpha + base;
                            const float* __restrict__ pm = clip_mask + base;
#pragma GCC ivdep
                            for (int l = 0; l < 8; ++l) {
                                float sa = fa * cov[l];
                                if (clip) sa *= pm[l];
                                const float da_ = pa[l];
                                const float inv_dst_a_src_a = (1.0f - da_) * sa;
                                const float inv_src_a = 1.0f - sa;
                                const float dst_a_src_a = da_ * sa;
                                const float cr = std::fmaf(sr, inv_dst_a_src_a, sr * dst_a_src_a);
                                const float cg = std::fmaf(sg, inv_dst_a_src_a, sg * dst_a_src_a);
                                const float cb = std::fmaf(sb, inv_dst_a_src_a, sb * dst_a_src_a);
                                pr[l] = std::fmaf(pr[l], inv_src_a, cr);
                                pg[l] = std::fmaf(pg[l], inv_src_a, cg);
                                pb[l] = std::fmaf(pb[l], inv_src_a, cb);
                                pa[l] = std::fmaf(da_, inv_src_a, sa);
                            }
                            continue;
                        }
                        float fx = (float)(px + tile_x * kTile);
                        float fy = (float)(half * 8 + tile_y * kTile);
                        for (int l = 0; l < 8; ++l) {
                            float fill[4];
                            if (props.fill_type == kSolid) {
                                fill[0] = props.color.r;
                                fill[1] = props.color.g;
                                fill[2] = props.color.b;
                                fill[3] = props.color.a;
                            } else if (props.fill_type == kGradient) {
                                gradient_color_at(props.gradient, fx, fy, l, fill);
                            } else {
                                texture_color_at(props.texture, fx, fy, l, fill);
                            }
                            blend_at(px * kTile + half * 8 + l, cov[l], apply_clip, fill, props.blend_mode);
                        }
                    } else {
                        // clip_at, cpu/painter/mod.rs:449-464
                        if (!clip_active) {
                            clip_active = true;
                            for (float& m : clip_mask) m = 0.0f;
                            clip_last = layer_id + props.clip_layers;
                        }
                        for (int l = 0; l < 8; ++l) clip_mask[px * kTile + half * 8 + l] = cov[l];
                    }
                }
            }
            for (int y = 0; y < kTile; ++y) run[y] = (int8_t)(run[y] + covers[x * kTile + y]);
        }
        Cover out;
        std::memcpy(out.c, run, sizeof(run));
        return out;
    }

    // cpu/painter/mod.rs:406-447
    void blend_at(int idx, float coverage, bool is_clipped, const float fill[4], BlendMode mode) {
        float dr = red[idx], dg = green[idx], db = blue[idx], da = alpha[idx];
        float sr = fill[0], sg = fill[1], sb = fill[2];
        float sa = fill[3] * coverage;
        if (is_clipped && clip_active) sa *= clip_mask[idx];
        float bl[3];
        lane_blend::blend(mode, dr, dg, db, sr, sg, sb, bl);
        float inv_dst_a = 1.0f - da;
        float inv_dst_a_src_a = inv_dst_a * sa;
        float inv_src_a = 1.0f - sa;
        float dst_a_src_a = da * sa;
        float cr = std::fmaf(sr, inv_dst_a_src_a, bl[0] * dst_a_src_a);
        float cg = std::fmaf(sg, inv_dst_a_src_a, bl[1] * dst_a_src_a);
        float cb = std::fmaf(sb, inv_dst_a_src_a, bl[2] * dst_a_src_a);
        red[idx] = std::fmaf(dr, inv_src_a, cr);
        green[idx] = std::fmaf(dg, inv_src_a, cg);
        blue[idx] = std::fmaf(db, inv_src_a, cb);
        alpha[idx] = std::fmaf(da, inv_src_a, sa);
    }

    // cpu/painter/mod.rs:466-483
    void compute_srgb(const Channel ch[4]) {
        // Plane by plane (vectorisable like the reference's f32x8 code), then interleaved in
        // the requested channel order.
        alignas(32) uint8_t plane[6][kTile * kTile];
        for (int i = 0; i < kTile * kTile; ++i) plane[0][i] = to_byte(linear_to_srgb(red[i]));
        for (int i = 0; i < kTile * kTile; ++i) plane[1][i] = to_byte(linear_to_srgb(green[i]));
        for (int i = 0; i < kTile * kTile; ++i) plane[2][i] = to_byte(linear_to_srgb(blue[i]));
        for (int i = 0; i < kTile * kTile; ++i) plane[3][i] = to_byte(alpha[i]);
        std::memset(plane[4], to_byte(0.0f), sizeof(plane[4]));
        std::memset(plane[5], to_byte(1.0f), sizeof(plane[5]));
        const uint8_t* src[4];
        for (int k = 0; k < 4; ++k) {
            switch (ch[k]) {
                case kRed: src[k] = plane[0]; break;
                case kGreen: src[k] = plane[1]; break;
                case kBlue: src[k] = plane[2]; break;
                case kAlpha: src[k] = plane[3]; break;
                case kZero: src[k] = plane[4]; break;
                default: src[k] = plane[5]; break;
            }
        }
        for (int i = 0; i < kTile * kTile; ++i) {
            srgb[i * 4 + 0] = src[0][i];
            srgb[i * 4 + 1] = src[1][i];
            srgb[i * 4 + 2] = src[2][i];
            srgb[i * 4 + 3] = src[3][i];
        }
    }
};

// LayerProps (cpu/painter/mod.rs:164-167). The reference answers `get` with a hash look-up
// (FxHashMap); here the layers are spread once per frame into an array indexed by order, so
// that the per-(tile, layer) look-ups of the painter cost no more than they do there.
struct PropsSource {
    std::vector<const Layer*> by_order;
    bool has_cache = false;
    uint8_t cache_id = 0;
    void index(const std::map<uint32_t, Layer*>& layers) {
        by_order.assign(layers.empty() ? 0 : (size_t)layers.rbegin()->first + 1, nullptr);
        for (auto& kv : layers) by_order[kv.first] = kv.second;
    }
    const Props& get(uint32_t id) const { return by_order.at(id)->props; }
    bool is_unchanged(uint32_t id) const {
        if (!has_cache) return false;
        return (by_order.at(id)->is_unchanged >> cache_id) & 1;
    }
};

struct TileContext {
    size_t tile_x, tile_y;
    const uint64_t* segs;
    size_t n_segs;
    const PropsSource* props;
    bool has_cached_clear;
    Color cached_clear;
    CachedTile* cached_tile;
    const Channel* channels;
    Color clear_color;
};

// layer_workbench/mod.rs:147-343 + passes/*.rs
struct Workbench {
    // The reference keeps `segment_ranges`, `queue_indices` and `skip_clipping` in per-tile
    // FxHashMaps keyed by layer id. Both the segments of a tile and the carry queue are ordered
    // by layer id, so here the same look-ups are binary searches over those two sorted arrays
    // and a flag next to each id — same answers, no per-tile node allocations.
    struct Id {
        uint32_t id;
        bool mask;
        bool skip_clipping;
    };
    struct SegRange {
        uint32_t id;
        size_t first, last;  // inclusive
    };
    std::vector<Id> ids;
    size_t skipped = 0;
    std::vector<SegRange> segment_ranges;        // ascending ids
    std::vector<CoverCarry> queue, next_queue;   // ascending layer ids
    bool layers_were_removed = true;

    const SegRange* segments_of(uint32_t id) const {
        auto it = std::lower_bound(segment_ranges.begin(), segment_ranges.end(), id,
                                   [](const SegRange& r, uint32_t v) { return r.id < v; });
        return it != segment_ranges.end() && it->id == id ? &*it : nullptr;
    }

    void init(std::vector<CoverCarry>&& carries) { queue = std::move(carries); }

    void next_tile() {
        ids.clear();
        skipped = 0;
        segment_ranges.clear();
        std::swap(queue, next_queue);
        next_queue.clear();
        layers_were_removed = true;
    }

    const Cover* cover(uint32_t id) const {
        auto it = std::lower_bound(queue.begin(), queue.end(), id,
                                   [](const CoverCarry& c, uint32_t v) { return c.layer_id < v; });
        return it != queue.end() && it->layer_id == id ? &it->cover : nullptr;
    }
    bool has_segments(uint32_t id) const { return segments_of(id) != nullptr; }
    bool layer_is_full(uint32_t id, FillRule fr) const {
        if (has_segments(id)) return false;
        const Cover* c = cover(id);
        return c ? c->is_full(fr) : false;
    }

    // layer_workbench/mod.rs:213-234
    bool cover_carry(const TileContext& ctx, uint32_t id, CoverCarry* out) const {
        Cover acc;
        if (const SegRange* r = segments_of(id)) {
            for (size_t i = r->first; i <= r->last; ++i) {
                int y = seg_local_y(ctx.segs[i]);
                acc.c[y] = (int8_t)(acc.c[y] + seg_cover(ctx.segs[i]));
            }
        }
        if (const Cover* c = cover(id)) {
            for (int y = 0; y < kTile; ++y) acc.c[y] = (int8_t)(acc.c[y] + c->c[y]);
        }
        if (acc.is_empty(ctx.props->get(id).fill_rule)) return false;
        out->cover = acc;
        out->layer_id = id;
        return true;
    }

    // layer_workbench/mod.rs:250-278
    void populate_layers(const TileContext& ctx) {
        size_t start = 0;
        while (start < ctx.n_segs) {
            uint32_t id = seg_layer(ctx.segs[start]);
            size_t end = start;
            while (end + 1 < ctx.n_segs && seg_layer(ctx.segs[end + 1]) == id) ++end;
            segment_ranges.push_back({id, start, end});
            start = end + 1;
        }
        // Union of the two ascending id lists (layers with segments, carried layers).
        size_t a = 0, b = 0;
        while (a < segment_ranges.size() || b < queue.size()) {
            uint32_t id;
            if (b == queue.size() || (a < segment_ranges.size() && segment_ranges[a].id <= queue[b].layer_id)) {
                id = segment_ranges[a].id;
                if (b < queue.size() && queue[b].layer_id == id) ++b;
                ++a;
            } else {
                id = queue[b++].layer_id;
            }
            ids.push_back({id, true, false});
        }
    }

    enum class Flow { Continue, BreakNone, BreakSolid };

    // passes/tile_unchanged.rs:24-57
    Flow tile_unchanged_pass(const TileContext& ctx) {
        bool clear_unchanged = ctx.has_cached_clear && ctx.cached_clear == ctx.clear_color;
        if (!ctx.cached_tile) return Flow::Continue;
        uint32_t layers = (uint32_t)ids.size();
        bool had = ctx.cached_tile->has_layer_count();
        uint32_t previous = ctx.cached_tile->layer_count_v;
        ctx.cached_tile->tags |= 2;
        ctx.cached_tile->layer_count_v = layers & 0xFFFFFF;
        bool is_unchanged = false;
        if (had) {
            layers_were_removed = layers < previous;
            is_unchanged = previous == layers;
            if (is_unchanged) {
                for (auto& e : ids)
                    if (!ctx.props->is_unchanged(e.id)) {
                        is_unchanged = false;
                        break;
                    }
            }
        }
        return (clear_unchanged && is_unchanged) ? Flow::BreakNone : Flow::Continue;
    }

    // passes/skip_trivial_clips.rs:27-112
    void skip_trivial_clips_pass(const TileContext& ctx) {
        struct Clip {
            bool is_full;
            uint32_t last_layer_id;
            size_t i;
            bool is_used;
        };
        bool has_clip = false;
        Clip clip{};
        for (size_t i = skipped; i < ids.size(); ++i) {
            if (!ids[i].mask) continue;
            uint32_t id = ids[i].id;
            const Props& props = ctx.props->get(id);
            if (props.func == kClip) {
                bool is_full = layer_is_full(id, props.fill_rule);
                clip = {is_full, id + props.clip_layers, i, false};
                has_clip = true;
                if (is_full) ids[i].mask = false;
            }
            if (props.func == kDraw && props.is_clipped) {
                if (has_clip && id <= clip.last_layer_id) {
                    if (clip.is_full) ids[i].skip_clipping = true;
                    else clip.is_used = true;
                } else {
                    ids[i].mask = false;
                }
            }
            if (has_clip && id > clip.last_layer_id) {
                has_clip = false;
                if (!clip.is_used) ids[clip.i].mask = false;
            }
        }
        if (has_clip && !clip.is_used) ids[clip.i].mask = false;
    }

    // passes/skip_fully_covered_layers.rs:27-119
    Flow skip_fully_covered_layers_pass(const TileContext& ctx, Color* solid) {
        enum { kNoneYet, kOpaque, kIncomplete } first = kNoneYet;
        Color opaque_color;
        bool visible_unchanged = !layers_were_removed;
        for (size_t k = ids.size(); k-- > skipped;) {
            if (!ids[k].mask) continue;
            uint32_t id = ids[k].id;
            const Props& props = ctx.props->get(id);
            if (!ctx.props->is_unchanged(id)) visible_unchanged = false;
            bool is_clipped = props.func == kDraw && props.is_clipped && !ids[k].skip_clipping;
            if (is_clipped || !layer_is_full(id, props.fill_rule)) {
                if (first == kNoneYet) first = kIncomplete;
            } else if (props.func == kDraw && props.fill_type == kSolid && props.blend_mode == kOver) {
                if (props.color.a == 1.0f) {
                    if (first == kNoneYet) {
                        first = kOpaque;
                        opaque_color = props.color;
                    }
                    skipped = k;
                    break;
                }
            }
        }
        size_t skip_n;
        Color bottom;
        if (first == kOpaque) {
            if (visible_unchanged) return Flow::BreakNone;
            skip_n = 1;
            bottom = opaque_color;
        } else if (first == kNoneYet) {
            skip_n = 0;
            bottom = ctx.clear_color;
        } else {
            return Flow::Continue;
        }
        Color dst = bottom;
        size_t seen = 0;
        for (size_t k = skipped; k < ids.size(); ++k) {
            if (!ids[k].mask) continue;
            if (seen++ < skip_n) continue;
            const Props& props = ctx.props->get(ids[k].id);
            if (props.func == kDraw && props.fill_type == kSolid) {
                dst = scalar_blend::blend(props.blend_mode, dst, props.color);
            } else {
                return Flow::Continue;
            }
        }
        *solid = dst;
        return Flow::BreakSolid;
    }

    // layer_workbench/mod.rs:280-342 + CachedTile::convert_optimizer_op
    // (cpu/painter/mod.rs:686-714).
    WriteOp drive_tile_painting(Painter& painter, const TileContext& ctx, uint8_t solid_out[4]) {
        populate_layers(ctx);

        Flow flow = tile_unchanged_pass(ctx);
        Color solid;
        if (flow == Flow::Continue) {
            skip_trivial_clips_pass(ctx);
            flow = skip_fully_covered_layers_pass(ctx, &solid);
        }

        bool brk = false;
        WriteOp op = WriteOp::ColorBuffer;
        if (flow == Flow::BreakSolid) {
            float sel[4];
            for (int k = 0; k < 4; ++k) sel[k] = color_channel(solid, ctx.channels[k]);
            uint8_t bytes[4];
            to_srgb_bytes(sel, bytes);
            bool unchanged = false;
            if (ctx.cached_tile) {
                bool had = ctx.cached_tile->has_solid();
                unchanged = had && std::memcmp(ctx.cached_tile->solid, bytes, 4) == 0;
                ctx.cached_tile->tags |= 1;
                std::memcpy(ctx.cached_tile->solid, bytes, 4);
            }
            std::memcpy(solid_out, bytes, 4);
            op = unchanged ? WriteOp::None : WriteOp::Solid;
            brk = true;
        } else if (flow == Flow::BreakNone) {
            op = WriteOp::None;
            brk = true;
        } else if (ctx.cached_tile) {
            ctx.cached_tile->tags &= 2;  // update_solid_color(None)
        }

        if (brk) {
            for (auto& e : ids) {
                CoverCarry cc;
                if (cover_carry(ctx, e.id, &cc)) next_queue.push_back(cc);
            }
            next_tile();
            return op;
        }

        painter.clear(ctx.clear_color);
        for (size_t k = 0; k < ids.size(); ++k) {
            uint32_t id = ids[k].id;
            bool mask = k >= skipped && ids[k].mask;
            if (mask) {
                painter.clear_cells();
                if (const SegRange* r = segments_of(id))
                    for (size_t i = r->first; i <= r->last; ++i) painter.acc_segment(ctx.segs[i]);
                if (const Cover* c = cover(id)) painter.acc_cover(*c);
                const Props& props = ctx.props->get(id);
                bool apply_clip = false;
                if (props.func == kDraw) apply_clip = props.is_clipped && !ids[k].skip_clipping;
                Cover out = painter.paint_layer(ctx.tile_x, ctx.tile_y, id, props, apply_clip);
                if (!out.is_empty(props.fill_rule)) next_queue.push_back({out, id});
            } else {
                CoverCarry cc;
                if (cover_carry(ctx, id, &cc)) next_queue.push_back(cc);
            }
        }
        next_tile();
        return WriteOp::ColorBuffer;
    }
};

// LinearLayout::write, cpu/buffer/layout/mod.rs:265-282
inline void write_tile(const RenderTarget& rt, size_t tile_x, size_t tile_y, const uint8_t* colors_col_major,
                       const uint8_t solid[4]) {
    size_t x0 = tile_x * kTile, y0 = tile_y * kTile;
    for (size_t y = 0; y < (size_t)kTile && y0 + y < rt.height; ++y) {
        uint8_t* row = rt.buffer + (y0 + y) * rt.stride;
        for (size_t x = 0; x < (size_t)kTile && x0 + x < rt.width; ++x) {
            const uint8_t* src = colors_col_major ? colors_col_major + (x * kTile + y) * 4 : solid;
            std::memcpy(row + (x0 + x) * 4, src, 4);
        }
    }
}

// cpu/painter/mod.rs:486-568 (paint_tile_row) for one tile row; `segs` is the
// sorted range with tile_y == row.
inline void paint_tile_row(Painter& painter, Workbench& wb, size_t tile_y, const uint64_t* segs, size_t n,
                           const PropsSource& props, const Channel ch[4], const Color& clear_color,
                           bool has_prev_clear, const Color& prev_clear, CachedTile* cached_tiles,
                           const RenderTarget& rt, const Rect* crop) {
    std::map<uint32_t, Cover> left;
    int16_t tile_x_start = crop ? (int16_t)crop->hor0 : 0;
    size_t pos = 0;
    while (pos < n && seg_tile_x(segs[pos]) < tile_x_start) {
        Cover& c = left[seg_layer(segs[pos])];
        int y = seg_local_y(segs[pos]);
        c.c[y] = (int8_t)(c.c[y] + seg_cover(segs[pos]));
        ++pos;
    }
    std::vector<CoverCarry> carries;
    for (auto& kv : left) carries.push_back({kv.second, kv.first});
    wb.init(std::move(carries));
    wb.next_queue.clear();
    wb.ids.clear();
    wb.skipped = 0;
    wb.segment_ranges.clear();
    wb.layers_were_removed = true;

    size_t width_in_tiles = (rt.width + kTile - 1) / kTile;
    for (size_t tile_x = 0; tile_x < width_in_tiles; ++tile_x) {
        if (crop && !(tile_x >= crop->hor0 && tile_x < crop->hor1)) continue;
        size_t begin = pos;
        while (pos < n && seg_tile_x(segs[pos]) == (int16_t)tile_x) ++pos;
        TileContext ctx;
        ctx.tile_x = tile_x;
        ctx.tile_y = tile_y;
        ctx.segs = segs + begin;
        ctx.n_segs = pos - begin;
        ctx.props = &props;
        ctx.has_cached_clear = has_prev_clear;
        ctx.cached_clear = prev_clear;
        ctx.cached_tile = cached_tiles ? cached_tiles + tile_x : nullptr;
        ctx.channels = ch;
        ctx.clear_color = clear_color;
        painter.clip_active = false;
        uint8_t solid[4];
        WriteOp op = wb.drive_tile_painting(painter, ctx, solid);
        if (op == WriteOp::Solid) {
            write_tile(rt, tile_x, tile_y, nullptr, solid);
        } else if (op == WriteOp::ColorBuffer) {
            painter.compute_srgb(ch);
            write_tile(rt, tile_x, tile_y, painter.srgb, nullptr);
        }
    }
}

}  // namespace fo
