// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of google/forma's CPU rendering path, used as the parity
// checker for the CUDA implementation in forma_b200/. Nothing under
// forma_b200/ may include, link or call anything in this directory; only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs do.
//
// Arithmetic discipline: every `mul_add` in the Rust
// source is an explicit fmaf()/fma() here; every other operation is a
// separately rounded IEEE op. The translation unit MUST be compiled with
// -ffp-contract=off and without -ffast-math. Vector code follows the portable
// shim forma/src/utils/simd/auto.rs (fused mul_add :732-738, exact recip
// :727-730, NaN-ignoring min/max :698-712).
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

namespace fo {

constexpr float kPi = 3.14159274101257324f;        // f32::consts::PI
constexpr float kFracPi2 = 1.57079637050628662f;   // f32::consts::FRAC_PI_2
constexpr float kEps = 1.1920928955078125e-7f;     // f32::EPSILON

inline float recip(float v) { return 1.0f / v; }

// Rust f32::min / f32::max: if exactly one operand is NaN the other is returned.
inline float rmin(float a, float b) { return std::fmin(a, b); }
inline float rmax(float a, float b) { return std::fmax(a, b); }

// Rust f32::clamp (NaN stays NaN).
inline float rclamp(float v, float lo, float hi) {
    if (v < lo) v = lo;
    if (v > hi) v = hi;
    return v;
}

inline uint32_t f2u(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}
inline float u2f(uint32_t u) {
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

// Rust `as u32` / `as usize` from f32: saturating, NaN -> 0.
inline uint32_t sat_u32(float v) {
    if (!(v > 0.0f)) return 0;  // negatives, -0, NaN
    if (v >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)v;
}
inline uint64_t sat_usize(float v) {
    if (!(v > 0.0f)) return 0;
    if (v >= 18446744073709551616.0f) return ~0ull;
    return (uint64_t)v;
}

// forma/src/path.rs:44-46
inline float lerp(float t, float a, float b) { return std::fmaf(t, b, std::fmaf(-t, a, a)); }

struct Point {
    float x = 0.0f, y = 0.0f;
};
inline Point operator+(Point a, Point b) { return {a.x + b.x, a.y + b.y}; }
inline Point operator-(Point a, Point b) { return {a.x - b.x, a.y - b.y}; }
inline bool operator==(Point a, Point b) { return a.x == b.x && a.y == b.y; }
inline bool operator!=(Point a, Point b) { return !(a == b); }

// forma/src/math/point.rs:83-85
inline float point_len(Point p) { return std::sqrt(p.x * p.x + p.y * p.y); }

// forma/src/math/point.rs:54-78
inline float approx_atan2(float y, float x) {
    float x_abs = std::fabs(x);
    float y_abs = std::fabs(y);
    float a = rmin(x_abs, y_abs) / rmax(x_abs, y_abs);
    float s = a * a;
    float r = std::fmaf(std::fmaf(std::fmaf(s, -0.046496473f, 0.15931422f), s, -0.32762277f), s * a, a);
    if (y_abs > x_abs) r = kFracPi2 - r;
    if (x < 0.0f) r = kPi - r;
    if (y < 0.0f) r = -r;
    return r;
}

// forma/src/math/point.rs:87-89 — returns false when the vector is too short.
inline bool point_angle(Point p, float* out) {
    if (point_len(p) >= kEps) {
        *out = approx_atan2(p.y, p.x);
        return true;
    }
    return false;
}

// forma/src/math/transform.rs:33-57
struct Affine {
    float ux = 1.0f, uy = 0.0f, vx = 0.0f, vy = 1.0f, tx = 0.0f, ty = 0.0f;
    Point apply(Point p) const {
        return {std::fmaf(ux, p.x, std::fmaf(vx, p.y, tx)), std::fmaf(uy, p.x, std::fmaf(vy, p.y, ty))};
    }
    bool is_identity() const {
        return ux == 1.0f && uy == 0.0f && vx == 0.0f && vy == 1.0f && tx == 0.0f && ty == 0.0f;
    }
    bool operator==(const Affine& o) const {
        return ux == o.ux && uy == o.uy && vx == o.vx && vy == o.vy && tx == o.tx && ty == o.ty;
    }
};

// forma/src/math/transform.rs:208-221 (GeomPresTransform::try_from)
inline bool geom_pres_ok(const Affine& t) {
    const float max_x = 1.0f + (1.0f / 16.0f) / 65536.0f;
    const float max_y = 1.0f + (1.0f / 16.0f) / 32768.0f;
    bool sx = t.ux * t.ux + t.uy * t.uy > max_x;
    bool sy = t.vx * t.vx + t.vy * t.vy > max_y;
    return !sx && !sy;
}

}  // namespace fo
