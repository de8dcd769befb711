// Levien flattening parameters of one quadratic (path.rs:271-347, push_quad),
// shared by the host (spline bookkeeping, host_path.cpp, compiled by g++ with
// -ffp-contract=off) and the device (quad_expand_kernel, nvcc --fmad=false):
// the same IEEE operations in the same order on both sides, so the device can
// recompute them instead of receiving them over PCIe.
#pragma once

#include <cmath>

#ifdef __CUDACC__
#define FORMA_HD __host__ __device__ __forceinline__
#else
#define FORMA_HD inline
#endif

namespace forma {

struct QuadParams {
    float x0, dx_recip, k0, dk, cur;  // cur = this quad's share of the spline's curvature (> 1)
};

FORMA_HD float quad_curvature(float x) {  // path.rs:48-51
    const float c = 0.67f;
    return x / (1.0f - c + sqrtf(sqrtf(fmaf(x * x, 0.25f, c * c * c * c))));
}

// Control points are given pre-multiplied by their weights (x, y, w), as in QuadRec.
FORMA_HD QuadParams quad_params(const float px[3], const float py[3], const float pw[3]) {
    const float kInvMaxError = 16.0f;  // 1 / MAX_ERROR, path.rs:40
    const float r0 = 1.0f / pw[0], r1 = 1.0f / pw[1], r2 = 1.0f / pw[2];  // WeightedPoint::applied, path.rs:65-72
    const float p0x = px[0] * r0, p0y = py[0] * r0;
    const float p1x = px[1] * r1, p1y = py[1] * r1;
    const float p2x = px[2] * r2, p2y = py[2] * r2;
    const float ax = p1x - p0x, ay = p1y - p0y, bx = p2x - p1x, by = p2y - p1y;
    const float hx = ax - bx, hy = ay - by;
    const float cross = fmaf(p2x - p0x, hy, -(p2y - p0y) * hx);
    const float cross_recip = 1.0f / cross;
    QuadParams q;
    q.x0 = fmaf(ax, hx, ay * hy) * cross_recip;
    const float x2 = fmaf(bx, hx, by * hy) * cross_recip;
    q.dx_recip = 1.0f / (x2 - q.x0);
    const float len_h = sqrtf(hx * hx + hy * hy);
    const float scale = fabsf(cross / (len_h * (x2 - q.x0)));
    q.k0 = quad_curvature(q.x0);
    q.dk = quad_curvature(x2) - q.k0;
    q.cur = 0.5f * fabsf(q.dk) * sqrtf(scale * kInvMaxError);
    const bool finite = fabsf(q.cur) <= 3.402823466e+38f;  // false for NaN and infinities
    if (!finite || q.cur <= 1.0f) {  // collinear, path.rs:322-332
        q.x0 = 0.03662467f;
        q.dx_recip = 1.0f;
        q.k0 = 0.0f;
        q.dk = 1.0f;
        q.cur = 2.0f;
    }
    return q;
}

}  // namespace forma
