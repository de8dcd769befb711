// ORACLE — TEST INFRASTRUCTURE ONLY (see fmath.hpp).
//
// Stage 1: PathBuilder command stream -> primitives (lines + rational quads,
// merged into splines) -> flattened points. Restates forma/src/path.rs.
#pragma once

#include <memory>
#include <vector>

#include "fmath.hpp"

namespace fo {

constexpr float kMaxError = 1.0f / 16.0f;      // path.rs:40
constexpr float kMaxAngleError = 0.001f;       // path.rs:41

// path.rs:48-51
inline float curvature(float x) {
    const float c = 0.67f;
    return x / (1.0f - c + std::sqrt(std::sqrt(std::fmaf(x * x, 0.25f, c * c * c * c))));
}
// path.rs:53-56
inline float inv_curvature(float k) {
    const float c = 0.39f;
    return k * (1.0f - c + std::sqrt(std::fmaf(k * k, 0.25f, c * c)));
}

struct WPoint {
    Point p;
    float w;
    // path.rs:65-72
    Point applied() const {
        float r = recip(w);
        return {p.x * r, p.y * r};
    }
};

// path.rs:75-120
inline WPoint eval_cubic(float t, const WPoint* q) {
    auto c3 = [t](float a, float b, float c, float d) {
        return lerp(t, lerp(t, lerp(t, a, b), lerp(t, b, c)), lerp(t, lerp(t, b, c), lerp(t, c, d)));
    };
    WPoint r;
    r.p.x = c3(q[0].p.x, q[1].p.x, q[2].p.x, q[3].p.x);
    r.p.y = c3(q[0].p.y, q[1].p.y, q[2].p.y, q[3].p.y);
    r.w = c3(q[0].w, q[1].w, q[2].w, q[3].w);
    return r;
}

struct Segments {
    std::vector<float> x, y;
    std::vector<uint8_t> start_new_contour;
};

// path.rs:174-188
struct Spline {
    float curvature = 0.0f;
    Point p0, p2;
    bool contour = false;  // Option<Contour>: true on the last spline of a contour
};

// path.rs:190-538
struct Primitives {
    bool has_last_angle = false;
    float last_angle = 0.0f;
    bool contour = true;  // Default: Some(Contour), path.rs:545
    std::vector<Spline> splines;
    std::vector<float> x, y, weight;  // 3 per quad
    std::vector<float> x0, dx_recip, k0, dk, curvatures_recip;
    std::vector<uint32_t> pc_spline;
    std::vector<float> pc_total;

    // path.rs:208-246. `make` builds the new spline when one is needed.
    template <class F>
    Spline& last_spline_or_insert_with(bool has_angle, float angle, Point point, F make) {
        bool got = false;
        if (contour) {
            contour = false;
            got = true;
        } else {
            bool angle_changed = false;
            if (has_last_angle && has_angle) {
                float diff = std::fabs(angle - last_angle);
                if (diff > kPi) diff -= kPi;
                if (diff > kFracPi2) diff = kPi - diff;
                angle_changed = diff > kMaxAngleError;
            }
            if (!splines.empty()) {
                Spline& last = splines.back();
                // Spline::new_spline_needed, path.rs:183-187
                bool needed = angle_changed || point_len(point - last.p2) >= kMaxError;
                if (needed && last.contour) {
                    last.contour = false;
                    got = true;
                }
            }
        }
        if (got) splines.push_back(make());
        return splines.back();
    }

    void push_contour() { contour = true; }

    // path.rs:252-269
    void push_line(WPoint a, WPoint b) {
        Point p0 = a.applied();
        Point p1 = b.applied();
        Point d = p1 - p0;
        float angle = 0.0f;
        bool has_angle = point_angle(d, &angle);
        Spline& s = last_spline_or_insert_with(has_angle, angle, p0, [&] {
            Spline n;
            n.curvature = 0.0f;
            n.p0 = p0;
            n.p2 = p1;
            n.contour = true;
            return n;
        });
        s.p2 = p1;
        has_last_angle = has_angle;
        last_angle = angle;
    }

    // path.rs:271-347
    void push_quad(WPoint q0, WPoint q1, WPoint q2) {
        const float pixel_accuracy_recip = 1.0f / kMaxError;
        Point p0 = q0.applied(), p1 = q1.applied(), p2 = q2.applied();
        Point a = p1 - p0;
        Point b = p2 - p1;
        float in_angle = 0.0f, out_angle = 0.0f;
        bool has_in = point_angle(a, &in_angle);
        bool has_out = point_angle(b, &out_angle);
        if (!has_in && !has_out) return;
        if (!has_in || !has_out) return push_line(q0, q2);

        for (const WPoint* q : {&q0, &q1, &q2}) {
            x.push_back(q->p.x);
            y.push_back(q->p.y);
            weight.push_back(q->w);
        }

        Spline& s = last_spline_or_insert_with(has_in, in_angle, p0, [&] {
            Spline n;
            n.curvature = 0.0f;
            n.p0 = p0;
            n.p2 = p2;
            n.contour = true;
            return n;
        });
        s.p2 = p2;

        Point h = a - b;
        float cross = std::fmaf(p2.x - p0.x, h.y, -(p2.y - p0.y) * h.x);
        float cross_recip = recip(cross);

        float vx0 = std::fmaf(a.x, h.x, a.y * h.y) * cross_recip;
        float vx2 = std::fmaf(b.x, h.x, b.y * h.y) * cross_recip;
        float vdx_recip = recip(vx2 - vx0);
        float scale = std::fabs(cross / (point_len(h) * (vx2 - vx0)));
        float vk0 = curvature(vx0);
        float vk2 = curvature(vx2);
        float vdk = vk2 - vk0;
        float cur = 0.5f * std::fabs(vdk) * std::sqrt(scale * pixel_accuracy_recip);

        if (!std::isfinite(cur) || cur <= 1.0f) {
            vx0 = 0.03662467f;
            vdx_recip = 1.0f;
            vk0 = 0.0f;
            vdk = 1.0f;
            cur = 2.0f;
        }

        float total = s.curvature + cur;
        s.curvature = total;

        has_last_angle = has_out;
        last_angle = out_angle;

        x0.push_back(vx0);
        dx_recip.push_back(vdx_recip);
        k0.push_back(vk0);
        dk.push_back(vdk);
        curvatures_recip.push_back(recip(cur));
        pc_spline.push_back((uint32_t)splines.size() - 1);
        pc_total.push_back(total);
    }

    // path.rs:349-398
    void push_cubic(const WPoint* q) {
        const float max_cubic_error_squared = (36.0f * 36.0f / 3.0f) * kMaxError * kMaxError;
        Point p0 = q[0].applied(), p1 = q[1].applied(), p2 = q[2].applied();
        float dx = std::fmaf(p2.x, 3.0f, -p0.x) - std::fmaf(p1.x, 3.0f, -p1.x);
        float dy = std::fmaf(p2.y, 3.0f, -p0.y) - std::fmaf(p1.y, 3.0f, -p1.y);
        float err = std::fmaf(dx, dx, dy * dy);
        float mult = rmax(rmax(q[1].w, q[2].w), 1.0f);
        uint64_t subdivisions =
            sat_usize(std::ceil(std::pow(err * recip(max_cubic_error_squared), 1.0f / 6.0f) * mult));
        if (subdivisions < 1) subdivisions = 1;
        float incr = recip((float)subdivisions);

        Point quad_p0 = p0;
        for (uint64_t i = 1; i <= subdivisions; ++i) {
            float t = (float)i * incr;
            Point quad_p2 = eval_cubic(t, q).applied();
            Point mid = eval_cubic(t - 0.5f * incr, q).applied();
            Point quad_p1 = {std::fmaf(mid.x, 2.0f, -0.5f * (quad_p0.x + quad_p2.x)),
                             std::fmaf(mid.y, 2.0f, -0.5f * (quad_p0.y + quad_p2.y))};
            push_quad({quad_p0, 1.0f}, {quad_p1, 1.0f}, {quad_p2, 1.0f});
            quad_p0 = quad_p2;
        }
    }

    // path.rs:447-471
    Point eval_quad(size_t qi, float t) const {
        size_t i0 = 3 * qi, i1 = i0 + 1, i2 = i0 + 2;
        float w = lerp(t, lerp(t, weight[i0], weight[i1]), lerp(t, weight[i1], weight[i2]));
        float w_recip = recip(w);
        float px = lerp(t, lerp(t, x[i0], x[i1]), lerp(t, x[i1], x[i2])) * w_recip;
        float py = lerp(t, lerp(t, y[i0], y[i1]), lerp(t, y[i1], y[i2])) * w_recip;
        return {px, py};
    }

    // path.rs:400-445 (populate_buffers) fused with the per-point evaluation of
    // path.rs:487-534; the point commands are consumed as they are produced.
    Segments into_segments() const {
        Segments out;
        auto emit = [&](Point p, bool c) {
            out.x.push_back(p.x);
            out.y.push_back(p.y);
            out.start_new_contour.push_back(c ? 1 : 0);
        };
        size_t i = 0;
        const Spline* last = nullptr;
        for (size_t si = 0; si < splines.size(); ++si) {
            const Spline& sp = splines[si];
            uint64_t subdivisions = sat_usize(std::ceil(sp.curvature));
            float point_command = sp.curvature / (float)subdivisions;
            bool needs_start = !last || last->contour || point_len(last->p2 - sp.p0) > kMaxError;
            if (needs_start) emit(sp.p0, false);
            for (uint64_t pi = 1; pi < subdivisions; ++pi) {
                if ((float)pi > pc_total[i]) i += 1;
                // path.rs:506-524
                size_t qi = i;
                uint32_t spline_i = pc_spline[qi];
                float previous = 0.0f;
                if (qi >= 1 && pc_spline[qi - 1] == spline_i) previous = pc_total[qi - 1];
                float ratio = std::fmaf(point_command, (float)pi, -previous) * curvatures_recip[qi];
                float xx = inv_curvature(std::fmaf(ratio, dk[qi], k0[qi]));
                float t = rclamp((xx - x0[qi]) * dx_recip[qi], 0.0f, 1.0f);
                emit(eval_quad(qi, t), false);
            }
            emit(sp.p2, sp.contour);
            last = &sp;
            if (subdivisions > 0) i += 1;
        }
        return out;
    }
};

enum PathCommand : uint8_t { kMove = 0, kLine = 1, kQuad = 2, kCubic = 3 };

// path.rs:574-668
struct PathData {
    std::vector<float> x{0.0f}, y{0.0f}, weight{1.0f};
    std::vector<uint8_t> commands{kMove};
    size_t open_point_index = 0;
    bool has_segments = false;
    Segments segs;

    WPoint at(size_t i) const { return {{x[i], y[i]}, weight[i]}; }

    // path.rs:596-615
    void close() {
        size_t len = x.size();
        WPoint last = at(len - 1);
        WPoint open = at(open_point_index);
        if (last.applied() != open.applied()) {
            x.push_back(open.p.x);
            y.push_back(open.p.y);
            weight.push_back(open.w);
            commands.push_back(kLine);
        }
    }

    // path.rs:617-654
    const Segments& segments() {
        if (!has_segments) {
            Primitives prim;
            size_t i = 0;
            for (uint8_t c : commands) {
                switch (c) {
                    case kMove:
                        i += 1;
                        prim.push_contour();
                        break;
                    case kLine:
                        i += 1;
                        prim.push_line(at(i - 2), at(i - 1));
                        break;
                    case kQuad:
                        i += 2;
                        prim.push_quad(at(i - 3), at(i - 2), at(i - 1));
                        break;
                    case kCubic: {
                        i += 3;
                        WPoint q[4] = {at(i - 4), at(i - 3), at(i - 2), at(i - 1)};
                        prim.push_cubic(q);
                        break;
                    }
                }
            }
            segs = prim.into_segments();
            has_segments = true;
        }
        return segs;
    }
};

// path.rs:670-766
struct Path {
    std::shared_ptr<PathData> inner = std::make_shared<PathData>();
    bool has_transform = false;
    Affine transform;

    // path.rs:726-765 + GeomPresTransform::new (math/transform.rs:161-182)
    Path transformed(const float m_in[9]) const {
        float m[9];
        for (int i = 0; i < 9; ++i) m[i] = m_in[i];
        if (std::fabs(m[6]) <= kEps && std::fabs(m[7]) <= kEps) {
            float a[6] = {m[0], m[1], m[2], m[3], m[4], m[5]};
            if (std::fabs(m[8] - 1.0f) > kEps) {
                float r = recip(m[8]);
                for (float& v : a) v *= r;
            }
            Affine t;
            t.ux = a[0];
            t.vx = a[1];
            t.tx = a[2];
            t.uy = a[3];
            t.vy = a[4];
            t.ty = a[5];
            if (geom_pres_ok(t)) {
                Path p;
                p.inner = inner;
                p.has_transform = true;
                p.transform = t;
                return p;
            }
        }
        auto data = std::make_shared<PathData>();
        data->x = inner->x;
        data->y = inner->y;
        data->weight = inner->weight;
        data->commands = inner->commands;
        data->open_point_index = inner->open_point_index;
        for (size_t i = 0; i < data->x.size(); ++i) {
            float px = data->x[i], py = data->y[i], pw = data->weight[i];
            data->x[i] = std::fmaf(m_in[0], px, std::fmaf(m_in[1], py, m_in[2] * pw));
            data->y[i] = std::fmaf(m_in[3], px, std::fmaf(m_in[4], py, m_in[5] * pw));
            data->weight[i] = std::fmaf(m_in[6], px, std::fmaf(m_in[7], py, m_in[8] * pw));
        }
        Path p;
        p.inner = data;
        return p;
    }
};

// path.rs:776-925
struct PathBuilder {
    std::shared_ptr<PathData> inner = std::make_shared<PathData>();

    void push(float px, float py, float w) {
        inner->x.push_back(px);
        inner->y.push_back(py);
        inner->weight.push_back(w);
    }
    void move_to(Point p) {
        PathData& d = *inner;
        size_t len = d.x.size();
        if (d.commands.back() == kMove) {
            d.x[len - 1] = p.x;
            d.y[len - 1] = p.y;
            d.weight[len - 1] = 1.0f;
        } else {
            d.close();
            size_t open = d.x.size();
            push(p.x, p.y, 1.0f);
            d.commands.push_back(kMove);
            d.open_point_index = open;
        }
    }
    void line_to(Point p) {
        push(p.x, p.y, 1.0f);
        inner->commands.push_back(kLine);
    }
    void quad_to(Point p1, Point p2) {
        push(p1.x, p1.y, 1.0f);
        push(p2.x, p2.y, 1.0f);
        inner->commands.push_back(kQuad);
    }
    void cubic_to(Point p1, Point p2, Point p3) {
        push(p1.x, p1.y, 1.0f);
        push(p2.x, p2.y, 1.0f);
        push(p3.x, p3.y, 1.0f);
        inner->commands.push_back(kCubic);
    }
    void rat_quad_to(Point p1, Point p2, float w) {
        push(p1.x * w, p1.y * w, w);
        push(p2.x, p2.y, 1.0f);
        inner->commands.push_back(kQuad);
    }
    void rat_cubic_to(Point p1, Point p2, Point p3, float w1, float w2) {
        push(p1.x * w1, p1.y * w1, w1);
        push(p2.x * w2, p2.y * w2, w2);
        push(p3.x, p3.y, 1.0f);
        inner->commands.push_back(kCubic);
    }
    Path build() {
        inner->close();
        Path p;
        p.inner = inner;
        return p;
    }
};

}  // namespace fo
