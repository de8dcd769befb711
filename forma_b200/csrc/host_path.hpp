// forma_b200 host side — PathBuilder / Path and the *flatten program*.
//
// Mirrors forma/src/path.rs. The reference flattens a path in two steps:
//   (1) a SERIAL walk over the path's commands that merges tangent-continuous
//       lines / quadratics into splines (path.rs:208-445), and
//   (2) a PARALLEL evaluation, one work item per output point (path.rs:487-534).
// Here (1) runs on the host once per PathData (cached like the reference's
// `segments: Option<Segments>`, path.rs:581,617-654) and produces a compact
// "flatten program"; (2) runs on the GPU (flatten_eval_kernel, kernels.cu)
// every time the path is inserted into a layer, writing straight into the
// composition's HBM-resident segment buffer.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include "device_types.h"

namespace forma {

struct Pt {
    float x, y;
};

// Separately rounded IEEE ops: this file is compiled with -ffp-contract=off
// (host) and every reference `mul_add` is an explicit fmaf.
inline float rcp(float v) { return 1.0f / v; }
inline float mix(float t, float a, float b) { return fmaf(t, b, fmaf(-t, a, a)); }  // path.rs:44-46

class PathData {
   public:
    // path.rs:657-668 — a fresh path starts with one Move at the origin.
    std::vector<float> x{0.0f}, y{0.0f}, w{1.0f};
    std::vector<uint8_t> cmd{0};  // 0 Move, 1 Line, 2 Quad, 3 Cubic
    size_t open_index = 0;

    void close();                        // path.rs:596-615
    const FlattenProgram& program();     // path.rs:617-654 (host part)

   private:
    bool built_ = false;
    FlattenProgram prog_;
};

struct Path {
    std::shared_ptr<PathData> data = std::make_shared<PathData>();
    bool has_xf = false;
    float xf[6] = {1, 0, 0, 1, 0, 0};  // ux, uy, vx, vy, tx, ty

    Path transformed(const float m[9]) const;  // path.rs:726-765
};

class PathBuilder {
   public:
    std::shared_ptr<PathData> data = std::make_shared<PathData>();
    void move_to(Pt p);
    void line_to(Pt p);
    void quad_to(Pt p1, Pt p2);
    void cubic_to(Pt p1, Pt p2, Pt p3);
    void rat_quad_to(Pt p1, Pt p2, float weight);
    void rat_cubic_to(Pt p1, Pt p2, Pt p3, float w1, float w2);
    Path build();
};

// GeomPresTransform::try_from (math/transform.rs:208-221).
bool geom_pres_ok(float ux, float uy, float vx, float vy);

}  // namespace forma
