"""Deterministic synthetic scenes shared by the parity tests and bench.py.

One splitmix64 generator drives everything so the CUDA library and the oracle
see bit-identical inputs: the reference's own demos use
Rust's StdRng, which cannot be reproduced here and need not be.
"""
from __future__ import annotations

import numpy as np

from forma_b200.binding import (BlendMode, Color, Fill, FillRule, Func, GradientBuilder, GradientType, Point,
                                Props, Style)

MASK = (1 << 64) - 1


class SplitMix64:
    def __init__(self, seed: int):
        self.s = seed & MASK

    def next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & MASK
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK
        return z ^ (z >> 31)

    def uniform(self, lo: float = 0.0, hi: float = 1.0) -> float:
        """float32-representable uniform in [lo, hi)."""
        u = (self.next() >> 40) / float(1 << 24)
        return float(np.float32(lo + (hi - lo) * u))

    def randint(self, n: int) -> int:
        return self.next() % n


def f32(v) -> float:
    return float(np.float32(v))


def splitmix_uniform_block(seed: int, n: int) -> np.ndarray:
    """The first n `SplitMix64(seed).uniform()` values (u in [0, 1), float64), vectorised:
    the k-th output only depends on seed + k * gamma."""
    with np.errstate(over="ignore"):
        k = np.arange(1, n + 1, dtype=np.uint64)
        z = np.uint64(seed & MASK) + k * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(40)).astype(np.float64) / float(1 << 24)


def circle_path(api, cx, cy, r):
    """Rational-quad circle (shape of e2e-tests/tests/tests.rs:80-105)."""
    w = f32(np.sqrt(np.float32(2.0)) / np.float32(2.0))
    return (api.PathBuilder().move_to(Point(f32(cx + r), cy))
            .rat_quad_to(Point(f32(cx + r), f32(cy - r)), Point(cx, f32(cy - r)), w)
            .rat_quad_to(Point(f32(cx - r), f32(cy - r)), Point(f32(cx - r), cy), w)
            .rat_quad_to(Point(f32(cx - r), f32(cy + r)), Point(cx, f32(cy + r)), w)
            .rat_quad_to(Point(f32(cx + r), f32(cy + r)), Point(f32(cx + r), cy), w).build())


def random_cubics(api, comp, n_layers: int, width: int, height: int, seed: int, extent=(20.0, 200.0),
                  opaque: bool = True, margin: float = 0.0):
    """BASELINE config 3 shape: each layer is one closed cubic
    `move_to(p0) cubic_to(p1, p2, p3)` + implicit close, solid fill."""
    rng = SplitMix64(seed)
    for i in range(n_layers):
        cx = rng.uniform(-margin, width + margin)
        cy = rng.uniform(-margin, height + margin)
        e = rng.uniform(*extent)
        pts = [Point(f32(cx + rng.uniform(-e, e)), f32(cy + rng.uniform(-e, e))) for _ in range(4)]
        path = api.PathBuilder().move_to(pts[0]).cubic_to(pts[1], pts[2], pts[3]).build()
        a = 1.0 if opaque else rng.uniform(0.3, 1.0)
        col = Color(rng.uniform(), rng.uniform(), rng.uniform(), a)
        comp.get_mut_or_insert_default(i).insert(path).set_props(Props(func=Func.Draw(Style(fill=Fill.Solid(col)))))


def random_mixed(api, comp, n_layers: int, width: int, height: int, seed: int):
    """Mixed content for parity: lines/quads/cubics/rational curves, solid and
    gradient fills, all 12 separable blend modes, both fill rules, translucent
    colours, geometry partly off-screen (left/top/right/bottom)."""
    rng = SplitMix64(seed)
    separable = [BlendMode.Over, BlendMode.Multiply, BlendMode.Screen, BlendMode.Overlay, BlendMode.Darken,
                 BlendMode.Lighten, BlendMode.ColorDodge, BlendMode.ColorBurn, BlendMode.HardLight,
                 BlendMode.SoftLight, BlendMode.Difference, BlendMode.Exclusion]
    for i in range(n_layers):
        cx = rng.uniform(-0.1 * width, 1.1 * width)
        cy = rng.uniform(-0.1 * height, 1.1 * height)
        e = rng.uniform(4.0, 0.35 * min(width, height))
        kind = rng.randint(5)
        pb = api.PathBuilder()

        def pt():
            return Point(f32(cx + rng.uniform(-e, e)), f32(cy + rng.uniform(-e, e)))
        if kind == 0:      # polygon
            pb.move_to(pt())
            for _ in range(3 + rng.randint(4)):
                pb.line_to(pt())
            path = pb.build()
        elif kind == 1:    # quads
            pb.move_to(pt())
            for _ in range(2 + rng.randint(3)):
                pb.quad_to(pt(), pt())
            path = pb.build()
        elif kind == 2:    # cubics, two contours
            pb.move_to(pt()).cubic_to(pt(), pt(), pt())
            pb.move_to(pt()).cubic_to(pt(), pt(), pt()).line_to(pt())
            path = pb.build()
        elif kind == 3:    # circle
            path = circle_path(api, cx, cy, f32(e * 0.5))
        else:              # rational cubic + axis-aligned rectangle (integer coordinates)
            pb.move_to(pt()).rat_cubic_to(pt(), pt(), pt(), rng.uniform(0.3, 2.0), rng.uniform(0.3, 2.0))
            x0, y0 = float(int(cx)), float(int(cy))
            pb.move_to(Point(x0, y0)).line_to(Point(x0, y0 + 7.0)).line_to(Point(x0 + 9.0, y0 + 7.0)) \
              .line_to(Point(x0 + 9.0, y0))
            path = pb.build()
        a = 1.0 if rng.randint(3) == 0 else rng.uniform(0.2, 1.0)
        col = Color(rng.uniform(), rng.uniform(), rng.uniform(), a)
        if rng.randint(3) == 0:
            gb = GradientBuilder(Point(f32(cx - e), f32(cy - e)), Point(f32(cx + e), f32(cy + 0.5 * e)))
            if rng.randint(2):
                gb.type(GradientType.Radial)
            for _ in range(2 + rng.randint(3)):
                gb.color(Color(rng.uniform(), rng.uniform(), rng.uniform(), rng.uniform(0.4, 1.0)))
            fill = Fill.Gradient(gb.build())
        else:
            fill = Fill.Solid(col)
        style = Style(fill=fill, blend_mode=separable[rng.randint(len(separable))] if rng.randint(2) else BlendMode.Over)
        rule = FillRule.EvenOdd if rng.randint(4) == 0 else FillRule.NonZero
        comp.get_mut_or_insert_default(i).insert(path).set_props(Props(fill_rule=rule, func=Func.Draw(style)))


def random_circles(api, comp, n_layers: int, width: int, height: int, seed: int, r=(4.0, 40.0)):
    """BASELINE config 5 shape: rational-quad circles, 3-stop radial gradients
    centred on the circle, blend mode = layer index mod 8 over separable modes.
    Same random stream as 13 `SplitMix64.uniform` calls per circle (cx, cy, radius, alpha,
    3 x rgb), drawn in one vectorised block: a 1 M-circle scene is built in seconds."""
    modes = [BlendMode.Over, BlendMode.Multiply, BlendMode.Screen, BlendMode.Overlay, BlendMode.Darken,
             BlendMode.Lighten, BlendMode.HardLight, BlendMode.Difference]
    u = splitmix_uniform_block(seed, 13 * n_layers).reshape(n_layers, 13)

    def rounded(v):  # float64 arithmetic, one rounding to f32 — what SplitMix64.uniform / f32() do
        return v.astype(np.float32).astype(np.float64)
    cx = rounded(0.0 + (float(width) - 0.0) * u[:, 0])
    cy = rounded(0.0 + (float(height) - 0.0) * u[:, 1])
    rad = rounded(r[0] + (r[1] - r[0]) * u[:, 2])
    alpha = rounded(0.3 + (1.0 - 0.3) * u[:, 3])
    rgb = rounded(u[:, 4:13])
    xp, xm, yp, ym = rounded(cx + rad), rounded(cx - rad), rounded(cy + rad), rounded(cy - rad)
    w = f32(np.sqrt(np.float32(2.0)) / np.float32(2.0))
    cols = np.stack([cx, cy, xp, xm, yp, ym, alpha], axis=1).tolist()
    rgb = rgb.tolist()
    for i in range(n_layers):
        x, y, x1, x0, y1, y0, a = cols[i]
        c = rgb[i]
        path = (api.PathBuilder().move_to(Point(x1, y))
                .rat_quad_to(Point(x1, y0), Point(x, y0), w)
                .rat_quad_to(Point(x0, y0), Point(x0, y), w)
                .rat_quad_to(Point(x0, y1), Point(x, y1), w)
                .rat_quad_to(Point(x1, y1), Point(x1, y), w).build())
        gb = GradientBuilder(Point(x, y), Point(x1, y)).type(GradientType.Radial)
        gb.color(Color(c[0], c[1], c[2], a))
        gb.color(Color(c[3], c[4], c[5], a))
        gb.color(Color(c[6], c[7], c[8], a))
        comp.get_mut_or_insert_default(i).insert(path).set_props(
            Props(func=Func.Draw(Style(fill=Fill.Gradient(gb.build()), blend_mode=modes[i % 8]))))


def random_circles_scalar(api, comp, n_layers: int, width: int, height: int, seed: int, r=(4.0, 40.0)):
    """The same scene drawn call by call (pins the vectorised generator in the CPU tests)."""
    rng = SplitMix64(seed)
    modes = [BlendMode.Over, BlendMode.Multiply, BlendMode.Screen, BlendMode.Overlay, BlendMode.Darken,
             BlendMode.Lighten, BlendMode.HardLight, BlendMode.Difference]
    for i in range(n_layers):
        cx, cy = rng.uniform(0.0, width), rng.uniform(0.0, height)
        rad = rng.uniform(*r)
        gb = GradientBuilder(Point(cx, cy), Point(f32(cx + rad), cy)).type(GradientType.Radial)
        a = rng.uniform(0.3, 1.0)
        for _ in range(3):
            gb.color(Color(rng.uniform(), rng.uniform(), rng.uniform(), a))
        comp.get_mut_or_insert_default(i).insert(circle_path(api, cx, cy, rad)).set_props(
            Props(func=Func.Draw(Style(fill=Fill.Gradient(gb.build()), blend_mode=modes[i % 8]))))


def spaceship_scene(api, comp, n_asteroids: int, width: int, height: int, seed: int):
    """BASELINE config 4 shape (demo/src/demos/spaceship.rs): an opaque backdrop, one
    small ship and `n_asteroids` potato-like closed quad paths drifting across the
    screen with a fixed dt; only the moving layers' transforms change per frame, so
    a layer cache skips most tiles. Returns `animate(frame)` which moves the layers."""
    rng = SplitMix64(seed)
    backdrop = (api.PathBuilder().move_to(Point(0.0, 0.0)).line_to(Point(float(width), 0.0))
                .line_to(Point(float(width), float(height))).line_to(Point(0.0, float(height))).build())
    comp.get_mut_or_insert_default(0).insert(backdrop).set_props(
        Props(func=Func.Draw(Style(fill=Fill.Solid(Color(0.02, 0.02, 0.05, 1.0))))))
    movers = []
    for i in range(n_asteroids + 1):
        ship = i == n_asteroids
        rad = 14.0 if ship else rng.uniform(8.0, 60.0)
        k = 3 if ship else 6 + rng.randint(5)
        pts = []
        for j in range(k):
            ang = 2.0 * np.pi * j / k
            rr = rad * (1.0 if ship else rng.uniform(0.7, 1.3))
            pts.append((f32(rr * np.cos(ang)), f32(rr * np.sin(ang))))
        pb = api.PathBuilder().move_to(Point(*pts[0]))
        for j in range(k):
            a, b = pts[j], pts[(j + 1) % k]
            ctrl = Point(f32((a[0] + b[0]) * 0.6), f32((a[1] + b[1]) * 0.6))
            if ship:
                pb.line_to(Point(*b))
            else:
                pb.quad_to(ctrl, Point(*b))
        g = rng.uniform(0.3, 0.8)
        col = Color(1.0, 0.9, 0.2, 1.0) if ship else Color(g, f32(g * 0.9), f32(g * 0.8), 1.0)
        layer = comp.get_mut_or_insert_default(1 + i)
        layer.insert(pb.build()).set_props(Props(func=Func.Draw(Style(fill=Fill.Solid(col)))))
        movers.append((1 + i, rng.uniform(0.0, width), rng.uniform(0.0, height),
                       rng.uniform(-120.0, 120.0), rng.uniform(-120.0, 120.0)))

    def animate(frame: int, dt: float = 1.0 / 60.0):
        for order, x0, y0, vx, vy in movers:
            x = (x0 + vx * dt * frame) % width
            y = (y0 + vy * dt * frame) % height
            comp.get(order).set_transform([1.0, 0.0, 0.0, 1.0, f32(x), f32(y)])
    animate(0)
    return animate
