"""The BASELINE.json configurations as scenes, shared by bench.py and the parity tests
(tests/test_gpu_bench_scale.py renders every one of them on the device and on the oracle
and compares the frames byte for byte, like the reference's own harness does per scene,
/root/reference/e2e-tests/tests/test_env.rs:262-290).

`build_scene(api, name)` works with the product binding and with the oracle binding, so both
renderers see the same Composition.
"""
from __future__ import annotations

import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKLOADS = {
    # name: (width, height, description)
    "circle256": (256, 256, "BASELINE config 1: one solid rational-quad circle r=100 at (128,128), 256x256"),
    "paris4k": (3840, 2160, "paris-30k.svg (50620 layers, solid fills) scaled 2160/1060 at 3840x2160"),
    "cubics100k": (3840, 2160, "100k random closed cubics, opaque solid fills, seed 3, 3840x2160"),
    "paris4k_grad": (3840, 2160, "paris-30k.svg, every 8th layer filled with a synthetic 3-stop linear gradient over its "
                                 "bounding box (the file itself has none; SURVEY.md C2 variant, seed 1), 3840x2160"),
    "circles8k": (7680, 4320, "200k rational-quad circles r in [4,40], radial gradients, 8 blend modes, seed 5, 7680x4320"),
    "spaceship1080p": (1920, 1080, "spaceship-like animation (backdrop + 1 ship + 400 drifting asteroids, seed 43, dt = 1/60 s), "
                                   "1920x1080, persistent layer cache (per-tile damage reuse), every step = next frame"),
    "circles8k_1m": (7680, 4320, "1M rational-quad circles r in [4,40], radial gradients, 8 blend modes, seed 5, 7680x4320"),
    "smoke": (640, 360, "400 mixed layers, 640x360 (plumbing check)"),
}


def build_scene(api, name):
    """-> (composition, width, height). Animated workloads attach `comp.animate(frame)`."""
    import synth
    from forma_b200 import svg
    from forma_b200.binding import Color, Fill, Func, GradientBuilder, Point, Props, Style
    comp = api.Composition()
    w, h, _ = WORKLOADS[name]
    if name == "circle256":
        comp.get_mut_or_insert_default(0).insert(synth.circle_path(api, 128.0, 128.0, 100.0)).set_props(
            Props(func=Func.Draw(Style(fill=Fill.Solid(Color(1.0, 0.0, 0.0, 1.0))))))
    elif name == "paris4k":
        paths = svg.PathList.load(os.path.join(ROOT, "tests", "data", "paris30k_paths.npz"))
        svg.compose(api, comp, paths, scale=2160.0 / 1060.0)
    elif name == "cubics100k":
        synth.random_cubics(api, comp, 100_000, w, h, 3)
    elif name == "paris4k_grad":
        paths = svg.PathList.load(os.path.join(ROOT, "tests", "data", "paris30k_paths.npz"))
        scale = 2160.0 / 1060.0
        rng = synth.SplitMix64(1)

        def fill_of(i, color):
            if i % 8 != 7:
                return Fill.Solid(color)
            p = paths.pts[int(paths.pt_off[i]):int(paths.pt_off[i + 1])].reshape(-1, 2) * scale
            lo, hi = p.min(axis=0), p.max(axis=0)
            gb = GradientBuilder(Point(float(lo[0]), float(lo[1])), Point(float(hi[0]), float(hi[1])))
            gb.color(color)
            gb.color(Color(rng.uniform(), rng.uniform(), rng.uniform(), color.a))
            gb.color(color)
            return Fill.Gradient(gb.build())
        svg.compose(api, comp, paths, scale=scale, fill_of=fill_of)
    elif name == "circles8k":
        synth.random_circles(api, comp, 200_000, w, h, 5)
    elif name == "circles8k_1m":
        synth.random_circles(api, comp, 1_000_000, w, h, 5)
    elif name == "spaceship1080p":
        comp.animate = synth.spaceship_scene(api, comp, 400, w, h, 43)  # animate(frame) moves the layers
    else:
        synth.random_mixed(api, comp, 400, w, h, 7)
    return comp, w, h
