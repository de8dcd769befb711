"""CPU-only checks of the product library: the C ABI surface and the host-side
logic that needs no device. No compute entry point is called here."""
import ctypes
import os
import re

import numpy as np
import pytest

import forma_b200
from forma_b200 import bands
from forma_b200.binding import (Color, Fill, FormaError, Func, GeomPresTransformError, GradientBuilder, OrderError, Point,
                                Props, SIGNATURES, Style)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "forma_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(forma_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(forma_b200.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_binding_covers_the_header():
    bound = {"forma_" + n for n in SIGNATURES}
    assert set(declared_symbols()) <= bound, sorted(set(declared_symbols()) - bound)


def test_no_oracle_dependency_in_product():
    """The product must not import, link or open anything under oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "forma_b200")):
        if os.sep + "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle/" not in text.replace("oracle/oracle.py, prefix", "") or f == "binding.py", (dirpath, f)
                assert "libforma_oracle" not in text and "import oracle" not in text and "from oracle" not in text, (dirpath, f)
    out = os.popen(f"ldd {forma_b200.LIB_PATH}").read()
    assert "oracle" not in out


def test_host_bookkeeping_without_device():
    api = forma_b200.load()
    comp = api.Composition()
    assert comp.is_empty() and len(comp) == 0
    layer = comp.get_mut_or_insert_default(0)  # composition/mod.rs:458-470 composition_len
    assert len(comp) == 1 and layer.is_enabled()
    path = api.PathBuilder().move_to(Point(0, 0)).line_to(Point(4, 0)).line_to(Point(4, 4)).build()
    g0 = layer.geom_id()
    layer.insert(path)
    assert layer.geom_id() == g0
    assert comp.point_count() == 4  # closed triangle: 3 corners + return to start
    layer.clear()
    assert layer.geom_id() != g0    # composition/mod.rs:972-1001 geom_id
    with pytest.raises(OrderError):
        comp.get_mut_or_insert_default(1 << 21)
    with pytest.raises(GeomPresTransformError):
        layer.set_transform([1.0, 0.0, 0.0, 1.5, 0.0, 0.0])
    layer.set_transform([1.0, 0.0, 0.0, 1.0, 3.0, 4.0])
    detached = comp.create_layer()
    assert comp.insert(0, detached) is not None and len(comp) == 1
    assert comp.remove(0) is not None and comp.remove(0) is None and comp.is_empty()
    with pytest.raises(FormaError):  # GradientBuilder::build -> None for < 2 stops
        from forma_b200.binding import Gradient
        comp.get_mut_or_insert_default(1).set_props(
            Props(func=Func.Draw(Style(fill=Fill.Gradient(Gradient(0, Point(0, 0), Point(1, 1), [(Color(), -1.0)]))))))
    assert GradientBuilder(Point(0, 0), Point(1, 0)).color(Color()).build() is None


def test_layer_bookkeeping_scales():
    """Dropping a layer (Drop for Layer, composition/layer.rs:355-363) is a lookup, not a
    scan: 60 k layers are created, attached, partly displaced and dropped in well under a
    second (paris-30k alone has 50 620 layers)."""
    import time
    api = forma_b200.load()
    comp = api.Composition()
    t0 = time.perf_counter()
    layers = [comp.create_layer() for _ in range(60_000)]
    for i, layer in enumerate(layers[:40_000]):
        assert comp.insert(i, layer) is None
    displaced = comp.insert(7, layers[50_000])           # 7 is taken: its layer comes back detached
    assert displaced is not None and len(comp) == 40_000
    displaced.drop()                                     # detached layer with a stale order: must not remove order 7
    assert len(comp) == 40_000 and comp.get(7) is not None
    for layer in layers[40_000:50_000] + layers[50_001:]:
        layer.drop()                                     # never attached
    assert len(comp) == 40_000
    for i, layer in enumerate(layers[:40_000]):
        if i != 7:
            layer.drop()
    assert len(comp) == 1
    comp.get(7).drop()
    assert comp.is_empty()
    assert time.perf_counter() - t0 < 10.0  # ~0.2 s here; a scan per drop is quadratic (tens of seconds at this size)


def test_renderer_requires_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(FormaError, match="no CPU fallback"):
        forma_b200.Renderer(0)
    # inspection of flattened points also runs on the device only
    api = forma_b200.load()
    path = api.PathBuilder().move_to(Point(0, 0)).quad_to(Point(5, 9), Point(10, 0)).build()
    with pytest.raises(FormaError):
        path.segments()
    # the multi-device renderer fails the same way, and refuses empty / oversized device lists before looking for devices
    with pytest.raises(FormaError, match="no CPU fallback"):
        api.MultiRenderer([0, 1])
    with pytest.raises(FormaError, match="1..64"):
        api.MultiRenderer([])


def test_flatten_program_matches_the_oracles_point_count(oracle_api):
    """Host half of the flattener (spline merge, subdivision counts, encoding choice):
    the program of every path announces exactly as many output points and contour
    ends as the oracle's flattener produces, and keeps the smaller encoding."""
    import synth
    api = forma_b200.load()
    rng_a, rng_b = synth.SplitMix64(21), synth.SplitMix64(21)

    def random_path(a, rng):
        pb = a.PathBuilder()

        def pt():
            return Point(rng.uniform(-200.0, 200.0), rng.uniform(-200.0, 200.0))
        pb.move_to(pt())
        for _ in range(1 + rng.randint(12)):
            k = rng.randint(6)
            if k <= 1:
                pb.line_to(pt())
            elif k == 2:
                pb.quad_to(pt(), pt())
            elif k == 3:
                pb.cubic_to(pt(), pt(), pt())
            elif k == 4:
                pb.rat_quad_to(pt(), pt(), rng.uniform(0.3, 2.5))
            else:
                pb.move_to(pt())
        return pb.build()

    seen = set()
    for i in range(300):
        p, q = random_path(api, rng_a), random_path(oracle_api, rng_b)
        s = p.program_stats()
        x, y, c = q.segments()
        assert s["points"] == len(x), f"path {i}"
        assert s["contour_ends"] == int(c[:-1].sum()) if len(c) else s["contour_ends"] == 0, f"path {i}"
        assert (s["splines"] == 0) != (s["point_records"] == 0) or s["points"] == 0, f"path {i}: both / no encodings"
        if s["point_records"]:
            assert s["point_records"] == s["points"]
        seen.add("points" if s["point_records"] else "splines")
        seen.add("rational" if s["rational"] else "polynomial")
    assert seen == {"points", "splines", "rational", "polynomial"}


@pytest.mark.parametrize("height,world", [(2160, 1), (2160, 2), (2160, 8), (4320, 8), (1080, 4), (17, 8), (16, 3)])
def test_bands_partition_the_frame(height, world):
    covered = np.zeros(height, np.int32)
    per = None
    for rank in range(world):
        b = bands.band_of(height, world, rank)
        covered[b.y0:b.y1] += 1
        assert b.y0 % 16 == 0 and (b.y1 % 16 == 0 or b.y1 == height)
        assert b.padded_height >= height and b.padded_height == world * b.rows_per_band * 16
        per = b.rows_per_band if per is None else per
        assert b.rows_per_band == per
    assert np.all(covered == 1)


def test_schedule_switches_round_trip_without_device():
    """forma_set_option / forma_get_option are host-side state: every switch DESIGN.md lists
    exists, keeps its documented default, takes its alternatives and refuses values outside its
    range (no device needed)."""
    from forma_b200.binding import FormaError
    api = forma_b200.load()
    defaults = {"speculate": 1, "sync_free": 1, "band_copy": 1, "copy_bands": 4, "host_slices": 1, "slice_bands": 2,
                "slice_chain": 1, "slice_min_points": 65536, "sort_full_key": 0, "sort_big_log2": 19, "sort_scan_log2": 22,
                "paint_lpt": 1, "paint_wide": 0, "band_filter": 1, "test_gap_cap": 0, "test_fast_shrink": 0}
    env_overrides = {k for k in defaults if os.environ.get("FORMA_" + k.upper()) is not None}
    for name, want in defaults.items():
        if name not in env_overrides:
            assert api.get_option(name) == want, name
    design = open(os.path.join(ROOT, "DESIGN.md")).read()
    for name in defaults:
        if not name.startswith("test_"):
            assert f"`{name}`" in design, f"{name} is not documented in DESIGN.md"
    for name, alt in (("host_slices", 16), ("slice_bands", 16), ("copy_bands", 1), ("slice_chain", 0), ("sync_free", 0)):
        saved = api.get_option(name)
        api.set_option(name, alt)
        assert api.get_option(name) == alt
        api.set_option(name, saved)
    for name, bad in (("host_slices", 0), ("host_slices", 17), ("slice_bands", 0), ("copy_bands", 17), ("sync_free", 2)):
        with pytest.raises(FormaError):
            api.set_option(name, bad)
