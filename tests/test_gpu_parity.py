"""Parity of the CUDA path against the CPU oracle, through the C ABI.

Bar (north star): bit-exact for integer / byte / index work and for solid
fills; the oracle follows the portable-SIMD semantics, and the CUDA kernels
keep the same operation order, so gradients and blends are expected — and
asserted — to be bit-exact as well (tolerance 0).
"""
import numpy as np
import pytest

import scenes
import synth
from forma_b200.binding import (BGRA, RGB1, RGBA, BlendMode, Color, Fill, FillRule, Func, Point, Props, Rect,
                                Style)

pytestmark = pytest.mark.gpu

GOLD = None


def gold():
    global GOLD
    if GOLD is None:
        import os
        GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "e2e_expected.npz"))
    return GOLD


def render(api, renderer, build, w, h, channels=RGBA, clear=scenes.E2E_CLEAR, crop=None, stride=None, fill=0):
    comp = api.Composition()
    build(api, comp)
    stride = w * 4 if stride is None else stride
    buf = np.full(h * stride, fill, np.uint8)
    t = renderer.render(comp, buf, w, h, channels, clear, crop, None, stride)
    return buf.reshape(h, stride), t


def assert_same(a, b, what):
    if not np.array_equal(a, b):
        d = np.abs(a.astype(int) - b.astype(int))
        ys, xs = np.nonzero(d)
        raise AssertionError(f"{what}: {len(ys)} bytes differ, max diff {d.max()}, first at (row {ys[0]}, byte {xs[0]})")


# --- stage 1: flatten ----------------------------------------------------------------
def _paths(api):
    rng = synth.SplitMix64(11)
    out = []
    for i in range(200):
        pb = api.PathBuilder()

        def pt(scale=100.0):
            return Point(rng.uniform(-scale, scale), rng.uniform(-scale, scale))
        pb.move_to(pt())
        for _ in range(1 + rng.randint(4)):
            k = rng.randint(5)
            if k == 0:
                pb.line_to(pt())
            elif k == 1:
                pb.quad_to(pt(), pt())
            elif k == 2:
                pb.cubic_to(pt(), pt(), pt())
            elif k == 3:
                pb.rat_quad_to(pt(), pt(), rng.uniform(0.2, 3.0))
            else:
                pb.rat_cubic_to(pt(), pt(), pt(), rng.uniform(0.2, 3.0), rng.uniform(0.2, 3.0))
            if rng.randint(6) == 0:
                pb.move_to(pt())
        out.append(pb.build())
    out.append(synth.circle_path(api, 128.0, 128.0, 100.0))
    out.append(api.PathBuilder().build())  # empty path
    return out


def test_flatten_points_bit_exact(cuda_api, oracle_api, cuda_renderer):
    for i, (pc, po) in enumerate(zip(_paths(cuda_api), _paths(oracle_api))):
        xc, yc, cc = pc.segments()
        xo, yo, co = po.segments()
        assert xc.shape == xo.shape, f"path {i}: {xc.shape} vs {xo.shape} points"
        assert np.array_equal(xc.view(np.uint32), xo.view(np.uint32)), f"path {i}: x differs"
        assert np.array_equal(yc.view(np.uint32), yo.view(np.uint32)), f"path {i}: y differs"
        assert np.array_equal(cc, co), f"path {i}: contour flags differ"
        # Path::transform: geometry-preserving (shared data) and projective (re-flatten)
        for m in ([1.0, 0.0, 5.0, 0.0, 1.0, 20.0, 0.0, 0.0, 1.0], [2.0, 0.0, 0.0, 0.0, 2.0, 0.0, 0.0, 0.0, 1.0],
                  [1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.001, 0.0, 1.0]):
            xc, yc, cc = pc.transform(m).segments()
            xo, yo, co = po.transform(m).segments()
            assert np.array_equal(xc.view(np.uint32), xo.view(np.uint32)) and np.array_equal(
                yc.view(np.uint32), yo.view(np.uint32)) and np.array_equal(cc, co), f"path {i} transform {m}"


# --- stages 1½ + 2: pixel-grid intersection -------------------------------------------
def test_rasterize_unsorted_segments_identical(cuda_api, oracle_api, cuda_renderer, oracle_renderer):
    """Same u64 words in the same (line, k) emission order as the reference."""
    for seed, n, w, h in [(1, 64, 256, 256), (2, 400, 640, 480), (3, 1500, 1920, 1080)]:
        res = []
        for api, r in ((cuda_api, cuda_renderer), (oracle_api, oracle_renderer)):
            comp = api.Composition()
            synth.random_mixed(api, comp, n, w, h, seed)
            res.append(r.rasterize_only(comp, w, h))
        assert res[0].shape == res[1].shape, f"seed {seed}: {res[0].size} vs {res[1].size} segments"
        assert np.array_equal(res[0], res[1]), f"seed {seed}: segments differ"


def test_rasterize_axis_aligned_and_degenerate(cuda_api, oracle_api, cuda_renderer, oracle_renderer):
    def build(api, comp):
        pb = api.PathBuilder()
        # vertical / horizontal / 45 degree / tiny / off-grid / far off-screen lines
        pts = [(0, 0), (0, 16), (16, 16), (16, 0), (3.5, 0.25), (4.0, 2.0), (-40.0, 5.0), (700.0, 9.0), (8.0, -30.0),
               (8.0, 300.0), (0.5, 0.5), (0.5000001, 0.5000001), (1e-3, 2e-3), (100.0, 100.0)]
        pb.move_to(Point(*pts[0]))
        for p in pts[1:]:
            pb.line_to(Point(*p))
        comp.get_mut_or_insert_default(7).insert(pb.build())
    res = []
    for api, r in ((cuda_api, cuda_renderer), (oracle_api, oracle_renderer)):
        comp = api.Composition()
        build(api, comp)
        res.append(r.rasterize_only(comp, 256, 256))
    assert np.array_equal(res[0], res[1])


# --- stage 3: sort ---------------------------------------------------------------------
# 4 194 303 / 4 194 305: either side of the single-sweep / reduce-then-scan switch (sort_scan_log2 = 22);
# 40 M: the bench's sort size (100 k cubics), all 44 key bits incl. the full 21-bit layer field in play.
@pytest.mark.parametrize("n", [0, 1, 2, 31, 4095, 4096, 4097, 100_000, 3_000_001, 4_194_303, 4_194_305, 40_000_000])
def test_sort_u64_stable_on_top_44_bits(cuda_renderer, n):
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 1 << 63, size=n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=n, dtype=np.uint64)
    if n > 100:
        # realistic skew: few tile rows / layers, many duplicates of the 44-bit key
        keys[: n // 2] = (keys[: n // 2] & np.uint64(0x000FFFFF)) | (rng.integers(0, 64, size=n // 2, dtype=np.uint64) << np.uint64(41))
    out = cuda_renderer.sort_u64(keys)
    order = np.argsort(keys >> np.uint64(20), kind="stable")
    assert np.array_equal(out, keys[order])  # LSD radix is stable: a legal (and unique) outcome


def test_pipeline_sorted_segments_match_oracle(cuda_api, oracle_api, cuda_renderer, oracle_renderer, unsliced):
    res = []
    for api, r in ((cuda_api, cuda_renderer), (oracle_api, oracle_renderer)):
        comp = api.Composition()
        synth.random_mixed(api, comp, 800, 1024, 768, 5)
        buf = np.zeros(1024 * 768 * 4, np.uint8)
        r.render(comp, buf, 1024, 768, RGBA, Color(1, 1, 1, 1))
        res.append(r.segments())
    a, b = res
    assert a.size == b.size
    ka = a >> np.uint64(20)
    assert np.all(ka[1:] >= ka[:-1]), "CUDA output is not sorted on bits [20, 64)"
    # crumsort is unstable: compare as multisets (full u64 order is a canonical form of both)
    assert np.array_equal(np.sort(a), np.sort(b))


# --- stage 4 + full pipeline ------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(scenes.E2E))
def test_e2e_scene_matches_oracle_and_golden(cuda_api, oracle_api, cuda_renderer, oracle_renderer, name):
    img_c = scenes.render_e2e(cuda_api, name, cuda_renderer)
    img_o = scenes.render_e2e(oracle_api, name, oracle_renderer)
    assert_same(img_c.reshape(64, -1), img_o.reshape(64, -1), f"{name} vs oracle")
    g = gold()[name + "__cpu"]
    if name in scenes.NON_SEPARABLE:
        # goldens were rendered through Arm's 8-bit reciprocal estimate (tests/test_oracle_pinned.py)
        assert np.abs(img_c.astype(int) - g.astype(int)).max() <= 8  # reference tolerance, test_env.rs:278
    else:
        assert_same(img_c.reshape(64, -1), g.reshape(64, -1), f"{name} vs reference golden")


@pytest.mark.parametrize("seed,n,w,h", [(21, 60, 200, 120), (22, 300, 513, 259), (23, 1200, 1280, 720),
                                        (24, 4000, 1920, 1080)])
def test_random_mixed_scene_bit_exact(cuda_api, oracle_api, cuda_renderer, oracle_renderer, seed, n, w, h):
    def build(api, comp):
        synth.random_mixed(api, comp, n, w, h, seed)
    a, _ = render(cuda_api, cuda_renderer, build, w, h)
    b, _ = render(oracle_api, oracle_renderer, build, w, h)
    assert_same(a, b, f"random_mixed seed {seed}")


@pytest.mark.parametrize("order", ["descending", "shuffled"])
def test_layers_inserted_out_of_order(cuda_api, oracle_api, cuda_renderer, oracle_renderer, order):
    """Layer orders that do not follow insertion order: the sort cannot rely on the
    rasterizer's emission order for the layer digits (Composition::layers_in_order)."""
    w, h, n = 640, 480, 300

    def build(api, comp):
        rng = synth.SplitMix64(77)
        ids = list(range(n))
        if order == "descending":
            ids.reverse()
        else:
            for i in range(n - 1, 0, -1):
                j = rng.randint(i + 1)
                ids[i], ids[j] = ids[j], ids[i]
        for k in ids:
            cx, cy, e = rng.uniform(0, w), rng.uniform(0, h), rng.uniform(10, 120)
            pb = api.PathBuilder()
            pb.move_to(Point(synth.f32(cx - e), synth.f32(cy - e))).line_to(Point(synth.f32(cx + e), synth.f32(cy - 0.5 * e)))
            pb.quad_to(Point(synth.f32(cx + e), synth.f32(cy + e)), Point(synth.f32(cx), synth.f32(cy + e)))
            col = Color(rng.uniform(), rng.uniform(), rng.uniform(), rng.uniform(0.3, 1.0))
            comp.get_mut_or_insert_default(k * 3).insert(pb.build()).set_props(Props(func=Func.Draw(Style(fill=Fill.Solid(col)))))
    a, _ = render(cuda_api, cuda_renderer, build, w, h)
    b, _ = render(oracle_api, oracle_renderer, build, w, h)
    assert_same(a, b, f"layers inserted in {order} order")


def test_layer_orders_up_to_the_limit(cuda_api, oracle_api, cuda_renderer, oracle_renderer):
    """Orders spread over the whole 21-bit range, inserted high to low: the segment
    sort needs every layer digit (5 passes with the tile digits)."""
    w, h = 400, 300
    top = (1 << 21) - 1

    def build(api, comp):
        rng = synth.SplitMix64(5)
        for k in range(60):
            order = top - k * 34567 if k % 2 == 0 else k * 17
            cx, cy, r = rng.uniform(0, w), rng.uniform(0, h), rng.uniform(10, 90)
            col = Color(rng.uniform(), rng.uniform(), rng.uniform(), rng.uniform(0.4, 1.0))
            comp.get_mut_or_insert_default(order).insert(synth.circle_path(api, synth.f32(cx), synth.f32(cy), synth.f32(r))).set_props(
                Props(func=Func.Draw(Style(fill=Fill.Solid(col)))))
    a, _ = render(cuda_api, cuda_renderer, build, w, h)
    b, _ = render(oracle_api, oracle_renderer, build, w, h)
    assert_same(a, b, "orders up to LAYER_LIMIT")


def test_8k_frame_bit_exact(cuda_api, oracle_api, cuda_renderer, oracle_renderer):
    """7680x4320: 9-bit tile coordinates (three 6-bit sort passes), band-wise copy-back."""
    w, h = 7680, 4320

    def build(api, comp):
        synth.random_circles(api, comp, 300, w, h, 8, r=(20.0, 400.0))
    a, _ = render(cuda_api, cuda_renderer, build, w, h)
    b, _ = render(oracle_api, oracle_renderer, build, w, h)
    assert_same(a, b, "8K frame")


def test_opaque_cubics_bit_exact(cuda_api, oracle_api, cuda_renderer, oracle_renderer):
    def build(api, comp):
        synth.random_cubics(api, comp, 3000, 1920, 1080, 3)
    a, _ = render(cuda_api, cuda_renderer, build, 1920, 1080, clear=Color(1, 1, 1, 1))
    b, _ = render(oracle_api, oracle_renderer, build, 1920, 1080, clear=Color(1, 1, 1, 1))
    assert_same(a, b, "random opaque cubics")


def test_gradient_circles_blend_modes_bit_exact(cuda_api, oracle_api, cuda_renderer, oracle_renderer):
    def build(api, comp):
        synth.random_circles(api, comp, 5000, 1024, 1024, 5)
    a, _ = render(cuda_api, cuda_renderer, build, 1024, 1024)
    b, _ = render(oracle_api, oracle_renderer, build, 1024, 1024)
    assert_same(a, b, "radial-gradient circles x 8 blend modes")


def test_non_separable_blend_modes_bit_exact(cuda_api, oracle_api, cuda_renderer, oracle_renderer):
    for mode in (BlendMode.Hue, BlendMode.Saturation, BlendMode.Color, BlendMode.Luminosity):
        def build(api, comp):
            synth.random_circles(api, comp, 300, 256, 256, 9)
            for i in range(0, 300, 3):
                comp.get(i).set_props(Props(func=Func.Draw(Style(
                    fill=Fill.Solid(Color(synth.f32(0.1 + i / 400.0), 0.5, synth.f32(0.9 - i / 400.0), 0.7)),
                    blend_mode=mode))))
        a, _ = render(cuda_api, cuda_renderer, build, 256, 256)
        b, _ = render(oracle_api, oracle_renderer, build, 256, 256)
        assert_same(a, b, f"non-separable mode {mode}")


def test_empty_composition_clears(cuda_api, cuda_renderer):
    # composition/mod.rs:496-518 background_color_clear
    buf, _ = render(cuda_api, cuda_renderer, lambda api, comp: None, 1, 1, clear=Color(1, 0, 0, 1), fill=7)
    assert buf.tolist() == [[255, 0, 0, 255]]
    buf, _ = render(cuda_api, cuda_renderer, lambda api, comp: None, 37, 21, clear=Color(0, 1, 0, 1), fill=7)
    assert np.all(buf.reshape(21, 37, 4) == np.array([0, 255, 0, 255], np.uint8))


def test_channels_stride_and_partial_tiles(cuda_api, oracle_api, cuda_renderer, oracle_renderer):
    def build(api, comp):
        synth.random_mixed(api, comp, 150, 333, 211, 31)
    for channels in (BGRA, RGB1, (3, 2, 1, 0)):
        for clear in (Color(0.2, 0.3, 0.4, 0.5), Color(1, 1, 1, 1)):
            a, _ = render(cuda_api, cuda_renderer, build, 333, 211, channels, clear, stride=333 * 4 + 20, fill=0xAB)
            b, _ = render(oracle_api, oracle_renderer, build, 333, 211, channels, clear, stride=333 * 4 + 20, fill=0xAB)
            assert_same(a, b, f"channels {channels} clear {clear}")
            assert np.all(a[:, 333 * 4:] == 0xAB), "stride padding was written"


def test_crop_touches_only_cropped_tiles(cuda_api, oracle_api, cuda_renderer, oracle_renderer):
    def build(api, comp):
        synth.random_mixed(api, comp, 200, 400, 300, 41)
    crop = Rect((50, 230), (40, 170))
    a, _ = render(cuda_api, cuda_renderer, build, 400, 300, crop=crop, fill=0x5A)
    b, _ = render(oracle_api, oracle_renderer, build, 400, 300, crop=crop, fill=0x5A)
    assert_same(a, b, "crop")


def test_inserts_with_and_without_path_transforms_in_one_batch(cuda_api, oracle_api, cuda_renderer, oracle_renderer):
    """Path::transform with a transform that preserves geometry (rotation / translation / shrink:
    math/transform.rs:161-221) keeps the path's data and applies the transform to the flattened
    points at insert time (path.rs:689-706). One upload batch mixes such inserts with plain ones
    and with up-scaled paths (re-flattened): every insert must pick its own transform."""
    import math
    w, h = 640, 400
    outs = []
    for api, r in ((cuda_api, cuda_renderer), (oracle_api, oracle_renderer)):
        comp = api.Composition()
        rng = synth.SplitMix64(23)
        for i in range(120):
            cx, cy, e = rng.uniform(60, w - 60), rng.uniform(60, h - 60), rng.uniform(10, 70)
            pb = api.PathBuilder().move_to(Point(synth.f32(cx - e), synth.f32(cy)))
            pb.cubic_to(Point(synth.f32(cx), synth.f32(cy - e)), Point(synth.f32(cx + e), synth.f32(cy - 0.3 * e)),
                        Point(synth.f32(cx + 0.6 * e), synth.f32(cy + e)))
            pb.quad_to(Point(synth.f32(cx), synth.f32(cy + 1.4 * e)), Point(synth.f32(cx - e), synth.f32(cy + 0.2 * e)))
            path = pb.build()
            kind = i % 4
            if kind == 1:    # rotation about the origin + translation back into the frame
                a = rng.uniform(-0.4, 0.4)
                c_, s_ = math.cos(a), math.sin(a)
                path = path.transform([c_, -s_, synth.f32(rng.uniform(-20, 60)), s_, c_, synth.f32(rng.uniform(-40, 40)), 0.0, 0.0, 1.0])
            elif kind == 2:  # shrink
                k = rng.uniform(0.4, 0.95)
                path = path.transform([k, 0.0, synth.f32(rng.uniform(0, 80)), 0.0, k, synth.f32(rng.uniform(0, 50)), 0.0, 0.0, 1.0])
            elif kind == 3:  # scale up: control points transformed, path re-flattened
                path = path.transform([1.3, 0.0, -90.0, 0.0, 1.2, -40.0, 0.0, 0.0, 1.0])
            col = Color(rng.uniform(), rng.uniform(), rng.uniform(), rng.uniform(0.3, 1.0))
            comp.get_mut_or_insert_default(i).insert(path).set_props(Props(func=Func.Draw(Style(fill=Fill.Solid(col)))))
        buf = np.zeros(w * h * 4, np.uint8)
        r.render(comp, buf, w, h, RGBA, Color(1, 1, 1, 1))
        outs.append(buf)
    assert_same(outs[0].reshape(h, -1), outs[1].reshape(h, -1), "inserts with path transforms")


def test_layer_ops_disable_transform_clear_remove(cuda_api, oracle_api, cuda_renderer, oracle_renderer):
    """composition/mod.rs:520-1000 style sequence: several renders of one composition."""
    outs = []
    for api, r in ((cuda_api, cuda_renderer), (oracle_api, oracle_renderer)):
        comp = api.Composition()
        synth.random_mixed(api, comp, 40, 160, 96, 51)
        frames = []

        def shot():
            buf = np.zeros(160 * 96 * 4, np.uint8)
            r.render(comp, buf, 160, 96, RGBA, Color(0.1, 0.1, 0.1, 1.0))
            frames.append(buf.copy())
        shot()
        comp.get(3).disable()
        comp.get(5).set_transform([1.0, 0.0, 0.0, 1.0, 12.5, -3.25])
        shot()
        comp.get(7).clear()
        comp.get(7).insert(synth.circle_path(api, 80.0, 48.0, 30.0))
        comp.remove(9)
        comp.get(3).enable()
        angle = np.float32(-np.pi / 2)
        comp.get(11).set_transform([float(np.cos(angle)), float(-np.sin(angle)), float(np.sin(angle)),
                                    float(np.cos(angle)), 40.0, 90.0])
        shot()
        layer = comp.create_layer()
        layer.insert(synth.circle_path(api, 20.0, 20.0, 15.0)).set_props(scenes.solid(Color(1, 0, 0, 0.5)))
        shot()  # detached layer is invisible
        comp.insert(2, layer)
        shot()
        outs.append(frames)
    for i, (a, b) in enumerate(zip(*outs)):
        assert_same(a.reshape(96, -1), b.reshape(96, -1), f"frame {i}")


def test_order_limit_and_bad_arguments(cuda_api, cuda_renderer):
    from forma_b200.binding import FormaError, GeomPresTransformError, OrderError
    comp = cuda_api.Composition()
    with pytest.raises(OrderError):
        comp.get_mut_or_insert_default((1 << 21))
    comp.get_mut_or_insert_default((1 << 21) - 1)  # LAYER_LIMIT itself is fine
    with pytest.raises(GeomPresTransformError):
        comp.get_mut_or_insert_default(0).set_transform([1.0, 0.0, 0.0, 2.0, 0.0, 0.0])
    buf = np.zeros(16, np.uint8)
    with pytest.raises(FormaError):
        cuda_renderer.render(comp, buf, 2, 2, RGBA, Color(), stride=4)  # width * 4 > stride


def test_large_frame_properties(cuda_api, cuda_renderer, unsliced):
    """Full-size checks that do not need the oracle: sortedness, idempotence,
    and a layer-order-independent checksum (opaque disjoint rectangles)."""
    w, h = 3840, 2160
    comp = cuda_api.Composition()
    synth.random_cubics(cuda_api, comp, 20000, w, h, 3)
    buf1 = np.zeros(w * h * 4, np.uint8)
    t = cuda_renderer.render(comp, buf1, w, h, RGBA, Color(1, 1, 1, 1))
    segs = cuda_renderer.segments()
    assert segs.size == t.n_segments > 100_000
    k = segs >> np.uint64(20)
    assert np.all(k[1:] >= k[:-1])
    buf2 = np.zeros_like(buf1)
    cuda_renderer.render(comp, buf2, w, h, RGBA, Color(1, 1, 1, 1))
    assert np.array_equal(buf1, buf2), "render is not idempotent"
    assert np.all(buf1.reshape(-1, 4)[:, 3] == 255)


# --- layer cache / damage reuse (cpu/buffer/mod.rs:114-197, composition/mod.rs:520-563,1038-1382) ---
import cache_scenarios  # noqa: E402


@pytest.mark.parametrize("scenario", cache_scenarios.SCENARIOS, ids=lambda f: f.__name__)
def test_layer_cache_scenarios(cuda_api, oracle_api, scenario):
    got = scenario(cuda_api)  # the reference's own asserts run inside
    want = scenario(oracle_api)
    assert len(got) == len(want)
    for i, (a, b) in enumerate(zip(got, want)):
        assert np.array_equal(a, b), f"{scenario.__name__}: frame {i} differs"


def test_layer_cache_animation_matches_oracle(cuda_api, oracle_api):
    got = cache_scenarios.animated_scene(cuda_api)
    want = cache_scenarios.animated_scene(oracle_api)
    for i, (a, b) in enumerate(zip(got, want)):
        assert_same(a.reshape(120, -1), b.reshape(120, -1), f"animated frame {i}")


def test_layer_cache_copies_back_only_damage(cuda_api):
    """The second frame of an unchanged scene writes no tile; moving one small
    layer damages only the tiles it leaves and enters."""
    w, h = 512, 256
    r = cuda_api.Renderer(0)
    cache = r.create_buffer_layer_cache()
    comp = cuda_api.Composition()
    synth.random_mixed(cuda_api, comp, 40, w, h, 9)
    comp.get_mut_or_insert_default(1000).insert(synth.circle_path(cuda_api, 100.0, 100.0, 10.0)).set_props(
        Props(func=Func.Draw(Style(fill=Fill.Solid(Color(1, 0, 0, 1))))))
    buf = np.zeros(w * h * 4, np.uint8)
    r.render(comp, buf, w, h, RGBA, Color(1, 1, 1, 1), None, cache)
    first = r.counters()["written_tiles"]
    assert first == (w // 16) * (h // 16)
    r.render(comp, buf, w, h, RGBA, Color(1, 1, 1, 1), None, cache)
    assert r.counters()["written_tiles"] == 0
    comp.get(1000).set_transform([1.0, 0.0, 0.0, 1.0, 64.0, 0.0])
    full = buf.copy()
    r.render(comp, buf, w, h, RGBA, Color(1, 1, 1, 1), None, cache)
    moved = r.counters()["written_tiles"]
    assert 0 < moved <= 16
    fresh = np.zeros_like(buf)
    r.render(comp, fresh, w, h, RGBA, Color(1, 1, 1, 1))
    assert np.array_equal(buf, fresh)
    assert not np.array_equal(buf, full)


def test_cleared_geometry_is_compacted(cuda_api, oracle_api, cuda_renderer, oracle_renderer):
    """Composition::compact_geom (composition/mod.rs:184-218): clearing and re-inserting
    a layer every frame must not grow the resident segment buffer without bound, and
    the frames stay identical to the oracle's."""
    w, h, frames = 320, 240, 60

    def big_path(api, k):
        rng = synth.SplitMix64(100 + k)
        pb = api.PathBuilder().move_to(Point(160.0, 120.0))
        for i in range(4000):
            ang = 0.05 * i
            r = 5.0 + 0.025 * i + rng.uniform(0.0, 1.0)
            pb.line_to(Point(synth.f32(160.0 + r * np.cos(ang)), synth.f32(120.0 + r * np.sin(ang))))
        return pb.build()

    outs = []
    for api, r in ((cuda_api, cuda_renderer), (oracle_api, oracle_renderer)):
        comp = api.Composition()
        synth.random_mixed(api, comp, 20, w, h, 3)
        layer = comp.get_mut_or_insert_default(500)
        layer.set_props(Props(fill_rule=FillRule.EvenOdd, func=Func.Draw(Style(fill=Fill.Solid(Color(0.1, 0.6, 0.3, 0.8))))))
        buf = np.zeros(w * h * 4, np.uint8)
        for k in range(frames):
            comp.get(500).clear()
            comp.get(500).insert(big_path(api, k))
            r.render(comp, buf, w, h, RGBA, Color(1, 1, 1, 1))
        outs.append(buf.copy())
        if api is cuda_api:
            assert comp.point_count() < 100_000, comp.point_count()  # 60 x 4001 points without compaction
    assert_same(outs[0].reshape(h, -1), outs[1].reshape(h, -1), "frame after 60 clear / insert cycles")


def test_cleared_ids_are_renumbered(cuda_api, oracle_api, cuda_renderer, oracle_renderer):
    """Every Layer::clear takes a new geometry id (layer.rs:131-146). The device's
    id -> layer table must follow the live layers, not the ids ever handed out: after
    70 000 clears the tables a frame uploads stay a few hundred bytes, the public
    geom_id keeps growing, and the frames stay identical to the oracle's."""
    w, h = 160, 96
    outs = []
    for api, r in ((cuda_api, cuda_renderer), (oracle_api, oracle_renderer)):
        comp = api.Composition()
        synth.random_mixed(api, comp, 6, w, h, 9)
        layer = comp.get_mut_or_insert_default(40)
        layer.set_props(Props(func=Func.Draw(Style(fill=Fill.Solid(Color(0.8, 0.2, 0.1, 0.7))))))
        tri = api.PathBuilder().move_to(Point(10, 10)).line_to(Point(150, 30)).line_to(Point(60, 90)).build()
        layer.insert(tri)
        buf = np.zeros(w * h * 4, np.uint8)
        r.render(comp, buf, w, h, RGBA, Color(1, 1, 1, 1))
        g0 = layer.geom_id()
        for _ in range(70_000):
            layer.clear()
        assert layer.geom_id() == g0 + 70_000
        layer.insert(tri)
        r.render(comp, buf, w, h, RGBA, Color(1, 1, 1, 1))   # renumbers (and re-evaluates) here
        if api is cuda_api:
            before = r.counters()["h2d_bytes"]
            comp.get(40).disable()                           # dirties the tables only
            r.render(comp, buf, w, h, RGBA, Color(1, 1, 1, 1))
            comp.get(40).enable()
            r.render(comp, buf, w, h, RGBA, Color(1, 1, 1, 1))
            assert r.counters()["h2d_bytes"] - before < 8192, r.counters()["h2d_bytes"] - before
        outs.append(buf.copy())
    assert_same(outs[0].reshape(h, -1), outs[1].reshape(h, -1), "frame after 70 000 clears")


def test_shared_frame_owner_side(cuda_api, cuda_renderer):
    """forma_shared_frame_create / _free and rendering into the shared allocation
    (the mapping side needs a second process; bench.py --gpus N exercises it)."""
    import torch
    w, h = 256, 128
    comp = cuda_api.Composition()
    synth.random_mixed(cuda_api, comp, 30, w, h, 12)
    host = np.zeros(w * h * 4, np.uint8)
    cuda_renderer.render(comp, host, w, h, RGBA, Color(0.2, 0.3, 0.4, 1))
    frame = cuda_api.SharedFrame(0, w * h * 4)
    assert len(frame.handle) == 64 and frame.ptr
    cuda_renderer.render_device(comp, frame.ptr, w, h, RGBA, Color(0.2, 0.3, 0.4, 1))
    view = torch.as_tensor(frame, device="cuda:0")
    assert np.array_equal(view.cpu().numpy(), host)
    del view
    frame.close()


def test_render_device_matches_host_buffer(cuda_api, cuda_renderer):
    import torch
    w, h = 300, 200
    comp = cuda_api.Composition()
    synth.random_mixed(cuda_api, comp, 50, w, h, 4)
    host = np.zeros(w * h * 4, np.uint8)
    cuda_renderer.render(comp, host, w, h, RGBA, Color(0.5, 0.5, 0.5, 1))
    dev = torch.zeros(w * h * 4, dtype=torch.uint8, device="cuda:0")
    torch.cuda.synchronize()
    cuda_renderer.render_device(comp, dev.data_ptr(), w, h, RGBA, Color(0.5, 0.5, 0.5, 1))
    assert np.array_equal(dev.cpu().numpy(), host)
    # With a layer cache the device buffer keeps the bytes of unwritten tiles.
    cache = cuda_renderer.create_buffer_layer_cache()
    cuda_renderer.render_device(comp, dev.data_ptr(), w, h, RGBA, Color(0.5, 0.5, 0.5, 1), None, cache)
    dev.fill_(7)
    torch.cuda.synchronize()
    cuda_renderer.render_device(comp, dev.data_ptr(), w, h, RGBA, Color(0.5, 0.5, 0.5, 1), None, cache)
    assert bool((dev == 7).all())
