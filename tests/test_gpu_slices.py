"""Host frames rendered as a pipeline of tile-row slices (forma_renderer_render, option
host_slices): every slice is a band render on its own stream with its own band-filtered
residency, so the frame must be the oracle's byte for byte for any number of slices and copy
bands, after the slices were re-balanced, after the composition changed, with a crop, a padded
stride or another channel order - and the frames the pipeline must leave alone (layer caches,
layer transforms, small compositions, short frames) must take the plain path."""
import numpy as np
import pytest

import synth
from forma_b200.binding import BGRA, RGBA, Color, Fill, Func, Point, Props, Rect, Style

pytestmark = pytest.mark.gpu

W, H = 3840, 2160
CLEAR = Color(1.0, 1.0, 1.0, 0.0)
N_LAYERS = 1200


def _scene(api, n=N_LAYERS, seed=97):
    comp = api.Composition()
    synth.random_mixed(api, comp, n, W, H, seed)
    return comp


@pytest.fixture(scope="module")
def expected(oracle_api):
    buf = np.zeros(W * H * 4, np.uint8)
    t = oracle_api.Renderer(0).render(_scene(oracle_api), buf, W, H, RGBA, CLEAR)
    return buf, int(t.n_segments)


@pytest.fixture()
def slice_options(cuda_api):
    names = ("host_slices", "slice_bands", "slice_min_points", "slice_chain", "sync_free", "band_filter")
    saved = {n: cuda_api.get_option(n) for n in names}
    cuda_api.set_option("slice_min_points", 0)
    yield
    for n, v in saved.items():
        cuda_api.set_option(n, v)


@pytest.mark.parametrize("slices,bands", [(2, 2), (4, 1), (4, 2), (4, 4), (8, 2), (3, 16), (16, 1)])
def test_sliced_host_frame_matches_oracle(cuda_api, expected, slice_options, slices, bands):
    want, n_want = expected
    cuda_api.set_option("host_slices", slices)
    cuda_api.set_option("slice_bands", bands)
    r = cuda_api.Renderer(0)
    comp = _scene(cuda_api)
    used = min(slices, (H // 16) // 16)
    for frame in range(4):  # 0: equal slices, first-frame tables; 1: re-balanced slices; 2: steady; 3: evicted
        if frame == 3:
            comp.evict()
        buf = np.zeros(W * H * 4, np.uint8)
        t = r.render(comp, buf, W, H, RGBA, CLEAR)
        assert np.array_equal(buf, want), f"{slices} slices x {bands} bands: frame {frame} differs from the oracle"
        assert len(r.host_slices()) == used
        assert int(t.n_segments) == n_want == r.counters()["segments"]


def test_sliced_counters_are_those_of_the_whole_frame(cuda_api, slice_options):
    comp = cuda_api.Composition()
    synth.random_cubics(cuda_api, comp, 6000, W, H, 3)  # shapes of 40-400 pixels: few cross a slice boundary
    r1, r4 = cuda_api.Renderer(0), cuda_api.Renderer(0)
    buf = np.zeros(W * H * 4, np.uint8)
    cuda_api.set_option("host_slices", 1)
    r1.render(comp, buf, W, H, RGBA, CLEAR)
    assert r1.host_slices() == []
    whole = r1.counters()
    cuda_api.set_option("host_slices", 4)
    r4.render(comp, buf, W, H, RGBA, CLEAR)
    c = r4.counters()
    assert c["segments"] == whole["segments"]  # every slice counts the segments of its own rows
    # a slice also forms the cells of the segments that boundary-crossing lines leave in its neighbours' rows
    assert whole["cells"] <= c["cells"] < 1.2 * whole["cells"] and 0.9 * whole["entries"] < c["entries"] < 1.2 * whole["entries"]
    assert c["d2h_bytes"] == whole["d2h_bytes"] == W * H * 4
    assert c["launches"] > whole["launches"]
    # every slice uploads the tables (one style record per layer here) and its band's geometry (shapes
    # crossing a boundary twice), never the whole composition four times
    assert whole["h2d_bytes"] <= c["h2d_bytes"] < 3.5 * whole["h2d_bytes"]
    assert r4.segments().size == 0  # no single sorted array after a sliced frame


def test_sliced_frames_follow_a_changing_composition(cuda_api, oracle_api, slice_options):
    cuda_api.set_option("host_slices", 4)
    outs = []
    for api in (cuda_api, oracle_api):
        r = api.Renderer(0)
        comp = _scene(api, 500, 5)
        frames = []

        def shot():
            buf = np.zeros(W * H * 4, np.uint8)
            r.render(comp, buf, W, H, RGBA, CLEAR)
            frames.append(buf)
            if api is cuda_api:
                sliced.append(len(r.host_slices()))
        sliced = []
        shot()
        tri = api.PathBuilder().move_to(Point(100, 100)).line_to(Point(3700, 400)).line_to(Point(1800, 2100)).build()
        layer = comp.get_mut_or_insert_default(600)
        layer.insert(tri).set_props(Props(func=Func.Draw(Style(fill=Fill.Solid(Color(0.1, 0.5, 0.9, 0.6))))))
        shot()                                   # a new insert reaches every slice
        for order in range(0, 500, 3):
            comp.get(order).disable()
        shot()                                   # tables only
        comp.remove(600)
        for order in range(0, 500, 5):
            comp.get(order).clear()
        shot()                                   # dead geometry
        comp.get(7).set_transform([1.0, 0.0, 0.0, 1.0, 40.0, -25.0])
        shot()                                   # a layer transform: no band filter, the frame is rendered in one piece
        comp.get(7).set_transform([1.0, 0.0, 0.0, 1.0, 0.0, 0.0])
        shot()
        outs.append(frames)
        if api is cuda_api:
            assert sliced[:4] == [4, 4, 4, 4] and sliced[4] == 0, sliced
    for k, (a, b) in enumerate(zip(*outs)):
        assert np.array_equal(a, b), f"frame {k} differs from the oracle"


def test_sliced_crop_stride_and_channel_order(cuda_api, oracle_api, slice_options):
    cuda_api.set_option("host_slices", 4)
    stride = W * 4 + 64
    crop = Rect((200, 3000), (300, 1900))  # 100 tile rows: four slices of 25
    outs = []
    for api in (cuda_api, oracle_api):
        comp = _scene(api, 600, 11)
        r = api.Renderer(0)
        buf = np.full(stride * H, 0x5A, np.uint8)
        r.render(comp, buf, W, H, BGRA, Color(0.2, 0.3, 0.4, 1.0), crop, None, stride)
        outs.append(buf)
        if api is cuda_api:
            assert len(r.host_slices()) == 4
    assert np.array_equal(outs[0], outs[1])


def test_frames_the_pipeline_leaves_alone(cuda_api, oracle_api, slice_options):
    cuda_api.set_option("host_slices", 4)
    r = cuda_api.Renderer(0)
    comp = _scene(cuda_api, 300, 3)
    buf = np.zeros(W * H * 4, np.uint8)
    r.render(comp, buf, W, H, RGBA, CLEAR)
    assert len(r.host_slices()) == 4
    want = buf.copy()
    cache = r.create_buffer_layer_cache()
    buf2 = np.zeros(W * H * 4, np.uint8)
    r.render(comp, buf2, W, H, RGBA, CLEAR, None, cache)        # layer cache
    assert r.host_slices() == [] and np.array_equal(buf2, want)
    small = np.zeros(W * 400 * 4, np.uint8)
    r.render(comp, small, W, 400, RGBA, CLEAR)                  # 25 tile rows: too short for two slices
    assert r.host_slices() == []
    assert np.array_equal(small.reshape(400, -1), want.reshape(H, -1)[:400])
    cuda_api.set_option("slice_min_points", 1 << 30)            # small composition
    r.render(comp, buf2, W, H, RGBA, CLEAR)
    assert r.host_slices() == [] and np.array_equal(buf2, want)
    cuda_api.set_option("slice_min_points", 0)
    cuda_api.set_option("band_filter", 0)                       # every slice keeps everything resident: still exact
    r.render(comp, buf2, W, H, RGBA, CLEAR)
    assert len(r.host_slices()) == 4 and np.array_equal(buf2, want)
    cuda_api.set_option("band_filter", 1)
    cuda_api.set_option("slice_chain", 0)                       # uploads issued by the slices' threads, all at once
    comp.evict()
    buf2[:] = 0
    r.render(comp, buf2, W, H, RGBA, CLEAR)
    assert len(r.host_slices()) == 4 and np.array_equal(buf2, want)
