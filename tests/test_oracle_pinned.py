"""Pins the CPU oracle against the reference's own known answers (no GPU).

Every expected value below is taken from a test in /root/reference (cited per
test); the 34 golden images are the reference's e2e-tests/expected/*.png,
committed as tests/golden/e2e_expected.npz by tests/golden/make_e2e_golden.py.
"""
import ctypes as C
import os

import numpy as np
import pytest

import scenes
from forma_b200.binding import (RGBA, BlendMode, Color, Fill, FillRule, Func, Point, Props, Style, unpack_segments)
from oracle import oracle

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "e2e_expected.npz"))


@pytest.fixture(scope="module")
def api():
    return oracle.load()


@pytest.fixture(scope="module")
def lib(api):
    lib = api.hooks
    fp, u8p = C.POINTER(C.c_float), C.POINTER(C.c_uint8)
    lib.fo_prim_new.restype = C.c_void_p
    lib.fo_prim_free.argtypes = [C.c_void_p]
    lib.fo_prim_contour.argtypes = [C.c_void_p]
    for n in ("fo_prim_line", "fo_prim_quad", "fo_prim_cubic"):
        getattr(lib, n).argtypes = [C.c_void_p, fp]
    lib.fo_prim_segments.restype = C.c_uint64
    lib.fo_prim_segments.argtypes = [C.c_void_p, C.c_uint64, fp, fp, u8p]
    lib.fo_set_recip_mode.argtypes = [C.c_int]
    return lib


class Prim:
    """forma/src/path.rs Primitives, driven like the reference's unit tests."""

    def __init__(self, lib):
        self.lib, self.h = lib, lib.fo_prim_new()

    def _call(self, fn, pts, weights=None):
        weights = weights or [1.0] * len(pts)
        flat = []
        for (x, y), w in zip(pts, weights):
            flat += [x, y, w]
        fn(self.h, (C.c_float * len(flat))(*flat))

    def line(self, *pts):
        self._call(self.lib.fo_prim_line, pts)

    def quad(self, *pts, weights=None):
        self._call(self.lib.fo_prim_quad, pts, weights)

    def cubic(self, *pts, weights=None):
        self._call(self.lib.fo_prim_cubic, pts, weights)

    def contour(self):
        self.lib.fo_prim_contour(self.h)

    def segments(self):
        n = self.lib.fo_prim_segments(self.h, 0, None, None, None)
        x, y, c = np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.uint8)
        self.lib.fo_prim_segments(self.h, n, x.ctypes.data_as(C.POINTER(C.c_float)),
                                  y.ctypes.data_as(C.POINTER(C.c_float)), c.ctypes.data_as(C.POINTER(C.c_uint8)))
        return x, y, c


# --- stage 1: forma/src/path.rs:1024-1492 -------------------------------------------------
def test_path_quads(lib):  # path.rs:1024-1073
    p = Prim(lib)
    p.quad((2.0, 0.0), (0.0, 1.0), (10.0, 1.0))
    p.quad((10.0, 1.0), (20.0, 1.0), (18.0, 0.0))
    x, y, _ = p.segments()
    assert len(x) == 9
    assert (x[0], y[0], x[8], y[8]) == (2.0, 0.0, 18.0, 0.0)
    assert np.hypot(x[3] - x[5], y[3] - y[5]) > 10.0


def test_path_two_splines(lib):  # path.rs:1075-1095
    p = Prim(lib)
    p.quad((0.0, 0.0), (1.0, 2.0), (2.0, 0.0))
    p.quad((3.0, 0.0), (4.0, 4.0), (5.0, 0.0))
    x, y, _ = p.segments()
    assert len(x) == 11
    assert [(x[i], y[i]) for i in (0, 4, 5, 10)] == [(0.0, 0.0), (2.0, 0.0), (3.0, 0.0), (5.0, 0.0)]


def test_path_collinear_and_overlapping(lib):  # path.rs:1097-1128
    p = Prim(lib)
    p.quad((0.0, 0.0), (2.0, 0.0001), (1.0, 0.0))
    x, y, _ = p.segments()
    assert len(x) == 3 and abs(x[1] - 1.25) < 0.01 and abs(y[1]) < 0.01
    p = Prim(lib)
    p.quad((0.0, 0.0), (0.0, 0.0), (1.0, 1.0))
    p.quad((1.0, 1.0), (1.0, 1.0), (1.0, 1.0))
    p.quad((1.0, 1.0), (2.0, 2.0), (2.0, 2.0))
    x, y, _ = p.segments()
    assert len(x) == 2 and abs(x[1] - 2.0) < 0.01 and abs(y[1] - 2.0) < 0.01


def test_path_rat_quad(lib):  # path.rs:1130-1173
    p = Prim(lib)
    w = 10.0
    p.quad((0.0, 0.0), (1.0 * w, 2.0 * w), (2.0, 0.0), weights=[1.0, w, 1.0])
    x, y, _ = p.segments()
    assert len(x) == 5 and abs(x[2] - 1.0) <= 0.001
    d = np.hypot(np.diff(x), np.diff(y))
    assert d[0] > 1.5 and d[1] < 0.2 and d[2] < 0.2 and d[3] > 1.5


def test_path_lines_and_quads(lib):  # path.rs:1175-1203
    p = Prim(lib)
    p.line((-1.0, -2.0), (0.0, 0.0))
    p.quad((0.0, 0.0), (1.0, 2.0), (2.0, 0.0))
    p.line((2.0, 0.0), (3.0, -2.0))
    p.line((3.0, -2.0), (4.0, 2.0))
    p.line((4.0, 2.0), (5.0, -4.0))
    p.line((5.0, -4.0), (6.0, 0.0))
    p.quad((6.0, 0.0), (7.0, 4.0), (8.0, 0.0))
    p.line((8.0, 0.0), (9.0, -4.0))
    x, y, _ = p.segments()
    assert len(x) == 12
    assert [(x[i], y[i]) for i in (0, 4, 5, 6, 11)] == [(-1.0, -2.0), (3.0, -2.0), (4.0, 2.0), (5.0, -4.0), (9.0, -4.0)]


def test_path_cubics(lib):  # path.rs:1205-1318
    p = Prim(lib)
    p.cubic((0.0, 0.0), (10.0, 6.0), (-2.0, 6.0), (8.0, 0.0))
    x, y, _ = p.segments()
    assert len(x) == 10
    assert x[2] > x[7] and x[3] > x[6] and x[4] > x[5]
    assert np.all(np.diff(y[:5]) > 0) and np.all(np.diff(y[5:]) < 0)
    for w, count in ((10.0, 45), (0.5, 7)):  # rat_cubic_high / rat_cubic_low
        p = Prim(lib)
        p.cubic((0.0, 0.0), (5.0 * w, 3.0 * w), (-1.0 * w, 3.0 * w), (4.0, 0.0), weights=[1.0, w, w, 1.0])
        assert len(p.segments()[0]) == count
    p = Prim(lib)  # collinear_cubic
    p.cubic((1.0, 0.0), (0.0, 0.0), (3.0, 0.0), (2.0, 0.0))
    x, y, _ = p.segments()
    assert len(x) == 5 and x[0] == 1.0 and x[4] == 2.0 and np.all(y == 0.0)
    assert 0.5 < x[1] < 1.0 < x[2] < 2.0 < x[3] < 2.5
    p = Prim(lib)  # overlapping_control_point_cubic_line
    p.cubic((0.0, 0.0), (0.0, 0.0), (1.0, 1.0), (1.0, 1.0))
    p.cubic((1.0, 1.0), (1.0, 1.0), (1.0, 1.0), (1.0, 1.0))
    p.cubic((1.0, 1.0), (1.0, 1.0), (2.0, 2.0), (2.0, 2.0))
    x, y, _ = p.segments()
    assert len(x) == 9 and np.all(np.diff(x) > 0) and np.array_equal(x, y)


def test_path_rings_contour_flags(lib):  # path.rs:1320-1384
    p = Prim(lib)
    for pts in [((0.0, 2.0), (2.0, 2.0), (2.0, 2.0), (2.0, 0.0)), ((2.0, 0.0), (2.0, -2.0), (2.0, -2.0), (0.0, -2.0)),
                ((0.0, -2.0), (-2.0, -2.0), (-2.0, -2.0), (-2.0, 0.0)), ((-2.0, 0.0), (-2.0, 2.0), (-2.0, 2.0), (0.0, 2.0))]:
        p.cubic(*pts)
    p.contour()
    for pts in [((0.0, 1.0), (-1.0, 1.0), (-1.0, 1.0), (-1.0, 0.0)), ((-1.0, 0.0), (-1.0, -1.0), (-1.0, -1.0), (0.0, -1.0)),
                ((0.0, -1.0), (1.0, -1.0), (1.0, -1.0), (1.0, 0.0)), ((1.0, 0.0), (1.0, 1.0), (1.0, 1.0), (0.0, 1.0))]:
        p.cubic(*pts)
    _, _, c = p.segments()
    assert len(c) == 30 and c.sum() == 2 and c[16] and c[29]
    p = Prim(lib)  # ring_overlapping_start
    for pts in [((0.0, 1.0), (-1.0, 1.0), (-1.0, 1.0), (-1.0, 0.0)), ((-1.0, 0.0), (-1.0, -1.0), (-1.0, -1.0), (0.0, -1.0)),
                ((0.0, -1.0), (1.0, -1.0), (1.0, -1.0), (1.0, 0.0)), ((1.0, 0.0), (1.0, 1.0), (1.0, 1.0), (0.0, 1.0))]:
        p.cubic(*pts)
    p.contour()
    for pts in [((0.0, 1.0), (1.0, 1.0), (1.0, 1.0), (1.0, 2.0)), ((1.0, 2.0), (1.0, 3.0), (1.0, 3.0), (0.0, 3.0)),
                ((0.0, 3.0), (-1.0, 3.0), (-1.0, 3.0), (-1.0, 2.0)), ((-1.0, 2.0), (-1.0, 1.0), (-1.0, 1.0), (0.0, 1.0))]:
        p.cubic(*pts)
    _, _, c = p.segments()
    assert len(c) == 26 and c.sum() == 2 and c[12] and c[25]


def test_path_circle_66_points(lib):  # path.rs:1386-1476
    r = 50.0
    w = float(np.sqrt(np.float32(2.0)) / np.float32(2.0))
    f = lambda v: float(np.float32(v))
    p = Prim(lib)
    p.quad((r, 0.0), (0.0, 0.0), (0.0, r), weights=[1.0, w, 1.0])
    p.quad((0.0, r), (0.0, f(np.float32(2.0 * r) * np.float32(w))), (r, 2.0 * r), weights=[1.0, w, 1.0])
    p.quad((r, 2.0 * r), (f(np.float32(2.0 * r) * np.float32(w)), f(np.float32(2.0 * r) * np.float32(w))), (2.0 * r, r),
           weights=[1.0, w, 1.0])
    p.quad((2.0 * r, r), (f(np.float32(2.0 * r) * np.float32(w)), 0.0), (r, 0.0), weights=[1.0, w, 1.0])
    x, y, _ = p.segments()
    assert len(x) == 66
    assert np.hypot(np.diff(x), np.diff(y)).max() < 5.0


def test_path_transform(api):  # path.rs:1495-1561
    path = scenes.custom_circle(api, 0.0, 0.0, 10.0)
    x, y, c = path.segments()
    assert np.all(np.abs(np.hypot(x, y) - 10.0) <= 0.1)
    xt, yt, _ = path.transform([1.0, 0.0, 5.0, 0.0, 1.0, 20.0, 0.0, 0.0, 1.0]).segments()
    assert np.all(np.abs(np.hypot(xt - 5.0, yt - 20.0) - 10.0) <= 0.1)
    xs, ys, _ = path.transform([2.0, 0.0, 0.0, 0.0, 2.0, 0.0, 0.0, 0.0, 1.0]).segments()
    assert np.all(np.abs(np.hypot(xs, ys) - 20.0) <= 0.1) and len(xs) > len(x)


def test_approx_atan2(lib):  # math/point.rs:172-180 (within 2e-3 of atan2)
    for y, x in [(0.0, 1.0), (1.0, 1.0), (1.0, 0.0), (1.0, -1.0), (0.0, -1.0), (-1.0, -1.0), (-1.0, 0.0), (-1.0, 1.0),
                 (0.3, 0.9), (-5.0, 0.2)]:
        assert abs(lib.fo_approx_atan2(y, x) - np.arctan2(y, x)) < 2e-3


# --- stage 2: forma/src/cpu/rasterizer.rs:205-557, cpu/pixel_segment.rs:221-355 ------------
def test_find_sequences(lib):  # rasterizer.rs:205-244
    got = [lib.fo_find(i - 1, 2.0, 3.0, np.float32(0.2), np.float32(0.1)) for i in range(7)]
    assert got == [float(np.float32(v)) for v in (0.1, 0.2, 2.2, 3.1, 4.2, 6.1, 6.2)]
    got = [lib.fo_find(i - 1, 16_777_216.0, np.float32(0.0001), 10.0, np.float32(0.00001)) for i in (2, 3)]
    assert got == [float(np.float32(0.00021)), float(np.float32(0.00031))]


def line_segments(api, p0, p1, order=0):
    """rasterizer.rs:179-195 `segments`: one raw line (the reference's
    #[cfg(test)] SegmentBuffer::push) on a huge canvas."""
    comp = api.Composition()
    layer = comp.get_mut_or_insert_default(order)
    api.hooks.fo_test_push_line.argtypes = [C.c_void_p, C.c_void_p] + [C.c_float] * 4
    api.hooks.fo_test_push_line(comp._h, layer._h, p0[0], p0[1], p1[0], p1[1])
    return api.Renderer().rasterize_only(comp, 1 << 24, 1 << 24)


OCTANTS = [  # rasterizer.rs:247-338
    ((0.0, 0.0), (3.0, 2.0), [(11 * 16, 11), (5 * 8 + 2 * (5 * 8), 5), (5 * 8, 5), (11 * 16, 11)]),
    ((0.0, 0.0), (2.0, 3.0), [(16 * 11 + 2 * (16 * 5), 16), (8 * 5, 8), (8 * 5 + 2 * (8 * 11), 8), (16 * 11, 16)]),
    ((0.0, 0.0), (-2.0, 3.0), [(16 * 11, 16), (8 * 5 + 2 * (8 * 11), 8), (8 * 5, 8), (16 * 11 + 2 * (16 * 5), 16)]),
    ((0.0, 0.0), (-3.0, 2.0), [(11 * 16, 11), (5 * 8, 5), (5 * 8 + 2 * (5 * 8), 5), (11 * 16, 11)]),
    ((3.0, 2.0), (0.0, 0.0), [(-(11 * 16), -11), (-(5 * 8), -5), (-(5 * 8 + 2 * (5 * 8)), -5), (-(11 * 16), -11)]),
    ((2.0, 3.0), (0.0, 0.0), [(-(16 * 11), -16), (-(8 * 5 + 2 * (8 * 11)), -8), (-(8 * 5), -8), (-(16 * 11 + 2 * (16 * 5)), -16)]),
    ((0.0, 3.0), (2.0, 0.0), [(-(16 * 11 + 2 * (16 * 5)), -16), (-(8 * 5), -8), (-(8 * 5 + 2 * (8 * 11)), -8), (-(16 * 11), -16)]),
    ((0.0, 2.0), (3.0, 0.0), [(-(11 * 16), -11), (-(5 * 8 + 2 * (5 * 8)), -5), (-(5 * 8), -5), (-(11 * 16), -11)]),
]
AXES = [  # rasterizer.rs:340-412
    ((0.0, 0.0), (1.0, 1.0), [(256, 16)]), ((0.0, 0.0), (0.0, 1.0), [(512, 16)]), ((0.0, 0.0), (-1.0, 1.0), [(256, 16)]),
    ((1.0, 1.0), (0.0, 0.0), [(-256, -16)]), ((0.0, 1.0), (0.0, 0.0), [(-512, -16)]), ((0.0, 1.0), (1.0, 0.0), [(-256, -16)]),
]


@pytest.mark.parametrize("p0,p1,expected", OCTANTS + AXES)
def test_area_cover(api, p0, p1, expected):
    u = unpack_segments(line_segments(api, p0, p1))
    assert list(zip(u["double_area"].tolist(), u["cover"].tolist())) == expected


def test_horizontal_lines_produce_nothing(api):  # rasterizer.rs:340-346,380-386
    assert line_segments(api, (0.0, 0.0), (1.0, 0.0)).size == 0
    assert line_segments(api, (0.0, 0.0), (-1.0, 0.0)).size == 0


TILES = [  # rasterizer.rs:429-544 (TILE_WIDTH = TILE_HEIGHT = 16)
    ((16.0, 16.0), (19.0, 18.0), [(1, 1, 0, 0), (1, 1, 1, 0), (1, 1, 1, 1), (1, 1, 2, 1)]),
    ((16.0, 16.0), (18.0, 19.0), [(1, 1, 0, 0), (1, 1, 0, 1), (1, 1, 1, 1), (1, 1, 1, 2)]),
    ((-16.0, 16.0), (-18.0, 19.0), [(-1, 1, 15, 0), (-1, 1, 15, 1), (-1, 1, 14, 1), (-1, 1, 14, 2)]),
    ((-16.0, 16.0), (-19.0, 18.0), [(-1, 1, 15, 0), (-1, 1, 14, 0), (-1, 1, 14, 1), (-1, 1, 13, 1)]),
    ((-16.0, 16.0), (-19.0, 14.0), [(-1, 0, 15, 15), (-1, 0, 14, 15), (-1, 0, 14, 14), (-1, 0, 13, 14)]),
    ((-16.0, 16.0), (-18.0, 13.0), [(-1, 0, 15, 15), (-1, 0, 15, 14), (-1, 0, 14, 14), (-1, 0, 14, 13)]),
    ((16.0, 16.0), (18.0, 13.0), [(1, 0, 0, 15), (1, 0, 0, 14), (1, 0, 1, 14), (1, 0, 1, 13)]),
    ((16.0, 16.0), (19.0, 14.0), [(1, 0, 0, 15), (1, 0, 1, 15), (1, 0, 1, 14), (1, 0, 2, 14)]),
]


@pytest.mark.parametrize("p0,p1,expected", TILES)
def test_tile_coordinates(api, p0, p1, expected):
    u = unpack_segments(line_segments(api, p0, p1))
    assert list(zip(u["tile_x"].tolist(), u["tile_y"].tolist(), u["local_x"].tolist(), u["local_y"].tolist())) == expected


def test_endpoints_off_pixel_border(api):  # rasterizer.rs:547-557
    u = unpack_segments(line_segments(api, (0.5, 0.25), (4.0, 2.0)))
    assert (u["double_area"][0], u["cover"][0]) == (4 * 8, 4)
    u = unpack_segments(line_segments(api, (0.0, 0.0), (3.5, 1.75)))
    assert (u["double_area"][4], u["cover"][4]) == (4 * 8 + 2 * (4 * 8), 4)


def test_rasterize_triangle_packed(api):  # gpu/rasterizer/mod.rs:323-352 (order 1, 16x16 tiles)
    comp = api.Composition()
    path = api.PathBuilder().move_to(Point(1.0, 1.0)).line_to(Point(1.0, 4.0)).line_to(Point(2.0, 4.0)).build()
    comp.get_mut_or_insert_default(1).insert(path)
    u = unpack_segments(api.Renderer().rasterize_only(comp, 1 << 24, 1 << 24))
    # expected_segments of the reference test: PixelSegment::new(layer, tile_x, tile_y, local_x, local_y, dam, cover)
    expected = [(1, 0, 0, 1, 1, 32, 16), (1, 0, 0, 1, 2, 32, 16), (1, 0, 0, 1, 3, 32, 16), (1, 0, 0, 1, 3, 5, -16),
                (1, 0, 0, 1, 2, 16, -16), (1, 0, 0, 1, 1, 27, -16)]
    got = list(zip(u["layer_id"].tolist(), u["tile_x"].tolist(), u["tile_y"].tolist(), u["local_x"].tolist(),
                   u["local_y"].tolist(), (u["double_area"] // u["cover"]).tolist(), u["cover"].tolist()))
    assert got == expected


def test_pixel_segment_bit_fields(lib):  # cpu/pixel_segment.rs:221-355
    s = lib.fo_pack_segment(0x15_5555, -1, -1, 15, 15, 32, -16)
    assert oracle.unpack(lib, s) == dict(layer_id=0x15_5555, tile_x=-1, tile_y=-1, local_x=15, local_y=15,
                                         double_area=32 * -16, cover=-16)
    s = lib.fo_pack_segment((1 << 21) - 1, 4094, 2046, 0, 0, 0, 16)
    assert oracle.unpack(lib, s)["tile_x"] == 4094 and oracle.unpack(lib, s)["tile_y"] == 2046
    # tiles left of / above -1 clamp to -1 (:47-52)
    s = lib.fo_pack_segment(3, -20, -7, 1, 2, 3, 4)
    assert oracle.unpack(lib, s)["tile_x"] == -1 and oracle.unpack(lib, s)["tile_y"] == -1
    # ordering ignores the low 20 bits (:161-171)
    a = lib.fo_pack_segment(5, 1, 1, 15, 15, 63, -1)
    b = lib.fo_pack_segment(6, 1, 1, 0, 0, 0, 0)
    assert (a >> 20) < (b >> 20)


# --- stage 4: forma/src/cpu/painter/mod.rs tests --------------------------------------------
def test_coverage_tables(lib):  # cpu/painter/mod.rs:1013-1040
    area = 512
    nz = [(-2 * area, 1.0), (-area * 3 // 2, 1.0), (-area, 1.0), (-area // 2, 0.5), (0, 0.0), (area // 2, 0.5), (area, 1.0),
          (area * 3 // 2, 1.0), (2 * area, 1.0)]
    for a, want in nz:
        assert lib.fo_coverage(a, 0) == want
    eo = [(-area * 3 // 2, 0.5), (-area, 1.0), (-area // 2, 0.5), (0, 0.0), (area // 2, 0.5), (area, 1.0), (area * 3 // 2, 0.5)]
    for a, want in eo:
        assert lib.fo_coverage(a, 1) == want


def test_f32_to_u8_all_values(lib):  # cpu/painter/mod.rs:1503-1515
    for i in range(256):
        assert lib.fo_to_byte(float(np.float32(i) / np.float32(255.0))) == i
    assert lib.fo_to_byte(-1.0) == 0 and lib.fo_to_byte(2.0) == 255


def test_to_srgb_bytes(lib):  # cpu/painter/mod.rs:1518-1531
    assert oracle.to_srgb_bytes(lib, [0.0005, 0.1, 0.25, 0.5]) == [2, 89, 137, 128]


def test_solid_fold_colors(lib):  # layer_workbench/mod.rs:907-919,965-977 (scalar BlendMode::blend)
    black_a0 = [0.0, 0.0, 0.0, 0.0]
    red_half = [0.5, 0.0, 0.0, 0.5]
    out = oracle.blend_scalar(lib, BlendMode.Over, [1.0, 1.0, 1.0, 1.0], [0.0, 0.0, 0.0, 0.25])
    out = oracle.blend_scalar(lib, BlendMode.Over, out, [0.0, 0.0, 0.0, 0.25])
    assert out.tolist() == [0.5625, 0.5625, 0.5625, 1.0]
    out = oracle.blend_scalar(lib, BlendMode.Over, black_a0, red_half)
    assert out.tolist() == [0.25, 0.0, 0.0, 0.5]


@pytest.mark.parametrize("mode", range(16))
def test_scalar_and_vector_blend_agree(lib, mode):  # cpu/painter/styling.rs:673-904 (EPSILON 1e-3)
    rng = np.random.default_rng(mode)
    for _ in range(200):
        d, s = rng.random(3, dtype=np.float32) * 0.98 + 0.01, rng.random(3, dtype=np.float32) * 0.98 + 0.01
        v = oracle.blend_lane(lib, mode, d, s)
        # scalar blend with dst.a = src.a = 1 returns the blended colour itself
        sc = oracle.blend_scalar(lib, mode, list(d) + [1.0], list(s) + [1.0])[:3]
        assert np.allclose(v, sc, atol=1e-3), (mode, d, s, v, sc)


# --- full pipeline: lib.rs doc test and composition/mod.rs tests -----------------------------
BLACK_SRGB, RED_SRGB, GREEN_SRGB = [0, 0, 0, 255], [255, 0, 0, 255], [0, 255, 0, 255]
GRAY_SRGB = [0xBB, 0xBB, 0xBB, 0xFF]
BLACK, RED, GREEN, GRAY = Color(0, 0, 0, 1), Color(1, 0, 0, 1), Color(0, 1, 0, 1), Color(0.5, 0.5, 0.5, 1)


def pixel_path(api, x, y):  # composition/mod.rs:446-456
    return (api.PathBuilder().move_to(Point(x, y)).line_to(Point(x, y + 1)).line_to(Point(x + 1, y + 1))
            .line_to(Point(x + 1, y)).line_to(Point(x, y)).build())


def render_row(api, comp, n, clear, renderer=None, prefill=GREEN_SRGB):
    buf = np.array(prefill * n, np.uint8)
    (renderer or api.Renderer()).render(comp, buf, n, 1, RGBA, clear)
    return buf.reshape(n, 4).tolist()


def test_lib_doc_example(api):  # forma/src/lib.rs:24-94
    comp = api.Composition()
    r = api.Renderer()
    w, h = 250, 150
    def rect(x0, y0, x1, y1):
        return (api.PathBuilder().move_to(Point(x0, y0)).line_to(Point(x1, y0)).line_to(Point(x1, y1))
                .line_to(Point(x0, y1)).build())
    comp.get_mut_or_insert_default(0).insert(rect(50.0, 50.0, 150.0, 100.0)).set_props(scenes.solid(Color(1, 0, 0, 1)))
    comp.get_mut_or_insert_default(1).insert(rect(100.0, 50.0, 200.0, 100.0)).set_props(scenes.solid(Color(0, 0, 1, 1)))
    buf = np.zeros(w * h * 4, np.uint8)
    r.render(comp, buf, w, h, RGBA, Color(1, 1, 1, 1))
    px = buf.reshape(h, w, 4)
    assert px[75, 75].tolist() == [255, 0, 0, 255]      # red only
    assert px[75, 125].tolist() == [0, 0, 255, 255]     # blue over red
    assert px[75, 175].tolist() == [0, 0, 255, 255]
    assert px[25, 25].tolist() == [255, 255, 255, 255]  # clear colour


def test_background_clear_and_one_pixel(api):  # composition/mod.rs:496-518,566-640
    comp = api.Composition()
    assert render_row(api, comp, 1, RED) == [RED_SRGB]
    comp = api.Composition()
    comp.get_mut_or_insert_default(0).insert(pixel_path(api, 1, 0)).set_props(scenes.solid(RED))
    assert render_row(api, comp, 3, GREEN) == [GREEN_SRGB, RED_SRGB, GREEN_SRGB]
    # translate by half a pixel -> two half-covered pixels (sRGB 0xBB of 0.5)
    comp = api.Composition()
    comp.get_mut_or_insert_default(0).insert(pixel_path(api, 1, 0)).set_props(scenes.solid(RED)) \
        .set_transform([1.0, 0.0, 0.0, 1.0, 0.5, 0.0])
    row = render_row(api, comp, 3, Color(0, 1, 0, 1))
    assert row[1] == [0xBB, 0xBB, 0, 0xFF] and row[2] == [0xBB, 0xBB, 0, 0xFF] and row[0] == GREEN_SRGB


def test_one_pixel_rotated(api):  # composition/mod.rs:642-679
    comp = api.Composition()
    angle = np.float32(-np.pi / 2)
    layer = comp.create_layer()
    layer.insert(pixel_path(api, -1, 1)).set_props(scenes.solid(RED)).set_transform(
        [float(np.cos(angle)), float(-np.sin(angle)), float(np.sin(angle)), float(np.cos(angle)), 0.0, 0.0])
    comp.insert(0, layer)
    assert render_row(api, comp, 3, GREEN) == [GREEN_SRGB, RED_SRGB, GREEN_SRGB]


def test_clear_insert_over_and_remove(api):  # composition/mod.rs:681-900
    r = api.Renderer()
    comp = api.Composition()
    for order, xs in ((0, [0]), (1, [1]), (2, [2, 3])):
        layer = comp.create_layer()
        for x in xs:
            layer.insert(pixel_path(api, x, 0))
        layer.set_props(scenes.solid(RED))
        comp.insert(order, layer)
    assert render_row(api, comp, 4, GREEN, r) == [RED_SRGB] * 4
    comp.get(0).clear()
    assert render_row(api, comp, 4, GREEN, r) == [GREEN_SRGB, RED_SRGB, RED_SRGB, RED_SRGB]
    comp.get(2).clear()
    assert render_row(api, comp, 4, GREEN, r) == [GREEN_SRGB, RED_SRGB, GREEN_SRGB, GREEN_SRGB]
    # insert_over_layer / layer_replace_remove
    comp = api.Composition()
    layer = comp.create_layer()
    layer.insert(pixel_path(api, 0, 0)).set_props(scenes.solid(RED))
    comp.insert(0, layer)
    assert render_row(api, comp, 3, BLACK, r, BLACK_SRGB) == [RED_SRGB, BLACK_SRGB, BLACK_SRGB]
    layer = comp.create_layer()
    layer.insert(pixel_path(api, 1, 0)).set_props(scenes.solid(GREEN))
    assert render_row(api, comp, 3, BLACK, r, BLACK_SRGB) == [RED_SRGB, BLACK_SRGB, BLACK_SRGB]  # detached: invisible
    old = comp.insert(0, layer)
    assert old is not None
    assert render_row(api, comp, 3, BLACK, r, BLACK_SRGB) == [BLACK_SRGB, GREEN_SRGB, BLACK_SRGB]
    comp.remove(0)
    assert render_row(api, comp, 3, BLACK, r, BLACK_SRGB) == [BLACK_SRGB] * 3


def test_geom_id_semantics(api):  # composition/mod.rs:972-1001
    comp = api.Composition()
    layer = comp.create_layer()
    layer.insert(api.PathBuilder().build())
    g0 = layer.geom_id()
    layer.insert(api.PathBuilder().build())
    assert layer.geom_id() == g0
    layer.clear()
    assert layer.geom_id() != g0


def test_srgb_alpha_blending_and_even_odd(api):  # composition/mod.rs:1003-1035,1386-1428
    comp = api.Composition()
    comp.get_mut_or_insert_default(0).insert(pixel_path(api, 0, 0)).set_props(scenes.solid(Color(0, 0, 0, 0.5)))
    comp.get_mut_or_insert_default(1).insert(pixel_path(api, 1, 0)).set_props(scenes.solid(GRAY))
    row = render_row(api, comp, 3, Color(1, 1, 1, 0), prefill=BLACK_SRGB)
    assert row == [[0xBB, 0xBB, 0xBB, 0x80], GRAY_SRGB, [0xFF, 0xFF, 0xFF, 0x00]]
    comp = api.Composition()
    pb = api.PathBuilder()
    for (a, b) in ((0, 2), (1, 3)):  # two overlapping 2-px rectangles in one path
        pb.move_to(Point(a, 0)).line_to(Point(a, 1)).line_to(Point(b, 1)).line_to(Point(b, 0)).line_to(Point(a, 0))
    comp.get_mut_or_insert_default(0).insert(pb.build()).set_props(
        Props(fill_rule=FillRule.EvenOdd, func=Func.Draw(Style(fill=Fill.Solid(BLACK)))))
    assert render_row(api, comp, 3, Color(1, 1, 1, 1)) == [BLACK_SRGB, [255, 255, 255, 255], BLACK_SRGB]


# --- the reference's golden images ----------------------------------------------------------
@pytest.mark.parametrize("name", sorted(scenes.E2E))
def test_e2e_golden(api, lib, name):
    """e2e-tests/tests/tests.rs scenes vs e2e-tests/expected/*__cpu.png. The four
    non-separable blend modes were rendered through Arm's vrecpeq_f32 8-bit
    reciprocal estimate (utils/simd/aarch64.rs:520-530); with that estimate
    emulated the oracle reproduces them bit for bit too."""
    lib.fo_set_recip_mode(2 if name in scenes.NON_SEPARABLE else 0)
    try:
        img = scenes.render_e2e(api, name)
    finally:
        lib.fo_set_recip_mode(0)
    assert np.array_equal(img, GOLD[name + "__cpu"]), name


# --- layer cache / damage reuse: composition/mod.rs:520-563,1038-1382 --------------------------
import cache_scenarios  # noqa: E402


@pytest.mark.parametrize("scenario", cache_scenarios.SCENARIOS, ids=lambda f: f.__name__)
def test_layer_cache_scenarios(api, scenario):
    scenario(api)  # the reference's asserts live inside the scenario
