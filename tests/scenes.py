"""Scenes used by the parity tests, written against the binding classes so the
same function builds the scene for the CUDA library and for the oracle.

`E2E` restates the scenes of the reference's end-to-end image tests
(/root/reference/e2e-tests/tests/tests.rs:219-742; harness constants
test_env.rs:35-37: 64x64, PADDING 8, clear = white with alpha 0, RGBA).
"""
from __future__ import annotations

import math

import numpy as np

from forma_b200.binding import (BlendMode, Color, Fill, FillRule, Func, GradientBuilder, GradientType, Point,
                                Props, Style, Texture)

WIDTH = 64.0
HEIGHT = 64.0
PADDING = 8.0
E2E_CLEAR = Color(1.0, 1.0, 1.0, 0.0)


def f32(v) -> float:
    return float(np.float32(v))


def triangle(api):
    return (api.PathBuilder().move_to(Point(PADDING, PADDING)).line_to(Point(WIDTH - PADDING, PADDING))
            .line_to(Point(WIDTH - PADDING, HEIGHT - PADDING)).build())


def custom_square(api, xmin, ymin, xmax, ymax):
    return (api.PathBuilder().move_to(Point(xmin, ymin)).line_to(Point(xmin, ymax)).line_to(Point(xmax, ymax))
            .line_to(Point(xmax, ymin)).build())


def square(api):
    return custom_square(api, PADDING, PADDING, WIDTH - PADDING, HEIGHT - PADDING)


def inner_square(api):
    return custom_square(api, PADDING * 2.0, PADDING * 2.0, WIDTH - PADDING * 2.0, HEIGHT - PADDING * 2.0)


def custom_circle(api, x, y, radius):
    weight = f32(np.sqrt(np.float32(2.0)) / np.float32(2.0))
    return (api.PathBuilder().move_to(Point(x + radius, y))
            .rat_quad_to(Point(x + radius, y - radius), Point(x, y - radius), weight)
            .rat_quad_to(Point(x - radius, y - radius), Point(x - radius, y), weight)
            .rat_quad_to(Point(x - radius, y + radius), Point(x, y + radius), weight)
            .rat_quad_to(Point(x + radius, y + radius), Point(x + radius, y), weight).build())


def circle(api):
    return custom_circle(api, WIDTH * 0.5, HEIGHT * 0.5, WIDTH * 0.5 - PADDING)


def inner_circle(api):
    return custom_circle(api, WIDTH * 0.5, HEIGHT * 0.5, WIDTH * 0.5 - PADDING * 2.0)


RAINBOW = [(1.00, 0.00, 0.00), (1.00, 0.32, 0.00), (0.63, 0.73, 0.02), (0.08, 0.72, 0.07), (0.05, 0.70, 0.69),
           (0.03, 0.58, 0.76), (0.01, 0.21, 0.85), (0.11, 0.01, 0.89), (0.49, 0.00, 0.94), (0.96, 0.00, 0.69),
           (1.00, 0.00, 0.00)]


def _rainbow(gb):
    for r, g, b in RAINBOW:
        gb.color(Color(f32(r), f32(g), f32(b), 1.0))
    return gb.build()


def vertical_rainbow():
    return _rainbow(GradientBuilder(Point(PADDING, 0.0), Point(WIDTH - PADDING, 0.0)))


def horizontal_rainbow():
    return _rainbow(GradientBuilder(Point(0.0, PADDING), Point(0.0, WIDTH - PADDING)))


def solid(color: Color) -> Props:
    return Props(func=Func.Draw(Style(fill=Fill.Solid(color))))


BLUE_WHITE_RED = [Color(0.0, 0.0, 1.0, 1.0), Color(1.0, 1.0, 1.0, 1.0), Color(1.0, 0.0, 0.0, 1.0)]


def linear_gradient(api, comp):
    gb = GradientBuilder(Point(PADDING, 0.0), Point(WIDTH - PADDING, 0.0))
    for c in BLUE_WHITE_RED:
        gb.color(c)
    comp.get_mut_or_insert_default(1).insert(triangle(api)).set_props(
        Props(func=Func.Draw(Style(fill=Fill.Gradient(gb.build())))))


def radial_gradient(api, comp):
    gb = GradientBuilder(Point(WIDTH * 0.5, HEIGHT * 0.5), Point(WIDTH - PADDING * 2.0, HEIGHT * 0.5))
    gb.type(GradientType.Radial)
    for c in BLUE_WHITE_RED:
        gb.color(c)
    comp.get_mut_or_insert_default(1).insert(circle(api)).set_props(
        Props(func=Func.Draw(Style(fill=Fill.Gradient(gb.build())))))


SOLID_COLORS = {
    "blue": Color(0.0, 0.0, 1.0, 1.0), "dark_blue": Color(0.0, 0.0, 0.5, 1.0), "red": Color(1.0, 0.0, 0.0, 1.0),
    "dark_red": Color(0.5, 0.0, 0.0, 1.0), "green": Color(0.0, 1.0, 0.0, 1.0),
    "dark_green": Color(0.0, 0.5, 0.0, 1.0), "transparent_black": Color(0.0, 0.0, 0.0, 0.5),
}


def solid_color(name):
    def build(api, comp):
        comp.get_mut_or_insert_default(1).insert(square(api)).set_props(solid(SOLID_COLORS[name]))
    return build


def pixel(api, comp):
    comp.get_mut_or_insert_default(1).insert(custom_square(api, PADDING, PADDING, PADDING + 1.0, PADDING + 1.0)) \
        .set_props(solid(Color(0.0, 0.0, 0.0, 1.0)))


def covers(api, comp):
    layer = comp.get_mut_or_insert_default(0).set_props(solid(Color(0.0, 0.0, 0.0, 1.0)))
    step = np.float32(2.0) + np.float32(1.0) / np.float32(32.0)
    for xi in range(32):
        for yi in range(32):
            x0 = f32(np.float32(xi) * step)
            y0 = f32(np.float32(yi) * step)
            layer.insert(custom_square(api, x0, y0, f32(np.float32(x0) + np.float32(1.0)),
                                       f32(np.float32(y0) + np.float32(1.0))))


def _srgb_to_linear(u8):
    # Image::from_srgba / to_linear, forma/src/styling.rs:250-258,302-316
    l = np.float32(u8) * (np.float32(1.0) / np.float32(255.0))
    if l <= np.float32(0.04045):
        return np.float32(l * (np.float32(1.0) / np.float32(12.92)))
    return np.float32(math.pow(float((l + np.float32(0.055)) * (np.float32(1.0) / np.float32(1.055))), 2.4))


def texture(api, comp):
    px = [[0, 0, 0, 255], [255, 0, 0, 255], [0, 255, 0, 255], [255, 255, 0, 255], [0, 0, 255, 255],
          [255, 0, 255, 255], [0, 255, 255, 255], [255, 255, 255, 255], [0, 0, 0, 255]]
    img = np.zeros((3, 3, 4), np.float32)
    for i, p in enumerate(px):
        img[i // 3, i % 3, :3] = [_srgb_to_linear(c) for c in p[:3]]
        img[i // 3, i % 3, 3] = np.float32(p[3]) * (np.float32(1.0) / np.float32(255.0))
    order = 0
    for xi in range(8):
        for yi in range(8):
            x0, y0 = xi * 8.0, yi * 8.0
            tx = -x0 - 2.0 + xi * 0.25
            ty = -y0 - 2.0 + yi * 0.25
            comp.get_mut_or_insert_default(order).insert(custom_square(api, x0, y0, x0 + 7.0, y0 + 7.0)).set_props(
                Props(fill_rule=FillRule.EvenOdd,
                      func=Func.Draw(Style(fill=Fill.Texture(Texture((1.0, 0.0, 0.0, 1.0, tx, ty), img))))))
            order += 1


def blend_modes(mode):
    def build(api, comp):
        comp.get_mut_or_insert_default(0).insert(square(api)).set_props(
            Props(func=Func.Draw(Style(fill=Fill.Gradient(horizontal_rainbow())))))
        comp.get_mut_or_insert_default(1).insert(triangle(api)).set_props(
            Props(func=Func.Draw(Style(fill=Fill.Gradient(vertical_rainbow()), blend_mode=mode))))
    return build


def fill_rules(rule):
    def build(api, comp):
        path = (api.PathBuilder().move_to(Point(PADDING, PADDING))
                .line_to(Point(WIDTH / 2.0 + PADDING, HEIGHT / 2.0 + PADDING))
                .line_to(Point(WIDTH / 2.0 - PADDING, HEIGHT / 2.0 + PADDING))
                .line_to(Point(WIDTH - PADDING, PADDING)).line_to(Point(WIDTH - PADDING, HEIGHT - PADDING))
                .line_to(Point(PADDING, HEIGHT - PADDING)).build())
        comp.get_mut_or_insert_default(0).insert(path).set_props(
            Props(fill_rule=rule, func=Func.Draw(Style(fill=Fill.Solid(Color(0.0, 0.0, 0.0, f32(0.8)))))))
    return build


def clipping(api, comp):
    comp.get_mut_or_insert_default(0).insert(square(api)).set_props(solid(Color(0.0, 0.0, 0.0, f32(0.7))))
    comp.get_mut_or_insert_default(1).insert(triangle(api)).set_props(Props(func=Func.Clip(4)))
    comp.get_mut_or_insert_default(2).insert(square(api)).set_props(
        Props(func=Func.Draw(Style(fill=Fill.Solid(Color(0.5, 0.5, 1.0, f32(0.7))), is_clipped=True))))
    comp.get_mut_or_insert_default(4).insert(circle(api)).set_props(
        Props(func=Func.Draw(Style(fill=Fill.Solid(Color(1.0, 0.5, 0.5, f32(0.7)))))))
    comp.get_mut_or_insert_default(5).insert(inner_square(api)).set_props(
        Props(func=Func.Draw(Style(fill=Fill.Solid(Color(0.5, 0.5, 1.0, f32(0.6))), is_clipped=True))))
    comp.get_mut_or_insert_default(6).insert(inner_circle(api)).set_props(
        Props(func=Func.Draw(Style(fill=Fill.Solid(Color(0.5, 1.0, 0.5, f32(0.6))), is_clipped=True))))


def clipping2(api, comp):
    comp.get_mut_or_insert_default(0).insert(square(api)).set_props(solid(Color(0.0, 0.0, 0.0, f32(0.7))))
    comp.get_mut_or_insert_default(1).insert(inner_circle(api)).set_props(Props(func=Func.Clip(1)))
    comp.get_mut_or_insert_default(2).insert(triangle(api)).set_props(
        Props(func=Func.Draw(Style(fill=Fill.Solid(Color(0.5, 0.5, 1.0, f32(0.7))), is_clipped=True))))


# name in tests/golden/e2e_expected.npz (without the __cpu suffix) -> builder
E2E = {"linear_gradient": linear_gradient, "radial_gradient": radial_gradient, "pixel": pixel, "covers": covers,
       "texture": texture, "clipping": clipping, "clipping2": clipping2}
for _n in SOLID_COLORS:
    E2E[f"solid_color__{_n}"] = solid_color(_n)
for _i, _n in enumerate(BlendMode.NAMES):
    E2E[f"blend_modes__{_n}"] = blend_modes(_i)
E2E["fill_rules__EvenOdd"] = fill_rules(FillRule.EvenOdd)
E2E["fill_rules__NonZero"] = fill_rules(FillRule.NonZero)

# Hue/Saturation/Color/Luminosity use `recip`, which the reference's x86 build
# evaluates with the ~12-bit _mm256_rcp_ps, so the goldens
# (rendered on an unknown SIMD path) cannot be matched bit-exactly there.
NON_SEPARABLE = {"blend_modes__Hue", "blend_modes__Saturation", "blend_modes__Color", "blend_modes__Luminosity"}


def render_e2e(api, name, renderer=None, size=64):
    """Renders one e2e scene like cpu_render in test_env.rs:40-59."""
    from forma_b200.binding import RGBA
    comp = api.Composition()
    E2E[name](api, comp)
    r = renderer if renderer is not None else api.Renderer()
    buf = np.zeros(size * size * 4, np.uint8)
    r.render(comp, buf, size, size, RGBA, E2E_CLEAR)
    return buf.reshape(size, size, 4)
