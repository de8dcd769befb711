"""Layer-cache (damage reuse) scenarios restating the reference's tests
forma/src/composition/mod.rs:520-563,1038-1382. Each scenario drives one API
(`api`) and returns the list of frame buffers plus the expectations the
reference asserts; the same functions run against the CPU oracle (pinning it)
and against the CUDA library (parity)."""
import numpy as np

from forma_b200.binding import RGBA, Color, Fill, Func, Point, Props, Style

T = 16
BLACK_SRGB, RED_SRGB, GREEN_SRGB = [0, 0, 0, 255], [255, 0, 0, 255], [0, 255, 0, 255]
BLACK, RED, GREEN = Color(0, 0, 0, 1), Color(1, 0, 0, 1), Color(0, 1, 0, 1)


def solid(c):
    return Props(func=Func.Draw(Style(fill=Fill.Solid(c))))


def pixel_path(api, x, y):
    return (api.PathBuilder().move_to(Point(x, y)).line_to(Point(x, y + 1)).line_to(Point(x + 1, y + 1))
            .line_to(Point(x + 1, y)).line_to(Point(x, y)).build())


class Frames:
    def __init__(self, api):
        self.api, self.r, self.shots = api, api.Renderer(), []

    def render(self, comp, w, h, clear, cache, prefill=BLACK_SRGB, buf=None):
        if buf is None:
            buf = np.array(prefill * (w * h), np.uint8)
        self.r.render(comp, buf, w, h, RGBA, clear, None, cache)
        self.shots.append(buf.copy())
        return buf.reshape(h, w, 4)


def px(img, x, y=0):
    return img[y, x].tolist()


def background_color_clear_when_changed(api):  # mod.rs:520-563
    f = Frames(api)
    comp = api.Composition()
    cache = f.r.create_buffer_layer_cache()
    img = f.render(comp, 1, 1, RED, cache, GREEN_SRGB)
    assert px(img, 0) == RED_SRGB
    img = f.render(comp, 1, 1, RED, cache, GREEN_SRGB)
    assert px(img, 0) == GREEN_SRGB      # unchanged tile: not written
    img = f.render(comp, 1, 1, BLACK, cache, GREEN_SRGB)
    assert px(img, 0) == BLACK_SRGB      # clear colour changed
    return f.shots


def render_changed_layers_only(api):  # mod.rs:1038-1105
    f = Frames(api)
    comp = api.Composition()
    cache = f.r.create_buffer_layer_cache()
    comp.insert(0, comp.create_layer().insert(pixel_path(api, 0, 0)).insert(pixel_path(api, T, 0)).set_props(solid(RED)))
    comp.insert(1, comp.create_layer().insert(pixel_path(api, T + 1, 0)).insert(pixel_path(api, 2 * T, 0))
                .set_props(solid(GREEN)))
    img = f.render(comp, 3 * T, T, BLACK, cache)
    assert px(img, 0) == RED_SRGB and px(img, T) == RED_SRGB and px(img, T + 1) == GREEN_SRGB and px(img, 2 * T) == GREEN_SRGB
    comp.get(1).set_props(solid(RED))
    img = f.render(comp, 3 * T, T, BLACK, cache)
    assert px(img, 0) == BLACK_SRGB      # first tile untouched
    assert px(img, T) == RED_SRGB and px(img, T + 1) == RED_SRGB and px(img, 2 * T) == RED_SRGB
    return f.shots


def insert_remove_same_order_will_not_render_again(api):  # mod.rs:1108-1150
    f = Frames(api)
    comp = api.Composition()
    cache = f.r.create_buffer_layer_cache()
    comp.insert(0, comp.create_layer().insert(pixel_path(api, 0, 0)).set_props(solid(RED)))
    img = f.render(comp, 3, 1, BLACK, cache)
    assert img[0].tolist() == [RED_SRGB, BLACK_SRGB, BLACK_SRGB]
    layer = comp.remove(0)
    comp.insert(0, layer)
    img = f.render(comp, 3, 1, BLACK, cache)
    assert img[0].tolist() == [BLACK_SRGB, BLACK_SRGB, BLACK_SRGB]
    return f.shots


def clear_emptied_tiles(api):  # mod.rs:1153-1228
    f = Frames(api)
    comp = api.Composition()
    cache = f.r.create_buffer_layer_cache()
    layer = comp.create_layer()
    layer.insert(pixel_path(api, 0, 0)).set_props(solid(RED)).insert(pixel_path(api, T, 0))
    comp.insert(0, layer)
    buf = np.array(BLACK_SRGB * (2 * T * T), np.uint8)
    img = f.render(comp, 2 * T, T, BLACK, cache, buf=buf)
    assert px(img, 0) == RED_SRGB
    for t, want in (([1.0, 0.0, 0.0, 1.0, float(T), 0.0], BLACK_SRGB), ([1.0, 0.0, 0.0, 1.0, -float(T), 0.0], RED_SRGB),
                    ([1.0, 0.0, 0.0, 1.0, 0.0, float(T)], BLACK_SRGB)):
        comp.get(0).set_transform(t)
        img = f.render(comp, 2 * T, T, BLACK, cache, buf=buf)
        assert px(img, 0) == want
    return f.shots


def separate_layer_caches(api):  # mod.rs:1232-1316
    f = Frames(api)
    comp = api.Composition()
    c0, c1 = f.r.create_buffer_layer_cache(), f.r.create_buffer_layer_cache()
    comp.insert(0, comp.create_layer().insert(pixel_path(api, 0, 0)).set_props(solid(RED)))
    assert px(f.render(comp, T, T, BLACK, c0), 0) == RED_SRGB
    buf = np.array(BLACK_SRGB * (T * T), np.uint8)
    assert px(f.render(comp, T, T, BLACK, c0, buf=buf), 0) == BLACK_SRGB
    assert px(f.render(comp, T, T, BLACK, c1, buf=buf), 0) == RED_SRGB
    comp.get(0).set_transform([1.0, 0.0, 0.0, 1.0, 1.0, 0.0])
    img = f.render(comp, T, T, BLACK, c0, buf=buf)
    assert px(img, 0) == BLACK_SRGB and px(img, 1) == RED_SRGB
    img = f.render(comp, T, T, BLACK, c1)
    assert px(img, 0) == BLACK_SRGB and px(img, 1) == RED_SRGB
    return f.shots


def draw_if_width_or_height_change(api):  # mod.rs:1320-1382
    f = Frames(api)
    comp = api.Composition()
    cache = f.r.create_buffer_layer_cache()
    assert px(f.render(comp, 1, 1, RED, cache), 0) == RED_SRGB
    assert px(f.render(comp, 1, 1, RED, cache), 0) == BLACK_SRGB
    assert f.render(comp, 2, 1, RED, cache)[0].tolist() == [RED_SRGB, RED_SRGB]
    assert f.render(comp, 1, 2, RED, cache)[:, 0].tolist() == [RED_SRGB, RED_SRGB]
    return f.shots


def animated_scene(api, frames=6, w=200, h=120):
    """Not from the reference: a small animation with a persistent cache; only
    used to compare the CUDA library with the oracle frame by frame."""
    import synth
    f = Frames(api)
    comp = api.Composition()
    cache = f.r.create_buffer_layer_cache()
    synth.random_mixed(api, comp, 60, w, h, 123)
    buf = np.full(w * h * 4, 0x33, np.uint8)
    for i in range(frames):
        if i == 1:
            comp.get(5).set_transform([1.0, 0.0, 0.0, 1.0, 7.0, 3.0])
        if i == 2:
            comp.get(9).disable()
            comp.get(20).set_props(solid(Color(0.2, 0.4, 0.9, 1.0)))
        if i == 3:
            comp.get(9).enable()
            comp.remove(30)
        if i == 4:
            comp.get(12).clear()
            comp.get(12).insert(synth.circle_path(api, 100.0, 60.0, 25.0))
        f.r.render(comp, buf, w, h, RGBA, Color(0.9, 0.9, 0.9, 1.0), None, cache)
        f.shots.append(buf.copy())
    return f.shots


def crop_with_layers_left_of_the_crop(api):
    """Not a reference test, but the reference's rule (cpu/painter/mod.rs:501-522): with a crop,
    every layer that has segments LEFT of the first cropped tile is queued for that tile even when
    its cover sums to zero there, and is counted in the tile's cached layer count. Removing such a
    layer therefore damages the first cropped tile of its tile row (and nothing else). The frames
    carry a prefill, so which tiles a frame wrote is part of what is compared with the oracle."""
    from forma_b200.binding import Point, Rect
    f = Frames(api)
    comp = api.Composition()
    cache = f.r.create_buffer_layer_cache()
    w, h = 4 * T, 2 * T
    box = lambda x0, y0, x1, y1: (api.PathBuilder().move_to(Point(x0, y0)).line_to(Point(x0, y1)).line_to(Point(x1, y1))
                                  .line_to(Point(x1, y0)).build())
    comp.insert(0, comp.create_layer().insert(box(0.0, 0.0, float(w), float(h))).set_props(solid(GREEN)))   # covers everything
    comp.insert(1, comp.create_layer().insert(box(3.0, 3.0, 9.0, 12.0)).set_props(solid(RED)))             # closed, inside tile (0, 0)
    comp.insert(2, comp.create_layer().insert(box(20.0, T + 2.0, 2.5 * T, T + 9.0)).set_props(solid(RED)))  # row 1: reaches into the crop
    crop = Rect((2 * T, 4 * T), (0, 2 * T))
    buf = np.array([0x11, 0x22, 0x33, 0x44] * (w * h), np.uint8)

    def frame():
        f.r.render(comp, buf, w, h, RGBA, BLACK, crop, cache)
        f.shots.append(buf.copy())
        buf[:] = np.array([0x11, 0x22, 0x33, 0x44] * (w * h), np.uint8)
        return f.shots[-1].reshape(h, w, 4)
    img = frame()
    assert px(img, 0) == [0x11, 0x22, 0x33, 0x44]            # left of the crop: untouched
    assert px(img, 2 * T) == GREEN_SRGB and px(img, 2 * T + 2, T + 4) == RED_SRGB
    img = frame()
    assert (img == np.array([0x11, 0x22, 0x33, 0x44], np.uint8)).all()   # nothing changed: nothing written
    comp.remove(1)                                             # the zero-cover layer left of the crop goes away
    frame()
    comp.remove(2)                                             # and the one that really reaches into the crop
    img = frame()
    assert px(img, 2 * T + 2, T + 4) == GREEN_SRGB
    frame()
    return f.shots


SCENARIOS = [background_color_clear_when_changed, render_changed_layers_only,
             insert_remove_same_order_will_not_render_again, clear_emptied_tiles, separate_layer_caches,
             draw_if_width_or_height_change, crop_with_layers_left_of_the_crop]
