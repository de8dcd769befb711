"""SVG subset loader — restates the reference's demo loader
(/root/reference/demo/src/demos/svg.rs:27-116,193-335,337-863):

* elements: <g transform= fill= opacity=>, <path d= fill= fill-opacity= opacity= fill-rule= style=>,
  <rect x= y= width= height=> (svg.rs:700-737), <linearGradient> / <radialGradient> with
  gradientUnits="userSpaceOnUse" and their <stop offset="N%" stop-color= stop-opacity=> children
  (svg.rs:738-857; other gradient units are ignored like the reference does), `fill="url(#id)"`
* path commands: M m L l H h V v C c S s Q q T t A a Z z (relative coordinates are accumulated in
  f32 like `add_diff`, svg.rs:371-373); elliptical arcs become rational quadratics of at most a
  quarter turn each (convert_to_center + push_rationals_from_arc, svg.rs:42-116,276-335 — including
  the reference's shadowed `x0` in the endpoint normalisation)
* the innermost group transform is applied in f64 and rounded to f32 (svg.rs:237-247)
* per-<path> `transform=` attributes are ignored, exactly like the reference
* colours: sRGB hex -> linear (demo/src/main.rs:134-151); opacity precedence `opacity`, then
  `fill-opacity`, else the product of the group opacities (svg.rs:129-139,258)
* `style="mix-blend-mode: ..."` selects the layer's blend mode (svg.rs:147-174)

The result is a flat `PathList` (numpy arrays) that can be inserted into any Composition-like
object; `tests/golden/make_paris_fixture.py` stores the list for paris-30k.svg so that the GPU
box does not need the SVG itself. `python -m forma_b200.render` is the headless front end.
"""
from __future__ import annotations

import re
import xml.etree.ElementTree as ET
from dataclasses import dataclass

import numpy as np

import math

from .binding import BlendMode, Color, Fill, FillRule, Func, GradientBuilder, GradientType, Point, Props, Style

F = np.float32

# command codes of PathList.cmd
MOVE, LINE, QUAD, CUBIC, RATQUAD = 0, 1, 2, 3, 4   # RATQUAD: two points + one weight (PathList.weights, in order)
_NPTS = {MOVE: 1, LINE: 1, QUAD: 2, CUBIC: 3, RATQUAD: 2}

_BLEND = {"normal": BlendMode.Over, "multiply": BlendMode.Multiply, "screen": BlendMode.Screen, "overlay": BlendMode.Overlay,
          "darken": BlendMode.Darken, "lighten": BlendMode.Lighten, "color-dodge": BlendMode.ColorDodge,
          "color-burn": BlendMode.ColorBurn, "hard-light": BlendMode.HardLight, "soft-light": BlendMode.SoftLight,
          "difference": BlendMode.Difference, "exclusion": BlendMode.Exclusion, "hue": BlendMode.Hue,
          "saturation": BlendMode.Saturation, "color": BlendMode.Color, "luminosity": BlendMode.Luminosity}


@dataclass
class PathList:
    cmd: np.ndarray        # uint8, one per path command
    pts: np.ndarray        # float32 (n, 2): the points consumed by the commands, in order
    cmd_off: np.ndarray    # int64 (n_paths + 1): command range of every path
    pt_off: np.ndarray     # int64 (n_paths + 1)
    color: np.ndarray      # float32 (n_paths, 4) linear RGBA (solid fills)
    fill_rule: np.ndarray  # uint8 (n_paths)
    weights: np.ndarray = None     # float32: one per RATQUAD command, in command order
    blend: np.ndarray = None       # uint8 (n_paths): BlendMode
    grad: np.ndarray = None        # int32 (n_paths): index into `gradients`, -1 = solid fill
    gradients: list = None         # dicts {type, start, end, stops: [((r, g, b, a) linear, stop)]}

    def __post_init__(self):
        n = len(self.color)
        if self.weights is None:
            self.weights = np.zeros(0, np.float32)
        if self.blend is None:
            self.blend = np.zeros(n, np.uint8)
        if self.grad is None:
            self.grad = np.full(n, -1, np.int32)
        if self.gradients is None:
            self.gradients = []

    def __len__(self):
        return len(self.color)

    def save(self, path):
        g = self.gradients
        np.savez_compressed(
            path, cmd=self.cmd, pts=self.pts, cmd_off=self.cmd_off, pt_off=self.pt_off, color=self.color,
            fill_rule=self.fill_rule, weights=self.weights, blend=self.blend, grad=self.grad,
            grad_type=np.array([x["type"] for x in g], np.uint8),
            grad_pts=np.array([list(x["start"]) + list(x["end"]) for x in g], np.float32).reshape(-1, 4),
            grad_stop_off=np.cumsum([0] + [len(x["stops"]) for x in g]).astype(np.int64),
            grad_stop_color=np.array([c for x in g for c, _ in x["stops"]], np.float32).reshape(-1, 4),
            grad_stop_pos=np.array([p for x in g for _, p in x["stops"]], np.float32))

    @staticmethod
    def load(path) -> "PathList":
        z = np.load(path)
        out = PathList(z["cmd"], z["pts"], z["cmd_off"], z["pt_off"], z["color"], z["fill_rule"])
        if "weights" in z.files:  # fixtures written before arcs / gradients existed lack these
            out.weights, out.blend, out.grad = z["weights"], z["blend"], z["grad"]
            off = z["grad_stop_off"]
            for i in range(len(z["grad_type"])):
                stops = [(tuple(float(v) for v in z["grad_stop_color"][k]), float(z["grad_stop_pos"][k]))
                         for k in range(int(off[i]), int(off[i + 1]))]
                p = z["grad_pts"][i]
                out.gradients.append({"type": int(z["grad_type"][i]), "start": (float(p[0]), float(p[1])),
                                      "end": (float(p[2]), float(p[3])), "stops": stops})
        return out


def to_linear(u8: int) -> float:
    """demo/src/main.rs:134-151"""
    l = F(u8) * (F(1.0) / F(255.0))
    if l <= F(0.04045):
        return float(F(l * (F(1.0) / F(12.92))))
    return float(np.power(F((l + F(0.055)) * (F(1.0) / F(1.055))), F(2.4), dtype=F))


def _parse_color(s):
    s = s.strip()
    if s.startswith("#"):
        h = s[1:]
        if len(h) == 3:
            h = "".join(c * 2 for c in h)
        if len(h) == 6:
            return tuple(int(h[i:i + 2], 16) for i in (0, 2, 4))
    return None


_NUM = re.compile(r"[-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?")
_TOK = re.compile(r"([MmLlHhVvCcSsQqTtZzAa])|([-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?)")
_ARGS = {"m": 2, "l": 2, "h": 1, "v": 1, "c": 6, "s": 4, "q": 4, "t": 2, "a": 7, "z": 0}


def _parse_transform(s):
    """`matrix(a b c d e f)` / translate / scale -> (a, b, c, d, e, f) in f64."""
    m = re.match(r"\s*(\w+)\s*\(([^)]*)\)", s or "")
    if not m:
        return None
    v = [float(x) for x in _NUM.findall(m.group(2))]
    kind = m.group(1)
    if kind == "matrix" and len(v) == 6:
        return tuple(v)
    if kind == "translate":
        return (1.0, 0.0, 0.0, 1.0, v[0], v[1] if len(v) > 1 else 0.0)
    if kind == "scale":
        return (v[0], 0.0, 0.0, v[1] if len(v) > 1 else v[0], 0.0, 0.0)
    return None


def convert_to_center(rx, ry, x_axis_rotation, large_arc, sweep, x0, y0, x1, y1):
    """svg.rs:42-116, in f32 like the reference. Returns (cx, cy, rx, ry, phi, angle, angle_delta)
    or None. Note the reference's shadowing: the normalised y0 / y1 are computed from the
    already normalised x0 / x1 (`let x0 = ...; let y0 = (-x0 * sin_phi + ...`)."""
    eps = float(np.finfo(np.float32).eps)
    rx, ry, phi, x0, y0, x1, y1 = (F(v) for v in (rx, ry, x_axis_rotation, x0, y0, x1, y1))
    if abs(F(x0 - x1)) < eps and abs(F(y0 - y1)) < eps:
        return None
    rx, ry = F(abs(rx)), F(abs(ry))
    if rx == 0.0 or ry == 0.0:
        return None
    cos_phi, sin_phi = F(math.cos(float(phi))), F(math.sin(float(phi)))  # f32::cos / sin: correctly rounded libm assumed
    x0 = F(F(F(x0 * cos_phi) + F(y0 * sin_phi)) / rx)
    y0 = F(F(F(-x0 * sin_phi) + F(y0 * cos_phi)) / ry)
    x1 = F(F(F(x1 * cos_phi) + F(y1 * sin_phi)) / rx)
    y1 = F(F(F(-x1 * sin_phi) + F(y1 * cos_phi)) / ry)
    lx, ly = F(F(x0 - x1) * F(0.5)), F(F(y0 - y1) * F(0.5))
    cx, cy = F(F(x0 + x1) * F(0.5)), F(F(y0 + y1) * F(0.5))
    len_squared = F(F(lx * lx) + F(ly * ly))
    if len_squared < 1.0:
        radicand = F(np.sqrt(F(F(F(1.0) - len_squared) / len_squared)))
        if large_arc != sweep:
            radicand = F(-radicand)
        cx = F(cx + F(-ly * radicand))
        cy = F(cy + F(lx * radicand))
    theta = F(math.atan2(float(F(y0 - cy)), float(F(x0 - cx))))
    delta_theta = F(F(math.atan2(float(F(y1 - cy)), float(F(x1 - cx)))) - theta)
    cxs, cys = F(cx * rx), F(cy * ry)
    cx = F(F(cxs * cos_phi) - F(cys * sin_phi))
    cy = F(F(cxs * sin_phi) + F(cys * cos_phi))
    two_pi = F(F(math.pi) * F(2.0))
    if sweep:
        if delta_theta < 0.0:
            delta_theta = F(delta_theta + two_pi)
    elif delta_theta > 0.0:
        delta_theta = F(delta_theta - two_pi)
    return cx, cy, rx, ry, phi, theta, delta_theta


def arc_to_rational_quads(arc):
    """push_rationals_from_arc (svg.rs:276-335): [(p1, p2, weight)], each piece at most a quarter
    turn, points before the group transform. Returns the pieces and the new end point."""
    cx, cy, rx, ry, phi, angle, angle_delta = arc
    cos_phi, sin_phi = F(math.cos(float(phi))), F(math.sin(float(phi)))
    angle_sweep = F(F(math.pi) / F(2.0))
    angle_incr = angle_sweep if angle_delta > 0.0 else F(-angle_sweep)
    out, end = [], None
    while angle_delta != 0.0:
        theta = angle
        sweep = angle_delta if abs(angle_delta) <= angle_sweep else angle_incr
        angle = F(angle + sweep)
        angle_delta = F(angle_delta - sweep)
        half = F(sweep * F(0.5))
        w = F(math.cos(float(half)))
        mid, far = F(theta + half), F(theta + sweep)
        p1x, p1y = F(F(math.cos(float(mid))) / w), F(F(math.sin(float(mid))) / w)
        p2x, p2y = F(math.cos(float(far))), F(math.sin(float(far)))
        p1x, p1y, p2x, p2y = F(p1x * rx), F(p1y * ry), F(p2x * rx), F(p2y * ry)
        p1 = (F(F(cx + F(p1x * cos_phi)) - F(p1y * sin_phi)), F(F(cy + F(p1x * sin_phi)) + F(p1y * cos_phi)))
        p2 = (F(F(cx + F(p2x * cos_phi)) - F(p2y * sin_phi)), F(F(cy + F(p2x * sin_phi)) + F(p2y * cos_phi)))
        out.append((p1, p2, w))
        end = p2
    return out, end


def _blend_of(attrib):
    """parse_blend_mode, svg.rs:147-174: `style="...; mix-blend-mode: X; ..."`."""
    for pair in (attrib.get("style") or "").split(";"):
        if ":" in pair:
            k, v = pair.split(":", 1)
            if k.strip() == "mix-blend-mode":
                return _BLEND.get(v.strip(), BlendMode.Over)
    return BlendMode.Over


def parse_svg(path: str) -> PathList:
    cmds, pts, cmd_off, pt_off, colors, rules = [], [], [0], [0], [], []
    weights, blends, grad_of, gradients, gradient_ids = [], [], [], [], {}
    open_gradient = [None]  # the <linearGradient> / <radialGradient> being read
    groups = []  # dicts: transform, fill, opacity

    def t(x: F, y: F):
        for g in reversed(groups):
            if g["transform"] is not None:
                a, b, c, d, e, f = g["transform"]
                xx, yy = float(x), float(y)
                return F(a * xx + c * yy + e), F(b * xx + d * yy + f)
        return x, y

    def emit(code, *points):
        cmds.append(code)
        for (x, y) in points:
            pts.append(t(x, y))

    def opacity_of(attrib):
        for key in ("opacity", "fill-opacity"):
            if key in attrib:
                try:
                    return F(float(attrib[key]))
                except ValueError:
                    pass
        return None

    def handle_path(attrib):
        if attrib.get("stroke", "none") != "none" or "d" not in attrib:
            return
        end = (F(0.0), F(0.0))
        start = None
        quad_cp = cubic_cp = None
        cur = None
        args = []
        n_cmd0, n_pt0 = len(cmds), len(pts)

        def run(c, a):
            nonlocal end, start, quad_cp, cubic_cp
            rel = c.islower()
            lc = c.lower()

            def P(i):
                x, y = F(a[i]), F(a[i + 1])
                return (F(end[0] + x), F(end[1] + y)) if rel else (x, y)
            if lc == "m":
                p = P(0)
                emit(MOVE, p)
                start, end, quad_cp, cubic_cp = None, p, None, None
            elif lc in "lhv":
                if lc == "l":
                    p = P(0)
                elif lc == "h":
                    p = (F(end[0] + F(a[0])) if rel else F(a[0]), end[1])
                else:
                    p = (end[0], F(end[1] + F(a[0])) if rel else F(a[0]))
                emit(LINE, p)
                start = start if start is not None else end
                end, quad_cp, cubic_cp = p, None, None
            elif lc == "c":
                p0, p1, p2 = P(0), P(2), P(4)
                emit(CUBIC, p0, p1, p2)
                start = start if start is not None else end
                end, quad_cp, cubic_cp = p2, None, p1
            elif lc == "s":
                p1, p2 = P(0), P(2)
                ref = cubic_cp if cubic_cp is not None else end
                cp = (F(end[0] * F(2.0) - ref[0]), F(end[1] * F(2.0) - ref[1]))  # reflect, svg.rs:27-29
                emit(CUBIC, cp, p1, p2)
                start = start if start is not None else end
                # the reference remembers the *reflected* point (svg.rs:578,599)
                end, quad_cp, cubic_cp = p2, None, cp
            elif lc == "q":
                p0, p1 = P(0), P(2)
                emit(QUAD, p0, p1)
                start = start if start is not None else end
                end, quad_cp, cubic_cp = p1, p0, None
            elif lc == "t":
                p1 = P(0)
                ref = quad_cp if quad_cp is not None else end
                cp = (F(end[0] * F(2.0) - ref[0]), F(end[1] * F(2.0) - ref[1]))
                emit(QUAD, cp, p1)
                start = start if start is not None else end
                end, quad_cp, cubic_cp = p1, cp, None
            elif lc == "a":
                p0 = P(5)
                arc = convert_to_center(a[0], a[1], a[2], a[3] != 0.0, a[4] != 0.0, end[0], end[1], p0[0], p0[1])
                if arc is not None:
                    pieces, new_end = arc_to_rational_quads(arc)
                    for p1, p2, w in pieces:
                        emit(RATQUAD, p1, p2)
                        weights.append(float(w))
                    start = start if start is not None else end
                    if new_end is not None:
                        end = new_end
                quad_cp, cubic_cp = None, None
            elif lc == "z":
                if start is not None:
                    end, start, quad_cp, cubic_cp = start, None, None, None
            else:
                raise ValueError(f"unsupported path command {c!r}")

        for m in _TOK.finditer(attrib["d"]):
            if m.group(1):
                if cur is not None and _ARGS.get(cur.lower(), 0) == 0:
                    run(cur, [])
                cur, args = m.group(1), []
                if cur.lower() == "z":
                    run(cur, [])
                    cur = None
            else:
                args.append(float(m.group(2)))
                need = _ARGS[cur.lower()]
                if len(args) == need:
                    run(cur, args)
                    args = []
                    if cur == "m":
                        cur = "l"  # implicit line-to after a move-to
                    elif cur == "M":
                        cur = "L"
        finish_shape(attrib)
        del n_cmd0, n_pt0

    def finish_shape(attrib):
        """parse_fill (svg.rs:249-274), parse_fill_rule, parse_blend_mode; closes the path's ranges."""
        fill = attrib.get("fill", "").strip()
        gid = -1
        if fill.startswith("url(#") and fill.endswith(")"):
            gid = gradient_ids.get(fill[5:-1], -1)
        rgb = _parse_color(attrib["fill"]) if "fill" in attrib else None
        if rgb is None:
            for g in reversed(groups):
                if g["fill"] is not None:
                    rgb = g["fill"]
                    break
        op = opacity_of(attrib)
        if op is None:
            op = F(1.0)
            for g in groups:
                if g["opacity"] is not None:
                    op = F(op * g["opacity"])
        if rgb is None:
            colors.append((0.0, 0.0, 0.0, 1.0))
        else:
            colors.append((to_linear(rgb[0]), to_linear(rgb[1]), to_linear(rgb[2]), float(op)))
        grad_of.append(gid)
        blends.append(_blend_of(attrib))
        rules.append(1 if attrib.get("fill-rule") == "evenodd" else 0)
        cmd_off.append(len(cmds))
        pt_off.append(len(pts))

    def handle_rect(attrib):  # svg.rs:700-737 (no group transform is applied to rectangles there either)
        if attrib.get("stroke", "none") != "none":
            return

        def num(key, default=None):
            try:
                return F(float(attrib[key]))
            except (KeyError, ValueError):
                if default is None:
                    raise ValueError(f"rect missing {key}")
                return F(default)
        x, y, w, h = num("x", 0.0), num("y", 0.0), num("width"), num("height")
        for code, p in ((MOVE, (x, y)), (LINE, (x, F(y + h))), (LINE, (F(x + w), F(y + h))), (LINE, (F(x + w), y)), (LINE, (x, y))):
            cmds.append(code)
            pts.append(p)
        finish_shape(attrib)

    def handle_gradient_start(tag, attrib):  # svg.rs:738-771,787-823
        if attrib.get("gradientUnits") != "userSpaceOnUse":
            return
        if tag == "linearGradient":
            x1, y1, x2, y2 = (F(float(attrib[k])) for k in ("x1", "y1", "x2", "y2"))
            open_gradient[0] = {"id": attrib["id"], "type": GradientType.Linear, "start": (float(x1), float(y1)),
                                "end": (float(x2), float(y2)), "stops": []}
        else:
            cx, cy, r = (F(float(attrib[k])) for k in ("cx", "cy", "r"))
            open_gradient[0] = {"id": attrib["id"], "type": GradientType.Radial, "start": (float(cx), float(cy)),
                                "end": (float(F(cx + r)), float(cy)), "stops": []}

    def handle_stop(attrib):  # svg.rs:824-855
        g = open_gradient[0]
        if g is None:
            return
        rgb = _parse_color(attrib.get("fill") or attrib.get("stop-color") or "") or (0, 0, 0)
        op = None
        for key in ("opacity", "stop-opacity", "fill-opacity"):
            if key in attrib:
                try:
                    op = F(float(attrib[key]))
                    break
                except ValueError:
                    pass
        off = attrib.get("offset", "")
        stop = F(F(float(off[:-1])) / F(100.0))  # "N%": the reference drops the last character
        g["stops"].append(((to_linear(rgb[0]), to_linear(rgb[1]), to_linear(rgb[2]), float(F(1.0) if op is None else op)), float(stop)))

    def handle_gradient_end():
        g = open_gradient[0]
        open_gradient[0] = None
        if g is None:
            return
        if len(g["stops"]) < 2:
            raise ValueError("a gradient requires at least 2 stops")
        gradient_ids[g.pop("id")] = len(gradients)
        gradients.append(g)

    for event, el in ET.iterparse(path, events=("start", "end")):
        tag = el.tag.rsplit("}", 1)[-1]
        if event == "start" and tag == "g":
            groups.append({"transform": _parse_transform(el.attrib.get("transform")),
                           "fill": _parse_color(el.attrib["fill"]) if "fill" in el.attrib else None,
                           "opacity": opacity_of(el.attrib)})
        elif event == "end" and tag == "g":
            groups.pop()
        elif event == "end" and tag == "path":
            handle_path(el.attrib)
            el.clear()
        elif event == "end" and tag == "rect":
            handle_rect(el.attrib)
        elif event == "start" and tag in ("linearGradient", "radialGradient"):
            handle_gradient_start(tag, el.attrib)
        elif event == "end" and tag in ("linearGradient", "radialGradient"):
            handle_gradient_end()
        elif event == "end" and tag == "stop":
            handle_stop(el.attrib)
    return PathList(np.array(cmds, np.uint8), np.array(pts, np.float32).reshape(-1, 2), np.array(cmd_off, np.int64),
                    np.array(pt_off, np.int64), np.array(colors, np.float32).reshape(-1, 4), np.array(rules, np.uint8),
                    np.array(weights, np.float32), np.array(blends, np.uint8), np.array(grad_of, np.int32), gradients)


def _gradient_of(g):
    gb = GradientBuilder(Point(*g["start"]), Point(*g["end"])).type(g["type"])
    for (r, gg, b, a), stop in g["stops"]:
        gb.color_with_stop(Color(r, gg, b, a), stop)
    return gb.build()


def compose(api, comp, paths: PathList, scale: float = 1.0, first_order: int = 0, limit=None, fill_of=None):
    """Svg::new + Svg::compose (svg.rs:193-213,904-920): every path becomes one
    layer, document order = layer order; `scale` goes through Path::transform
    like the demo's --scale (control points are transformed and re-flattened).
    Gradient coordinates are NOT scaled, like in the reference (user space of the file)."""
    n = len(paths) if limit is None else min(limit, len(paths))
    m = [scale, 0.0, 0.0, 0.0, scale, 0.0, 0.0, 0.0, 1.0]
    cmd, pts = paths.cmd, paths.pts
    has_rat = bool(len(paths.weights))
    wi = 0
    built = [None] * len(paths.gradients)
    for i in range(n):
        c0, c1, p0 = int(paths.cmd_off[i]), int(paths.cmd_off[i + 1]), int(paths.pt_off[i])
        pb = api.PathBuilder()
        if has_rat and (cmd[c0:c1] == RATQUAD).any():
            k = p0
            for c in cmd[c0:c1]:
                if c == MOVE:
                    pb.move_to(Point(float(pts[k][0]), float(pts[k][1])))
                elif c == LINE:
                    pb.line_to(Point(float(pts[k][0]), float(pts[k][1])))
                elif c == QUAD:
                    pb.quad_to(Point(float(pts[k][0]), float(pts[k][1])), Point(float(pts[k + 1][0]), float(pts[k + 1][1])))
                elif c == CUBIC:
                    pb.cubic_to(Point(float(pts[k][0]), float(pts[k][1])), Point(float(pts[k + 1][0]), float(pts[k + 1][1])),
                                Point(float(pts[k + 2][0]), float(pts[k + 2][1])))
                else:
                    pb.rat_quad_to(Point(float(pts[k][0]), float(pts[k][1])), Point(float(pts[k + 1][0]), float(pts[k + 1][1])),
                                   float(paths.weights[wi]))
                    wi += 1
                k += _NPTS[int(c)]
        else:
            pb.extend(cmd[c0:c1], pts[p0:int(paths.pt_off[i + 1])])
        path = pb.build()
        if scale != 1.0:
            path = path.transform(m)
        r, g, b, a = (float(v) for v in paths.color[i])
        gi = int(paths.grad[i])
        if fill_of is not None:
            fill = fill_of(i, Color(r, g, b, a))
        elif gi >= 0:
            if built[gi] is None:
                built[gi] = _gradient_of(paths.gradients[gi])
            fill = Fill.Gradient(built[gi])
        else:
            fill = Fill.Solid(Color(r, g, b, a))
        layer = comp.create_layer()
        layer.insert(path).set_props(Props(fill_rule=int(paths.fill_rule[i]),
                                           func=Func.Draw(Style(fill=fill, blend_mode=int(paths.blend[i])))))
        comp.insert(first_order + i, layer)
    return n
