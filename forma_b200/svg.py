"""SVG subset loader — restates the part of the
reference's demo loader that its benchmark asset needs
(/root/reference/demo/src/demos/svg.rs:193-260,337-690):

* elements: <g transform= fill= opacity=>, <path d= fill= fill-opacity= opacity= fill-rule=>
* path commands: M m L l H h V v C c S s Q q T t Z z (relative coordinates are
  accumulated in f32 like `add_diff`, svg.rs:371-373)
* the innermost group transform is applied in f64 and rounded to f32 (svg.rs:237-247)
* per-<path> `transform=` attributes are ignored, exactly like the reference
* colours: sRGB hex -> linear (demo/src/main.rs:134-151); opacity precedence
  `opacity`, then `fill-opacity`, else the product of the group opacities
  (svg.rs:129-139,258)

The result is a flat `PathList` (numpy arrays) that can be inserted into any
Composition-like object; `tests/golden/make_paris_fixture.py` stores the list
for paris-30k.svg so that the GPU box does not need the SVG itself.
"""
from __future__ import annotations

import re
import xml.etree.ElementTree as ET
from dataclasses import dataclass

import numpy as np

from .binding import Color, Fill, FillRule, Func, Point, Props, Style

F = np.float32

# command codes of PathList.cmd
MOVE, LINE, QUAD, CUBIC = 0, 1, 2, 3
_NPTS = {MOVE: 1, LINE: 1, QUAD: 2, CUBIC: 3}


@dataclass
class PathList:
    cmd: np.ndarray        # uint8, one per path command
    pts: np.ndarray        # float32 (n, 2): the points consumed by the commands, in order
    cmd_off: np.ndarray    # int64 (n_paths + 1): command range of every path
    pt_off: np.ndarray     # int64 (n_paths + 1)
    color: np.ndarray      # float32 (n_paths, 4) linear RGBA
    fill_rule: np.ndarray  # uint8 (n_paths)

    def __len__(self):
        return len(self.color)

    def save(self, path):
        np.savez_compressed(path, cmd=self.cmd, pts=self.pts, cmd_off=self.cmd_off, pt_off=self.pt_off,
                            color=self.color, fill_rule=self.fill_rule)

    @staticmethod
    def load(path) -> "PathList":
        z = np.load(path)
        return PathList(z["cmd"], z["pts"], z["cmd_off"], z["pt_off"], z["color"], z["fill_rule"])


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
_ARGS = {"m": 2, "l": 2, "h": 1, "v": 1, "c": 6, "s": 4, "q": 4, "t": 2, "z": 0}


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


def parse_svg(path: str) -> PathList:
    cmds, pts, cmd_off, pt_off, colors, rules = [], [], [0], [0], [], []
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
        rules.append(1 if attrib.get("fill-rule") == "evenodd" else 0)
        cmd_off.append(len(cmds))
        pt_off.append(len(pts))
        del n_cmd0, n_pt0

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
    return PathList(np.array(cmds, np.uint8), np.array(pts, np.float32).reshape(-1, 2), np.array(cmd_off, np.int64),
                    np.array(pt_off, np.int64), np.array(colors, np.float32).reshape(-1, 4), np.array(rules, np.uint8))


def compose(api, comp, paths: PathList, scale: float = 1.0, first_order: int = 0, limit=None, fill_of=None):
    """Svg::new + Svg::compose (svg.rs:193-213,904-920): every path becomes one
    layer, document order = layer order; `scale` goes through Path::transform
    like the demo's --scale (control points are transformed and re-flattened)."""
    n = len(paths) if limit is None else min(limit, len(paths))
    m = [scale, 0.0, 0.0, 0.0, scale, 0.0, 0.0, 0.0, 1.0]
    cmd, pts = paths.cmd, paths.pts
    for i in range(n):
        pb = api.PathBuilder()
        pb.extend(cmd[int(paths.cmd_off[i]):int(paths.cmd_off[i + 1])], pts[int(paths.pt_off[i]):int(paths.pt_off[i + 1])])
        path = pb.build()
        if scale != 1.0:
            path = path.transform(m)
        r, g, b, a = (float(v) for v in paths.color[i])
        fill = Fill.Solid(Color(r, g, b, a)) if fill_of is None else fill_of(i, Color(r, g, b, a))
        layer = comp.create_layer()
        layer.insert(path).set_props(Props(fill_rule=int(paths.fill_rule[i]), func=Func.Draw(Style(fill=fill))))
        comp.insert(first_order + i, layer)
    return n
