#!/usr/bin/env python
"""Prints the markdown tables of profiles/README.md from the committed bench lines:
    python profiles/make_tables.py [tag, default r2]"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r2"


def load(name):
    path = os.path.join(HERE, name)
    if not os.path.exists(path):
        return None
    lines = [l for l in open(path) if l.startswith("{")]
    return json.loads(lines[-1]) if lines else None


def row(name, d, wl=None):
    st = d["stage_ms"]
    ws = d.get("workload_stats", {})
    e = d["e2e"]
    return (f"| `{name}` | {ws.get('pixel_segments', 0) / 1e6:.2f} M | **{d['value']:.1f}** | {1e3 / d['value']:.3f} | "
            f"{d['mpixel_segments_per_s']:.0f} | **{e['value']:.1f}** | {e['h2d_bytes_per_step'] / 1e6:.1f} MB / {e['d2h_bytes_per_step'] / 1e6:.1f} MB |"), \
           (f"| {name} | {st['line_setup']:.3f} | {st['rasterize']:.3f} | {st['sort']:.3f} | {st['paint_tables']:.3f} | {st['paint_kernel']:.3f} | {st['total']:.3f} |")


print("## single GPU\n")
print("| workload | pixel segments | frames/s (frame in HBM) | ms | M segments/s | e2e frames/s | H2D / D2H per step |\n|---|---|---|---|---|---|---|")
stages = []
main = load(f"{TAG}_bench_paris4k.json")
items = [("paris4k", main)]
if main and main.get("extra"):
    for k, e in main["extra"].items():
        items.append((k + " (extra of the default run)", e))
for w in ("paris4k_grad", "cubics100k", "circles8k", "spaceship1080p"):
    items.append((w, load(f"{TAG}_bench_{w}.json")))
for name, d in items:
    if d:
        a, b = row(name, d)
        print(a)
        stages.append(b)
print("\n| workload | line setup | rasterize | sort | painter tables | paint kernel | total |\n|---|---|---|---|---|---|---|")
print("\n".join(stages))
if main:
    print("\ncpu_baseline:", json.dumps(main.get("cpu_baseline")))
    print("frame_matches_oracle:", main.get("frame_matches_oracle"), " clocks:", main.get("clocks"))
    print("roofline:", json.dumps({k: v for k, v in main["roofline"].items() if k != "kernels"}))
    for k, v in main["roofline"]["kernels"].items():
        print("  ", k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items()})
    for k, e in (main.get("extra") or {}).items():
        print("extra", k, "roofline kernels:")
        for kk, v in e["roofline"]["kernels"].items():
            print("  ", kk, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items()})
        print("   sort_stage", e["roofline"]["sort_stage"])
ref = load(f"{TAG}_bench_reference_paris4k.json")
if ref:
    print("\nreference arm:", round(ref["value"], 2), "frames/s", ref["cpu_baseline"]["cores"], "threads;",
          {k: round(v["value"], 3) for k, v in (ref.get("extra") or {}).items()})
for w in ("cubics100k", "circles8k"):
    d = load(f"{TAG}_bench_{w}.json")
    if d:
        print(f"\n{w} roofline kernels:")
        for kk, v in d["roofline"]["kernels"].items():
            print("  ", kk, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items()})
        print("   sort_stage", d["roofline"]["sort_stage"])
print("\n## multi GPU\n")
print("| ranks | paris4k frames/s | paris4k e2e | 1 M circles @ 8K frames/s | 1 M circles e2e | slowest / fastest rank (paris, ms) | assembly ms |\n|---|---|---|---|---|---|---|")
if main:
    ex = (main.get("extra") or {}).get("circles8k_1m")
    print(f"| 1 | {main['value']:.1f} | {main['e2e']['value']:.1f} | {ex['value']:.2f} | {ex['e2e']['value']:.2f} | | |" if ex else "")
for n in (2, 4, 8):
    d = load(f"{TAG}_mgpu_n{n}.json")
    if d:
        ex = (d.get("extra") or {}).get("circles8k_1m")
        mg = d["multi_gpu"]
        print(f"| {n} | {d['value']:.1f} | {d['e2e']['value']:.1f} | {ex['value']:.2f} | {ex['e2e']['value']:.2f} | "
              f"{mg['render_ms_slowest_rank']:.3f} / {mg['render_ms_fastest_rank']:.3f} | {mg['assembly_ms']:.3f} |")
