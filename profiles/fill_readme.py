#!/usr/bin/env python
"""Writes profiles/README.md from profiles/README.tmpl.md and the committed bench lines:
    python profiles/fill_readme.py [tag, default r2]
Every @name@ of the template is a number (or a table) computed here from the JSON files, so
the prose cannot drift from the evidence."""
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r2"


def load(name):
    path = os.path.join(HERE, name)
    if not os.path.exists(path):
        return None
    lines = [l for l in open(path) if l.startswith("{")]
    return json.loads(lines[-1]) if lines else None


def f1(v):
    return f"{v:,.1f}".replace(",", " ")


main = load(f"{TAG}_bench_paris4k.json")
ref = load(f"{TAG}_bench_reference_paris4k.json")
others = {k: load(f"{TAG}_bench_{w}.json") for k, w in (("grad", "paris4k_grad"), ("cub", "cubics100k"), ("cir", "circles8k"),
                                                        ("sp", "spaceship1080p"))}
vals = {}


def put(prefix, d):
    if not d:
        for k in ("v", "ms", "seg", "e2e", "io"):
            vals[f"{prefix}_{k}"] = "n/a"
        return
    vals[f"{prefix}_v"] = f1(d["value"])
    vals[f"{prefix}_ms"] = f"{1e3 / d['value']:.3f}"
    vals[f"{prefix}_seg"] = f"{d['mpixel_segments_per_s']:,.0f}".replace(",", " ")
    vals[f"{prefix}_e2e"] = f1(d["e2e"]["value"])
    vals[f"{prefix}_io"] = f"{d['e2e']['h2d_bytes_per_step'] / 1e6:.1f} MB / {d['e2e']['d2h_bytes_per_step'] / 1e6:.1f} MB"


put("paris", main)
put("c1m", (main.get("extra") or {}).get("circles8k_1m"))
for k, d in others.items():
    put(k, d)

cb = main.get("cpu_baseline") or {}
vals["cpu_v"] = f"{cb.get('value', 0):.1f}"
st = cb.get("stage_ms", {})
vals["cpu_ls"], vals["cpu_ra"], vals["cpu_so"], vals["cpu_pa"] = (f"{st.get(k, 0):.1f}" for k in ("line_setup", "rasterize", "sort", "paint"))
vals["cpu_threads"] = ", ".join(f"{k} → {v:.1f} ms" for k, v in (cb.get("thread_candidates_ms") or {}).items())
vals["ref_v"] = f"{ref['value']:.1f}" if ref else "n/a"
base = ref["value"] if ref else cb.get("value", 1.0)
vals["ratio_hbm"] = f"{main['value'] / base:.0f}"
vals["ratio_e2e"] = f"{main['e2e']['value'] / base:.0f}"
vals["tab_paris"] = f"{main['stage_ms']['paint_tables']:.3f}"

rows = ["| workload | line setup | rasterize | sort | painter tables | paint kernel | total | end to end: upload | … paint + copy-back | … total |",
        "|---|---|---|---|---|---|---|---|---|---|"]
items = [("paris4k", main), ("circles8k_1m", (main.get("extra") or {}).get("circles8k_1m")), ("paris4k_grad", others["grad"]),
         ("cubics100k", others["cub"]), ("circles8k", others["cir"]), ("spaceship1080p", others["sp"])]
for name, d in items:
    if not d:
        continue
    s, e = d["stage_ms"], d["e2e"]["stage_ms"]
    rows.append(f"| {name} | {s['line_setup']:.3f} | {s['rasterize']:.3f} | {s['sort']:.3f} | {s['paint_tables']:.3f} | {s['paint_kernel']:.3f} | "
                f"{s['total']:.3f} | {e['upload']:.3f} | {e['paint_kernel'] + e['d2h']:.3f} | {e['total']:.3f} |")
vals["stage_table"] = "\n".join(rows)

rows = ["| kernel | algorithmic bytes / launch | " + " | ".join(n for n, d in items[:5] if d and n != "paris4k_grad") + " | ncu DRAM traffic / launch (paris4k) |",
        "|---|---|" + "---|" * (len([1 for n, d in items[:5] if d and n != "paris4k_grad"]) + 1)]
names = {"radix_downsweep": ("`radix_downsweep_wide_kernel`", "16·N"), "radix_upsweep_scan": ("`radix_upsweep_kernel` + tile scan", "8·N"),
         "paint": ("`paint_kernel`", "8·N + 4·W·H")}
traffic = {}
try:
    traffic = json.load(open(os.path.join(HERE, "ncu_traffic.json"))).get("paris4k", {})
except Exception:
    pass
for key, (label, alg) in names.items():
    cells = []
    for n, d in items[:5]:
        if not d or n == "paris4k_grad":
            continue
        k = d["roofline"]["kernels"].get(key)
        peak = d["roofline"]["peak"]
        cells.append(f"{k['ms_per_launch']:.4f} ms → {k['GBps']:,.0f} GB/s = {100 * k['GBps'] / peak:.1f} %".replace(",", " ") if k else "—")
    t = traffic.get(key)
    rows.append(f"| {label} | {alg} | " + " | ".join(cells) + f" | {t / 1e6:.1f} MB |" if t else f"| {label} | {alg} | " + " | ".join(cells) + " | — |")
cells = []
for n, d in items[:5]:
    if not d or n == "paris4k_grad":
        continue
    ss = d["roofline"]["sort_stage"]
    cells.append(f"{ss['ms']:.3f} ms → {ss['GBps_vs_16N']:,.0f} GB/s = {100 * ss['frac_vs_16N']:.1f} %".replace(",", " "))
rows.append("| whole sort stage vs the algorithm-independent 16·N | 16·N | " + " | ".join(cells) + " | |")
vals["roof_table"] = "\n".join(rows)
c1m = (main.get("extra") or {}).get("circles8k_1m")
vals["ds_frac"] = f"{100 * c1m['roofline']['kernels']['radix_downsweep']['GBps'] / c1m['roofline']['peak']:.0f} %" if c1m else "n/a"

rows = ["| workload | N=1 | N=2 | N=4 | N=8 | speed-up at 8 | e2e N=1 → N=8 |", "|---|---|---|---|---|---|---|"]
mg = {n: load(f"{TAG}_mgpu_n{n}.json") for n in (2, 4, 8)}
m1 = load(f"{TAG}_mgpu_n1.json") or load(f"{TAG}_bench_paris4k_before_slices.json") or main
for label, pick in (("`paris4k`", lambda d: d), ("`circles8k_1m` (BASELINE config 5)", lambda d: (d.get("extra") or {}).get("circles8k_1m"))):
    one = pick(m1)
    cells, last = [f1(one["value"]) if one else "—"], None
    for n in (2, 4, 8):
        d = pick(mg[n]) if mg[n] else None
        cells.append(f1(d["value"]) if d else "—")
        last = d or last
    d8 = pick(mg[8]) if mg[8] else None
    sp = f"{d8['value'] / one['value']:.2f}×" if d8 and one else "—"
    e = f"{f1(one['e2e']['value'])} → {f1(d8['e2e']['value'])}" if d8 and one else "—"
    rows.append(f"| {label} | " + " | ".join(cells) + f" | {sp} | {e} |")
closing2 = load(f"{TAG}_mgpu_n2_closing.json")
if closing2:
    rows.append(f"| `paris4k`, closing build | {f1(main['value'])} | {f1(closing2['value'])} | — | — | — | "
                f"{f1(main['e2e']['value'])} → {f1(closing2['e2e']['value'])} (N = 2) |")
vals["mgpu_table"] = "\n".join(rows)
modes = main.get("library_options", {})
vals["sync_free_note"] = ("`sync_free` (default): from a renderer's second frame on the table kernels are sized by the previous frame's counts and read "
                          "the real ones on the device — no count read-back between the sort and the paint kernel "
                          f"(A/B on the same box, `r2_slices_ab_chained.jsonl`: 0.217 → 0.199 ms, 1507 → 1556 frames/s). Options of this run: {json.dumps(modes)}.")

tmpl = open(os.path.join(HERE, "README.tmpl.md")).read()
missing = sorted(set(re.findall(r"@([a-z_0-9]+)@", tmpl)) - set(vals))
assert not missing, missing
out = re.sub(r"@([a-z_0-9]+)@", lambda m: str(vals[m.group(1)]), tmpl)
open(os.path.join(HERE, "README.md"), "w").write(out)
print("wrote profiles/README.md;", len(vals), "values")
