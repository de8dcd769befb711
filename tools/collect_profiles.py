#!/usr/bin/env python
"""Turns the files tools/final_profiles.sh left in gpurun_out/ into the committed
evidence under profiles/: bench JSON lines, launch-list summaries, `--set full`
summaries and profiles/ncu_traffic.json (DRAM bytes per launch of the kernels
bench.py reports a roofline for).

    python tools/collect_profiles.py [round-tag, default r1]
"""
import csv
import glob
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC, DST = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r1"
SUMM = os.path.join(DST, "summarize.py")
ROOF = {"radix_downsweep": "radix_downsweep_wide_kernel", "radix_upsweep_scan": "radix_upsweep_kernel", "paint": "paint_kernel"}


def run(args):
    return subprocess.run([sys.executable, SUMM] + args, capture_output=True, text=True).stdout


for f in sorted(glob.glob(os.path.join(SRC, f"{TAG}_bench_*.json")) + glob.glob(os.path.join(SRC, f"{TAG}_gpu_tests.txt")) +
                (glob.glob(os.path.join(SRC, "mgpu_*.json")) if TAG == "r1" else [])):  # r1's multi-GPU files had no tag
    if os.path.getsize(f):
        name = os.path.basename(f)
        shutil.copy(f, os.path.join(DST, name if name.startswith(TAG) else f"{TAG}_{name}"))
for f in sorted(glob.glob(os.path.join(SRC, f"{TAG}_launches_*.csv"))):
    open(os.path.join(DST, os.path.basename(f).replace(".csv", ".txt")), "w").write(run(["launches", f]))
traffic = {}
for f in sorted(glob.glob(os.path.join(SRC, f"{TAG}_full_*.ncu-rep"))):
    workload = os.path.basename(f)[len(TAG) + 6:-8]
    open(os.path.join(DST, os.path.basename(f).replace(".ncu-rep", ".txt")), "w").write(run(["full", f]))
    out = subprocess.run(["ncu", "-i", f, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        continue
    idx = {h: i for i, h in enumerate(rows[0])}
    units = rows[1]

    def to_bytes(r, key):
        v, u = float(r[idx[key]].replace(",", "")), units[idx[key]]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    for key, kern in ROOF.items():
        vals = [to_bytes(r, "dram__bytes_read.sum") + to_bytes(r, "dram__bytes_write.sum") for r in rows[2:]
                if kern in r[idx["Kernel Name"]]]
        if vals:
            traffic.setdefault(workload, {})[key] = sum(vals) / len(vals)
if traffic:
    path = os.path.join(DST, "ncu_traffic.json")
    old = json.load(open(path)) if os.path.exists(path) else {}
    old.update(traffic)
    json.dump(old, open(path, "w"), indent=1, sort_keys=True)
print("collected into", DST, "traffic:", traffic)
