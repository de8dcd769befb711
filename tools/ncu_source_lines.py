#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares of one kernel of an .ncu-rep.

ncu's source page (csv) lists the kernel's SASS with its counters; `nvdisasm -g` of the same
cubin lists the same SASS with file/line markers. Both are in the same order, so the rows are
joined by index and summed per (file, line).

usage: python tools/ncu_source_lines.py <file.ncu-rep> <kernel regex> <lib.so> <mangled-name substring> [top_n]
e.g.   python tools/ncu_source_lines.py gpurun_out/r1_full_paris4k.ncu-rep paint_kernel \
           forma_b200/libforma_b200.so paint_kernelILi8E 40
"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

rep, kernel, lib, mangled = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kernel}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr, body = rows[1], rows[2:]
i_s, i_e = hdr.index("# Samples"), hdr.index("Instructions Executed")

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
listing = None
for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
    text = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    if mangled in text:
        listing = text
        break
assert listing, "kernel not found in any cubin of the library"
insts, cur, f, l = [], None, None, None
for line in listing.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)', line)
    if m:
        f, l = os.path.basename(m.group(1)), int(m.group(2))
        continue
    if cur and mangled in cur and re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+.*;", line):
        insts.append((f, l))
assert len(insts) == len(body), f"{len(insts)} SASS instructions in the cubin, {len(body)} in the report: different builds"
by = collections.defaultdict(lambda: [0, 0])
for (f, l), r in zip(insts, body):
    by[(f, l)][0] += int(r[i_s] or 0)
    by[(f, l)][1] += int(r[i_e] or 0)
ts, te = sum(v[0] for v in by.values()) or 1, sum(v[1] for v in by.values()) or 1
print(f"{len(body)} SASS instructions, {te} warp instructions executed, {ts} stall samples")
for (f, l), v in sorted(by.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{f}:{l:<5d} inst {100 * v[1] / te:5.1f} %   samples {100 * v[0] / ts:5.1f} %")
