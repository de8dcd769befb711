#!/usr/bin/env python
"""Brief per-kernel summary of an .ncu-rep (raw page): duration, DRAM bytes, stalls.
usage: python tools/ncu_brief.py <file.ncu-rep>"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
def g(r, name):
    try: return float(r[col[name]])
    except Exception: return float("nan")
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0]
    dur = g(r, "gpu__time_duration.sum")
    rd, wr = g(r, "dram__bytes_read.sum"), g(r, "dram__bytes_write.sum")
    print(f"{name[:44]:44s} {dur:8.1f}us dram r/w {rd:7.1f}/{wr:7.1f} MB  dram%={g(r,'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):5.1f} "
          f"sm%={g(r,'sm__throughput.avg.pct_of_peak_sustained_elapsed'):5.1f} issue%={g(r,'smsp__issue_active.avg.pct_of_peak_sustained_active'):5.1f} "
          f"occ%={g(r,'sm__warps_active.avg.pct_of_peak_sustained_active'):5.1f} regs={int(g(r,'launch__registers_per_thread'))} "
          f"inst={g(r,'smsp__inst_executed.sum')/1e6:7.1f}M")
    st = sorted(((g(r, h), h.split('stalled_')[1]) for h in hdr if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h and r[col[h]]), reverse=True)
    tot = sum(v for v, _ in st if v == v) or 1
    print("     stalls: " + ", ".join(f"{n} {100*v/tot:.0f}%" for v, n in st[:6]))
