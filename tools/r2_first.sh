#!/bin/bash
# Round 2, first GPU call: the slab paint kernel's first run (parity suite + A/B), the 1 M-circle
# scene's first run on a device, and the host's CPU topology for the CPU arm's thread tuning.
mkdir -p gpurun_out
R=r2a
(nproc; lscpu | head -30; cat /sys/fs/cgroup/cpu.max 2>/dev/null; python -c "import os;print(len(os.sched_getaffinity(0)))") > gpurun_out/${R}_cpu.txt 2>&1
FORMA_PAINT_KERNEL=slab timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/${R}_gpu_tests_slab.txt
for w in paris4k circles8k paris4k_grad; do
  timeout 90 python bench.py --no-cpu --workload $w > gpurun_out/${R}_bench_${w}_default.json 2>/dev/null
  FORMA_PAINT_KERNEL=slab timeout 90 python bench.py --no-cpu --workload $w > gpurun_out/${R}_bench_${w}_slab.json 2>/dev/null
done
timeout 240 python bench.py --no-cpu --workload circles8k_1m --steps 5 > gpurun_out/${R}_bench_circles8k_1m_default.json 2> gpurun_out/${R}_bench_circles8k_1m_default.err
cat gpurun_out/${R}_gpu_tests_slab.txt
for f in gpurun_out/${R}_bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1].split('/')[-1], round(d["value"], 1), "stage", d["stage_ms"], "e2e", round(d["e2e"]["value"], 1))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
