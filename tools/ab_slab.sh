#!/bin/bash
# First GPU check of the slab-mapped paint kernel (FORMA_PAINT_KERNEL=slab, see paint_kernel in
# kernels_painter.cu and profiles/r1_paint_kernel_analysis.md): full parity suite, then A/B bench.
mkdir -p gpurun_out
R=${1:-slab}
FORMA_PAINT_KERNEL=slab timeout 200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/${R}_gpu_tests_slab.txt
for w in paris4k cubics100k circles8k paris4k_grad; do
  timeout 90 python bench.py --no-cpu --workload $w > gpurun_out/${R}_bench_${w}_default.json 2>/dev/null
  FORMA_PAINT_KERNEL=slab timeout 90 python bench.py --no-cpu --workload $w > gpurun_out/${R}_bench_${w}_slab.json 2>/dev/null
done
cat gpurun_out/${R}_gpu_tests_slab.txt
for f in gpurun_out/${R}_bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1].split('/')[-1], round(d["value"], 1), "paint_ms", d["stage_ms"]["paint_kernel"], "e2e", round(d["e2e"]["value"], 1))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
