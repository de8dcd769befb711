#!/bin/bash
# usage: tools/r2_batch2.sh <tag>   (under gpurun, 1 GPU): full parity suite, band diagnosis, bench lines
mkdir -p gpurun_out
R=$1
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -8 > gpurun_out/${R}_gpu_tests.txt
timeout 120 python tools/band_diag.py paris4k > gpurun_out/${R}_band_diag.txt 2>&1
for w in paris4k cubics100k circles8k; do
  timeout 120 python bench.py --no-cpu --no-extra --workload $w > gpurun_out/${R}_bench_${w}.json 2>/dev/null
done
cat gpurun_out/${R}_gpu_tests.txt gpurun_out/${R}_band_diag.txt
for f in gpurun_out/${R}_bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1].split('/')[-1], round(d["value"], 1), "stage", d["stage_ms"], "e2e", round(d["e2e"]["value"], 1), d["e2e"]["stage_ms"])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
