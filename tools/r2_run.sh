#!/bin/bash
# usage: tools/r2_run.sh <tag> [bench workloads...]   (under gpurun, 1 GPU)
# GPU parity suite (all failures listed), then one bench line per workload (no CPU leg, no extras).
mkdir -p gpurun_out
R=$1; shift
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -40 > gpurun_out/${R}_gpu_tests.txt
for w in "$@"; do
  timeout 120 python bench.py --no-cpu --no-extra --workload $w > gpurun_out/${R}_bench_${w}.json 2> gpurun_out/${R}_bench_${w}.err
done
cat gpurun_out/${R}_gpu_tests.txt
for f in gpurun_out/${R}_bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1].split('/')[-1], round(d["value"], 1), "stage", d["stage_ms"], "e2e", round(d["e2e"]["value"], 1), d["e2e"]["stage_ms"])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
    print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
done
