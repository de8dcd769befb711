#!/bin/bash
# usage: tools/r2_prof2.sh <tag> <kernel-regex> <workload>...   (under gpurun, 1 GPU)
# Quick parity subset, bench lines, then one `ncu --set full` capture of the named kernel per workload.
mkdir -p gpurun_out
R=$1; K=$2; shift; shift
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_options.py -m gpu -q --timeout 300 -x 2>&1 | tail -15 > gpurun_out/${R}_gpu_tests.txt
for w in "$@"; do
  timeout 120 python bench.py --no-cpu --no-extra --workload $w > gpurun_out/${R}_bench_${w}.json 2> gpurun_out/${R}_bench_${w}.err
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 4 -c 1 -o gpurun_out/${R}_full_$w -f \
      python bench.py --workload $w --steps 2 --warmup 1 --no-cpu --no-extra > /dev/null 2> gpurun_out/${R}_ncu_$w.err
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${R}_launches_$1.csv python bench.py --workload $1 --steps 2 --warmup 1 --no-cpu --no-extra > /dev/null 2>&1
cat gpurun_out/${R}_gpu_tests.txt
for f in gpurun_out/${R}_bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1].split('/')[-1], round(d["value"], 1), "stage", d["stage_ms"], "e2e", round(d["e2e"]["value"], 1), d["e2e"]["stage_ms"])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
