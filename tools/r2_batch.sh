#!/bin/bash
# usage: tools/r2_batch.sh <tag>   (under gpurun, 1 GPU): option tests, paint register A/B, ncu captures, launch lists
mkdir -p gpurun_out
R=$1
timeout 600 python -m pytest tests/test_gpu_options.py tests/test_gpu_multi.py -m gpu -q --timeout 300 2>&1 | tail -6 > gpurun_out/${R}_gpu_tests.txt
for w in paris4k circles8k; do
  for v in 0 1; do
    FORMA_PAINT_WIDE=$v timeout 120 python bench.py --no-cpu --no-extra --workload $w > gpurun_out/${R}_bench_${w}_wide$v.json 2>/dev/null
  done
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:paint_kernel -s 4 -c 1 -o gpurun_out/${R}_full_$w -f \
      python bench.py --workload $w --steps 2 --warmup 1 --no-cpu --no-extra > /dev/null 2>&1
done
for w in cubics100k paris4k; do
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_$w.csv python bench.py --workload $w --steps 2 --warmup 1 --no-cpu --no-extra > /dev/null 2>&1
done
cat gpurun_out/${R}_gpu_tests.txt
for f in gpurun_out/${R}_bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1].split('/')[-1], round(d["value"], 1), "stage", d["stage_ms"], "e2e", round(d["e2e"]["value"], 1))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
