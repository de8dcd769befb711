#!/bin/bash
# usage: tools/r2_launches.sh <tag> <workload>...  (under gpurun, 1 GPU): ncu launch lists + copy-band A/B on the first workload
mkdir -p gpurun_out
R=$1; shift
for w in "$@"; do
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_$w.csv python bench.py --workload $w --steps 2 --warmup 1 --no-cpu --no-extra > /dev/null 2>&1
done
for b in 1 2 8; do
  FORMA_COPY_BANDS=$b timeout 60 python bench.py --no-cpu --no-extra --workload $1 > gpurun_out/${R}_bench_$1_bands$b.json 2>/dev/null
done
for f in gpurun_out/${R}_bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1].split('/')[-1], round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), d["e2e"]["stage_ms"])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
