#!/bin/bash
# usage: tools/launches.sh <workload>...   (run under gpurun) -> gpurun_out/launches_<workload>.csv
mkdir -p gpurun_out
for w in "$@"; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$w.csv python bench.py --workload $w --steps 2 --warmup 1 --no-cpu > /dev/null 2>&1
done
ls gpurun_out
