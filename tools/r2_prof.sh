#!/bin/bash
# usage: tools/r2_prof.sh <tag> <kernel-regex> <workload>...   (under gpurun, 1 GPU)
# Full GPU parity suite (no -x), then one `ncu --set full` capture of the named kernel per workload.
mkdir -p gpurun_out
R=$1; K=$2; shift; shift
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -60 > gpurun_out/${R}_gpu_tests.txt
for w in "$@"; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 4 -c 1 -o gpurun_out/${R}_full_$w -f \
      python bench.py --workload $w --steps 2 --warmup 1 --no-cpu --no-extra > /dev/null 2> gpurun_out/${R}_ncu_$w.err
done
tail -30 gpurun_out/${R}_gpu_tests.txt
ls -la gpurun_out/${R}_*
