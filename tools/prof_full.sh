#!/bin/bash
# usage: tools/prof_full.sh <workload> <kernel-regex> <skip> <count> <out-name>   (run under gpurun)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o gpurun_out/$5 -f python bench.py --workload $1 --steps 2 --warmup 1 --no-cpu > /dev/null 2>&1
ls -la gpurun_out/$5.ncu-rep
