#!/bin/bash
# usage: tools/prof_sort.sh  (run under gpurun)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in paris4k cubics100k; do
python bench.py --workload $w --steps 20 --warmup 5 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_$w.json
python -c "import json; d=json.load(open('gpurun_out/bench_$w.json')); print('$w', round(d['value'],1), round(d['e2e']['value'],1), d['stage_ms'], d['gpu_launches'])"
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cubics.csv python bench.py --workload cubics100k --steps 2 --warmup 1 --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'radix_downsweep|radix_upsweep|radix_tile_scan' -s 6 -c 6 -o gpurun_out/prof_sort -f python bench.py --workload cubics100k --steps 2 --warmup 1 --no-cpu > /dev/null 2>&1
ls -la gpurun_out
