#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in paris4k cubics100k; do
for cfg in "A=1" "FORMA_SORT_DS=simple"; do
echo "== $w $cfg"
env $cfg python bench.py --workload $w --steps 20 --warmup 5 --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['e2e']['value'],1), d['stage_ms'], d['gpu_launches'])"
done; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cubics.csv python bench.py --workload cubics100k --steps 2 --warmup 1 --no-cpu > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_paris.csv python bench.py --workload paris4k --steps 2 --warmup 1 --no-cpu > /dev/null 2>&1
