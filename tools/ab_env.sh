#!/bin/bash
# usage: tools/ab_env.sh "<workloads>" "ENV1=a" "ENV2=b" ...   (one bench line per workload x env; run under gpurun)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
W="$1"; shift
for cfg in "$@"; do
for w in $W; do
env $cfg python bench.py --workload $w --steps 20 --warmup 5 --no-cpu 2>&1 | tail -1 > gpurun_out/b.json
python -c "import json; d=json.load(open('gpurun_out/b.json')); print('$cfg $w', round(d['value'],1), round(d['e2e']['value'],1), d['stage_ms'])"
done; done
