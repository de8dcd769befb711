#!/bin/bash
# usage: tools/bw.sh <workload>...   GPU tests, then one bench line per workload (run under gpurun)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in "$@"; do
echo "== $w"
python bench.py --workload $w --steps 20 --warmup 5 --no-cpu 2>&1 | tail -1 > gpurun_out/b_$w.json
python -c "import json,sys; d=json.load(open('gpurun_out/b_$w.json')); print(round(d['value'],1), round(d['e2e']['value'],1), d['stage_ms'], d['gpu_launches']); print('  e2e', d['e2e'].get('stage_ms'), d['e2e']['h2d_bytes_per_step']); print('  roof', {k:(round(v['ms_per_launch'],4), round(v['GBps'])) for k,v in d['roofline']['kernels'].items()}, d['roofline']['kernel'], round(d['roofline']['frac'],3))"
done
