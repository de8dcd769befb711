#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for cfg in "A=1" "FORMA_PAINT_REGS=96" "FORMA_PAINT_REGS=80"; do
for w in paris4k cubics100k circles8k; do
env $cfg python bench.py --workload $w --steps 20 --warmup 5 --no-cpu 2>&1 | tail -1 > gpurun_out/b.json
python -c "import json; d=json.load(open('gpurun_out/b.json')); print('$cfg $w', round(d['value'],1), d['stage_ms']['paint_kernel'])"
done; done
