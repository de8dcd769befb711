#!/bin/bash
# usage: tools/mgpu.sh <n_gpus> <workload> [steps] [assembly: p2p|gather]   (run under gpurun --gpus N)
N=$1; W=$2; K=${3:-10}; A=${4:-p2p}
mkdir -p gpurun_out
OUT=gpurun_out/mgpu_${W}_${N}_${A}
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --steps $K --warmup 3 --workload $W --no-cpu 2>$OUT.err | tail -1 > $OUT.json
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps $K --warmup 3 --workload $W --no-cpu --assembly $A 2>$OUT.err | tail -1 > $OUT.json
fi
python -c "
import json; d=json.load(open('$OUT.json'))
print('$W N=$N $A', 'fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'segs', d['config']['pixel_segments'], d['stage_ms'], d.get('multi_gpu'), d['clocks'])" || tail -8 $OUT.err
