#!/bin/bash
# usage: tools/mgpu.sh <n_gpus> <workload> [steps]   (run under gpurun --gpus N)
N=$1; W=$2; K=${3:-10}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --steps $K --warmup 3 --workload $W --no-cpu 2>&1 | tail -1 > gpurun_out/mgpu_${W}_$N.json
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps $K --warmup 3 --workload $W --no-cpu 2>gpurun_out/mgpu_${W}_$N.err | tail -1 > gpurun_out/mgpu_${W}_$N.json
fi
python -c "
import json; d=json.load(open('gpurun_out/mgpu_${W}_$N.json'))
print('$W N=$N', 'fps', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'segs', d['config']['pixel_segments'], d['stage_ms'], d['clocks'])" || tail -5 gpurun_out/mgpu_${W}_$N.err
