#!/bin/bash
# usage: tools/r2_mgpu.sh <tag> <n_gpus> [extra bench args]   (under gpurun --gpus N)
# Multi-GPU renderer tests (C ABI, one process), then bench.py under torchrun at N ranks.
mkdir -p gpurun_out
R=$1; N=$2; shift; shift
nvidia-smi -L > gpurun_out/${R}_gpus.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 300 2>&1 | tail -15 > gpurun_out/${R}_multi_tests.txt
cat gpurun_out/${R}_multi_tests.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N "$@" > gpurun_out/${R}_bench_n$N.json 2> gpurun_out/${R}_bench_n$N.err
python - gpurun_out/${R}_bench_n$N.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print("N", d["n_gpus"], "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), d["stage_ms"])
    print(" multi_gpu", {k: v for k, v in d["multi_gpu"].items() if k != "stage_ms_per_rank"})
    for r in d["multi_gpu"]["stage_ms_per_rank"] or []:
        print("   rank stages", r)
    print(" e2e", d["e2e"])
    for k, e in (d.get("extra") or {}).items():
        print(" extra", k, round(e["value"], 2), "e2e", round(e["e2e"]["value"], 2), e["stage_ms"], e["multi_gpu"]["band_rows"], e["multi_gpu"]["assembled_frame_equals_single_gpu_frame"])
        for r in e["multi_gpu"]["stage_ms_per_rank"] or []:
            print("   rank stages", r)
except Exception as ex:
    print("unreadable", ex)
    print(open(sys.argv[1].replace(".json", ".err")).read()[-3000:])
PY
