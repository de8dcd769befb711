#!/bin/bash
# usage: tools/r2_mgpu8.sh <tag>   (under gpurun --gpus 8): multi-GPU renderer tests, bench at 8 and 4 ranks
mkdir -p gpurun_out
R=$1
nvidia-smi -L > gpurun_out/${R}_gpus.txt 2>&1
nvidia-smi topo -m >> gpurun_out/${R}_gpus.txt 2>&1
(nproc; cat /sys/fs/cgroup/cpu.max) >> gpurun_out/${R}_gpus.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 300 2>&1 | tail -8 > gpurun_out/${R}_multi_tests.txt
cat gpurun_out/${R}_multi_tests.txt
for N in ${NS:-8 4}; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
    bench.py --gpus $N > gpurun_out/${R}_bench_n$N.json 2> gpurun_out/${R}_bench_n$N.err
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
        print(" extra", k, round(e["value"], 2), "e2e", round(e["e2e"]["value"], 2), e["stage_ms"], e["multi_gpu"]["band_rows"], e["multi_gpu"]["assembled_frame_equals_single_gpu_frame"], e["multi_gpu"]["assembly_ms"])
        for r in e["multi_gpu"]["stage_ms_per_rank"] or []:
            print("   rank stages", r)
        print("  e2e", e["e2e"])
except Exception as ex:
    print("unreadable", ex)
    print(open(sys.argv[1].replace(".json", ".err")).read()[-3000:])
PY
done
