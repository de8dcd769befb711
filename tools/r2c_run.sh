#!/bin/bash
# Host-frame pipeline, second run: chained uploads. Tests of the pipeline, A/B, bench-scale parity.
T=${1:-r2c}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_slices.py tests/test_gpu_options.py -m gpu -q --timeout 300 > gpurun_out/${T}_tests_new.txt 2>&1
echo "new tests rc=$?"; tail -12 gpurun_out/${T}_tests_new.txt
timeout 400 python tools/ab_slices.py paris4k cubics100k circles8k > gpurun_out/${T}_ab.jsonl 2> gpurun_out/${T}_ab.err
echo "ab rc=$?"; tail -3 gpurun_out/${T}_ab.err
python - <<PY
import json
for l in open("gpurun_out/${T}_ab.jsonl"):
    d = json.loads(l)
    print(d["workload"], d["tag"], d["opts"], "e2e" if d["e2e"] else "dev", d["ms_mean"], d["ms_min"], d["fps_mean"], d["same_frame_as_first"], d["slices"], {k: round(v, 3) for k, v in d["stage_ms"].items()})
PY
timeout 600 python -m pytest tests/test_gpu_bench_scale.py tests/test_gpu_multi.py -m gpu -q --timeout 500 -k "not one_million and not spaceship" > gpurun_out/${T}_scale_tests.txt 2>&1
echo "scale rc=$?"; tail -6 gpurun_out/${T}_scale_tests.txt
