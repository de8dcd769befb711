#!/bin/bash
# Host-frame pipeline (host_slices): its tests, the A/B of its settings on three workloads, then the whole GPU suite.
T=${1:-r2b}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_slices.py tests/test_gpu_options.py tests/test_gpu_multi.py -m gpu -q -x --timeout 300 > gpurun_out/${T}_tests_new.txt 2>&1
echo "new tests rc=$?"; tail -15 gpurun_out/${T}_tests_new.txt
timeout 400 python tools/ab_slices.py paris4k cubics100k circles8k > gpurun_out/${T}_ab.jsonl 2> gpurun_out/${T}_ab.err
echo "ab rc=$?"; tail -3 gpurun_out/${T}_ab.err
python - <<PY
import json
for l in open("gpurun_out/${T}_ab.jsonl"):
    d = json.loads(l)
    print(d["workload"], d["tag"], d["opts"], "e2e" if d["e2e"] else "dev", d["ms_mean"], d["ms_min"], d["fps_mean"], d["same_frame_as_first"], d["tables_mode"], d["slices"], d["stage_ms"])
PY
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/${T}_gpu_tests.txt 2>&1
echo "suite rc=$?"; tail -6 gpurun_out/${T}_gpu_tests.txt
