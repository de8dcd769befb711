#!/bin/bash
# Host-frame pipeline, fourth run: the slices' threads queue their work in turn (IssueGate).
T=${1:-r2e}
mkdir -p gpurun_out
timeout 300 python tools/ab_slices.py --quick paris4k cubics100k > gpurun_out/${T}_ab.jsonl 2> gpurun_out/${T}_ab.err
echo "ab rc=$?"; tail -2 gpurun_out/${T}_ab.err
python - <<PY
import json
for l in open("gpurun_out/${T}_ab.jsonl"):
    d = json.loads(l)
    print(d["workload"], d["tag"], d["opts"], d["ms_mean"], d["ms_min"], d["fps_mean"], d["same_frame_as_first"], d["slices"])
    for row in d["slice_stages"]:
        print("      ", row)
PY
timeout 300 python -m pytest tests/test_gpu_slices.py -m gpu -q -x --timeout 300 2>&1 | tail -4
