#!/bin/bash
# Host-frame pipeline, third run: do the slices' kernel chains overlap once the streams stop aliasing
# onto 8 hardware queues? (CUDA_DEVICE_MAX_CONNECTIONS, fewer streams per slice)
T=${1:-r2d}
mkdir -p gpurun_out
for mc in 8 32; do
  CUDA_DEVICE_MAX_CONNECTIONS=$mc timeout 300 python tools/ab_slices.py --quick paris4k cubics100k > gpurun_out/${T}_ab_mc$mc.jsonl 2> gpurun_out/${T}_ab_mc$mc.err
  echo "ab mc=$mc rc=$?"; tail -2 gpurun_out/${T}_ab_mc$mc.err
done
python - <<PY
import json
for mc in (8, 32):
    for l in open("gpurun_out/${T}_ab_mc%d.jsonl" % mc):
        d = json.loads(l)
        print("mc", mc, d["workload"], d["tag"], d["opts"], d["ms_mean"], d["ms_min"], d["fps_mean"], d["same_frame_as_first"], d["slices"])
        for row in d["slice_stages"]:
            print("      ", row)
PY
timeout 300 python -m pytest tests/test_gpu_slices.py -m gpu -q --timeout 300 2>&1 | tail -4
