#!/bin/bash
# Quick end-of-change validation under gpurun (1 GPU), most important first: GPU parity tests,
# the default bench line, then A/Bs of the host-frame copy-back scheme
# (FORMA_BAND_SIGNAL=0: one paint launch per band; FORMA_COPY_BANDS=n).
mkdir -p gpurun_out
R=${1:-val}
timeout 150 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${R}_gpu_tests.txt
timeout 60 python bench.py > gpurun_out/${R}_bench_paris4k.json 2> gpurun_out/${R}_bench_paris4k.err
FORMA_BAND_SIGNAL=0 timeout 40 python bench.py --no-cpu > gpurun_out/${R}_bench_paris4k_nosignal.json 2>/dev/null
FORMA_COPY_BANDS=16 timeout 40 python bench.py --no-cpu > gpurun_out/${R}_bench_paris4k_bands16.json 2>/dev/null
FORMA_COPY_BANDS=4 timeout 40 python bench.py --no-cpu > gpurun_out/${R}_bench_paris4k_bands4.json 2>/dev/null
FORMA_BAND_SIGNAL=0 timeout 60 python -m pytest tests -m gpu -x -q -k "random_mixed or large_frame or 8k or channels or crop" 2>&1 | tail -3 > gpurun_out/${R}_gpu_tests_nosignal.txt
timeout 60 python bench.py --no-cpu --workload circles8k > gpurun_out/${R}_bench_circles8k.json 2>/dev/null
cat gpurun_out/${R}_gpu_tests.txt gpurun_out/${R}_gpu_tests_nosignal.txt
for f in gpurun_out/${R}_bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1].split('/')[-1], round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), d["e2e"]["gpu_launches"], d["e2e"]["stage_ms"])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
