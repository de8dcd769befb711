#!/bin/bash
# Quick end-of-change validation under gpurun (1 GPU), most important first: GPU parity tests,
# smoke(), the default bench line, then an A/B of the number of copy-back bands of host
# frames (FORMA_COPY_BANDS=n, default 4). Every command has its own timeout.
mkdir -p gpurun_out
R=${1:-val}
timeout 150 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${R}_gpu_tests.txt
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${R}_smoke.txt 2>&1
timeout 60 python bench.py > gpurun_out/${R}_bench_paris4k.json 2> gpurun_out/${R}_bench_paris4k.err
for b in 2 8; do
  FORMA_COPY_BANDS=$b timeout 40 python bench.py --no-cpu > gpurun_out/${R}_bench_paris4k_bands$b.json 2>/dev/null
done
cat gpurun_out/${R}_gpu_tests.txt gpurun_out/${R}_smoke.txt
for f in gpurun_out/${R}_bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1].split('/')[-1], round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), d["e2e"]["gpu_launches"], d["e2e"]["stage_ms"])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
