#!/bin/bash
# Quick end-of-change validation under gpurun (1 GPU): GPU parity tests, smoke(), the default
# bench line, and an A/B of the number of copy-back bands of host frames (FORMA_COPY_BANDS).
mkdir -p gpurun_out
R=${1:-val}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${R}_gpu_tests.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${R}_smoke.txt 2>&1
python bench.py > gpurun_out/${R}_bench_paris4k.json 2> gpurun_out/${R}_bench_paris4k.err
for b in 2 8 16; do
  FORMA_COPY_BANDS=$b python bench.py --no-cpu > gpurun_out/${R}_bench_paris4k_bands$b.json 2>/dev/null
done
FORMA_COPY_BANDS=8 python bench.py --no-cpu --workload circles8k > gpurun_out/${R}_bench_circles8k_bands8.json 2>/dev/null
python bench.py --no-cpu --workload spaceship1080p --steps 100 --warmup 5 > gpurun_out/${R}_bench_spaceship1080p.json 2>/dev/null
cat gpurun_out/${R}_gpu_tests.txt gpurun_out/${R}_smoke.txt
for f in gpurun_out/${R}_bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1].split('/')[-1], round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), d["e2e"]["stage_ms"])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
