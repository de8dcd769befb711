#!/bin/bash
# Stage times resolved lazily (outside bench.py's timed interval): parity / option tests and the bench lines again.
R=${1:-r2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_options.py tests/test_gpu_slices.py -m gpu -q --timeout 300 -k "not 40_000_000" 2>&1 | tail -4
python bench.py > gpurun_out/${R}_bench_paris4k.json 2> gpurun_out/${R}_bench_paris4k.err
for w in cubics100k circles8k paris4k_grad; do
  python bench.py --workload $w --no-cpu > gpurun_out/${R}_bench_$w.json 2>/dev/null
done
python bench.py --workload spaceship1080p --steps 100 --warmup 5 > gpurun_out/${R}_bench_spaceship1080p.json 2>/dev/null
python - <<PY
import json
for w in ("paris4k", "cubics100k", "circles8k", "paris4k_grad", "spaceship1080p"):
    try:
        d = json.loads([l for l in open("gpurun_out/${R}_bench_%s.json" % w) if l.startswith("{")][-1])
        print(w, round(d["value"], 1), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), d.get("stage_ms"), d.get("frame_matches_oracle"), d.get("cpu_baseline", {}).get("value"))
    except Exception as e:
        print(w, "failed:", e)
PY
