#!/bin/bash
# usage: tools/r2_default.sh <tag>   (under gpurun, 1 GPU): what the driver runs at round end:
# GPU tests, smoke(), `python bench.py`, `python bench.py --impl reference`.
mkdir -p gpurun_out
R=$1
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -8 > gpurun_out/${R}_gpu_tests.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.txt 2>&1
( time timeout 900 python bench.py > gpurun_out/${R}_bench_default.json 2> gpurun_out/${R}_bench_default.err ) 2> gpurun_out/${R}_bench_default.time
( time timeout 900 python bench.py --impl reference > gpurun_out/${R}_bench_reference.json 2> gpurun_out/${R}_bench_reference.err ) 2> gpurun_out/${R}_bench_reference.time
cat gpurun_out/${R}_gpu_tests.txt gpurun_out/${R}_smoke.txt gpurun_out/${R}_bench_default.time gpurun_out/${R}_bench_reference.time
python - gpurun_out/${R}_bench_default.json gpurun_out/${R}_bench_reference.json <<'PY'
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "config", d["config"], "data", d["data"])
        print("  stage", d.get("stage_ms"), "frame_matches_oracle", d.get("frame_matches_oracle"))
        print("  cpu_baseline", d.get("cpu_baseline"))
        print("  roofline", {k: v for k, v in (d.get("roofline") or {}).items() if k != "kernels"})
        for k, e in (d.get("extra") or {}).items():
            print("  extra", k, round(e["value"], 2), "e2e", round(e["e2e"]["value"], 2), e.get("stage_ms"), e.get("cpu_baseline", {}).get("cores"))
    except Exception as ex:
        print(f, "unreadable", ex)
        print(open(f.replace(".json", ".err")).read()[-2000:])
PY
