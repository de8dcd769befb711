#!/bin/bash
# Round-2 closing evidence (1 GPU) after the host-pipeline work: GPU tests, smoke, bench lines of all
# workloads, the reference arm, ncu launch lists, one `--set full` capture (paris4k; the kernels of
# the other workloads are unchanged since the captures of tools/final_profiles.sh).
mkdir -p gpurun_out
R=${1:-r2}
python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -3 > gpurun_out/${R}_gpu_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/${R}_gpu_tests.txt 2>&1
cat gpurun_out/${R}_gpu_tests.txt
if grep -q "failed\|error" gpurun_out/${R}_gpu_tests.txt; then
  python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -40
  echo "GPU tests failed: stopping before the benches"; exit 1
fi
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/${R}_bench_reference_paris4k.json 2>/dev/null
python bench.py > gpurun_out/${R}_bench_paris4k.json 2> gpurun_out/${R}_bench_paris4k.err
for w in cubics100k circles8k paris4k_grad; do
  python bench.py --workload $w --no-cpu > gpurun_out/${R}_bench_$w.json 2>/dev/null
done
python bench.py --workload spaceship1080p --steps 100 --warmup 5 > gpurun_out/${R}_bench_spaceship1080p.json 2>/dev/null
for w in paris4k cubics100k circles8k; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_$w.csv \
      python bench.py --workload $w --steps 2 --warmup 1 --no-cpu --no-extra > /dev/null 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:'radix_downsweep_wide|radix_upsweep|paint_kernel|raster_emit|cells_kernel' \
    -s 7 -c 7 -o gpurun_out/${R}_full_paris4k -f python bench.py --workload paris4k --steps 2 --warmup 1 --no-cpu --no-extra > /dev/null 2>&1
python - <<PY
import json
for w in ("paris4k", "cubics100k", "circles8k", "paris4k_grad", "spaceship1080p", "reference_paris4k"):
    try:
        d = json.loads([l for l in open("gpurun_out/${R}_bench_%s.json" % w) if l.startswith("{")][-1])
        print(w, round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), d.get("stage_ms"), d.get("frame_matches_oracle"), d.get("cpu_baseline", {}).get("value"))
    except Exception as e:
        print(w, "failed:", e)
PY
ls -la gpurun_out | grep ${R}_ | tail -30
