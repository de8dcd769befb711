#!/bin/bash
# Sync-free painter tables: option tests, memcheck of the redo path, whole GPU suite, bench lines.
T=${1:-r2s}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_options.py -m gpu -q -x --timeout 300 2>&1 | tail -5
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_options.py -m gpu -q -x --timeout 500 \
    -k "read_backs or grows or test_fast_shrink or sync_free" > gpurun_out/${T}_memcheck.txt 2>&1
echo "memcheck rc=$?"; tail -4 gpurun_out/${T}_memcheck.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/${T}_gpu_tests.txt 2>&1; tail -3 gpurun_out/${T}_gpu_tests.txt
for w in paris4k cubics100k circles8k spaceship1080p; do
  python bench.py --workload $w --no-cpu > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench_$w.json"))
print("$w", round(d["value"],1), d["stage_ms"], "e2e", round(d["e2e"]["value"],1), d["e2e"]["stage_ms"])
x=d.get("extra") or {}
for k,v in x.items():
    print("  extra",k, round(v["value"],2), v.get("stage_ms"), "e2e", round(v["e2e"]["value"],2))
PY
done
FORMA_SYNC_FREE=0 python bench.py --workload paris4k --no-cpu --no-extra > gpurun_out/${T}_bench_paris4k_sync.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_paris4k_sync.json')); print('paris4k sync_free=0', round(d['value'],1), d['stage_ms'])"
