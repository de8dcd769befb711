#!/bin/bash
# Round-end evidence, run under gpurun (1 GPU):
#   bench lines (default run with the CPU baseline, the reference arm, the other workloads),
#   ncu launch lists of the same commands, one `--set full` capture of the top kernels.
# Everything lands in gpurun_out/; tools/collect_profiles.py turns it into profiles/.
mkdir -p gpurun_out
R=${1:-r1}
python -m pytest tests -m gpu -x -q 2>&1 | tail -2 > gpurun_out/${R}_gpu_tests.txt
python bench.py > gpurun_out/${R}_bench_paris4k.json 2> gpurun_out/${R}_bench_paris4k.err
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/${R}_bench_reference_paris4k.json 2>/dev/null
for w in cubics100k circles8k paris4k_grad; do
  python bench.py --workload $w --no-cpu > gpurun_out/${R}_bench_$w.json 2>/dev/null
done
python bench.py --workload spaceship1080p --steps 100 --warmup 5 > gpurun_out/${R}_bench_spaceship1080p.json 2>/dev/null
for w in paris4k cubics100k circles8k; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_$w.csv \
      python bench.py --workload $w --steps 2 --warmup 1 --no-cpu > /dev/null 2>&1
done
ncu --set full --clock-control none --import-source on \
    -k regex:'radix_downsweep_wide|radix_upsweep|radix_tile_scan|paint_kernel|cell_cover|raster_emit|merge_entries|line_count' \
    -s 12 -c 12 -o gpurun_out/${R}_full_cubics100k -f python bench.py --workload cubics100k --steps 2 --warmup 1 --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'radix_downsweep_wide|radix_upsweep|paint_kernel|raster_emit' \
    -s 8 -c 8 -o gpurun_out/${R}_full_paris4k -f python bench.py --workload paris4k --steps 2 --warmup 1 --no-cpu > /dev/null 2>&1
ls -la gpurun_out | tail -20
