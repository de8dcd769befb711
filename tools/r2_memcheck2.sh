#!/bin/bash
mkdir -p gpurun_out
timeout 170 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -m gpu -q -x --timeout 160 \
    "tests/test_gpu_slices.py::test_sliced_host_frame_matches_oracle[4-2]" "tests/test_gpu_slices.py::test_sliced_counters_are_those_of_the_whole_frame" \
    > gpurun_out/r2_memcheck_slices.txt 2>&1
echo "memcheck rc=$?"; tail -5 gpurun_out/r2_memcheck_slices.txt
