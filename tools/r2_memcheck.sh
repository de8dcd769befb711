#!/bin/bash
# compute-sanitizer memcheck over the code paths touched at the end of round 2: slim flatten jobs with path
# transforms, row costs / segment counts of band renders, the host-frame pipeline.
mkdir -p gpurun_out
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -m gpu -q -x --timeout 400 \
    tests/test_gpu_parity.py tests/test_gpu_slices.py tests/test_gpu_multi.py \
    -k "path_transforms or flatten_points or crop_stride or leaves_alone or host_frame_matches_oracle and 4-2 or multi_renderer_host" \
    > gpurun_out/r2_memcheck.txt 2>&1
echo "memcheck rc=$?"; tail -6 gpurun_out/r2_memcheck.txt; grep -c "Invalid\|out of bounds" gpurun_out/r2_memcheck.txt
