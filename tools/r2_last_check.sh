#!/bin/bash
# Last sanity check of the closing tree: smoke + the tests that touch the wide sort, the paint grid and the slices.
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 80 python -m pytest -m gpu -q -x --timeout 60 tests/test_gpu_parity.py tests/test_gpu_slices.py \
    -k "large_frame_properties or random_mixed_scene or leaves_alone or path_transforms or sorted_segments" 2>&1 | tail -3
