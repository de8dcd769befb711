"""Parses the reference's benchmark asset into a committed fixture.

Run in the authoring container (needs /root/reference):
    python tests/golden/make_paris_fixture.py
Source: /root/reference/assets/svgs/paris-30k.svg (BASELINE config 2), parsed
by forma_b200/svg.py (the subset loader restating demo/src/demos/svg.rs).
Output: tests/data/paris30k_paths.npz — path commands + f32 points as handed to
PathBuilder (group transform applied), linear fill colours, fill rules. The GPU
box has no /root/reference, so bench.py and the tests read this file.
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from forma_b200 import svg  # noqa: E402

SRC = "/root/reference/assets/svgs/paris-30k.svg"
OUT = os.path.join(ROOT, "tests", "data", "paris30k_paths.npz")

t0 = time.time()
paths = svg.parse_svg(SRC)
paths.save(OUT)
print(f"{len(paths)} paths, {len(paths.cmd)} commands, {len(paths.pts)} points in {time.time() - t0:.1f}s -> {OUT} "
      f"({os.path.getsize(OUT) / 1e6:.1f} MB)")
import numpy as np  # noqa: E402
print("commands:", dict(zip(*np.unique(paths.cmd, return_counts=True))))
print("opaque layers:", int((paths.color[:, 3] == 1.0).sum()), "bbox:", paths.pts.min(0), paths.pts.max(0))
