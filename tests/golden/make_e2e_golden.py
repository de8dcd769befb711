"""Converts the reference's e2e golden PNGs into one committed fixture.

Run in the authoring container (needs /root/reference and PIL):
    python tests/golden/make_e2e_golden.py
Source: /root/reference/e2e-tests/expected/*.png (34 images, 64x64 RGBA8),
produced by the scenes of e2e-tests/tests/tests.rs:219-742. The output
tests/golden/e2e_expected.npz maps "<scene>[__<param>]__<cpu|gpu>" -> uint8
array (64, 64, 4); it is data, not reference source code.
"""
import glob
import os

import numpy as np
from PIL import Image

SRC = "/root/reference/e2e-tests/expected"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "e2e_expected.npz")

arrays = {}
for path in sorted(glob.glob(os.path.join(SRC, "*.png"))):
    name = os.path.basename(path)[len("tests__"):-len(".png")]
    img = np.array(Image.open(path).convert("RGBA"), dtype=np.uint8)
    assert img.shape == (64, 64, 4), (name, img.shape)
    arrays[name] = img
np.savez_compressed(OUT, **arrays)
print(f"wrote {OUT}: {len(arrays)} images")
