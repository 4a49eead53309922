"""Generates tests/golden/resize_golden.npz from the CPU oracle (oracle/imageops_oracle.c).

The reference (/root/reference) has no Rust toolchain here and ships no pixel fixtures, so these
vectors are SELF-GENERATED regression pins of the restated algorithm ("parity unpinned" with respect
to the real image 0.25.8 binary).  Re-run: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
from conftest import checker, photo_like, splitmix_noise  # noqa: E402
from oracle import oracle  # noqa: E402

CASES = [  # name, generator, (h, w, c), (dw, dh), filter
    ("noise_rgb_down_l3", splitmix_noise, (60, 80, 3), (40, 30), 4),
    ("noise_rgba_down2_l3", splitmix_noise, (64, 96, 4), (48, 32), 4),
    ("noise_rgb_thumb_l3", splitmix_noise, (121, 161, 3), (16, 12), 4),
    ("noise_rgb_up_catmull", splitmix_noise, (24, 32, 3), (64, 48), 2),
    ("checker_rgb_down_l3", checker, (70, 70, 3), (33, 33), 4),
    ("photo_rgba_down_gauss", photo_like, (50, 66, 4), (25, 19), 3),
    ("photo_luma_down_tri", photo_like, (45, 45, 1), (20, 17), 1),
    ("noise_la_nearest", splitmix_noise, (33, 47, 2), (19, 13), 0),
    ("noise_rgb_to_1x1", splitmix_noise, (30, 40, 3), (1, 1), 4),
    ("tiny_2x2_up_l3", splitmix_noise, (2, 2, 3), (20, 20), 4),
]

out = {}
for name, gen, shape, (dw, dh), filt in CASES:
    src = gen(shape)
    out[name + "/src"] = src
    out[name + "/meta"] = np.array([dw, dh, filt], np.int32)
    out[name + "/out"] = oracle.resize_exact(src, dw, dh, filt)
np.savez_compressed(os.path.join(HERE, "resize_golden.npz"), **out)
print("wrote", len(CASES), "cases")
