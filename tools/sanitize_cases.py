"""Dev tool: a handful of small resizes that reach every kernel (tensor-core downscale kernels general / uniform /
converting / row-band, tile, up2 vector + ragged paths, generic), checked against the oracle.  Meant to be run under compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/sanitize_cases.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rust-image-transform_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import imagekit_cuda as ik
from oracle import oracle
from conftest import splitmix_noise

ctx = ik.Context([0])
cases = [  # (h, w, c, dw, dh, filter, out_channels)
    (480, 640, 3, 200, 150, 4, None), (600, 800, 4, 400, 300, 4, None), (600, 800, 4, 400, 300, 4, 3),
    (300, 400, 3, 200, 150, 4, 4), (512, 1024, 4, 256, 128, 4, None), (777, 1031, 3, 515, 388, 4, None),
    (240, 320, 3, 640, 480, 2, None), (120, 161, 4, 322, 240, 4, None), (64, 64, 3, 128, 128, 1, None),
    (100, 90, 2, 45, 50, 2, 4), (90, 100, 1, 300, 270, 3, None), (33, 17, 3, 34, 66, 0, None),
    # the row-band tensor-core kernel (2:1 Rgba8): partial last band, image borders, one block, odd vertical ratio
    (384, 512, 4, 256, 192, 4, None), (130, 264, 4, 132, 65, 4, None), (64, 96, 4, 48, 32, 4, None), (700, 1000, 4, 500, 400, 4, None),
]
worst = 0
for h, w, c, dw, dh, filt, co in cases:
    src = splitmix_noise((h, w, c), image_id=h)
    before = ctx.kernel_launches
    got = ctx.resize(src, dw, dh, filt, out_channels=co)
    want = oracle.resize_exact(src, dw, dh, filt)
    if co == 3: want = oracle.to_rgb8(want)
    if co == 4: want = oracle.to_rgba8(want)
    d = int(np.abs(got.astype(np.int32) - want.astype(np.int32)).max())
    worst = max(worst, d)
    print((h, w, c, dw, dh, filt, co), "launches", ctx.kernel_launches - before, "max |delta|", d)
assert worst <= 1, worst
print("ok")
