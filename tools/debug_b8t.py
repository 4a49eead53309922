"""Dev tool: parity of the 2:1 Rgba8 row-band kernel (banded8t) against the CPU oracle on a few shapes; prints where the
differences are."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rust-image-transform_b200"))
import numpy as np
import imagekit_cuda as ik
from oracle import oracle
ctx = ik.Context([0])
rng = np.random.default_rng(7)
shapes = [(512, 384), (640, 300), (3840, 2160), (1000, 700), (258, 130)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for sw, sh in shapes:
    dw, dh = sw // 2, sh // 2
    src = rng.integers(0, 256, (sh, sw, 4), dtype=np.uint8)
    got = ctx.resize(src, dw, dh, ik.FILTER_LANCZOS3)
    want = oracle.resize_exact(src, dw, dh, oracle.LANCZOS3)
    d = got.astype(np.int16) - want.astype(np.int16)
    bad = np.argwhere(np.abs(d) > 1)
    print(f"{sw}x{sh} -> {dw}x{dh}: max |d| = {np.abs(d).max()}, hist = {dict(zip(*np.unique(d, return_counts=True)))}" if np.abs(d).max() <= 3 else
          f"{sw}x{sh} -> {dw}x{dh}: max |d| = {np.abs(d).max()}, bad = {len(bad)} of {d.size}")
    if len(bad):
        ys, xs = np.unique(bad[:, 0]), np.unique(bad[:, 1])
        print("  bad rows", ys[:12], "...", ys[-4:], " bad cols", xs[:16], "...", xs[-6:])
        y, x, c = bad[0]
        print("  first bad", (y, x, c), "got", got[y, x], "want", want[y, x])
