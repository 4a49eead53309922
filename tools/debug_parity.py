"""Dev tool: print where the fused path differs from the oracle."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rust-image-transform_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import imagekit_cuda as ik
from oracle import oracle
from conftest import splitmix_noise

ctx = ik.Context([0])
cases = [(480, 640, 3, 200, 150, 4, "const"), (480, 640, 3, 200, 150, 4, "noise"), (480, 640, 4, 320, 240, 4, "noise"),
         (64, 96, 4, 48, 32, 4, "noise"), (60, 80, 3, 40, 30, 4, "noise"), (2160, 3840, 4, 1920, 1080, 4, "noise")]
for (h, w, c, dw, dh, f, kind) in cases:
    src = splitmix_noise((h, w, c), image_id=10 + f) if kind == "noise" else np.full((h, w, c), 37, np.uint8)
    got = ctx.resize(src, dw, dh, f)
    want = oracle.resize_exact(src, dw, dh, f)
    d = got.astype(int) - want.astype(int)
    bad = np.argwhere(np.abs(d) > 1)
    print((h, w, c, dw, dh, f, kind), "max|d|", np.abs(d).max(), "n_bad", len(bad), "of", d.size)
    if len(bad):
        rows = np.unique(bad[:, 0]); cols = np.unique(bad[:, 1])
        print("  bad rows:", rows[:60].tolist())
        print("  bad cols:", cols[:80].tolist())
        print("  sample:", [(int(y), int(x), int(ch), int(got[y, x, ch]), int(want[y, x, ch])) for y, x, ch in bad[:10]])
