"""Dev tool: one resize against the oracle, printing where it differs.  usage: debug_case.py h w c dw dh [filter]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rust-image-transform_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import imagekit_cuda as ik
from oracle import oracle
from conftest import splitmix_noise
h, w, c, dw, dh = (int(v) for v in sys.argv[1:6])
filt = int(sys.argv[6]) if len(sys.argv) > 6 else 4
ctx = ik.Context([0])
src = splitmix_noise((h, w, c))
got = ctx.resize(src, dw, dh, filt)
want = oracle.resize_exact(src, dw, dh, filt)
d = np.abs(got.astype(int) - want.astype(int))
bad = np.argwhere(d > 1)
print("launch families", {k: v for k, v in ctx.stats().items() if k.startswith("launches_") and v}, "max |d|", d.max(), "bad", len(bad), "of", d.size)
if len(bad):
    ys, xs = np.unique(bad[:, 0]), np.unique(bad[:, 1])
    print("bad rows", ys[:20], "...", ys[-5:], "n", len(ys))
    print("bad cols", xs[:20], "...", xs[-5:], "n", len(xs))
    y, x, ch = bad[0]
    print("first bad", (y, x, ch), "got", got[y, x], "want", want[y, x])
from imagekit_cuda import engine
print("v pass", engine.pass_info(filt, h, dh), "band8t", (lambda b: None if b is None else (b[0], b[3], b[1][:8]))(engine.pass_band8t(filt, h, dh)))
