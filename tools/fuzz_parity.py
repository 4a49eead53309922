"""Dev tool: randomised parity sweep on a GPU box -- random shapes, ratios (with a bias towards the exact
integer ratios the specialised kernels detect), channel counts, filters and fused channel conversions,
each checked against the CPU oracle (max |delta| <= 1 in FAST mode, 0 in EXACT mode).
    python tools/fuzz_parity.py [cases=300] [seed=1]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rust-image-transform_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import imagekit_cuda as ik
from oracle import oracle
from conftest import splitmix_noise, checker

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
ctx = ik.Context([0])
bad = 0
t0 = time.time()
for case in range(n_cases):
    kind = rng.choice(["down_int", "down_any", "up2", "up_any", "mixed", "down2_rgba"], p=[0.22, 0.25, 0.13, 0.1, 0.1, 0.2])
    c = int(rng.choice([1, 2, 3, 4], p=[0.1, 0.1, 0.4, 0.4]))
    filt = int(rng.choice([0, 1, 2, 3, 4], p=[0.05, 0.1, 0.2, 0.1, 0.55]))
    if kind == "down_int":
        r = int(rng.choice([2, 2, 2, 3, 4, 4, 5]))
        dw, dh = int(rng.integers(8, 700)), int(rng.integers(8, 500))
        w, h = dw * r, dh * r
        if rng.random() < 0.3: h = dh * int(rng.choice([2, 3, 4]))       # different integer ratios per axis
    elif kind == "down2_rgba":   # the row-band tensor-core kernel: Rgba8, exactly 2:1 horizontally, ~1.5 .. 2.3 : 1 vertically
        c = 4
        filt = int(rng.choice([4, 4, 4, 3]))
        dw = int(rng.integers(6, 800)) * (4 if rng.random() < 0.8 else 1)
        dh = int(rng.integers(8, 700))
        w = 2 * dw
        h = 2 * dh if rng.random() < 0.6 else int(dh * rng.uniform(1.5, 2.3))
    elif kind == "down_any":
        w, h = int(rng.integers(16, 2200)), int(rng.integers(16, 1600))
        dw, dh = int(rng.integers(1, w + 1)), int(rng.integers(1, h + 1))
    elif kind == "up2":
        w, h = int(rng.integers(1, 500)), int(rng.integers(1, 400))
        dw, dh = 2 * w, 2 * h
    elif kind == "up_any":
        w, h = int(rng.integers(1, 300)), int(rng.integers(1, 300))
        dw, dh = int(rng.integers(w, 3 * w + 2)), int(rng.integers(h, 3 * h + 2))
    else:
        w, h = int(rng.integers(8, 1200)), int(rng.integers(8, 1200))
        dw, dh = int(rng.integers(1, 2 * w)), int(rng.integers(1, 2 * h))
    if (w, h) == (dw, dh): dw += 1
    co = None
    if rng.random() < 0.25: co = int(rng.choice([3, 4]))
    exact = rng.random() < 0.15
    src = (checker if rng.random() < 0.2 else splitmix_noise)((h, w, c))
    if co is None and rng.random() < 0.15:  # 16-bit rasters (tile / generic kernels)
        src = (src.astype(np.uint16) * 257) ^ rng.integers(0, 256, src.shape, dtype=np.uint16)
    ctx.set_mode(ik.MODE_EXACT if exact else ik.MODE_FAST)
    how = rng.random()
    if src.dtype == np.uint8 and how < 0.15:
        got = ctx.submit(src, dw, dh, filt, out_channels=co)                       # coalescing queue
    elif src.dtype == np.uint8 and how < 0.3:
        got = ctx.resize_begin(src, dw, dh, filt, out_channels=co).end()           # split call
    else:
        got = ctx.resize(src, dw, dh, filt, out_channels=co)
    want = oracle.resize_exact(src, dw, dh, filt)
    if co == 3: want = oracle.to_rgb8(want)
    if co == 4: want = oracle.to_rgba8(want)
    d = int(np.abs(got.astype(np.int32) - want.astype(np.int32)).max()) if got.shape == want.shape else 999
    if d > (0 if exact else 1):
        bad += 1
        print("FAIL", dict(kind=kind, h=h, w=w, c=c, dw=dw, dh=dh, filt=filt, co=co, exact=exact, max_delta=d), flush=True)
print(f"{n_cases} cases, {bad} failures, {time.time() - t0:.1f} s; launches per family: { {k: v for k, v in ctx.stats().items() if k.startswith('launches_')} }")
sys.exit(1 if bad else 0)
