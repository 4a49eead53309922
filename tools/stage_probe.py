"""Dev tool: per-call time of ikc_resize_u8 against source size, pageable vs pinned (slope = staging copy rate,
intercept = fixed cost of the staged path).  python tools/stage_probe.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rust-image-transform_b200"))
import numpy as np
import imagekit_cuda as ik
ctx = ik.Context([0])
print("chunk_kb", os.environ.get("IKC_STAGE_CHUNK_KB"), "piece_kb", os.environ.get("IKC_COPY_PIECE_KB"))
for h in (135, 270, 540, 1080, 2160, 4320):
    sw, ch = 1920, 3
    dw, dh = 400, max(1, h * 400 // 1920)
    pin_s, pin_d = ik.PinnedArray((h, sw, ch)), ik.PinnedArray((dh, dw, ch))
    pin_s.array[...] = 77
    page_s, page_d = np.array(pin_s.array), np.zeros((dh, dw, ch), np.uint8)
    row = [f"{h * sw * ch / 1e6:6.2f} MB"]
    for src, dst in ((page_s, page_d), (pin_s.array, pin_d.array), (page_s, pin_d.array), (pin_s.array, page_d)):
        for _ in range(10):
            ctx.resize(src, dw, dh, 4, out=dst)
        t = []
        for _ in range(100):
            t0 = time.perf_counter(); ctx.resize(src, dw, dh, 4, out=dst); t.append(time.perf_counter() - t0)
        row.append(f"{np.median(t) * 1e6:8.1f}")
    print(" ".join(row), "  (us: page->page, pin->pin, page->pin, pin->page)")
