"""Dev tool: one source size, pageable -> pinned, per-call time under the current env knobs."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rust-image-transform_b200"))
import numpy as np
import imagekit_cuda as ik
ctx = ik.Context([0])
out = []
for h in ((int(os.environ["PROBE_H"]),) if "PROBE_H" in os.environ else (270, 1080, 4320)):
    sw, ch = 1920, 3
    dw, dh = 400, max(1, h * 400 // 1920)
    pin_d = ik.PinnedArray((dh, dw, ch))
    page_s = np.full((h, sw, ch), 77, np.uint8)
    for _ in range(10):
        ctx.resize(page_s, dw, dh, 4, out=pin_d.array)
    t = []
    for _ in range(100):
        t0 = time.perf_counter(); ctx.resize(page_s, dw, dh, 4, out=pin_d.array); t.append(time.perf_counter() - t0)
    out.append(f"{h * sw * ch / 1e6:.2f}MB {np.median(t) * 1e6:.0f}us")
print({k: os.environ.get(k) for k in ("IKC_COPY_HELPERS", "IKC_STAGE_CHUNK_KB", "IKC_COPY_PIECE_KB")}, " | ".join(out), flush=True)
ctx.close()
