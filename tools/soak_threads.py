"""Dev tool: multi-threaded soak of the host entry points on one context -- N threads, each a random mix of
ikc_resize_u8 / ikc_submit_u8 / begin-end / small batches over random shapes (pageable and pinned buffers), every result
checked against the CPU oracle (max |delta| <= 1).  Exercises the copy pool, the lanes, the submit queue and the
persistent kernels under contention.     python tools/soak_threads.py [threads=16] [jobs_per_thread=150] [seed=1]"""
import os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rust-image-transform_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import imagekit_cuda as ik
from oracle import oracle

n_threads = int(sys.argv[1]) if len(sys.argv) > 1 else 16
n_jobs = int(sys.argv[2]) if len(sys.argv) > 2 else 150
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 1
ctx = ik.Context([0])
errors, done = [], [0] * n_threads


def shape(rng):
    c = int(rng.choice([1, 3, 3, 4, 4]))
    kind = rng.integers(0, 5)
    if kind == 0:      # thumbnails of mid-sized rasters (banded8, persistent CTAs)
        w, h = int(rng.integers(600, 2400)), int(rng.integers(400, 1600))
        dw = int(rng.integers(60, 400)); dh = max(1, h * dw // w)
    elif kind == 1:    # exact 2:1 rgba (banded8t)
        c = 4; dw, dh = int(rng.integers(40, 500)) * 4, int(rng.integers(40, 400)); w, h = 2 * dw, 2 * dh
    elif kind == 2:    # exact 2x upscale (banded8u / up2)
        w, h = int(rng.integers(40, 500)) * 4, int(rng.integers(40, 400)); dw, dh = 2 * w, 2 * h
    elif kind == 3:    # anything
        w, h, dw, dh = (int(rng.integers(8, 900)) for _ in range(4))
    else:              # large source: the staged upload runs over many pieces
        w, h = int(rng.integers(2500, 4200)), int(rng.integers(1500, 3000)); dw = int(rng.integers(100, 800)); dh = max(1, h * dw // w)
    return h, w, c, dw, dh


def check(got, src, dw, dh, filt, tag):
    want = oracle.resize_exact(src, dw, dh, filt)
    d = np.abs(got.astype(np.int16) - want.astype(np.int16)).max() if got.size else 0
    if got.shape != want.shape or d > 1:
        raise AssertionError(f"{tag}: shape {src.shape} -> {(dh, dw)} filt {filt}: max delta {d}")


def worker(t):
    rng = np.random.default_rng(seed * 1000 + t)
    try:
        for j in range(n_jobs):
            h, w, c, dw, dh = shape(rng)
            filt = int(rng.choice([4, 4, 4, 2, 1, 3]))
            src = rng.integers(0, 256, (h, w, c), dtype=np.uint8)
            how = rng.integers(0, 4)
            if how == 0:
                check(ctx.resize(src, dw, dh, filt), src, dw, dh, filt, "resize")
            elif how == 1:
                check(ctx.submit(src, dw, dh, filt), src, dw, dh, filt, "submit")
            elif how == 2:
                tk = ctx.resize_begin(src, dw, dh, filt)
                check(tk.end(), src, dw, dh, filt, "begin/end")
            else:
                more = [rng.integers(0, 256, (max(8, h // 3), max(8, w // 3), c), dtype=np.uint8) for _ in range(3)]
                srcs = [src] + more
                sizes = [(dw, dh)] + [(max(1, dw // 2), max(1, dh // 2))] * 3
                outs, _ = ctx.resize_batch(srcs, sizes, filt)
                for s, (a, b), o in zip(srcs, sizes, outs):
                    check(o, s, a, b, filt, "batch")
            done[t] = j + 1
    except Exception as e:  # noqa: BLE001
        errors.append(f"thread {t} job {done[t]}: {e!r}")


t0 = time.time()
th = [threading.Thread(target=worker, args=(t,)) for t in range(n_threads)]
for x in th: x.start()
for x in th: x.join()
st = ctx.stats()
print(f"{n_threads} threads x {n_jobs} jobs: {sum(done)} done, {len(errors)} errors, {time.time() - t0:.1f} s; calls {st['calls']} failed {st['failed']} "
      f"launches {st['launches']} submit groups/jobs {st['submit_batches']}/{st['submit_jobs']} table misses {st['table_misses']}")
for e in errors[:5]: print(e)
ctx.close()
sys.exit(1 if errors else 0)
