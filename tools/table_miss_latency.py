"""Dev tool / record: call latency of ikc_resize_u8 when every call misses the weight-table cache (a service fed
arbitrary target sizes).  8 threads x N random target sizes on one context; prints p50 / p99 per call as JSON.
    python tools/table_miss_latency.py [calls_per_thread]"""
import json, os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rust-image-transform_b200"))
import numpy as np
import imagekit_cuda as ik


def run(calls=500, threads=8, seed=1):
    ctx = ik.Context([0])
    src = np.random.default_rng(seed).integers(0, 256, (480, 640, 3), dtype=np.uint8)
    ctx.resize(src, 100, 75, ik.FILTER_LANCZOS3)  # warm the context (streams, lane buffers)
    lat = [[] for _ in range(threads)]

    def worker(t):
        rng = np.random.default_rng(seed * 1000 + t)
        for _ in range(calls):
            dw, dh = int(rng.integers(40, 600)), int(rng.integers(40, 440))   # almost always a size nobody asked for before
            t0 = time.perf_counter()
            ctx.resize(src, dw, dh, ik.FILTER_LANCZOS3)
            lat[t].append(time.perf_counter() - t0)

    th = [threading.Thread(target=worker, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    for x in th: x.start()
    for x in th: x.join()
    wall = time.perf_counter() - t0
    a = np.sort(np.concatenate([np.asarray(v) for v in lat]))
    ctx.close()
    return {"threads": threads, "calls_per_thread": calls, "p50_ms": float(a[len(a) // 2] * 1e3), "p90_ms": float(a[int(len(a) * 0.9)] * 1e3),
            "p99_ms": float(a[int(len(a) * 0.99)] * 1e3), "max_ms": float(a[-1] * 1e3), "calls_per_s": threads * calls / wall,
            "workload": "640x480 RGB8 -> random (40..600) x (40..440) Lanczos3, pageable numpy buffers, one context, every call a weight-table miss"}


if __name__ == "__main__":
    print(json.dumps(run(int(sys.argv[1]) if len(sys.argv) > 1 else 500)))
