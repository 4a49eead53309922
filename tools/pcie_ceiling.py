"""Dev tool: the host <-> device DMA ceiling of the box, which bounds every host-resident (drop-in) number.
Plain cudaMemcpyAsync H2D + D2H from / to pinned memory on N GPUs at once, driven from ONE process (one stream pair
per GPU) -- the regime of ikc_resize_batch's in-process sharding.  bench.py measures the N-process regime itself
(e2e.dma_ceiling).  usage: python tools/pcie_ceiling.py [n_gpus] [MB per copy]"""
import json
import sys
import time

import torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
mb = int(sys.argv[2]) if len(sys.argv) > 2 else 256
res = {}
for g in sorted({1, n} | {k for k in (2, 4, 8) if k <= n}):
    bufs = []
    for d in range(g):
        dev = torch.device("cuda", d)
        bufs.append((torch.empty(mb << 20, dtype=torch.uint8).pin_memory(), torch.empty(mb << 20, dtype=torch.uint8, device=dev),
                     torch.empty(mb << 18, dtype=torch.uint8).pin_memory(), torch.empty(mb << 18, dtype=torch.uint8, device=dev),
                     torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)))

    def step():
        for h_in, d_in, h_out, d_out, s1, s2 in bufs:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)     # H2D: an image's worth
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)   # D2H: a quarter of it (2:1 downscale)

    for _ in range(3):
        step()
    for d in range(g):
        torch.cuda.synchronize(d)
    t0 = time.perf_counter()
    reps = 10
    for _ in range(reps):
        step()
    for d in range(g):
        torch.cuda.synchronize(d)
    dt = time.perf_counter() - t0
    res[f"{g}_gpus"] = {"h2d_gbs_total": g * reps * (mb << 20) / dt / 1e9, "d2h_gbs_total": g * reps * (mb << 18) / dt / 1e9,
                        "h2d_gbs_per_gpu": reps * (mb << 20) / dt / 1e9}
    del bufs
print(json.dumps({"tool": "pcie_ceiling", "process_model": "one process, one stream pair per GPU", "mb_per_h2d_copy": mb, "results": res}))
