#!/usr/bin/env python
"""bench.py -- resize output MP/s on B200, with the HBM roofline and the CPU path beside it.

One "step" = one pass of the hot path (resize_image's resampling, /root/reference/src/transform.rs:85-89)
over one batch of synthetic rasters.  Default workload = BASELINE.json configs[1]:
3840x2160 RGBA8 -> 1920x1080 Lanczos3, as a batch of distinct images per step (working set >> L2).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg3|cfg4|cfg1] [--batch B]
  python bench.py --impl reference ...     # the CPU path (oracle port of image 0.25.8) on the host cores

Multi-GPU (torchrun, one rank per GPU): every rank resizes its own batch (independent images, no
collective on the data path -> weak scaling); time = max over ranks; value = all ranks' output MP / time.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "rust-image-transform_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

LANCZOS3, CATMULLROM = 4, 2
WORKLOADS = {
    # name: (sw, sh, channels, dw, dh, filter, default batch, description)
    "cfg2": (3840, 2160, 4, 1920, 1080, LANCZOS3, 32, "3840x2160 RGBA8 -> 1920x1080 Lanczos3"),
    "cfg3": (4032, 3024, 3, 400, 300, LANCZOS3, 64, "4032x3024 RGB8 -> 400x300 Lanczos3 thumbnails"),
    "cfg4": (1920, 1080, 3, 3840, 2160, CATMULLROM, 32, "1920x1080 RGB8 -> 3840x2160 CatmullRom upscale"),
    "cfg1": (1920, 1080, 3, 400, 225, LANCZOS3, 64, "1920x1080 RGB8 -> 400x225 Lanczos3"),
}
FILTER_NAMES = {LANCZOS3: "lanczos3", CATMULLROM: "catmullrom"}
METRIC = "resize output MP/s"
UNIT = "MP/s"


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def recorded_traffic(workload: str, batch: int):
    """DRAM bytes per launch from the committed ncu capture (profiles/roofline_traffic.json), if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            t = json.load(f).get(workload)
        if t and t.get("batch"):
            return float(t["dram_bytes_per_launch"]) * batch / float(t["batch"])
    except Exception:  # noqa: BLE001
        pass
    return None


class ClockSampler:
    """Samples SM clock / throttle reasons with NVML every 20 ms while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self._nv = None

    def _run(self):
        nv = self._nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:  # noqa: BLE001
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        if self._nv:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_port_rate(sw, sh, ch, dw, dh, filt, images: int, threads: int):
    """Output MP/s of the CPU port (oracle) over `images` independent rasters on `threads` threads."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle
    oracle.build()
    rng = np.random.default_rng(0)
    srcs = [rng.integers(0, 256, (sh, sw, ch), dtype=np.uint8) for _ in range(min(images, threads))]
    oracle.resize_exact(srcs[0][: max(8, sh // 8), : max(8, sw // 8)], max(1, dw // 8), max(1, dh // 8), filt)
    t0 = time.perf_counter()
    if threads == 1:
        for i in range(images):
            oracle.resize_exact(srcs[i % len(srcs)], dw, dh, filt)
    else:
        with ThreadPoolExecutor(threads) as ex:  # ctypes releases the GIL inside the C oracle
            list(ex.map(lambda i: oracle.resize_exact(srcs[i % len(srcs)], dw, dh, filt), range(images)))
    dt = time.perf_counter() - t0
    return images * dw * dh / 1e6 / dt, dt


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  The reference itself (Rust,
    crate image 0.25.8) cannot be built in this image, so this is the oracle port, with all host
    threads it can use (independent images per thread, as a saturated reference server would)."""
    if rank != 0:
        return
    sw, sh, ch, dw, dh, filt, batch, desc = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, 64))
    per_step = threads  # one image per thread per step: a bounded sample of the workload
    for _ in range(min(args.warmup, 1)):
        cpu_port_rate(sw, sh, ch, dw, dh, filt, per_step, threads)
    rates, t_total = [], 0.0
    for _ in range(args.steps):
        r, dt = cpu_port_rate(sw, sh, ch, dw, dh, filt, per_step, threads)
        rates.append(r)
        t_total += dt
    value = per_step * args.steps * dw * dh / 1e6 / t_total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "filter": FILTER_NAMES[filt]},
        "run": {"images_per_step": per_step,
                "note": "CPU port of image 0.25.8 imageops::resize (oracle/imageops_oracle.c); the Rust "
                        "reference cannot be built here (no cargo/rustc, crate not vendored)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{per_step} images/step x {args.steps} steps, one image per thread"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_cfg3_shard(ik, ctx, torch, dev, dist, rank, world, barrier, args):
    """BASELINE configs[2]: a batch of 1024 synthetic 4032x3024 RGB8 rasters -> 400x300 Lanczos3 thumbnails, sharded
    round robin over the ranks (image i -> rank i mod N, no collective): strong scaling.  Device-resident time
    (CUDA events, max over ranks) and end to end from pinned host memory (H2D + kernel + D2H)."""
    from imagekit_cuda.sharding import aggregate_throughput, shard_indices
    sw, sh, ch, dw, dh, filt = 4032, 3024, 3, 400, 300, LANCZOS3
    total = args.shard_images
    mine = shard_indices(total, world, rank)
    n = len(mine)
    free_b, _ = torch.cuda.mem_get_info(dev)
    distinct = n if n * sw * sh * ch < 0.6 * free_b else max(8, int(0.4 * free_b) // (sw * sh * ch))
    g = torch.Generator(device=dev)
    g.manual_seed(0xC0FFEE + rank)
    src = torch.empty((distinct, sh, sw, ch), dtype=torch.uint8, device=dev)
    for a in range(0, distinct, 16):
        b = min(distinct, a + 16)
        src[a:b] = torch.randint(0, 256, (b - a, sh, sw, ch), dtype=torch.uint8, device=dev, generator=g)
    dst = torch.zeros((n, dh, dw, ch), dtype=torch.uint8, device=dev)
    jobs = [(src[i % distinct].data_ptr(), sw, sh, sw * ch, dst[i].data_ptr(), dw, dh, dw * ch, ch, filt) for i in range(n)]
    prepared = ctx.prepare_batch(0, jobs)
    stream = torch.cuda.Stream(device=dev)
    for _ in range(2):
        prepared.launch(stream.cuda_stream)
    barrier()
    steps = 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        prepared.launch(stream.cuda_stream)
    e1.record(stream)
    stream.synchronize()
    ms = e0.elapsed_time(e1)
    barrier()
    _, ms_max, value = aggregate_throughput(n * steps * dw * dh / 1e6, ms, dist, dev)
    kernel = prepared.describe()
    per_img_us = ms / steps / max(1, n) * 1e3
    prepared.free()
    # end to end: the rank's share pushed through ikc_resize_batch from 8 pinned source buffers used in turn
    h_src = [ik.PinnedArray((sh, sw, ch)) for _ in range(8)]
    rng = np.random.default_rng(100 + rank)
    for a in h_src:
        a.array[...] = rng.integers(0, 256, a.shape, dtype=np.uint8)
    h_dst = ik.PinnedArray((n, dh, dw, ch))
    srcs = [h_src[i % 8].array for i in range(n)]
    outs = [h_dst.array[i] for i in range(n)]
    ctx.resize_batch(srcs[:16], [(dw, dh)] * min(n, 16), filt, outs=outs[:16])
    barrier()
    t0 = time.perf_counter()
    ctx.resize_batch(srcs, [(dw, dh)] * n, filt, outs=outs)
    dt = time.perf_counter() - t0
    _, _, e2e = aggregate_throughput(n * dw * dh / 1e6, dt * 1e3, dist, dev)
    algo = sw * sh * ch + dw * dh * ch
    peak, _ = measured_peak()
    del src, dst
    torch.cuda.empty_cache()
    return {"workload": f"cfg3: batch of {total} x 4032x3024 RGB8 -> 400x300 Lanczos3, sharded over {world} GPU(s)",
            "scaling": "strong", "images_total": total, "images_this_rank": n, "distinct_sources_this_rank": distinct,
            "value": value, "unit": UNIT, "thumbnails_per_s": value / (dw * dh / 1e6), "ms_batch": ms_max / steps,
            "us_per_image_this_rank": per_img_us, "kernel": kernel,
            "roofline_frac_this_rank": algo / (per_img_us * 1e-6) / 1e9 / peak,
            "e2e": {"value": e2e, "unit": UNIT, "thumbnails_per_s": e2e / (dw * dh / 1e6), "s_batch": dt,
                    "h2d_bytes": n * sw * sh * ch, "d2h_bytes": n * dw * dh * ch,
                    "api": "one ikc_resize_batch call per rank over its shard, pinned host buffers"}}


def run_other_workloads(ik, ctx, torch, dev, dist, barrier, names):
    """The other BASELINE workloads (configs 1 and 4; config 3 has its own leg), device-resident, their default batch per launch, the
    same timing rules as the headline (inputs larger than L2, CUDA events on the launch stream, max over ranks): so that
    the driver's record carries their roofline fractions too, not only the builder's runs."""
    from imagekit_cuda.sharding import aggregate_throughput
    peak, _ = measured_peak()
    out = {}
    for name in names:
        sw, sh, ch, dw, dh, filt, batch, desc = WORKLOADS[name]   # (the workload's own default batch, as `--workload name` runs it)
        g = torch.Generator(device=dev)
        g.manual_seed(0xBEEF + len(name))
        src = torch.randint(0, 256, (batch, sh, sw, ch), dtype=torch.uint8, device=dev, generator=g)
        dst = torch.zeros((batch, dh, dw, ch), dtype=torch.uint8, device=dev)
        jobs = [(src[i].data_ptr(), sw, sh, sw * ch, dst[i].data_ptr(), dw, dh, dw * ch, ch, filt) for i in range(batch)]
        prepared = ctx.prepare_batch(0, jobs)
        stream = torch.cuda.Stream(device=dev)
        for _ in range(3):
            prepared.launch(stream.cuda_stream)
        stream.synchronize()
        barrier()
        steps = 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            prepared.launch(stream.cuda_stream)
        e1.record(stream)
        stream.synchronize()
        ms = e0.elapsed_time(e1)
        barrier()
        _, ms_max, value = aggregate_throughput(batch * steps * dw * dh / 1e6, ms, dist, dev)
        algo = batch * (sw * sh * ch + dw * dh * ch)
        out[name] = {"workload": f"{name}: {desc}", "value": value, "unit": UNIT, "images_per_launch_per_gpu": batch, "steps": steps,
                     "us_per_image_this_rank": ms / steps / batch * 1e3, "kernel": prepared.describe(),
                     "roofline_frac_this_rank": algo / (ms / steps * 1e-3) / 1e9 / peak,
                     "working_set_mb": algo / 1e6}
        prepared.free()
        del src, dst
        torch.cuda.empty_cache()
    return out


def run_call_latency(ik, ctx, names, no_cpu):
    """Rank 0, one thread: what ONE request sees (config 1 is latency-bound, report us per call).  Each call
    is one synchronous ikc_resize_u8 -- host buffer in, H2D, kernels, D2H, host buffer out -- from pageable memory (the Rust
    wrapper's Vec<u8>) and from pinned memory, beside the CPU port's single-thread time for the same resize (what the same
    request costs in the reference, which resizes inline on one worker)."""
    out = {}
    for name in names:
        sw, sh, ch, dw, dh, filt, _, desc = WORKLOADS[name]
        rng = np.random.default_rng(len(name))
        pin_s, pin_d = ik.PinnedArray((sh, sw, ch)), ik.PinnedArray((dh, dw, ch))
        pin_s.array[...] = rng.integers(0, 256, pin_s.shape, dtype=np.uint8)
        page_s, page_d = np.array(pin_s.array), np.empty((dh, dw, ch), np.uint8)
        calls = 200 if sw * sh <= 4_000_000 else 60
        rec = {"workload": f"{name}: {desc}", "api": "ikc_resize_u8 (C ABI), one synchronous call per image, one thread", "calls": calls}
        for label, src, dst in (("pageable", page_s, page_d), ("pinned", pin_s.array, pin_d.array)):
            for _ in range(10):
                ctx.resize(src, dw, dh, filt, out=dst)
            t = np.empty(calls)
            for i in range(calls):
                t0 = time.perf_counter()
                ctx.resize(src, dw, dh, filt, out=dst)
                t[i] = time.perf_counter() - t0
            rec[label] = {"p50_us": float(np.percentile(t, 50) * 1e6), "p99_us": float(np.percentile(t, 99) * 1e6),
                          "mean_us": float(t.mean() * 1e6)}
        # the call the Rust wrapper's resize_image makes (ikc_submit_u8: coalescing queue; a lone request runs on the caller)
        for _ in range(10):
            ctx.submit(page_s, dw, dh, filt)
        t = np.empty(calls)
        for i in range(calls):
            t0 = time.perf_counter()
            ctx.submit(page_s, dw, dh, filt)
            t[i] = time.perf_counter() - t0
        rec["submit_pageable"] = {"p50_us": float(np.percentile(t, 50) * 1e6), "p99_us": float(np.percentile(t, 99) * 1e6),
                                  "mean_us": float(t.mean() * 1e6), "note": "ikc_submit_u8, result buffer allocated per call"}
        if not no_cpu:
            n = 12 if sw * sh <= 4_000_000 else 4
            v, dt = cpu_port_rate(sw, sh, ch, dw, dh, filt, n, 1)
            rec["cpu_port_1_thread_us"] = dt / n * 1e6
        out[name] = rec
        pin_s.free()
        pin_d.free()
    return out


def run_cfg5_upload(ik, ctx, dist, dev, rank, world, barrier, args):
    """BASELINE config 5, the /upload shape (reference src/lib.rs:246-309): CPU decode -> GPU resize -> CPU webp q=80 encode
    over 256 synthetic 8 MP JPEGs, split over the ranks (one GPU each) and, inside a rank, over its share of the host
    cores.  Every worker thread runs the library's split call: decode(i + 1) and encode(i - 1) happen between
    ikc_resize_begin_u8(i) and ikc_resize_end(i), so the upload, kernel and download of image i hide behind the codecs;
    the resize stores rgb8 directly (to_rgb8() fused).  Pillow stands in for the reference's image / webp crates."""
    import io
    from concurrent.futures import ThreadPoolExecutor
    from PIL import Image
    from imagekit_cuda.sharding import aggregate_throughput, shard_indices
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import photo_like
    total, width, w8, h8 = args.upload_images, 800, 3264, 2448
    base = photo_like((h8, w8, 3), seed=3)
    jpegs = []
    for i in range(4):
        buf = io.BytesIO()
        Image.fromarray(np.roll(base, 211 * i, axis=1)).save(buf, "JPEG", quality=90)
        jpegs.append(buf.getvalue())
    mine = shard_indices(total, world, rank)
    threads = max(2, (os.cpu_count() or 8) // world)
    tw, th, _ = ik.target_dims(w8, h8, width, None)
    stage = {"decode": 0.0, "resize_queue": 0.0, "resize_wait": 0.0, "encode": 0.0}
    lock = __import__("threading").Lock()

    def decode(b):
        im = Image.open(io.BytesIO(b))
        im.draft("RGB", (w8, h8))
        return np.asarray(im.convert("RGB"))

    def encode(a):
        buf = io.BytesIO()
        Image.fromarray(a).save(buf, "WEBP", quality=80, method=4)
        return buf.tell()

    def worker(idx):
        # one ticket per thread at a time (a ticket holds a lane): decode(i + 1) runs while image i is on the GPU, then
        # end(i), begin(i + 1), and encode(i) runs while image i + 1 is on the GPU
        acc = dict.fromkeys(stage, 0.0)
        prev, nbytes = None, 0
        for i in idx:
            t0 = time.perf_counter()
            d = decode(jpegs[i % len(jpegs)])
            t1 = time.perf_counter()
            out = prev.end() if prev is not None else None
            t2 = time.perf_counter()
            prev = ctx.resize_begin(d, tw, th, ik.FILTER_LANCZOS3, out_channels=3)
            t3 = time.perf_counter()
            if out is not None:
                nbytes += encode(out)
            t4 = time.perf_counter()
            acc["decode"] += t1 - t0
            acc["resize_wait"] += t2 - t1
            acc["resize_queue"] += t3 - t2
            acc["encode"] += t4 - t3
        if prev is not None:
            t2 = time.perf_counter()
            out = prev.end()
            t3 = time.perf_counter()
            nbytes += encode(out)
            acc["resize_wait"] += t3 - t2
            acc["encode"] += time.perf_counter() - t3
        with lock:
            for k in stage:
                stage[k] += acc[k]
        return nbytes

    pool = ThreadPoolExecutor(threads)
    parts = [mine[k::threads] for k in range(threads)]
    list(pool.map(worker, [p[:1] for p in parts if len(p)]))   # warm-up: lane buffers, tables, codec imports
    for k in stage:
        stage[k] = 0.0
    barrier()
    t0 = time.perf_counter()
    out_bytes = sum(pool.map(worker, parts))
    dt = time.perf_counter() - t0
    barrier()
    _, ms_max, per_s = aggregate_throughput(float(len(mine)), dt * 1e3, dist, dev)
    busy = sum(stage.values()) or 1.0
    res = {"workload": f"cfg5: {total} x 8 MP JPEG (3264x2448) -> w={width} Lanczos3 -> webp q=80, {world} GPU(s)",
           "uploads_per_s": per_s, "seconds": ms_max / 1e3, "uploads_this_rank": len(mine), "worker_threads_this_rank": threads,
           "stage_seconds_per_upload_this_rank": {k: v / max(1, len(mine)) for k, v in stage.items()},
           "stage_share_of_worker_time_this_rank": {k: v / busy for k, v in stage.items()},
           "out_bytes_mean": out_bytes / max(1, len(mine)),
           "api": "ikc_resize_begin_u8 / ikc_resize_end (split call, pageable decoder output, rgb8 store fused), one upload in flight per worker"}
    if rank == 0:   # the same pipeline with the reference's CPU resize (oracle port), bounded sample
        from oracle import oracle
        sample = min(len(mine), 2 * threads)

        def cpu_one(i):
            return encode(oracle.resize_image(decode(jpegs[i % len(jpegs)]), width, None))
        t0 = time.perf_counter()
        list(pool.map(cpu_one, range(sample)))
        res["cpu_resize_uploads_per_s"] = {"value": sample / (time.perf_counter() - t0), "threads": threads, "sample": f"{sample} uploads",
                                           "kind": "port"}
    return res


def run_inprocess(ik, n_dev, sw, sh, ch, dw, dh, filt, args):
    """Rank 0 only, the other ranks idle at a barrier: ONE context over all the box's GPUs, one ikc_resize_batch over
    8 images per device.  Checks the round-robin placement (job i -> device i mod N) and that every device returns
    the oracle's pixels (+-1) for the same source."""
    from oracle import oracle
    ctx = ik.Context(list(range(n_dev)))
    per_dev = 8
    n = per_dev * n_dev
    uniq = [ik.PinnedArray((sh, sw, ch)) for _ in range(per_dev)]   # image i uses source i // n_dev: every device sees all 8
    rng = np.random.default_rng(7)
    for a in uniq:
        a.array[...] = rng.integers(0, 256, a.shape, dtype=np.uint8)
    h_dst = [ik.PinnedArray((dh, dw, ch)) for _ in range(n)]
    srcs = [uniq[i // n_dev].array for i in range(n)]
    outs = [a.array for a in h_dst]
    sizes = [(dw, dh)] * n
    _, jobs = ctx.resize_batch(srcs, sizes, filt, outs=outs)
    placement_ok = all(jobs[i].device == i % n_dev and jobs[i].status == 0 for i in range(n))
    want = oracle.resize_exact(uniq[0].array, dw, dh, filt).astype(np.int32)
    max_delta = max(int(np.abs(outs[d].astype(np.int32) - want).max()) for d in range(n_dev))   # jobs 0..N-1: source 0 on each device
    same_across_devices = all(np.array_equal(outs[d], outs[0]) for d in range(n_dev))
    steps = max(3, min(args.steps, 10))
    for _ in range(2):
        ctx.resize_batch(srcs, sizes, filt, outs=outs)
    t0 = time.perf_counter()
    for _ in range(steps):
        ctx.resize_batch(srcs, sizes, filt, outs=outs)
    dt = time.perf_counter() - t0
    ctx.close()
    return {"value": n * steps * dw * dh / 1e6 / dt, "unit": UNIT, "devices": n_dev, "images_per_step": n, "steps": steps,
            "placement_round_robin_ok": bool(placement_ok), "max_abs_delta_vs_oracle_per_device": max_delta,
            "bit_identical_across_devices": bool(same_across_devices),
            "api": "one process, Context() over all devices, ikc_resize_batch (job i -> device i mod N), pinned host buffers"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--e2e-batch", type=int, default=8)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-shard", action="store_true", help="skip the cfg3_shard leg (1024 thumbnails, strong scaling)")
    ap.add_argument("--shard-images", type=int, default=1024)
    ap.add_argument("--no-others", action="store_true", help="skip the other_workloads leg (cfg1, cfg4 device-resident)")
    ap.add_argument("--no-latency", action="store_true", help="skip the call_latency leg (one ikc_resize_u8 call at a time)")
    ap.add_argument("--no-upload", action="store_true", help="skip the cfg5_upload leg (decode -> resize -> webp encode of 8 MP JPEGs)")
    ap.add_argument("--upload-images", type=int, default=256)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        args.steps = args.steps if args.steps is not None else 3
        args.warmup = args.warmup if args.warmup is not None else 1
        run_reference(args, rank, world)
        return
    args.steps = args.steps if args.steps is not None else 200
    args.warmup = max(3, args.warmup if args.warmup is not None else 20)

    import torch
    import imagekit_cuda as ik

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        host_group = dist.new_group(backend="gloo")   # host-side waits that must not occupy the GPUs (the in-process leg)

    sw, sh, ch, dw, dh, filt, batch_default, desc = WORKLOADS[args.workload]
    batch = args.batch or batch_default
    ctx = ik.Context([local_rank])
    ctx.set_mode(ik.MODE_FAST)

    # ---- synthetic inputs, resident in HBM before the timed region (uniform u8 noise)
    g = torch.Generator(device=dev)
    g.manual_seed(0x1234ABCD + rank)
    src = torch.randint(0, 256, (batch, sh, sw, ch), dtype=torch.uint8, device=dev, generator=g)
    dst = torch.zeros((batch, dh, dw, ch), dtype=torch.uint8, device=dev)
    jobs = [(src[i].data_ptr(), sw, sh, sw * ch, dst[i].data_ptr(), dw, dh, dw * ch, ch, filt) for i in range(batch)]
    prepared = ctx.prepare_batch(0, jobs)
    launches_per_step = prepared.launch_count
    stream = torch.cuda.Stream(device=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- parity of what is about to be timed (outside the timed region): image 0 vs the CPU oracle
    parity = None
    if not args.no_parity and rank == 0:
        from oracle import oracle
        prepared.launch(stream.cuda_stream)
        stream.synchronize()
        want = oracle.resize_exact(src[0].cpu().numpy(), dw, dh, filt)
        d = dst[0].cpu().numpy().astype(np.int32) - want.astype(np.int32)
        vals, counts = np.unique(d, return_counts=True)
        parity = {"checked": "image 0 of the batch vs CPU oracle", "max_abs_delta": int(np.abs(d).max()),
                  "delta_histogram": {str(int(v)): int(c) for v, c in zip(vals, counts)}}

    # ---- device-resident timed region
    for _ in range(args.warmup):
        prepared.launch(stream.cuda_stream)
    barrier()
    launches0 = ctx.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(args.steps):
                prepared.launch(stream.cuda_stream)
            e1.record(stream)
        stream.synchronize()
        ms_total = e0.elapsed_time(e1)
        launches = ctx.kernel_launches - launches0
        if ms_total < 300.0:  # keep the GPU under the same load long enough for the clock sampler
            t_end = time.perf_counter() + 0.3
            while time.perf_counter() < t_end:
                prepared.launch(stream.cuda_stream)
                stream.synchronize()
    barrier()
    from imagekit_cuda.sharding import aggregate_throughput
    out_mp_step = batch * dw * dh / 1e6
    # units all ranks processed / max-over-ranks device time
    total_mp, ms_total_max, value = aggregate_throughput(out_mp_step * args.steps, ms_total, dist, dev)
    ms_per_step = ms_total_max / args.steps

    algo_bytes = batch * (sw * sh * ch + dw * dh * ch)
    kernel_ms = (ms_total / args.steps) / max(1, launches_per_step)  # this rank's average launch duration
    peak, peak_src = measured_peak()
    achieved = algo_bytes / max(1, launches_per_step) / (kernel_ms * 1e-3) / 1e9
    # The same launch against the FP32 pipe: algorithmic FMAs = one per tap, channel and sample of each pass
    # (vertical pass over sw columns, then horizontal over dh rows), tap counts from the planner's own tables.
    _, cnt_v, _ = ik.pass_table(filt, sh, dh)
    _, cnt_h, _ = ik.pass_table(filt, sw, dw)
    algo_fma = batch * ch * (sw * int(cnt_v.sum()) + dh * int(cnt_h.sum()))

    # ---- end to end through the C ABI with HOST buffers (H2D + kernel + D2H inside the timed region)
    eb = max(1, min(args.e2e_batch, batch))
    h_src = [ik.PinnedArray((sh, sw, ch)) for _ in range(eb)]
    h_dst = [ik.PinnedArray((dh, dw, ch)) for _ in range(eb)]
    rng = np.random.default_rng(rank)
    for a in h_src:
        a.array[...] = rng.integers(0, 256, a.shape, dtype=np.uint8)
    sizes = [(dw, dh)] * eb
    e2e_steps = max(3, min(args.steps, 20))

    def e2e_leg(srcs, outs, steps):
        """Wall clock around `steps` synchronous ikc_resize_batch calls, all ranks at once; whole-job MP/s."""
        for _ in range(3):
            ctx.resize_batch(srcs, sizes, filt, outs=outs)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            ctx.resize_batch(srcs, sizes, filt, outs=outs)   # synchronous: returns with results in host memory
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        return aggregate_throughput(eb * steps * dw * dh / 1e6, dt * 1e3, dist, dev)[2]

    e2e_value = e2e_leg([a.array for a in h_src], [a.array for a in h_dst], e2e_steps)
    # The same through PAGEABLE host memory -- what the Rust wrapper's Vec<u8> is (crate/src/lib.rs): the library
    # stages it through pinned buffers in chunks, the copy pool sharing each chunk's memcpy.
    p_src = [np.array(a.array) for a in h_src]
    p_dst = [np.empty((dh, dw, ch), np.uint8) for _ in range(eb)]
    e2e_pageable = e2e_leg(p_src, p_dst, max(3, e2e_steps // 2))

    # ---- what the host <-> device link allows: the step's H2D and D2H bytes as plain concurrent cudaMemcpyAsync,
    # every rank at once (the ceiling of any host-resident drop-in on this box at this rank count)
    t_in = torch.empty((eb, sh, sw, ch), dtype=torch.uint8).pin_memory()
    t_out = torch.empty((eb, dh, dw, ch), dtype=torch.uint8).pin_memory()
    d_in = torch.empty_like(t_in, device=dev)
    d_out = torch.empty_like(t_out, device=dev)
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def dma_step():
        with torch.cuda.stream(s_in):
            d_in.copy_(t_in, non_blocking=True)
        with torch.cuda.stream(s_out):
            t_out.copy_(d_out, non_blocking=True)

    for _ in range(3):
        dma_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        dma_step()
    torch.cuda.synchronize(dev)
    dma_s = time.perf_counter() - t0
    dma_ceiling = aggregate_throughput(eb * e2e_steps * dw * dh / 1e6, dma_s * 1e3, dist, dev)[2]
    dma_gbs = eb * e2e_steps * (sw * sh * ch + dw * dh * ch) / dma_s / 1e9  # this rank's H2D + D2H rate
    del t_in, t_out, d_in, d_out

    # ---- BASELINE config 3 as it is stated: ONE batch of 1024 thumbnails (4032x3024 RGB8 -> 400x300) sharded
    # round robin over the ranks (strong scaling; shard_indices is the rule ikc_resize_batch applies in-process)
    cfg3_shard = None
    if not args.no_shard:
        cfg3_shard = run_cfg3_shard(ik, ctx, torch, dev, dist, rank, world, barrier, args)

    # ---- the other device-resident BASELINE workloads, briefly (the headline stays the one named by --workload)
    other_workloads = None
    if not args.no_others:
        other_workloads = run_other_workloads(ik, ctx, torch, dev, dist, barrier, [n for n in ("cfg1", "cfg4") if n != args.workload])

    # ---- BASELINE config 5 as it is stated: 256 x 8 MP JPEG uploads, CPU decode -> GPU resize -> CPU webp encode,
    # split over the ranks; the library's begin / end call lets every worker hide the GPU leg behind the codecs
    cfg5_upload = None
    if not args.no_upload:
        try:
            cfg5_upload = run_cfg5_upload(ik, ctx, dist, dev, rank, world, barrier, args)
        except ImportError as e:   # no Pillow on this box: the codecs are the stand-ins, not the product
            cfg5_upload = {"unavailable": f"{e}"}

    # ---- the product's own multi-GPU entry point: ONE context over all N devices in ONE process (rank 0),
    # ikc_resize_batch shards job i -> device i mod N with one worker thread per device
    e2e_inprocess = None
    if world > 1:
        barrier()
        if rank == 0:
            e2e_inprocess = run_inprocess(ik, world, sw, sh, ch, dw, dh, filt, args)
        # the idle ranks wait on the host: an NCCL barrier would park a spinning kernel on their GPUs, which are the
        # very devices rank 0's context is driving in this leg
        dist.barrier(group=host_group)
        barrier()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- one request at a time: per-call latency through ikc_resize_u8 (config 1 is the latency-bound one)
    call_latency = None
    if not args.no_latency:
        call_latency = run_call_latency(ik, ctx, sorted({"cfg1", args.workload}), args.no_cpu)

    # ---- CPU baseline on the host cores (bounded sample; reported, not the target)
    cpu = None
    cpu_mt = None
    if not args.no_cpu:
        per_img_guess = {"cfg2": 0.45, "cfg3": 0.22, "cfg4": 0.6, "cfg1": 0.05}[args.workload]
        n1 = max(2, int(8.0 / per_img_guess))
        v1, dt1 = cpu_port_rate(sw, sh, ch, dw, dh, filt, n1, 1)
        cpu = {"value": v1, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"{n1} images of this workload, single thread ({dt1:.1f} s): what one reference request gets "
                         f"(resize runs inline on one tokio worker)"}
        cores = os.cpu_count() or 1
        th = max(1, min(cores, 64))
        vm, dtm = cpu_port_rate(sw, sh, ch, dw, dh, filt, 2 * th, th)
        cpu_mt = {"value": vm, "unit": UNIT, "cores": th, "kind": "port",
                  "sample": f"{2 * th} images over {th} threads ({dtm:.1f} s): a saturated reference server"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        # the arithmetic the path computes in: the tensor-core kernels run the vertical pass as u8 x s8 -> s32 integer
        # products (exact) and the horizontal pass in f32; the CUDA-core kernels are f32 throughout
        "dtype": "s32+f32" if "banded8" in prepared.describe() else "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "filter": FILTER_NAMES[filt]},
        "run": {"images_per_step_per_gpu": batch, "layout": "u8 interleaved, tight pitch, device-resident",
                "l2": f"working set {(algo_bytes) / 1e6:.0f} MB per step per GPU >> 126 MB L2 (inputs larger than L2)",
                "timing": "CUDA events on the launch stream, max over ranks"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": recorded_traffic(args.workload, batch), "peak_source": peak_src,
                     "kernel": prepared.describe(),
                     "algorithmic_bytes_per_launch": algo_bytes // max(1, launches_per_step),
                     "kernel_ms": kernel_ms},
        "roofline_fp32": {"note": "secondary: all of the resize's FMAs against the CUDA cores' FP32 pipe (round 1's ceiling). "
                                  "Since round 2 the vertical pass of downscales runs on the tensor cores instead, so this "
                                  "fraction overstates what the FP32 pipe does; peak = 128 FMA/clk/SM x SMs x max SM clock",
                          "achieved": algo_fma / max(1, launches_per_step) / (kernel_ms * 1e-3) / 1e12,
                          "peak": 128 * torch.cuda.get_device_properties(dev).multi_processor_count *
                                  (clocks.max_mhz or 1965) * 1e6 / 1e12,
                          "unit": "TFMA/s", "algorithmic_fma_per_launch": algo_fma // max(1, launches_per_step)},
        "cpu_baseline": cpu,
        "cpu_baseline_all_cores": cpu_mt,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": eb * sw * sh * ch,
                "d2h_bytes_per_step": eb * dw * dh * ch, "images_per_step_per_gpu": eb, "steps": e2e_steps,
                "api": "ikc_resize_batch (C ABI), pinned host buffers, wall clock around the synchronous call",
                "dma_ceiling": {"value": dma_ceiling, "unit": UNIT, "this_rank_gbs": dma_gbs,
                                "how": "the same H2D + D2H bytes per step as plain concurrent cudaMemcpyAsync from / to "
                                       "pinned memory, all ranks at once, no kernel"},
                "frac_of_dma_ceiling": e2e_value / dma_ceiling},
        "e2e_pageable": {"value": e2e_pageable, "unit": UNIT,
                         "api": "ikc_resize_batch with pageable numpy buffers (what the Rust wrapper's Vec<u8> is): "
                                "chunked staging through pinned memory, copy pool",
                         "effective_h2d_gbs_per_gpu": e2e_pageable / world * 1e6 / (dw * dh) * sw * sh * ch / 1e9},
        "e2e_inprocess": e2e_inprocess,
        "cfg3_shard": cfg3_shard,
        "cfg5_upload": cfg5_upload,
        "other_workloads": other_workloads,
        "call_latency": call_latency,
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
        "parity": parity,
    }
    line["roofline_fp32"]["frac"] = line["roofline_fp32"]["achieved"] / line["roofline_fp32"]["peak"]
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
