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
        "config": {"workload": f"{args.workload}: {desc}", "images_per_step": per_step, "filter": FILTER_NAMES[filt],
                   "note": "CPU port of image 0.25.8 imageops::resize (oracle/imageops_oracle.c); the Rust "
                           "reference cannot be built here (no cargo/rustc, crate not vendored)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{per_step} images/step x {args.steps} steps, one image per thread"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


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

    # ---- end to end through the C ABI with pinned HOST buffers (H2D + kernel + D2H inside the timed region)
    eb = max(1, min(args.e2e_batch, batch))
    h_src = [ik.PinnedArray((sh, sw, ch)) for _ in range(eb)]
    h_dst = [ik.PinnedArray((dh, dw, ch)) for _ in range(eb)]
    rng = np.random.default_rng(rank)
    for a in h_src:
        a.array[...] = rng.integers(0, 256, a.shape, dtype=np.uint8)
    srcs = [a.array for a in h_src]
    outs = [a.array for a in h_dst]
    sizes = [(dw, dh)] * eb
    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(3):
        ctx.resize_batch(srcs, sizes, filt, outs=outs)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.resize_batch(srcs, sizes, filt, outs=outs)   # synchronous: returns with results in host memory
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    _, _, e2e_value = aggregate_throughput(eb * e2e_steps * dw * dh / 1e6, e2e_s * 1e3, dist, dev)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

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
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "images_per_step_per_gpu": batch,
                   "filter": FILTER_NAMES[filt], "layout": "u8 interleaved, tight pitch, device-resident",
                   "l2": f"working set {(algo_bytes) / 1e6:.0f} MB per step per GPU >> 126 MB L2 (inputs larger than L2)",
                   "timing": "CUDA events on the launch stream, max over ranks"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": recorded_traffic(args.workload, batch), "peak_source": peak_src,
                     "kernel": prepared.describe(),
                     "algorithmic_bytes_per_launch": algo_bytes // max(1, launches_per_step),
                     "kernel_ms": kernel_ms},
        "roofline_fp32": {"note": "secondary: Lanczos3 downscales sit above the FP32 ridge, so the FMA pipe, not HBM, "
                                  "is their ceiling (DESIGN.md 4.1); peak = 128 FMA/clk/SM x SMs x max SM clock",
                          "achieved": algo_fma / max(1, launches_per_step) / (kernel_ms * 1e-3) / 1e12,
                          "peak": 128 * torch.cuda.get_device_properties(dev).multi_processor_count *
                                  (clocks.max_mhz or 1965) * 1e6 / 1e12,
                          "unit": "TFMA/s", "algorithmic_fma_per_launch": algo_fma // max(1, launches_per_step)},
        "cpu_baseline": cpu,
        "cpu_baseline_all_cores": cpu_mt,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": eb * sw * sh * ch,
                "d2h_bytes_per_step": eb * dw * dh * ch, "images_per_step_per_gpu": eb, "steps": e2e_steps,
                "api": "ikc_resize_batch (C ABI), pinned host buffers, wall clock around the synchronous call"},
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
