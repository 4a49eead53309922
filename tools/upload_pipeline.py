"""BASELINE config 5: the /upload-shaped path (src/lib.rs:281-294) end to end:
    decode_image (CPU) -> resize_image (GPU, all visible B200s) -> encode_image webp q=80 (CPU)
over a batch of synthetic 8 MP JPEGs, with a per-stage breakdown.  Decode and encode stay on the CPU
(north-star); Pillow stands in for the reference's image/webp crates.

usage: python tools/upload_pipeline.py [--images 64] [--width 800] [--threads N] [--cpu-resize] [--pipelined]

Default: the three stages run one after the other over the whole batch (per-stage breakdown).
--pipelined: every worker thread takes one upload through decode -> resize -> encode, as the reference's
tokio workers do (src/lib.rs:281-294), so decode(i+1), H2D+resize(i) and encode(i-1) of different uploads
overlap; the resize stores rgb8 directly (encode_as=webp: to_rgb8() fused).
"""
import argparse
import io
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rust-image-transform_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from PIL import Image


def synth_jpegs(n, w=3264, h=2448):
    from conftest import photo_like
    base = photo_like((h, w, 3), seed=3)
    out = []
    for i in range(n):
        img = np.roll(base, 37 * i, axis=1)
        buf = io.BytesIO()
        Image.fromarray(img).save(buf, "JPEG", quality=90)
        out.append(buf.getvalue())
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=64)
    ap.add_argument("--width", type=int, default=800)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 8)
    ap.add_argument("--cpu-resize", action="store_true", help="resize with the CPU oracle port instead (baseline)")
    ap.add_argument("--pipelined", action="store_true", help="one upload per worker thread, stages overlapped")
    args = ap.parse_args()
    import imagekit_cuda as ik
    jpegs = synth_jpegs(min(args.images, 8))
    jpegs = [jpegs[i % len(jpegs)] for i in range(args.images)]
    pool = ThreadPoolExecutor(args.threads)
    if not args.cpu_resize:
        # a server creates the context once and reuses its staging buffers: do that outside the timing
        ctx = ik.default_context()
        warm = ik.decode_image(jpegs[0])[0]
        tw, th, _ = ik.target_dims(warm.width(), warm.height(), args.width, None)
        ctx.resize_batch([warm.pixels] * 8, [(tw, th)] * 8)

    if args.pipelined:
        if args.cpu_resize:
            from oracle import oracle

            def one(b):
                d = ik.decode_image(b)[0]
                r = ik.DynamicImage(oracle.resize_image(d.pixels, args.width, None))
                return len(ik.encode_image(r, ik.ImageFormat.webp, 80))
        else:
            def one(b):
                d = ik.decode_image(b)[0]
                r = ik.resize_image(d, args.width, None, ctx=ctx, encode_as=ik.ImageFormat.webp)
                return len(ik.encode_image(r, ik.ImageFormat.webp, 80))
        t0 = time.perf_counter()
        sizes = list(pool.map(one, jpegs))
        t3 = time.perf_counter()
        print(json.dumps({
            "workload": f"cfg5: {args.images} x 8 MP JPEG -> w={args.width} Lanczos3 -> webp q=80, one upload per worker thread",
            "images_per_s": args.images / (t3 - t0), "threads": args.threads,
            "gpus": 0 if args.cpu_resize else ctx.device_count,
            "resize": "cpu oracle port" if args.cpu_resize else "gpu (ikc_resize_convert_u8, pageable host buffers)",
            "seconds": t3 - t0, "out_bytes_mean": sum(sizes) / args.images,
        }))
        return

    t0 = time.perf_counter()
    decoded = list(pool.map(lambda b: ik.decode_image(b)[0], jpegs))           # CPU decode, threaded
    t1 = time.perf_counter()
    if args.cpu_resize:
        from oracle import oracle
        resized = list(pool.map(lambda d: ik.DynamicImage(oracle.resize_image(d.pixels, args.width, None)), decoded))
        devices = 0
    else:
        ctx = ik.default_context()
        devices = ctx.device_count
        sizes = []
        for d in decoded:
            tw, th, _ = ik.target_dims(d.width(), d.height(), args.width, None)
            sizes.append((tw, th))
        outs, jobs = ctx.resize_batch([d.pixels for d in decoded], sizes)          # GPU resize, sharded over devices
        resized = [ik.DynamicImage(o) for o in outs]
    t2 = time.perf_counter()
    encoded = list(pool.map(lambda r: ik.encode_image(r, ik.ImageFormat.webp, 80), resized))  # CPU encode, threaded
    t3 = time.perf_counter()
    n = args.images
    print(json.dumps({
        "workload": f"cfg5: {n} x 8 MP JPEG -> w={args.width} Lanczos3 -> webp q=80",
        "images_per_s": n / (t3 - t0), "threads": args.threads, "gpus": devices,
        "resize": "cpu oracle port" if args.cpu_resize else "gpu (ikc_resize_batch, pageable host buffers)",
        "stage_seconds": {"decode": t1 - t0, "resize": t2 - t1, "encode": t3 - t2},
        "stage_share": {"decode": (t1 - t0) / (t3 - t0), "resize": (t2 - t1) / (t3 - t0), "encode": (t3 - t2) / (t3 - t0)},
        "out_bytes_mean": sum(map(len, encoded)) / n,
    }))


if __name__ == "__main__":
    main()
