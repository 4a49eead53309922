"""Device-resident entry points (what the roofline bench times): prepared batches replayed on a
torch stream, checked against the oracle; plus full-size size-independent properties."""
import numpy as np
import pytest

from conftest import delta_histogram, splitmix_noise

pytestmark = pytest.mark.gpu


def test_prepared_batch_on_torch_stream(ctx, ik, oracle):
    import torch
    ctx.set_mode(ik.MODE_FAST)
    dev = torch.device("cuda:0")
    shapes = [(480, 640, 3, 200, 150), (600, 800, 4, 400, 300), (1080, 1920, 3, 400, 225), (96, 128, 4, 61, 47)]
    srcs, dsts, jobs = [], [], []
    for i, (h, w, c, dw, dh) in enumerate(shapes):
        s = splitmix_noise((h, w, c), image_id=i)
        ts = torch.from_numpy(s).to(dev)
        td = torch.zeros((dh, dw, c), dtype=torch.uint8, device=dev)
        srcs.append((s, ts)); dsts.append(td)
        jobs.append((ts.data_ptr(), w, h, w * c, td.data_ptr(), dw, dh, dw * c, c, ik.FILTER_LANCZOS3))
    batch = ctx.prepare_batch(0, jobs)
    assert all(j.status == 0 for j in batch.jobs)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        batch.launch(stream.cuda_stream)
        batch.launch(stream.cuda_stream)                 # replayable
    stream.synchronize()
    for (s, _), td, (h, w, c, dw, dh) in zip(srcs, dsts, shapes):
        got = td.cpu().numpy()
        hist = delta_histogram(got, oracle.resize_exact(s, dw, dh, oracle.LANCZOS3))
        assert max(abs(k) for k in hist) <= 1, hist
    batch.free()


@pytest.mark.parametrize("mode", ["tc", "f16", "fp32"])
def test_prepared_batch_mixes_every_kernel(ctx, ik, oracle, mode):
    """One prepared batch whose jobs need the downscale kernel (banded tensor-core kernel, or the CUDA-core ring
    kernel in FAST_FP32 mode; plain, uniform, converting), the 2x upscale kernel, the tile kernel and the generic
    kernels: one launch per kernel variant, every job within +-1."""
    import torch
    ctx.set_mode({"tc": ik.MODE_FAST, "f16": ik.MODE_FAST_F16, "fp32": ik.MODE_FAST_FP32}[mode])
    dev = torch.device("cuda:0")
    cases = [  # (h, w, c, dw, dh, filter, out_channels)
        (480, 640, 3, 200, 150, 4, 3), (600, 800, 4, 400, 300, 4, 4), (600, 800, 4, 400, 300, 4, 3),
        (240, 320, 3, 640, 480, 2, 3), (240, 320, 4, 640, 480, 2, 4), (200, 300, 1, 150, 100, 4, 1),
        (100, 120, 3, 333, 222, 1, 4), (300, 400, 3, 200, 150, 4, 4),
    ]
    keep, jobs, want = [], [], []
    for i, (h, w, c, dw, dh, filt, co) in enumerate(cases):
        s = splitmix_noise((h, w, c), image_id=40 + i)
        ts = torch.from_numpy(s).to(dev)
        td = torch.zeros((dh, dw, co), dtype=torch.uint8, device=dev)
        keep.append((ts, td))
        jobs.append((ts.data_ptr(), w, h, w * c, td.data_ptr(), dw, dh, dw * co, c | (co << 8) if co != c else c, filt))
        r = oracle.resize_exact(s, dw, dh, filt)
        want.append(r if co == c else (oracle.to_rgb8(r) if co == 3 else oracle.to_rgba8(r)))
    batch = ctx.prepare_batch(0, jobs)
    assert all(j.status == 0 for j in batch.jobs), [j.status for j in batch.jobs]
    desc = batch.describe()
    ctx.set_mode(ik.MODE_FAST)
    # (default mode: the 2x upscales take the tensor-core kernel too; the other modes keep the CUDA-core upscale kernel)
    for name in ({"tc": "banded8_kernel", "f16": "banded_kernel", "fp32": "fused_ring_kernel"}[mode],
                 "banded8u_kernel" if mode == "tc" else "up2_kernel", "tile_kernel"):
        assert name in desc, desc
    assert ("banded8_kernel" in desc) == (mode == "tc") and ("fused_ring_kernel" in desc) == (mode == "fp32")
    stream = torch.cuda.Stream()
    batch.launch(stream.cuda_stream)
    stream.synchronize()
    for (_, td), w_, case in zip(keep, want, cases):
        hist = delta_histogram(td.cpu().numpy().reshape(w_.shape), w_)
        assert max(abs(k) for k in hist) <= 1, (case, hist)
    batch.free()


def test_device_entry_point(ctx, ik, oracle):
    import torch
    ctx.set_mode(ik.MODE_FAST)
    s = splitmix_noise((600, 800, 4))
    ts = torch.from_numpy(s).cuda()
    td = torch.zeros((300, 400, 4), dtype=torch.uint8, device="cuda")
    ctx.resize_device(0, torch.cuda.current_stream().cuda_stream, ts.data_ptr(), 800, 600, 3200, 4, td.data_ptr(),
                      400, 300, 1600)
    torch.cuda.synchronize()
    hist = delta_histogram(td.cpu().numpy(), oracle.resize_exact(s, 400, 300, oracle.LANCZOS3))
    assert max(abs(k) for k in hist) <= 1, hist


def test_full_size_properties_4k(ctx, ik):
    """BASELINE config 2 at full size through size-independent properties: a constant image stays
    constant, and resizing is linear up to rounding: resize(a) + resize(255 - a) ~= 255."""
    ctx.set_mode(ik.MODE_FAST)
    c = np.full((2160, 3840, 4), 201, np.uint8)
    assert (ctx.resize(c, 1920, 1080) == 201).all()
    a = splitmix_noise((2160, 3840, 4))
    ra = ctx.resize(a, 1920, 1080).astype(np.int32)
    rb = ctx.resize(255 - a, 1920, 1080).astype(np.int32)
    s = ra + rb
    inner = (ra > 0) & (ra < 255) & (rb > 0) & (rb < 255)   # unclamped samples
    assert np.abs(s[inner] - 255).max() <= 1


def test_host_batch_shards_round_robin_over_all_devices(ik, oracle):
    """ikc_resize_batch over every visible GPU: job i runs on device i mod G, no collective."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs")
    ctx = ik.Context()                     # all visible devices
    try:
        g = ctx.device_count
        assert g == torch.cuda.device_count()
        srcs = [splitmix_noise((300 + 8 * i, 400, 3 + (i & 1)), image_id=i) for i in range(4 * g + 1)]
        sizes = [(200, 150 + 4 * i) for i in range(len(srcs))]
        outs, jobs = ctx.resize_batch(srcs, sizes)
        for i, (s, (dw, dh), o_, j) in enumerate(zip(srcs, sizes, outs, jobs)):
            assert j.status == 0 and j.device == i % g
            hist = delta_histogram(o_, oracle.resize_exact(s, dw, dh, oracle.LANCZOS3))
            assert max(abs(k) for k in hist) <= 1, hist
    finally:
        ctx.close()


def test_randomised_prepared_batches(ctx, ik, oracle):
    """Random jobs sharing launches: prepared batches of 8 mixed shapes / ratios / channels / conversions, so that
    jobs with different strip and tile geometries meet in one launch of each kernel variant."""
    import torch
    from conftest import checker
    ctx.set_mode(ik.MODE_FAST)
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(77)
    failures = []
    for _ in range(25):
        keep, jobs, want, descs = [], [], [], []
        filt = int(rng.choice([1, 2, 3, 4], p=[0.1, 0.2, 0.1, 0.6]))
        for _ in range(8):
            kind = rng.choice(["down2", "down_any", "up2", "mixed"], p=[0.3, 0.4, 0.15, 0.15])
            c = int(rng.choice([1, 2, 3, 4], p=[0.1, 0.1, 0.4, 0.4]))
            if kind == "down2":
                r = int(rng.choice([2, 2, 4]))
                dw, dh = int(rng.integers(8, 500)), int(rng.integers(8, 300))
                w, h = dw * r, dh * r
            elif kind == "down_any":
                w, h = int(rng.integers(16, 1800)), int(rng.integers(16, 1200))
                dw, dh = int(rng.integers(1, w)), int(rng.integers(1, h))
            elif kind == "up2":
                w, h = int(rng.integers(1, 300)), int(rng.integers(1, 200))
                dw, dh = 2 * w, 2 * h
            else:
                w, h = int(rng.integers(8, 800)), int(rng.integers(8, 800))
                dw, dh = int(rng.integers(1, 2 * w)), int(rng.integers(1, 2 * h))
            if (w, h) == (dw, dh):
                dw += 1
            co = int(rng.choice([3, 4])) if rng.random() < 0.25 else c
            s = (checker if rng.random() < 0.2 else splitmix_noise)((h, w, c))
            ts = torch.from_numpy(s).to(dev)
            td = torch.zeros((dh, dw, co), dtype=torch.uint8, device=dev)
            keep.append((ts, td))
            jobs.append((ts.data_ptr(), w, h, w * c, td.data_ptr(), dw, dh, dw * co, c | (co << 8) if co != c else c, filt))
            r_ = oracle.resize_exact(s, dw, dh, filt)
            r_ = r_[:, :, None] if r_.ndim == 2 else r_
            want.append(r_ if co == c else (oracle.to_rgb8(r_) if co == 3 else oracle.to_rgba8(r_)))
            descs.append(dict(h=h, w=w, c=c, dw=dw, dh=dh, filt=filt, co=co))
        batch = ctx.prepare_batch(0, jobs)
        assert all(j.status == 0 for j in batch.jobs), [j.status for j in batch.jobs]
        stream = torch.cuda.Stream()
        batch.launch(stream.cuda_stream)
        stream.synchronize()
        for (_, td), w_, dsc in zip(keep, want, descs):
            d = int(np.abs(td.cpu().numpy().astype(np.int32) - w_.astype(np.int32)).max())
            if d > 1:
                failures.append((dsc, d, batch.describe()))
        batch.free()
    assert not failures, failures[:4]


ROW_BAND_SHAPES = [  # Rgba8, exactly 2:1 horizontally, at most ~2:1 vertically, 32-byte aligned destination rows: (h, w, dw, dh)
    (384, 512, 256, 192), (130, 272, 136, 65), (700, 1008, 504, 350), (300, 640, 320, 150), (514, 2064, 1032, 257),
    (700, 1008, 504, 400), (2160, 3840, 1920, 1080), (1090, 768, 384, 545), (64, 96, 48, 32),
]


@pytest.mark.parametrize("shape", [(1236, 730, 365, 549), (300, 202, 101, 150), (200, 1034, 517, 100), (90, 94, 47, 45)])
def test_row_band_kernel_widths_not_a_multiple_of_four(ctx, ik, oracle, shape):
    """Host buffers are staged with a 256-byte pitch, so any width reaches banded8t: the last, partial group of four outputs
    of a row must still be written (a randomised sweep found the last column missing for dw = 365)."""
    h, w, dw, dh = shape
    ctx.set_mode(ik.MODE_FAST)
    s = splitmix_noise((h, w, 4), image_id=w)
    before = ctx.stats()["launches_banded8t"]
    got = ctx.resize(s, dw, dh, ik.FILTER_LANCZOS3)
    assert ctx.stats()["launches_banded8t"] == before + 1
    hist = delta_histogram(got, oracle.resize_exact(s, dw, dh, oracle.LANCZOS3))
    assert max(abs(k) for k in hist) <= 1, hist


@pytest.mark.parametrize("shape", ROW_BAND_SHAPES)
@pytest.mark.parametrize("content", ["noise", "edges"])
def test_row_band_kernel(ctx, ik, oracle, shape, content):
    """banded8t.cu: accumulator lanes are output rows, the horizontal pass runs from registers.  Single images are cut
    into column segments (one CTA each), so segment seams, the image borders (looked-up weights) and the last, partial
    band are all on the path."""
    import torch
    from conftest import checker
    h, w, dw, dh = shape
    if content != "noise" and h * w > 2_000_000:
        pytest.skip("large shapes run on noise only")
    ctx.set_mode(ik.MODE_FAST)
    dev = torch.device("cuda:0")
    s = splitmix_noise((h, w, 4), image_id=h) if content == "noise" else checker((h, w, 4))
    ts = torch.from_numpy(s).to(dev)
    td = torch.zeros((dh, dw, 4), dtype=torch.uint8, device=dev)
    batch = ctx.prepare_batch(0, [(ts.data_ptr(), w, h, w * 4, td.data_ptr(), dw, dh, dw * 4, 4, ik.FILTER_LANCZOS3)])
    assert batch.jobs[0].status == 0
    assert "banded8t_kernel" in batch.describe(), batch.describe()
    stream = torch.cuda.Stream()
    batch.launch(stream.cuda_stream)
    stream.synchronize()
    hist = delta_histogram(td.cpu().numpy(), oracle.resize_exact(s, dw, dh, oracle.LANCZOS3))
    assert max(abs(k) for k in hist) <= 1, hist
    off = sum(v for k, v in hist.items() if k != 0) / td.numel()
    assert off < (0.02 if content == "noise" else 0.5), hist
    batch.free()


def test_row_band_kernel_batches_and_fallbacks(ctx, ik, oracle):
    """A batch of 2:1 Rgba8 jobs of different sizes shares one banded8t launch (whole-width items); a destination that is
    not 16-byte aligned, another channel count or a fused conversion take the other downscale kernel in the same batch."""
    import torch
    ctx.set_mode(ik.MODE_FAST)
    dev = torch.device("cuda:0")
    cases = [  # (h, w, c, dw, dh, co, dst pitch slack)
        (384, 512, 4, 256, 192, 4, 0), (1080, 1920, 4, 960, 540, 4, 0), (300, 640, 4, 320, 150, 4, 0), (700, 1008, 4, 504, 350, 4, 0),
        (1236, 736, 4, 368, 549, 4, 0),   # 2.25:1 vertically: ten weight tiles per band, the others nine, in one launch
        (300, 640, 4, 320, 150, 4, 4), (300, 640, 3, 320, 150, 3, 0), (300, 640, 4, 320, 150, 3, 0),
    ]
    keep, jobs, want = [], [], []
    for i, (h, w, c, dw, dh, co, slack) in enumerate(cases):
        s = splitmix_noise((h, w, c), image_id=70 + i)
        ts = torch.from_numpy(s).to(dev)
        pitch = dw * co + slack
        td = torch.zeros((dh, pitch), dtype=torch.uint8, device=dev)
        keep.append((ts, td))
        jobs.append((ts.data_ptr(), w, h, w * c, td.data_ptr(), dw, dh, pitch, c | (co << 8) if co != c else c, ik.FILTER_LANCZOS3))
        r = oracle.resize_exact(s, dw, dh, oracle.LANCZOS3)
        want.append(r if co == c else oracle.to_rgb8(r))
    batch = ctx.prepare_batch(0, jobs)
    assert all(j.status == 0 for j in batch.jobs), [j.status for j in batch.jobs]
    desc = batch.describe()
    assert "banded8t_kernel" in desc and "banded8_kernel" in desc, desc
    stream = torch.cuda.Stream()
    batch.launch(stream.cuda_stream)
    stream.synchronize()
    for (h, w, c, dw, dh, co, slack), (_, td), exp in zip(cases, keep, want):
        got = td.cpu().numpy()[:, :dw * co].reshape(dh, dw, co)
        hist = delta_histogram(got, exp)
        assert max(abs(k) for k in hist) <= 1, ((h, w, c, co, slack), hist)
    batch.free()
