"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle on the same inputs.
Bars: EXACT mode (two-launch generic kernels) is bit-equal; FAST mode (fused kernels: banded tensor-core
kernel for downscales, FMA) and FAST_FP32 mode (the same without tensor cores: CUDA-core ring kernel) are
within max |delta| <= 1 per u8 sample (north-star tolerance), with the delta histogram checked."""
import numpy as np
import pytest

from conftest import checker, delta_histogram, photo_like, splitmix_noise

pytestmark = pytest.mark.gpu
TOL = 1  # max |delta| per u8 channel, BASELINE.json north_star


def _fast(ctx, ik, mode="tc"):
    ctx.set_mode({"tc": ik.MODE_FAST, "f16": ik.MODE_FAST_F16, "fp32": ik.MODE_FAST_FP32}[mode])


# downscales: integer tensor-core kernel (banded8) / f16 tensor-core kernel (banded) / CUDA-core ring kernel
FAST_MODES = ["tc", "f16", "fp32"]


def _check_fast(got, want, label, max_off=0.016):
    hist = delta_histogram(got, want)
    assert max(abs(k) for k in hist) <= TOL, (label, hist)
    off = sum(v for k, v in hist.items() if k != 0) / got.size
    # FMA vs mul+add only flips values sitting on a .5 boundary: rare on noise, a percent or so on
    # synthetic 0/255 checker patterns whose exact sums land on many ties
    assert off < max_off, (label, hist)
    return hist


SMALL_SHAPES = [  # (h, w, c, dw, dh)
    (60, 80, 3, 40, 30), (64, 96, 4, 48, 32), (121, 161, 3, 16, 12), (24, 32, 3, 64, 48), (70, 70, 3, 33, 33),
    (33, 47, 2, 19, 13), (45, 45, 1, 20, 17), (30, 40, 3, 1, 1), (2, 2, 3, 200, 200), (333, 1000, 3, 99, 33),
    (17, 1, 3, 1, 5), (1, 19, 4, 7, 1), (600, 800, 3, 400, 300),
]


@pytest.mark.parametrize("shape", SMALL_SHAPES)
@pytest.mark.parametrize("filt", [0, 1, 2, 3, 4])
def test_exact_mode_bit_equal(ctx, ik, oracle, shape, filt):
    h, w, c, dw, dh = shape
    src = splitmix_noise((h, w, c), image_id=filt)
    ctx.set_mode(ik.MODE_EXACT)
    got = ctx.resize(src, dw, dh, filt)
    want = oracle.resize_exact(src, dw, dh, filt)
    assert np.array_equal(got, want), delta_histogram(got, want)


@pytest.mark.parametrize("shape", SMALL_SHAPES)
@pytest.mark.parametrize("filt", [0, 1, 2, 3, 4])
def test_fast_mode_within_one_lsb(ctx, ik, oracle, shape, filt):
    h, w, c, dw, dh = shape
    src = splitmix_noise((h, w, c), image_id=10 + filt)
    _fast(ctx, ik)
    got = ctx.resize(src, dw, dh, filt)
    want = oracle.resize_exact(src, dw, dh, filt)
    _check_fast(got, want, (shape, filt))


FUSED_SHAPES = [  # Lanczos3 downscales that take the fused ring kernel: (h, w, c, dw, dh)
    (480, 640, 3, 200, 150), (480, 640, 4, 320, 240), (1080, 1920, 3, 400, 225), (1000, 1504, 4, 752, 500),
    (777, 1031, 3, 515, 388), (901, 1200, 4, 411, 309), (2160, 3840, 4, 1920, 1080), (3024, 4032, 3, 400, 300),
    (512, 512, 4, 511, 511), (300, 5000, 3, 2500, 150), (4000, 304, 4, 152, 2000), (768, 1024, 3, 1000, 750),
    # integer ratios (the kernel's uniform-stretch loops): 4x both channel counts, 2x with ragged strips, 2x by 4x
    (1024, 2048, 4, 512, 256), (960, 1280, 3, 320, 240), (1500, 2002, 3, 1001, 750), (1216, 1216, 4, 608, 304),
    (600, 800, 4, 400, 300), (602, 802, 4, 401, 301), (300, 400, 3, 200, 150), (128, 96, 4, 48, 64),
]


@pytest.mark.parametrize("shape", FUSED_SHAPES)
@pytest.mark.parametrize("content", ["noise", "edges", "photo"])
@pytest.mark.parametrize("mode", FAST_MODES)
def test_fused_kernel_parity(ctx, ik, oracle, shape, content, mode):
    h, w, c, dw, dh = shape
    if content != "noise" and h * w > 2_000_000:
        pytest.skip("large shapes run on noise only")
    src = {"noise": splitmix_noise, "edges": checker, "photo": photo_like}[content]((h, w, c))
    _fast(ctx, ik, mode)
    before = ctx.kernel_launches
    got = ctx.resize(src, dw, dh, ik.FILTER_LANCZOS3)
    _fast(ctx, ik)
    assert ctx.kernel_launches - before == 1, "expected the single-launch fused kernel"
    want = oracle.resize_exact(src, dw, dh, oracle.LANCZOS3)
    # checkerboards on integer ratios put many sums exactly on x.5, where FMA contraction decides the tie
    # (the integer kernel's 15-bit weights move more samples across a .5 boundary than 22+-bit weights do)
    scale = 8 if mode == "tc" else 1
    _check_fast(got, want, (shape, content), max_off=scale * {"noise": 0.002, "photo": 0.03, "edges": 0.06}[content])


def test_fused_gaussian_downscale(ctx, ik, oracle):
    src = splitmix_noise((700, 900, 4))
    _fast(ctx, ik)
    got = ctx.resize(src, 450, 350, ik.FILTER_GAUSSIAN)
    _check_fast(got, oracle.resize_exact(src, 450, 350, oracle.GAUSSIAN), "gaussian")


def test_constants_and_zero(ctx, ik, oracle):
    _fast(ctx, ik)
    for v in (0, 37, 255):
        src = np.full((480, 640, 3), v, np.uint8)
        assert (ctx.resize(src, 200, 150) == v).all()
        src4 = np.full((300, 400, 4), v, np.uint8)
        assert (ctx.resize(src4, 800, 600, ik.FILTER_CATMULLROM) == v).all()


def test_reference_semantics_empty_and_same_size(ctx, ik):
    _fast(ctx, ik)
    out = ctx.resize(np.zeros((0, 0, 3), np.uint8), 5, 4)
    assert out.shape == (4, 5, 3) and not out.any()
    n = splitmix_noise((40, 50, 3))
    assert np.array_equal(ctx.resize(n, 50, 40), n)


def test_channel_independence_on_gpu(ctx, ik):
    _fast(ctx, ik)
    rgba = splitmix_noise((480, 640, 4))
    full = ctx.resize(rgba, 320, 240)
    rgb = ctx.resize(np.ascontiguousarray(rgba[:, :, :3]), 320, 240)
    assert np.abs(full[:, :, :3].astype(int) - rgb.astype(int)).max() <= 1


def test_pitched_host_buffers(ctx, ik, oracle):
    _fast(ctx, ik)
    big = splitmix_noise((500, 700, 3))
    view = big[10:490, 20:660, :]                      # non-contiguous rows (pitch > row bytes)
    out_big = np.zeros((200, 300, 3), np.uint8)
    out_view = out_big[5:155, 7:207, :]
    ctx.resize(view, 200, 150, ik.FILTER_LANCZOS3, out=out_view)
    want = oracle.resize_exact(np.ascontiguousarray(view), 200, 150, oracle.LANCZOS3)
    _check_fast(out_view, want, "pitched")
    assert not out_big[:5].any() and not out_big[:, :7].any() and not out_big[155:].any()


def test_u16_samples(ctx, ik, oracle):
    rng = np.random.default_rng(5)
    src = rng.integers(0, 65536, (90, 120, 3), dtype=np.uint16)
    ctx.set_mode(ik.MODE_EXACT)
    assert np.array_equal(ctx.resize(src, 50, 41), oracle.resize_exact(src, 50, 41, oracle.LANCZOS3))
    _fast(ctx, ik)
    got = ctx.resize(src, 50, 41)
    assert np.abs(got.astype(int) - oracle.resize_exact(src, 50, 41, oracle.LANCZOS3).astype(int)).max() <= 1


def test_resize_image_matches_reference_tests(ctx, ik, oracle):
    """The reference's own resize tests (tests/transform.rs:11-96, 239-257), run on the GPU path."""
    _fast(ctx, ik)
    cases = [((800, 600), (400, None), (400, 300)), ((800, 600), (None, 300), (400, 300)),
             ((800, 600), (400, 300), (400, 300)), ((1920, 1080), (960, None), (960, 540)),
             ((800, 600), (None, None), (800, 600)), ((100, 100), (200, 200), (200, 200)),
             ((800, 600), (1, 1), (1, 1)), ((2, 2), (200, 200), (200, 200)), ((1920, 1080), (640, 480), (640, 360))]
    for (ow, oh), (w, h), expect in cases:
        img = ik.DynamicImage.new_rgb8(ow, oh)
        out = ik.resize_image(img, w, h, ctx=ctx)
        assert out.dimensions() == expect
        assert not out.pixels.any()                       # zero image stays zero
    noise = ik.DynamicImage(splitmix_noise((333, 1000, 3)))
    out = ik.resize_image(noise, 100, None, ctx=ctx)
    assert out.dimensions() == (99, 33)
    _check_fast(out.pixels, oracle.resize_image(noise.pixels, 100, None), "99x33")
    got = ctx.resize_image(noise.pixels, 100, None)      # same through ikc_resize_image_u8
    assert np.array_equal(got, out.pixels)


def test_host_batch(ctx, ik, oracle):
    _fast(ctx, ik)
    srcs = [splitmix_noise((480 + 16 * i, 640, 3), image_id=i) for i in range(9)]
    sizes = [(200, 150 + 5 * i) for i in range(9)]
    outs, jobs = ctx.resize_batch(srcs, sizes)
    for s, (dw, dh), o_, j in zip(srcs, sizes, outs, jobs):
        assert j.status == 0
        _check_fast(o_, oracle.resize_exact(s, dw, dh, oracle.LANCZOS3), "batch")


def test_error_codes(ctx, ik):
    L = ik._lib.load()
    src = np.zeros((8, 8, 3), np.uint8)
    dst = np.zeros((4, 4, 3), np.uint8)
    h = ctx.handle
    assert L.ikc_resize_u8(h, src.ctypes.data, 8, 8, 24, 3, dst.ctypes.data, 4, 4, 12, 9) == ik._lib.ERR_INVALID_ARG
    assert L.ikc_resize_u8(h, src.ctypes.data, 8, 8, 24, 5, dst.ctypes.data, 4, 4, 20, 4) == ik._lib.ERR_UNSUPPORTED
    assert L.ikc_resize_u8(h, src.ctypes.data, 8, 8, 24, 3, dst.ctypes.data, 70000, 4, 210000, 4) == ik._lib.ERR_TOO_LARGE
    assert L.ikc_resize_u8(h, src.ctypes.data, 8, 8, 10, 3, dst.ctypes.data, 4, 4, 12, 4) == ik._lib.ERR_INVALID_ARG
    assert L.ikc_resize_u8(h, None, 8, 8, 24, 3, dst.ctypes.data, 4, 4, 12, 4) == ik._lib.ERR_INVALID_ARG
    with pytest.raises(ik.ImageKitError):
        ik.resize_image(ik.DynamicImage.new_rgb8(10, 10), 4_000_000_000, None, ctx=ctx)  # reference would allocate


def test_threads_share_one_context(ctx, ik, oracle):
    """The reference calls resize_image concurrently from its tokio workers (src/lib.rs:180,286)."""
    import threading
    _fast(ctx, ik)
    src = splitmix_noise((480, 640, 3))
    want = oracle.resize_exact(src, 200, 150, oracle.LANCZOS3)
    errs = []

    def work():
        try:
            for _ in range(5):
                _check_fast(ctx.resize(src, 200, 150), want, "thread")
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=work) for _ in range(8)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs


# ---- to_rgb8() / to_rgba8() fused into the store (src/transform.rs:123,131,140) ----------------
CONVERT_CASES = [  # (h, w, c, dw, dh, filter, out_channels): ring kernel, tile kernel, both conversions, all variants
    (2160, 3840, 4, 1920, 1080, 4, 3), (1080, 1920, 3, 400, 225, 4, 4), (600, 800, 4, 400, 300, 4, 3),
    (600, 800, 3, 400, 300, 4, 4), (777, 1031, 3, 515, 388, 4, 4), (901, 1200, 4, 411, 309, 4, 3),
    (240, 320, 3, 640, 480, 2, 4), (240, 320, 4, 640, 480, 2, 3), (300, 200, 1, 100, 150, 4, 3),
    (300, 200, 1, 100, 150, 4, 4), (300, 200, 2, 100, 150, 1, 3), (300, 200, 2, 100, 150, 1, 4),
    (64, 64, 2, 128, 128, 0, 4), (480, 640, 3, 200, 150, 3, 4),
]


@pytest.mark.parametrize("case", CONVERT_CASES)
@pytest.mark.parametrize("mode", ["fast", "fast_f16", "fast_fp32", "exact"])
def test_resize_with_fused_channel_conversion(ctx, ik, oracle, case, mode):
    h, w, c, dw, dh, filt, co = case
    if mode == "exact" and h * w > 2_000_000:
        pytest.skip("large shapes run in fast mode only")
    src = splitmix_noise((h, w, c), image_id=co)
    ctx.set_mode({"exact": ik.MODE_EXACT, "fast": ik.MODE_FAST, "fast_f16": ik.MODE_FAST_F16, "fast_fp32": ik.MODE_FAST_FP32}[mode])
    got = ctx.resize(src, dw, dh, filt, out_channels=co)
    resized = oracle.resize_exact(src, dw, dh, filt)
    want = oracle.to_rgb8(resized) if co == 3 else oracle.to_rgba8(resized)
    assert got.shape == want.shape == (dh, dw, co)
    if mode == "exact":
        assert np.array_equal(got, want), delta_histogram(got, want)
    else:
        _check_fast(got, want, (case, mode), max_off=0.016 if mode == "fast" else 0.002)
    ctx.set_mode(ik.MODE_FAST)


def test_channel_conversion_trivial_cases(ctx, ik, oracle):
    src = splitmix_noise((40, 50, 3), image_id=5)
    same = ctx.resize(src, 50, 40, ik.FILTER_LANCZOS3, out_channels=4)      # same size: conversion only
    assert np.array_equal(same, oracle.to_rgba8(src))
    grey = splitmix_noise((40, 50, 2), image_id=6)
    assert np.array_equal(ctx.resize(grey, 50, 40, ik.FILTER_LANCZOS3, out_channels=3), oracle.to_rgb8(grey))
    with pytest.raises(ik.ImageKitError):
        ctx.resize(src, 20, 20, ik.FILTER_LANCZOS3, out_channels=2)         # only rgb / rgba destinations
    with pytest.raises(ik.ImageKitError):
        ctx.resize(src.astype(np.uint16), 20, 20, ik.FILTER_LANCZOS3, out_channels=4)


def test_resize_image_encode_as_matches_unfused_path(ctx, ik):
    """resize_image(..., encode_as=f) must hand encode_image the bytes to_rgb8()/to_rgba8() would have produced."""
    rgba = ik.DynamicImage(splitmix_noise((300, 400, 4), image_id=11))
    grey = ik.DynamicImage(splitmix_noise((300, 400, 1), image_id=12))
    for img in (rgba, grey):
        plain = ik.resize_image(img, 200, None, ctx=ctx)
        for fmt, conv in ((ik.ImageFormat.jpeg, "to_rgb8"), (ik.ImageFormat.webp, "to_rgb8"), (ik.ImageFormat.avif, "to_rgba8")):
            fused = ik.resize_image(img, 200, None, ctx=ctx, encode_as=fmt)
            assert np.array_equal(getattr(fused, conv)(), getattr(plain, conv)())
            assert fused.pixels.shape[2] == (4 if fmt == ik.ImageFormat.avif else 3)


# ---- exact 2x upscales (csrc/up2.cu; BASELINE config 4) ------------------------------------------------------
UP2_CASES = [  # (h, w, c, filter): output is (2w, 2h)
    (1080, 1920, 3, 2), (540, 960, 4, 2), (300, 400, 4, 4), (300, 400, 3, 4), (123, 77, 3, 1), (77, 123, 4, 1),
    (64, 64, 3, 0), (65, 33, 4, 0), (200, 336, 3, 3), (97, 224, 4, 3), (16, 8, 3, 2), (5, 3, 4, 4), (1, 1, 3, 2),
]


@pytest.mark.parametrize("case", UP2_CASES)
@pytest.mark.parametrize("content", ["noise", "edges"])
@pytest.mark.parametrize("mode", ["tc", "fp32"])
def test_exact_2x_upscale_kernel(ctx, ik, oracle, case, content, mode):
    """Default mode: banded8u.cu (vertical pass as an integer tensor-core product, 16-bit weights) where its layout applies
    (Rgb8 CatmullRom, Rgba8), else up2.cu; FAST_FP32: always up2.cu (CUDA cores, f32 weights)."""
    h, w, c, filt = case
    src = {"noise": splitmix_noise, "edges": checker}[content]((h, w, c))
    _fast(ctx, ik, mode)
    before = ctx.kernel_launches
    got = ctx.resize(src, 2 * w, 2 * h, filt)
    _fast(ctx, ik)
    assert ctx.kernel_launches - before == 1, "expected a single fused launch"
    want = oracle.resize_exact(src, 2 * w, 2 * h, filt)
    scale = 8 if mode == "tc" else 1
    _check_fast(got, want, (case, content), max_off=scale * (0.002 if content == "noise" else 0.06))


# ---- Luma8 / LumaA8 downscales on the ring kernel -------------------------------------------------
LUMA_RING_CASES = [  # (h, w, c, dw, dh, out_channels or None)
    (3024, 4032, 1, 400, 300, None), (1080, 1920, 2, 400, 225, None), (600, 800, 1, 400, 300, None),
    (777, 1031, 2, 515, 388, None), (1080, 1920, 1, 400, 225, 3), (901, 1200, 2, 411, 309, 4),
    (480, 640, 1, 200, 150, 4), (480, 640, 2, 200, 150, 3),
]


@pytest.mark.parametrize("case", LUMA_RING_CASES)
@pytest.mark.parametrize("mode", FAST_MODES)
def test_luma_downscales_on_the_ring_kernel(ctx, ik, oracle, case, mode):
    h, w, c, dw, dh, co = case
    src = splitmix_noise((h, w, c), image_id=c)
    _fast(ctx, ik, mode)
    before = ctx.kernel_launches
    got = ctx.resize(src, dw, dh, ik.FILTER_LANCZOS3, out_channels=co)
    _fast(ctx, ik)
    assert ctx.kernel_launches - before == 1, "expected the single-launch fused kernel"
    want = oracle.resize_exact(src, dw, dh, oracle.LANCZOS3)
    if co == 3: want = oracle.to_rgb8(want)
    if co == 4: want = oracle.to_rgba8(want)
    if want.ndim == 2: want = want[:, :, None]
    _check_fast(got.reshape(want.shape), want, case, max_off=0.016 if mode == "tc" else 0.002)


def test_randomised_parity_sweep(ctx, ik, oracle):
    """200 random shapes / ratios / channel counts / filters / conversions (seeded), with a bias towards the exact
    integer ratios the specialised loops detect: the cases a fixed list does not think of."""
    rng = np.random.default_rng(20261018)
    failures = []
    for _ in range(200):
        kind = rng.choice(["down_int", "down_any", "up2", "up_any", "mixed"], p=[0.3, 0.3, 0.15, 0.1, 0.15])
        c = int(rng.choice([1, 2, 3, 4], p=[0.1, 0.1, 0.4, 0.4]))
        filt = int(rng.choice([0, 1, 2, 3, 4], p=[0.05, 0.1, 0.2, 0.1, 0.55]))
        if kind == "down_int":
            r = int(rng.choice([2, 2, 2, 3, 4, 4, 5]))
            dw, dh = int(rng.integers(8, 600)), int(rng.integers(8, 400))
            w, h = dw * r, dh * r
            if rng.random() < 0.3:
                h = dh * int(rng.choice([2, 3, 4]))
        elif kind == "down_any":
            w, h = int(rng.integers(16, 2000)), int(rng.integers(16, 1400))
            dw, dh = int(rng.integers(1, w + 1)), int(rng.integers(1, h + 1))
        elif kind == "up2":
            w, h = int(rng.integers(1, 400)), int(rng.integers(1, 300))
            dw, dh = 2 * w, 2 * h
        elif kind == "up_any":
            w, h = int(rng.integers(1, 250)), int(rng.integers(1, 250))
            dw, dh = int(rng.integers(w, 3 * w + 2)), int(rng.integers(h, 3 * h + 2))
        else:
            w, h = int(rng.integers(8, 1000)), int(rng.integers(8, 1000))
            dw, dh = int(rng.integers(1, 2 * w)), int(rng.integers(1, 2 * h))
        if (w, h) == (dw, dh):
            dw += 1
        co = int(rng.choice([3, 4])) if rng.random() < 0.25 else None
        exact = rng.random() < 0.15
        src = (checker if rng.random() < 0.2 else splitmix_noise)((h, w, c))
        ctx.set_mode(ik.MODE_EXACT if exact else [ik.MODE_FAST, ik.MODE_FAST_F16, ik.MODE_FAST_FP32][int(rng.choice([0, 0, 1, 2]))])
        got = ctx.resize(src, dw, dh, filt, out_channels=co)
        want = oracle.resize_exact(src, dw, dh, filt)
        if co == 3:
            want = oracle.to_rgb8(want)
        if co == 4:
            want = oracle.to_rgba8(want)
        d = int(np.abs(got.astype(np.int32).reshape(want.shape) - want.astype(np.int32)).max())
        if d > (0 if exact else 1):
            failures.append(dict(kind=str(kind), h=h, w=w, c=c, dw=dw, dh=dh, filt=filt, co=co, exact=exact, max_delta=d))
    ctx.set_mode(ik.MODE_FAST)
    assert not failures, failures[:5]


def test_registered_host_memory_is_used_in_place(ctx, ik, oracle):
    """ikc_host_register: the caller's own (pageable) buffers are page-locked and DMA'd directly; same result."""
    L = ik._lib.load()
    src = splitmix_noise((600, 800, 3), image_id=21)
    dst = np.empty((300, 400, 3), np.uint8)
    assert L.ikc_host_register(src.ctypes.data, src.nbytes) == 0
    assert L.ikc_host_register(dst.ctypes.data, dst.nbytes) == 0
    try:
        _fast(ctx, ik)
        ctx.resize(src, 400, 300, ik.FILTER_LANCZOS3, out=dst)
        _check_fast(dst, oracle.resize_exact(src, 400, 300, oracle.LANCZOS3), "registered")
    finally:
        assert L.ikc_host_unregister(src.ctypes.data) == 0
        assert L.ikc_host_unregister(dst.ctypes.data) == 0
    assert L.ikc_host_register(None, 16) != 0


# ---- 16-bit rasters on the tile kernel (Luma16 / LumaA16 / Rgb16 / Rgba16 from PNG) ---------------
@pytest.mark.parametrize("shape", [(90, 120, 3, 50, 41), (200, 300, 1, 150, 100), (64, 48, 4, 96, 128), (120, 160, 2, 161, 119),
                                   (300, 400, 3, 200, 150), (37, 53, 4, 100, 80)])
@pytest.mark.parametrize("filt", [1, 2, 4])
def test_u16_rasters_single_launch(ctx, ik, oracle, shape, filt):
    h, w, c, dw, dh = shape
    rng = np.random.default_rng(h * 7 + filt)
    src = rng.integers(0, 65536, (h, w, c), dtype=np.uint16)
    src[: h // 4] = 65535                      # saturated and black bands: overshoot must clamp at both ends
    src[h // 4: h // 2, : w // 2] = 0
    _fast(ctx, ik)
    before = ctx.kernel_launches
    got = ctx.resize(src, dw, dh, filt)
    assert ctx.kernel_launches - before == 1, "expected the single-launch tile kernel"
    want = oracle.resize_exact(src, dw, dh, filt)
    assert np.abs(got.astype(np.int64) - want.astype(np.int64)).max() <= 1
    ctx.set_mode(ik.MODE_EXACT)
    assert np.array_equal(ctx.resize(src, dw, dh, filt), want)
    ctx.set_mode(ik.MODE_FAST)


def test_table_misses_do_not_stall_the_other_callers(ik, oracle, tmp_path):
    """A weight-table miss uploads on its own stream from pinned staging and frees stream-ordered (no device-wide
    synchronisation), and lanes are handed out first come, first served.  8 threads x 300 random target sizes through
    the C ABI (tools/table_miss_latency.c: every call a miss): p99 < 3 x p50.  (The old path -- cudaMalloc + blocking
    copies + cudaDeviceSynchronize per miss, a synchronising cudaFree per eviction, unfair lane hand-out -- had a p99 of
    12 ms against a p50 of 0.4 ms.)"""
    import json, os, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "rust-image-transform_b200")
    exe = str(tmp_path / "table_miss_latency")
    subprocess.check_call(["gcc", "-O2", "-std=c99", "-I" + os.path.join(root, "include"), os.path.join(root, "tools", "table_miss_latency.c"),
                           "-o", exe, "-L" + libdir, "-limagekit_cuda", "-lpthread", "-Wl,-rpath," + libdir])
    r = json.loads(subprocess.check_output([exe, "300"], timeout=300).decode().strip().splitlines()[-1])
    print(r)
    assert r["p99_ms"] < 3.0 * r["p50_ms"], r
    # and the results are still right when tables are brand new
    ctx = ik.Context([0])
    src = splitmix_noise((300, 400, 3))
    for dw, dh in [(173, 131), (201, 97), (88, 211)]:
        got = ctx.resize(src, dw, dh, ik.FILTER_LANCZOS3)
        _check_fast(got, oracle.resize_exact(src, dw, dh, oracle.LANCZOS3), (dw, dh))
    ctx.close()


def test_submit_queue_coalesces_concurrent_callers(ik, oracle):
    """ikc_submit_u8: handler threads that each bring one image share uploads, plans and launches.  Results are those of
    ikc_resize_u8; the counters (ikc_get_stats) show fewer launch groups than images under concurrency, and one group per
    image for a lone caller."""
    import threading
    ctx = ik.Context([0])
    shapes = [(480, 640, 3, 200, 150), (300, 400, 4, 200, 150), (240, 320, 3, 100, 75), (600, 800, 3, 160, 120), (96, 128, 1, 61, 47),
              (240, 320, 3, 640, 480), (64, 64, 3, 64, 64), (200, 300, 2, 150, 100)]
    srcs = [splitmix_noise((h, w, c), image_id=90 + i) for i, (h, w, c, _, _) in enumerate(shapes)]
    want = [oracle.resize_exact(s, dw, dh, oracle.LANCZOS3) for s, (_, _, _, dw, dh) in zip(srcs, shapes)]
    # a lone caller: every image its own group
    for s, (h, w, c, dw, dh), exp in zip(srcs, shapes, want):
        got = ctx.submit(s, dw, dh, ik.FILTER_LANCZOS3)
        hist = delta_histogram(got, exp)
        assert max(abs(k) for k in hist) <= TOL, ((h, w, c, dw, dh), hist)
    st0 = ctx.stats()
    assert st0["submit_jobs"] == len(shapes) == st0["submit_batches"] and st0["calls"] == len(shapes) and st0["failed"] == 0
    assert st0["trivial"] == 1 and st0["launches"] >= len(shapes) - 1 and st0["src_bytes"] > 0 and st0["busy_ns"] > 0
    # 16 threads x 12 images at once
    errors, rounds = [], 12
    def worker(t):
        try:
            for r in range(rounds):
                i = (t + r) % len(shapes)
                h, w, c, dw, dh = shapes[i]
                got = ctx.submit(srcs[i], dw, dh, ik.FILTER_LANCZOS3)
                hist = delta_histogram(got, want[i])
                assert max(abs(k) for k in hist) <= TOL, (shapes[i], hist)
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))
    th = [threading.Thread(target=worker, args=(t,)) for t in range(16)]
    for x in th: x.start()
    for x in th: x.join()
    assert not errors, errors[:3]
    st1 = ctx.stats()
    jobs = st1["submit_jobs"] - st0["submit_jobs"]
    groups = st1["submit_batches"] - st0["submit_batches"]
    assert jobs == 16 * rounds and groups < jobs, (jobs, groups)
    # a bad request fails alone
    with pytest.raises(ik.ImageKitError):
        ctx.submit(srcs[0], 100, 10, 9)      # unknown filter: rejected inside the library
    assert ctx.stats()["failed"] >= 1
    ctx.close()


@pytest.mark.skipif(not __import__("os").path.exists(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "reference_image_0_25_8.npz")),
                    reason="reference_image_0_25_8.npz not generated yet: needs cargo (oracle/ref_harness)")
def test_reference_pinned_vectors_on_the_gpu(ctx, ik):
    """The GPU twin of tests/test_oracle_props.py::test_reference_pinned_vectors: EXACT mode reproduces the real
    `image` 0.25.8 bit for bit, FAST mode within +-1."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_image_0_25_8.npz"))
    for nme in sorted({k.split("/")[0] for k in g.files}):
        src, ref = g[nme + "/src"], g[nme + "/ref"]
        filt = int(g[nme + "/meta"][2])
        dh, dw = ref.shape[:2]
        ctx.set_mode(ik.MODE_EXACT)
        assert np.array_equal(ctx.resize(src, dw, dh, filt), ref), nme
        ctx.set_mode(ik.MODE_FAST)
        assert np.abs(ctx.resize(src, dw, dh, filt).astype(int) - ref.astype(int)).max() <= TOL, nme


def test_split_call_begin_end(ik, oracle):
    """ikc_resize_begin_u8 / ikc_resize_end: the call split in two so that a worker can decode its next upload while this
    one is on the GPU.  Same pixels as ikc_resize_u8; one thread may hold several tickets (the lane pool grows); many
    threads, one ticket each, run concurrently."""
    import threading
    ctx = ik.Context([0])
    src = splitmix_noise((480, 640, 3), image_id=5)
    want = oracle.resize_exact(src, 200, 150, oracle.LANCZOS3)
    want_rgba = oracle.to_rgba8(want)
    tickets = [ctx.resize_begin(src, 200, 150, ik.FILTER_LANCZOS3, out_channels=4 if i % 2 else None) for i in range(6)]
    for i, t in enumerate(tickets):
        got = t.end()
        assert np.abs(got.astype(int) - (want_rgba if i % 2 else want).astype(int)).max() <= TOL
    assert ctx.resize_begin(src, 640, 480).end().tobytes() == src.tobytes()       # same size: answered at begin
    errors = []
    def worker(t):
        try:
            prev = None
            for r in range(10):
                s = np.roll(src, t + r, axis=1)
                if prev is not None:
                    got, exp = prev[0].end(), prev[1]
                    assert np.abs(got.astype(int) - exp.astype(int)).max() <= TOL
                prev = (ctx.resize_begin(s, 200, 150), np.roll(want, 0, axis=1) if (t + r) == 0 else None)
                if prev[1] is None:
                    prev = (prev[0], oracle.resize_exact(s, 200, 150, oracle.LANCZOS3))
            prev[0].end()
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))
    th = [threading.Thread(target=worker, args=(t,)) for t in range(12)]
    for x in th: x.start()
    for x in th: x.join(120)
    assert not any(x.is_alive() for x in th), "a worker is stuck"
    assert not errors, errors[:3]
    with pytest.raises(ik.ImageKitError):
        ctx.resize_begin(src, 100, 10, 9)
    ctx.close()


def test_host_batch_groups_small_images(ik, oracle):
    """ikc_resize_batch: runs of small images (<= 1 MB of pixels) share one staged upload, plan, launch per kernel variant and
    download; big ones keep their per-image pipeline; a bad job in a run fails alone."""
    ctx = ik.Context([0])
    rng = np.random.default_rng(5)
    srcs, sizes = [], []
    for i in range(70):
        if i % 23 == 11:
            h, w, c = 1200, 1600, 3                                    # a big one in the middle of the run
        else:
            h, w, c = int(rng.integers(100, 300)), int(rng.integers(100, 400)), int(rng.choice([3, 3, 4]))
        srcs.append(splitmix_noise((h, w, c), image_id=300 + i))
        r = float(rng.uniform(2.0, 4.0))                               # thumbnails of small sources (same-size: a few)
        sizes.append((w, h) if i % 17 == 3 else (max(8, int(w / r)), max(8, int(h / r))))
    before = ctx.stats()
    outs, jobs = ctx.resize_batch(srcs, sizes)
    after = ctx.stats()
    assert all(j.status == 0 for j in jobs)
    for s, (dw, dh), o_ in zip(srcs, sizes, outs):
        want = oracle.resize_exact(s, dw, dh, oracle.LANCZOS3) if (dw, dh) != (s.shape[1], s.shape[0]) else s
        assert np.abs(o_.astype(int) - want.reshape(o_.shape).astype(int)).max() <= TOL, (s.shape, dw, dh)
    assert after["calls"] - before["calls"] == 70 and after["trivial"] - before["trivial"] == sum(1 for i in range(70) if i % 17 == 3)
    assert after["launches"] - before["launches"] <= 33, (before, after)    # 66 resizes: groups of up to 32 share their launches
    # one bad job (unknown filter is per batch; use an absurd pitch instead) fails alone
    L = ik._lib.load()
    n = 5
    arr = (ik._lib.Job * n)()
    keep = []
    for i in range(n):
        s = splitmix_noise((64, 64, 3), image_id=i)
        d = np.zeros((32, 32, 3), np.uint8)
        keep.append((s, d))
        arr[i] = ik._lib.Job(s.ctypes.data, d.ctypes.data, 64, 64, 32, 32, 64 * 3 if i != 2 else 5, 32 * 3, 3, ik.FILTER_LANCZOS3, 0, 0)
    rc = L.ikc_resize_batch(ctx._h, arr, n)
    assert rc != 0 and [arr[i].status for i in range(n)] == [0, 0, ik._lib.ERR_INVALID_ARG, 0, 0]
    for i in (0, 1, 3, 4):
        assert np.abs(keep[i][1].astype(int) - oracle.resize_exact(keep[i][0], 32, 32, oracle.LANCZOS3).astype(int)).max() <= TOL
    ctx.close()


def test_fast_mode_delta_on_the_headline_config(ctx, ik, oracle):
    """BASELINE config 2 at full size (4K Rgba8 -> 1080p Lanczos3, the row-band tensor-core kernel), on three kinds of
    content: how many samples FAST mode moves by one level, never more.  The measured histograms are written to
    gpurun_out/ (copied to profiles/r02_fast_mode_delta.json) -- the source of the figures DESIGN.md / INTEGRATION.md
    quote."""
    import json
    import os
    h, w, c, dw, dh = 2160, 3840, 4, 1920, 1080
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    smooth = np.stack([127 + 90 * np.sin(xx / (211.0 + 17 * k)) * np.cos(yy / (173.0 - 11 * k)) for k in range(c)], axis=2)
    contents = {"noise": splitmix_noise((h, w, c), image_id=77), "photo_like": photo_like((h, w, c)),
                "smooth": np.clip(smooth, 0, 255).astype(np.uint8)}
    _fast(ctx, ik)
    record = {}
    before = ctx.stats()["launches_banded8t"]
    for name, src in contents.items():
        got = ctx.resize(src, dw, dh, 4)
        want = oracle.resize_exact(src, dw, dh, 4)
        hist = _check_fast(got, want, name)
        record[name] = {"delta_histogram": {str(k): v for k, v in hist.items()},
                        "fraction_off_by_one": sum(v for k, v in hist.items() if k != 0) / got.size}
    assert ctx.stats()["launches_banded8t"] - before == len(contents)
    out = os.path.join(os.path.dirname(os.path.dirname(__file__)), "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "fast_mode_delta.json"), "w") as f:
            json.dump({"workload": "cfg2: 3840x2160 RGBA8 -> 1920x1080 Lanczos3, IKC_MODE_FAST vs CPU oracle", "contents": record}, f, indent=1)


def test_oversized_request_gives_its_staging_back(ik, oracle):
    """One huge raster must not pin hundreds of megabytes on a lane for the life of the process: lane buffers above
    256 MB are freed when the lane is released (ikc_stats_t.staging_trims counts it), and the next ordinary request
    on that lane simply allocates ordinary buffers again."""
    c = ik.Context([0])
    try:
        c.set_mode(ik.MODE_FAST)
        big = np.full((7500, 10000, 4), 131, np.uint8)          # 300 MB: above the keep size
        big[::97, ::89] = 17
        got = c.resize(big, 100, 75, 4)
        assert c.stats()["staging_trims"] >= 1
        assert got.shape == (75, 100, 4) and int(got.min()) >= 16 and int(got.max()) <= 132
        t0 = c.stats()["staging_trims"]
        for i in range(6):                                     # ordinary traffic afterwards: no trims, right answers
            src = splitmix_noise((240, 320, 4), image_id=900 + i)
            _check_fast(c.resize(src, 160, 120, 4), oracle.resize_exact(src, 160, 120, 4), ("after trim", i))
        assert c.stats()["staging_trims"] == t0 and c.stats()["failed"] == 0
    finally:
        c.close()
