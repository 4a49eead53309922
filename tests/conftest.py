import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "rust-image-transform_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.build()
    return o


@pytest.fixture(scope="session")
def ik():
    import imagekit_cuda
    return imagekit_cuda


@pytest.fixture(scope="session")
def ctx(ik):
    """One context over cuda:0 for the whole GPU session (fails loudly if the .so or GPU is missing)."""
    c = ik.Context([0])
    yield c
    c.close()


# ---- deterministic synthetic rasters -------------------------------------------

def splitmix_noise(shape, seed=0x1234ABCD, image_id=0):
    """u8 noise from a counter hash: splitmix64(seed ^ (image_id << 40 | byte_index)) >> 56."""
    n = int(np.prod(shape))
    z = (np.arange(n, dtype=np.uint64) | (np.uint64(image_id) << np.uint64(40))) ^ np.uint64(seed)
    z = z + np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(56)).astype(np.uint8).reshape(shape)


def photo_like(shape, seed=7):
    h, w, c = shape
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.zeros(shape, np.float64)
    for k in range(c):
        img[:, :, k] = (127 + 60 * np.sin(xx / (37.0 + 5 * k)) + 40 * np.sin(yy / (53.0 - 3 * k)) +
                        25 * np.sin((xx + yy) / 91.0))
    rng = np.random.default_rng(seed)
    img += rng.integers(-8, 9, shape)
    return np.clip(img, 0, 255).astype(np.uint8)


def checker(shape, period=7):
    h, w, c = shape
    yy, xx = np.mgrid[0:h, 0:w]
    m = (((xx // period) + (yy // period)) & 1).astype(np.uint8) * 255
    return np.repeat(m[:, :, None], c, axis=2)


def delta_histogram(got, want):
    d = got.astype(np.int32) - want.astype(np.int32)
    vals, counts = np.unique(d, return_counts=True)
    return {int(v): int(n) for v, n in zip(vals, counts)}
