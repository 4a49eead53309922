"""Dims rule: the nine dimension known-answers the reference's own tests pin
(/root/reference/tests/transform.rs), checked against the oracle AND the product's ikc_target_dims."""
import numpy as np
import pytest

# (orig_w, orig_h, w, h) -> (tw, th); file:line of the reference assertion
REFERENCE_KNOWN_ANSWERS = [
    ((800, 600, 400, None), (400, 300)),     # tests/transform.rs:11-19  width only
    ((800, 600, None, 300), (400, 300)),     # :22-30  height only
    ((800, 600, 400, 300), (400, 300)),      # :33-40  both
    ((1920, 1080, 960, None), (960, 540)),   # :43-51  16:9
    ((800, 600, None, None), (800, 600)),    # :58-66  none -> unchanged
    ((100, 100, 200, 200), (200, 200)),      # :69-76  upscale
    ((800, 600, 1, 1), (1, 1)),              # :79-86  minimum
    ((2, 2, 200, 200), (200, 200)),          # :89-96  extreme upscale
    ((1920, 1080, 640, 480), (640, 360)),    # :239-257 fit-within, not exact
    ((800, 600, 400, None), (400, 300)),     # :224-236 resize_and_encode_jpeg
    ((1000, 1000, 100, 100), (100, 100)),    # :276-288 resize shrinks jpeg
]


@pytest.mark.parametrize("args,expect", REFERENCE_KNOWN_ANSWERS)
def test_reference_dimension_known_answers_oracle(oracle, args, expect):
    ow, oh, w, h = args
    tw, th, _ = oracle.target_dims(ow, oh, w, h)
    assert (tw, th) == expect


@pytest.mark.parametrize("args,expect", REFERENCE_KNOWN_ANSWERS)
def test_reference_dimension_known_answers_product(ik, args, expect):
    ow, oh, w, h = args
    tw, th, _ = ik.target_dims(ow, oh, w, h)
    assert (tw, th) == expect


def test_dims_codes(oracle, ik):
    for fn in (oracle.target_dims, ik.target_dims):
        assert fn(800, 600, None, None)[2] == 1          # passthrough
        assert fn(800, 600, 800, 600)[2] == 2            # clone
        assert fn(800, 600, 800, 700)[2] == 3            # fit-within lands on the same size -> copy
        assert fn(800, 600, 400, None)[2] == 0
        assert fn(1000, 333, 100, None)[:2] == (99, 33)  # smaller than asked in both axes


def test_dims_product_equals_oracle_random(oracle, ik):
    rng = np.random.default_rng(11)
    for _ in range(4000):
        ow, oh = int(rng.integers(1, 9000)), int(rng.integers(1, 9000))
        mode = rng.integers(0, 3)
        w = int(rng.integers(0, 12000)) if mode != 1 else None
        h = int(rng.integers(0, 12000)) if mode != 0 else None
        assert ik.target_dims(ow, oh, w, h) == oracle.target_dims(ow, oh, w, h), (ow, oh, w, h)


def test_dims_extremes(oracle, ik):
    cases = [(1, 1, 0xFFFFFFFF, None), (3, 1, 0xFFFFFFFF, 0xFFFFFFFF), (65535, 1, None, 7),
             (1, 65535, 4000000000, None), (7, 3, 0, 0), (5, 5, 0, None)]
    for ow, oh, w, h in cases:
        assert ik.target_dims(ow, oh, w, h) == oracle.target_dims(ow, oh, w, h), (ow, oh, w, h)
