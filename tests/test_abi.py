"""C-ABI checks that need no GPU: the library loads, exports every symbol include/imagekit_cuda.h
declares, its host-built weight tables equal the oracle's bit for bit, and it fails loudly (no CPU
fallback) when there is no CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "imagekit_cuda.h")


def _declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"IKC_API[^;(]*?\b(ikc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(ik):
    L = ik._lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert sorted(ik._lib.EXPORTS) == declared
    assert L.ikc_version() == 1


def test_job_struct_layout_matches_header(ik):
    # struct ikc_job: 2 pointers, 4 u32, 2 size_t, 4 i32 = 64 bytes on LP64
    assert ctypes.sizeof(ik._lib.Job) == 64
    assert ik._lib.Job.sw.offset == 16 and ik._lib.Job.src_pitch.offset == 32 and ik._lib.Job.channels.offset == 48


@pytest.mark.parametrize("filt,n_in,n_out", [(4, 2160, 1080), (4, 3024, 300), (4, 1080, 225), (2, 1080, 2160),
                                             (3, 777, 123), (1, 50, 500), (0, 100, 37), (4, 600, 1), (4, 2, 200),
                                             (4, 4032, 400), (2, 640, 640), (4, 1, 1)])
def test_weight_tables_bit_identical_to_oracle(ik, oracle, filt, n_in, n_out):
    from imagekit_cuda import engine
    l1, c1, w1 = engine.pass_table(filt, n_in, n_out)
    l2, c2, w2 = oracle.pass_table(filt, n_in, n_out)
    m = min(w1.shape[1], w2.shape[1])
    assert np.array_equal(l1, l2) and np.array_equal(c1, c2)
    assert c1.max() <= m
    assert np.array_equal(w1[:, :m].view(np.uint32), w2[:, :m].view(np.uint32))


def test_no_cpu_fallback_without_gpu(ik):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(ik.ImageKitError) as e:
        ik.Context()
    assert "no CPU fallback" in str(e.value)
    img = ik.DynamicImage.new_rgb8(80, 60)
    assert ik.resize_image(img, None, None) is img          # transform.rs:67-69 needs no device
    with pytest.raises(ik.ImageKitError):
        ik.resize_image(img, 40, None)


def test_error_paths_without_context(ik):
    L = ik._lib.load()
    assert L.ikc_resize_u8(None, None, 1, 1, 1, 1, None, 1, 1, 1, 4) == ik._lib.ERR_INVALID_ARG
    assert L.ikc_device_count(None) == 0
    assert L.ikc_batch_launch(None, None) == ik._lib.ERR_INVALID_ARG
    assert b"null" in L.ikc_last_error()
