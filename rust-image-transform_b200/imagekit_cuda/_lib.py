"""ctypes binding of libimagekit_cuda.so (C ABI: include/imagekit_cuda.h).

This is what a maintainer's FFI stub looks like from Python; the Rust equivalent is
rust-image-transform_b200/crate/src/ffi.rs.  There is no fallback: if the shared library is
missing, importing the symbols fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libimagekit_cuda.so")

OK, ERR_INVALID_ARG, ERR_UNSUPPORTED, ERR_TOO_LARGE, ERR_CUDA, ERR_OOM = range(6)
STATUS_NAMES = ["ok", "invalid-arg", "unsupported-layout", "too-large", "cuda-error", "oom"]
FILTER_NEAREST, FILTER_TRIANGLE, FILTER_CATMULLROM, FILTER_GAUSSIAN, FILTER_LANCZOS3 = range(5)
MODE_FAST, MODE_EXACT, MODE_FAST_FP32, MODE_FAST_F16 = 0, 1, 2, 3
DIMS_RESAMPLE, DIMS_PASSTHROUGH, DIMS_CLONE, DIMS_COPY = range(4)
MAX_DIM, MAX_PIXELS = 65535, 1 << 28  # IKC_MAX_DIM, IKC_MAX_PIXELS

# every symbol include/imagekit_cuda.h declares
EXPORTS = [
    "ikc_create", "ikc_destroy", "ikc_device_count", "ikc_set_mode", "ikc_get_mode", "ikc_kernel_launches",
    "ikc_last_error", "ikc_version", "ikc_target_dims", "ikc_check_dims", "ikc_pass_table", "ikc_pass_info", "ikc_pass_band", "ikc_pass_band8", "ikc_pass_band8t", "ikc_resize_u8", "ikc_submit_u8", "ikc_get_stats", "ikc_resize_begin_u8", "ikc_resize_end", "ikc_resize_u16",
    "ikc_resize_convert_u8", "ikc_resize_image_u8", "ikc_resize_batch", "ikc_host_alloc", "ikc_host_free", "ikc_host_register", "ikc_host_unregister", "ikc_resize_u8_device",
    "ikc_batch_prepare", "ikc_batch_launch", "ikc_batch_launch_count", "ikc_batch_describe", "ikc_batch_free",
]


class Job(C.Structure):
    """struct ikc_job"""
    _fields_ = [
        ("src", C.c_void_p), ("dst", C.c_void_p),
        ("sw", C.c_uint32), ("sh", C.c_uint32), ("dw", C.c_uint32), ("dh", C.c_uint32),
        ("src_pitch", C.c_size_t), ("dst_pitch", C.c_size_t),
        ("channels", C.c_int32), ("filter", C.c_int32), ("status", C.c_int32), ("device", C.c_int32),
    ]


class Stats(C.Structure):
    """struct ikc_stats_t"""
    _fields_ = [(n, C.c_uint64) for n in (
        "calls", "failed", "trivial", "launches", "launches_banded8t", "launches_banded8", "launches_banded_f16", "launches_ring",
        "launches_up2", "launches_tile", "launches_generic", "src_bytes", "dst_bytes", "busy_ns", "table_hits", "table_misses",
        "submit_batches", "submit_jobs", "launches_banded8u", "staging_trims")]


class PassInfo(C.Structure):
    """struct ikc_pass_info_t"""
    _fields_ = [("stride", C.c_uint32), ("max_count", C.c_uint32), ("ring_k", C.c_int32),
                ("uni_step", C.c_int32), ("uni_lo", C.c_int32), ("uni_hi", C.c_int32),
                ("up2_taps", C.c_int32), ("up2_off", C.c_int32), ("up2_uni_lo", C.c_int32), ("up2_uni_hi", C.c_int32)]


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C rust-image-transform_b200/csrc` "
            "(or __graft_entry__.build()). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    u32, sz, i32, vp = C.c_uint32, C.c_size_t, C.c_int, C.c_void_p
    pu32 = C.POINTER(C.c_uint32)
    L.ikc_create.argtypes = [C.POINTER(C.c_int), i32, C.POINTER(vp)]
    L.ikc_create.restype = i32
    L.ikc_destroy.argtypes = [vp]
    L.ikc_destroy.restype = None
    L.ikc_device_count.argtypes = [vp]
    L.ikc_device_count.restype = i32
    L.ikc_set_mode.argtypes = [vp, i32]
    L.ikc_set_mode.restype = i32
    L.ikc_get_mode.argtypes = [vp]
    L.ikc_get_mode.restype = i32
    L.ikc_kernel_launches.argtypes = [vp]
    L.ikc_kernel_launches.restype = C.c_uint64
    L.ikc_last_error.argtypes = []
    L.ikc_last_error.restype = C.c_char_p
    L.ikc_version.argtypes = []
    L.ikc_version.restype = i32
    L.ikc_target_dims.argtypes = [u32, u32, i32, u32, i32, u32, pu32, pu32]
    L.ikc_target_dims.restype = i32
    L.ikc_check_dims.argtypes = [u32, u32, u32, u32]
    L.ikc_check_dims.restype = i32
    L.ikc_pass_table.argtypes = [i32, u32, u32, pu32, pu32, C.POINTER(C.c_float), u32]
    L.ikc_pass_table.restype = u32
    L.ikc_pass_band.argtypes = [i32, u32, u32, pu32, C.POINTER(C.c_int32), C.POINTER(C.c_uint16), sz]
    L.ikc_pass_band.restype = u32
    L.ikc_pass_band8.argtypes = [i32, u32, u32, pu32, pu32, C.POINTER(C.c_int32), C.POINTER(C.c_int8), sz]
    L.ikc_pass_band8.restype = u32
    L.ikc_pass_band8t.argtypes = [i32, u32, u32, pu32, pu32, C.POINTER(C.c_int32), C.POINTER(C.c_int8), sz]
    L.ikc_pass_band8t.restype = u32
    L.ikc_pass_info.argtypes = [i32, u32, u32, C.POINTER(PassInfo)]
    L.ikc_pass_info.restype = i32
    L.ikc_resize_u8.argtypes = [vp, vp, u32, u32, sz, i32, vp, u32, u32, sz, i32]
    L.ikc_resize_u8.restype = i32
    L.ikc_submit_u8.argtypes = [vp, vp, u32, u32, sz, i32, vp, u32, u32, sz, i32]
    L.ikc_submit_u8.restype = i32
    L.ikc_resize_begin_u8.argtypes = [vp, vp, u32, u32, sz, i32, vp, u32, u32, sz, i32, C.POINTER(vp)]
    L.ikc_resize_begin_u8.restype = i32
    L.ikc_resize_end.argtypes = [vp]
    L.ikc_resize_end.restype = i32
    L.ikc_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.ikc_get_stats.restype = i32
    L.ikc_resize_convert_u8.argtypes = [vp, vp, u32, u32, sz, i32, vp, u32, u32, sz, i32, i32]
    L.ikc_resize_convert_u8.restype = i32
    L.ikc_resize_u16.argtypes = [vp, vp, u32, u32, sz, i32, vp, u32, u32, sz, i32]
    L.ikc_resize_u16.restype = i32
    L.ikc_resize_image_u8.argtypes = [vp, vp, u32, u32, i32, i32, u32, i32, u32, vp, sz, pu32, pu32]
    L.ikc_resize_image_u8.restype = i32
    L.ikc_resize_batch.argtypes = [vp, C.POINTER(Job), sz]
    L.ikc_resize_batch.restype = i32
    L.ikc_host_alloc.argtypes = [sz, C.POINTER(vp)]
    L.ikc_host_alloc.restype = i32
    L.ikc_host_free.argtypes = [vp]
    L.ikc_host_free.restype = None
    L.ikc_host_register.argtypes = [vp, sz]
    L.ikc_host_register.restype = i32
    L.ikc_host_unregister.argtypes = [vp]
    L.ikc_host_unregister.restype = i32
    L.ikc_resize_u8_device.argtypes = [vp, i32, vp, vp, u32, u32, sz, i32, vp, u32, u32, sz, i32]
    L.ikc_resize_u8_device.restype = i32
    L.ikc_batch_prepare.argtypes = [vp, i32, C.POINTER(Job), sz, C.POINTER(vp)]
    L.ikc_batch_prepare.restype = i32
    L.ikc_batch_launch.argtypes = [vp, vp]
    L.ikc_batch_launch.restype = i32
    L.ikc_batch_launch_count.argtypes = [vp]
    L.ikc_batch_launch_count.restype = i32
    L.ikc_batch_describe.argtypes = [vp, C.c_char_p, sz]
    L.ikc_batch_describe.restype = i32
    L.ikc_batch_free.argtypes = [vp]
    L.ikc_batch_free.restype = None
    _lib = L
    return L


def last_error() -> str:
    return (load().ikc_last_error() or b"").decode("utf-8", "replace")
