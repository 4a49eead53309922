"""ctypes loader for the CPU oracle (oracle/imageops_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importers allowed: tests/, __graft_entry__.smoke(), and
bench.py's cpu_baseline / --impl reference legs.  The product package
(rust-image-transform_b200/) must never import this module.

Parity status: "parity unpinned" for pixel values (see the header of imageops_oracle.c);
the reference pins dimensions only (/root/reference/tests/transform.rs:11-96, 239-257).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liborc.so")

NEAREST, TRIANGLE, CATMULLROM, GAUSSIAN, LANCZOS3 = range(5)
FILTER_NAMES = ["nearest", "triangle", "catmullrom", "gaussian", "lanczos3"]

_lib = None
_lock = threading.Lock()


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "imageops_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liborc.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    with _lock:
        if _lib is None:
            build()
            L = C.CDLL(_SO)
            u32, sz, i32 = C.c_uint32, C.c_size_t, C.c_int
            p8, p16, pf, pu32 = (C.POINTER(C.c_uint8), C.POINTER(C.c_uint16), C.POINTER(C.c_float),
                                 C.POINTER(C.c_uint32))
            L.orc_target_dims.argtypes = [u32, u32, i32, u32, i32, u32, pu32, pu32]
            L.orc_target_dims.restype = i32
            L.orc_kernel.argtypes = [i32, C.c_float]
            L.orc_kernel.restype = C.c_float
            L.orc_pass_table.argtypes = [i32, u32, u32, pu32, pu32, pf, u32]
            L.orc_pass_table.restype = u32
            L.orc_resize_u8.argtypes = [p8, u32, u32, sz, i32, p8, u32, u32, sz, i32]
            L.orc_resize_u8.restype = i32
            L.orc_resize_u16.argtypes = [p16, u32, u32, sz, i32, p16, u32, u32, sz, i32]
            L.orc_resize_u16.restype = i32
            L.orc_vertical_f32_u8.argtypes = [p8, u32, u32, sz, i32, pf, u32, i32]
            L.orc_vertical_f32_u8.restype = i32
            L.orc_resize_image_u8.argtypes = [p8, u32, u32, i32, i32, u32, i32, u32, p8]
            L.orc_resize_image_u8.restype = i32
            _lib = L
    return _lib


def target_dims(ow: int, oh: int, w: int | None, h: int | None):
    """(tw, th, code) per transform.rs:62-90 + DynamicImage::resize + resize_dimensions.
    code: 0 resample, 1 (None,None) passthrough, 2 clone (same dims asked), 3 copy (fit == dims)."""
    tw, th = C.c_uint32(), C.c_uint32()
    code = lib().orc_target_dims(ow, oh, w is not None, w or 0, h is not None, h or 0,
                                 C.byref(tw), C.byref(th))
    return tw.value, th.value, code


def kernel(filt: int, x: float) -> float:
    return float(lib().orc_kernel(filt, x))


def pass_table(filt: int, n_in: int, n_out: int):
    """(left[u32 n_out], count[u32 n_out], weights[f32 n_out x stride]) for one pass."""
    L = lib()
    stride = L.orc_pass_table(filt, n_in, n_out, None, None, None, 0)
    left = np.zeros(n_out, np.uint32)
    cnt = np.zeros(n_out, np.uint32)
    w = np.zeros((n_out, stride), np.float32)
    L.orc_pass_table(filt, n_in, n_out, left.ctypes.data_as(C.POINTER(C.c_uint32)),
                     cnt.ctypes.data_as(C.POINTER(C.c_uint32)),
                     w.ctypes.data_as(C.POINTER(C.c_float)), stride)
    return left, cnt, w


def _as_hwc(a: np.ndarray) -> np.ndarray:
    if a.ndim == 2:
        a = a[:, :, None]
    assert a.ndim == 3 and 1 <= a.shape[2] <= 4
    return np.ascontiguousarray(a)


def resize_exact(src: np.ndarray, dw: int, dh: int, filt: int = LANCZOS3) -> np.ndarray:
    """imageops::resize(src, dw, dh, filt): src is HxWxC (or HxW) u8/u16; returns dh x dw x C."""
    squeeze = src.ndim == 2
    s = _as_hwc(src)
    sh, sw, ch = s.shape
    dst = np.empty((dh, dw, ch), s.dtype)
    if s.dtype == np.uint8:
        rc = lib().orc_resize_u8(s.ctypes.data_as(C.POINTER(C.c_uint8)), sw, sh, sw * ch, ch,
                                 dst.ctypes.data_as(C.POINTER(C.c_uint8)), dw, dh, dw * ch, filt)
    elif s.dtype == np.uint16:
        rc = lib().orc_resize_u16(s.ctypes.data_as(C.POINTER(C.c_uint16)), sw, sh, sw * ch * 2, ch,
                                  dst.ctypes.data_as(C.POINTER(C.c_uint16)), dw, dh, dw * ch * 2, filt)
    else:
        raise TypeError(s.dtype)
    if rc != 0:
        raise RuntimeError(f"oracle resize failed rc={rc}")
    return dst[:, :, 0] if squeeze else dst


def to_rgb8(img: np.ndarray) -> np.ndarray:
    """DynamicImage::to_rgb8() of an 8-bit HxW / HxWxC raster (image 0.25.8 color conversions; used by
    encode_image, src/transform.rs:123,131): grey replicated, alpha dropped."""
    a = _as_hwc(img)
    c = a.shape[2]
    return np.ascontiguousarray(a[:, :, [0, 0, 0]] if c <= 2 else a[:, :, :3])


def to_rgba8(img: np.ndarray) -> np.ndarray:
    """DynamicImage::to_rgba8() (src/transform.rs:140): as to_rgb8 plus the source's alpha, or 255."""
    a = _as_hwc(img)
    c = a.shape[2]
    alpha = a[:, :, c - 1:c] if c in (2, 4) else np.full(a.shape[:2] + (1,), 255, np.uint8)
    return np.ascontiguousarray(np.concatenate([to_rgb8(a), alpha], axis=2))


def vertical_f32(src: np.ndarray, dh: int, filt: int = LANCZOS3) -> np.ndarray:
    """The unclamped f32 intermediate of vertical_sample: dh x sw x C float32."""
    s = _as_hwc(src)
    assert s.dtype == np.uint8
    sh, sw, ch = s.shape
    tmp = np.empty((dh, sw, ch), np.float32)
    rc = lib().orc_vertical_f32_u8(s.ctypes.data_as(C.POINTER(C.c_uint8)), sw, sh, sw * ch, ch,
                                   tmp.ctypes.data_as(C.POINTER(C.c_float)), dh, filt)
    if rc != 0:
        raise RuntimeError(f"oracle vertical pass failed rc={rc}")
    return tmp


def resize_image(src: np.ndarray, w: int | None, h: int | None) -> np.ndarray:
    """resize_image(img, w, h) (transform.rs:62-90): dims rule + Lanczos3."""
    s = _as_hwc(src)
    sh, sw, _ = s.shape
    tw, th, code = target_dims(sw, sh, w, h)
    if code != 0:
        out = s.copy()
    else:
        out = resize_exact(s, tw, th, LANCZOS3)
    return out[:, :, 0] if src.ndim == 2 else out
