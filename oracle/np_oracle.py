"""numpy twin of the CPU oracle -- an INDEPENDENT second restatement of image 0.25.8's
`imageops::resize` (src/imageops/sample.rs; call site /root/reference/src/transform.rs:85-89)
used only to cross-check oracle/imageops_oracle.c on small cases.

TEST INFRASTRUCTURE ONLY (same import rules as oracle/oracle.py).  "parity unpinned".

Every scalar step is done in np.float32 so each operation rounds exactly like Rust's f32;
sin/exp go through glibc's sinf/expf via ctypes (numpy's own float32 sin is a different
SIMD implementation and may differ in the last ulp).
"""
from __future__ import annotations

import ctypes
import ctypes.util

import numpy as np

f32 = np.float32
_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.sinf.argtypes = [ctypes.c_float]
_libm.sinf.restype = ctypes.c_float
_libm.expf.argtypes = [ctypes.c_float]
_libm.expf.restype = ctypes.c_float
PI = f32(3.1415927410125732)

NEAREST, TRIANGLE, CATMULLROM, GAUSSIAN, LANCZOS3 = range(5)
SUPPORT = {NEAREST: f32(0.0), TRIANGLE: f32(1.0), CATMULLROM: f32(2.0), GAUSSIAN: f32(3.0),
           LANCZOS3: f32(3.0)}


def _sinc(t: f32) -> f32:
    a = f32(t * PI)
    if t == 0.0:
        return f32(1.0)
    return f32(f32(_libm.sinf(a)) / a)


def _lanczos3(x: f32) -> f32:
    if abs(x) < f32(3.0):
        return f32(_sinc(x) * _sinc(f32(x / f32(3.0))))
    return f32(0.0)


def _catmullrom(x: f32) -> f32:
    b, c = f32(0.0), f32(0.5)
    a = f32(abs(x))
    a2 = f32(a * a)
    a3 = f32(a * a2)
    if a < 1.0:
        c3 = f32(f32(f32(12.0) - f32(f32(9.0) * b)) - f32(f32(6.0) * c))
        c2 = f32(f32(f32(-18.0) + f32(f32(12.0) * b)) + f32(f32(6.0) * c))
        c0 = f32(f32(6.0) - f32(f32(2.0) * b))
        k = f32(f32(f32(c3 * a3) + f32(c2 * a2)) + c0)
    elif a < 2.0:
        c3 = f32(f32(-b) - f32(f32(6.0) * c))
        c2 = f32(f32(f32(6.0) * b) + f32(f32(30.0) * c))
        c1 = f32(f32(f32(-12.0) * b) - f32(f32(48.0) * c))
        c0 = f32(f32(f32(8.0) * b) + f32(f32(24.0) * c))
        k = f32(f32(f32(f32(c3 * a3) + f32(c2 * a2)) + f32(c1 * a)) + c0)
    else:
        k = f32(0.0)
    return f32(k / f32(6.0))


def _gaussian(x: f32) -> f32:
    r = f32(0.5)
    norm = f32(f32(1.0) / f32(np.sqrt(f32(f32(2.0) * PI)) * r))
    e = f32(_libm.expf(f32(f32(-f32(x * x)) / f32(f32(2.0) * f32(r * r)))))
    return f32(norm * e)


def _triangle(x: f32) -> f32:
    return f32(f32(1.0) - f32(abs(x))) if abs(x) < 1.0 else f32(0.0)


KERNEL = {NEAREST: lambda x: f32(1.0), TRIANGLE: _triangle, CATMULLROM: _catmullrom,
          GAUSSIAN: _gaussian, LANCZOS3: _lanczos3}


def windows(filt: int, n_in: int, n_out: int):
    """[(left, weights float32[n])] for every output index of one pass."""
    kern, support = KERNEL[filt], SUPPORT[filt]
    ratio = f32(f32(n_in) / f32(n_out))
    sratio = f32(1.0) if ratio < 1.0 else ratio
    ssup = f32(support * sratio)
    out = []
    for o in range(n_out):
        c = f32(f32(f32(o) + f32(0.5)) * ratio)
        left = int(np.floor(f32(c - ssup)))
        left = min(max(left, 0), n_in - 1)
        right = int(np.ceil(f32(c + ssup)))
        right = min(max(right, left + 1), n_in)
        c = f32(c - f32(0.5))
        ws = np.array([kern(f32(f32(f32(i) - c) / sratio)) for i in range(left, right)], f32)
        s = f32(0.0)
        for w in ws:
            s = f32(s + w)
        ws = (ws / s).astype(f32)
        out.append((left, ws))
    return out


def _pass(data: np.ndarray, filt: int, n_out: int) -> np.ndarray:
    """Resample axis 0 of a float32 array (n_in, ...) -> (n_out, ...), mul-then-add, ascending taps."""
    n_in = data.shape[0]
    res = np.empty((n_out,) + data.shape[1:], f32)
    for o, (left, ws) in enumerate(windows(filt, n_in, n_out)):
        acc = np.zeros(data.shape[1:], f32)
        for i, w in enumerate(ws):
            acc = (acc + (data[left + i] * w).astype(f32)).astype(f32)
        res[o] = acc
    return res


def resize_exact(src: np.ndarray, dw: int, dh: int, filt: int = LANCZOS3) -> np.ndarray:
    """src HxWxC (or HxW) u8/u16 -> dh x dw x C, vertical pass first, f32 intermediate."""
    squeeze = src.ndim == 2
    s = src[:, :, None] if squeeze else src
    sh, sw, _ = s.shape
    if sh == 0 or sw == 0:
        out = np.zeros((dh, dw, s.shape[2]), s.dtype)
    elif (sw, sh) == (dw, dh):
        out = s.copy()
    else:
        maxv = f32(np.iinfo(s.dtype).max)
        tmp = _pass(s.astype(f32), filt, dh)                        # vertical_sample
        hor = _pass(np.ascontiguousarray(tmp.transpose(1, 0, 2)), filt, dw)  # horizontal_sample
        hor = np.clip(hor.transpose(1, 0, 2), f32(0.0), maxv)
        # f32::round = half away from zero (values are >= 0 after the clamp)
        out = np.floor(hor.astype(np.float64) + 0.5).astype(s.dtype)
    return out[:, :, 0] if squeeze else out
