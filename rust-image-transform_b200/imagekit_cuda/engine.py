"""Thin object layer over the C ABI: context, pinned host buffers, prepared device batches."""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from . import _lib
from ._lib import Job


class ImageKitError(Exception):
    """Mirror of ImageKitError::TransformError(String) (/root/reference/src/lib.rs:38-39)."""

    def __init__(self, status: int, message: str):
        self.status = status
        name = _lib.STATUS_NAMES[status] if 0 <= status < len(_lib.STATUS_NAMES) else str(status)
        super().__init__(f"Transformation error: {message} [{name}]")


def _check(rc: int) -> None:
    if rc != _lib.OK:
        raise ImageKitError(rc, _lib.last_error())


def target_dims(ow: int, oh: int, w: int | None, h: int | None):
    """(tw, th, code) -- the dims rule of resize_image (transform.rs:62-90 + image 0.25.8)."""
    tw, th = C.c_uint32(), C.c_uint32()
    code = _lib.load().ikc_target_dims(ow, oh, w is not None, w or 0, h is not None, h or 0,
                                       C.byref(tw), C.byref(th))
    if code < 0:
        raise ImageKitError(-code, _lib.last_error())
    return tw.value, th.value, code


def check_dims(sw: int, sh: int, dw: int, dh: int) -> None:
    """The library's size guard (ikc_check_dims): raises too-large before anything is allocated."""
    _check(_lib.load().ikc_check_dims(sw, sh, dw, dh))


def pass_table(filt: int, n_in: int, n_out: int):
    """(left, count, weights[n_out, stride]) of one pass, as the kernels consume it."""
    L = _lib.load()
    stride = L.ikc_pass_table(filt, n_in, n_out, None, None, None, 0)
    if stride == 0:
        raise ImageKitError(_lib.ERR_INVALID_ARG, "cannot plan pass")
    left = np.zeros(n_out, np.uint32)
    cnt = np.zeros(n_out, np.uint32)
    w = np.zeros((n_out, stride), np.float32)
    pu32 = C.POINTER(C.c_uint32)
    L.ikc_pass_table(filt, n_in, n_out, left.ctypes.data_as(pu32), cnt.ctypes.data_as(pu32),
                     w.ctypes.data_as(C.POINTER(C.c_float)), stride)
    return left, cnt, w


def pass_band(filt: int, n_in: int, n_out: int):
    """(band_n, gbase[n_chunks + 1], hi[n_chunks, band_n, 16], lo[...]) of a downscale pass: the f16 weight tiles of
    the tensor-core vertical pass, un-laid-out to [chunk][output][index]; None if the pass has no band form."""
    L = _lib.load()
    bn = C.c_uint32()
    chunks = L.ikc_pass_band(filt, n_in, n_out, C.byref(bn), None, None, 0)
    if chunks == 0:
        return None
    n = bn.value
    gbase = np.zeros(chunks + 1, np.int32)
    raw = np.zeros(chunks * 2 * n * 16, np.uint16)
    got = L.ikc_pass_band(filt, n_in, n_out, C.byref(bn), gbase.ctypes.data_as(C.POINTER(C.c_int32)),
                          raw.ctypes.data_as(C.POINTER(C.c_uint16)), raw.size)
    assert got == chunks
    t = raw.view(np.float16).reshape(chunks, 2, 2, n // 8, 8, 8)   # [chunk][hi/lo][k / 8][n / 8][n % 8][k % 8]
    t = t.transpose(0, 1, 3, 4, 2, 5).reshape(chunks, 2, n, 16)    # [chunk][hi/lo][n][k]
    return n, gbase, t[:, 0], t[:, 1]


def pass_band8(filt: int, n_in: int, n_out: int):
    """(limbs, shift, gbase[n_chunks + 1], digits[n_chunks, limbs, 32 window positions, 32 indices]) of a downscale pass:
    the s8 weight tiles of the integer tensor-core vertical pass, un-laid-out (window position p of chunk k is output
    8 * gbase[k] + p); None if the pass has no such form."""
    L = _lib.load()
    limbs, shift = C.c_uint32(), C.c_uint32()
    chunks = L.ikc_pass_band8(filt, n_in, n_out, C.byref(limbs), C.byref(shift), None, None, 0)
    if chunks == 0:
        return None
    nl = limbs.value
    gbase = np.zeros(chunks + 1, np.int32)
    raw = np.zeros(chunks * nl * 32 * 32, np.int8)
    got = L.ikc_pass_band8(filt, n_in, n_out, C.byref(limbs), C.byref(shift), gbase.ctypes.data_as(C.POINTER(C.c_int32)),
                           raw.ctypes.data_as(C.POINTER(C.c_int8)), raw.size)
    assert got == chunks
    t = raw.reshape(chunks, 2, nl * 4, 8, 16)                      # [chunk][k / 16][n / 8][n % 8][k % 16]
    t = t.transpose(0, 2, 3, 1, 4).reshape(chunks, 32, nl, 32)     # [chunk][window position][digit][k]
    t = t.transpose(0, 2, 1, 3)                                    # [chunk][digit][window position][k]
    return nl, shift.value, gbase, t


def pass_band8t(filt: int, n_in: int, n_out: int):
    """(chunks, k_lo[n_bands], digits[n_bands, chunks, 2, 128 rows, 32 indices], rows) of a pass: the s8 weight tiles of the
    row-band integer tensor-core kernels, un-laid-out (row m of band r is output rows * r + m, index k of chunk c is source
    index k_lo[r] + 32 c + k); None if the pass has no such form."""
    L = _lib.load()
    chunks, rows = C.c_uint32(), C.c_uint32()
    bands = L.ikc_pass_band8t(filt, n_in, n_out, C.byref(chunks), C.byref(rows), None, None, 0)
    if bands == 0:
        return None
    nc = chunks.value
    k_lo = np.zeros(bands, np.int32)
    raw = np.zeros(bands * nc * 2 * 128 * 32, np.int8)
    got = L.ikc_pass_band8t(filt, n_in, n_out, C.byref(chunks), C.byref(rows), k_lo.ctypes.data_as(C.POINTER(C.c_int32)),
                            raw.ctypes.data_as(C.POINTER(C.c_int8)), raw.size)
    assert got == bands
    t = raw.reshape(bands, nc, 2, 2, 16, 8, 16)                     # [band][chunk][digit][k / 16][m / 8][m % 8][k % 16]
    t = t.transpose(0, 1, 2, 4, 5, 3, 6).reshape(bands, nc, 2, 128, 32)
    return nc, k_lo, t, rows.value


def pass_info(filt: int, n_in: int, n_out: int) -> dict:
    """What the planner derived for one pass (ring size, uniform stretch, 2x-upscale frame): ikc_pass_info."""
    info = _lib.PassInfo()
    _check(_lib.load().ikc_pass_info(filt, n_in, n_out, C.byref(info)))
    return {name: getattr(info, name) for name, _ in _lib.PassInfo._fields_}


class PinnedArray:
    """A numpy view over cudaMallocHost memory obtained through ikc_host_alloc."""

    def __init__(self, shape, dtype=np.uint8):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        n = int(np.prod(self.shape)) * self.dtype.itemsize
        p = C.c_void_p()
        _check(_lib.load().ikc_host_alloc(max(n, 1), C.byref(p)))
        self.ptr = p.value
        buf = (C.c_uint8 * max(n, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def free(self):
        if self.ptr:
            self.array = None
            _lib.load().ikc_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Ticket:
    """A resize in flight (Context.resize_begin).  end() blocks until the result is complete and returns it."""

    def __init__(self, handle, src, dst):
        self._h, self._src, self._dst = handle, src, dst

    def end(self) -> np.ndarray:
        if self._h is not None:
            h, self._h = self._h, None
            _check(_lib.load().ikc_resize_end(h))
        self._src = None
        return self._dst

    def __del__(self):
        if getattr(self, "_h", None) is not None:
            try:
                _lib.load().ikc_resize_end(self._h)
            except Exception:  # noqa: BLE001
                pass


class Context:
    """ikc_ctx: one per process (the Rust side keeps it in a OnceLock)."""

    def __init__(self, device_ids=None):
        L = _lib.load()
        self._h = C.c_void_p()
        if device_ids:
            arr = (C.c_int * len(device_ids))(*device_ids)
            rc = L.ikc_create(arr, len(device_ids), C.byref(self._h))
        else:
            rc = L.ikc_create(None, 0, C.byref(self._h))
        _check(rc)

    @property
    def handle(self):
        return self._h

    def close(self):
        if self._h:
            _lib.load().ikc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def device_count(self) -> int:
        return _lib.load().ikc_device_count(self._h)

    @property
    def kernel_launches(self) -> int:
        return int(_lib.load().ikc_kernel_launches(self._h))

    def set_mode(self, mode: int) -> None:
        _check(_lib.load().ikc_set_mode(self._h, mode))

    # ---- host-buffer entry points -------------------------------------------------------------
    def resize(self, src: np.ndarray, dw: int, dh: int, filt: int = _lib.FILTER_LANCZOS3,
               out: np.ndarray | None = None, out_channels: int | None = None) -> np.ndarray:
        """imageops::resize(src, dw, dh, filt) on HxWxC (or HxW) u8/u16 host arrays.  out_channels = 3 / 4
        additionally applies to_rgb8() / to_rgba8() to the result inside the kernels' store (u8 only)."""
        if out_channels is not None:
            return self._resize_convert(src, dw, dh, filt, out, out_channels)
        squeeze = src.ndim == 2
        s = src[:, :, None] if squeeze else src
        if s.ndim != 3:
            raise ImageKitError(_lib.ERR_INVALID_ARG, "expected HxWxC or HxW array")
        if not (s.strides[2] == s.itemsize and s.strides[1] == s.itemsize * s.shape[2]):
            s = np.ascontiguousarray(s)
        sh, sw, ch = s.shape
        if out is None:
            check_dims(min(sw, 0xFFFFFFFF), min(sh, 0xFFFFFFFF), min(dw, 0xFFFFFFFF), min(dh, 0xFFFFFFFF))  # before allocating the result
        dst = out if out is not None else np.empty((dh, dw, ch), s.dtype)
        d3 = dst[:, :, None] if dst.ndim == 2 else dst
        assert d3.shape == (dh, dw, ch) and d3.dtype == s.dtype
        L = _lib.load()
        fn = {1: L.ikc_resize_u8, 2: L.ikc_resize_u16}.get(s.itemsize)
        if fn is None:
            raise ImageKitError(_lib.ERR_UNSUPPORTED, f"unsupported sample type {s.dtype}")
        src_pitch = s.strides[0] if sh > 1 else sw * ch * s.itemsize
        dst_pitch = d3.strides[0] if dh > 1 else dw * ch * s.itemsize
        _check(fn(self._h, s.ctypes.data, sw, sh, src_pitch, ch, d3.ctypes.data, dw, dh, dst_pitch, filt))
        if out is not None:
            return out
        return dst[:, :, 0] if squeeze else dst

    def submit(self, src: np.ndarray, dw: int, dh: int, filt: int = _lib.FILTER_LANCZOS3, out_channels: int | None = None) -> np.ndarray:
        """resize() for handler threads that each bring one 8-bit image: concurrent calls are coalesced into shared
        uploads and launches (ikc_submit_u8).  Same result as resize()."""
        s = np.ascontiguousarray(src[:, :, None] if src.ndim == 2 else src)
        if s.ndim != 3 or s.dtype != np.uint8:
            raise ImageKitError(_lib.ERR_UNSUPPORTED, "submit() takes an 8-bit HxW or HxWxC array")
        sh, sw, ch = s.shape
        co = ch if out_channels is None else out_channels
        check_dims(sw, sh, min(dw, 0xFFFFFFFF), min(dh, 0xFFFFFFFF))  # before allocating the result
        dst = np.empty((dh, dw, co), np.uint8)
        _check(_lib.load().ikc_submit_u8(self._h, s.ctypes.data, sw, sh, sw * ch, ch | (co << 8) if co != ch else ch, dst.ctypes.data,
                                         dw, dh, dw * co, filt))
        return dst[:, :, 0] if src.ndim == 2 and co == 1 else dst

    def resize_begin(self, src: np.ndarray, dw: int, dh: int, filt: int = _lib.FILTER_LANCZOS3, out_channels: int | None = None,
                     out: np.ndarray | None = None):
        """First half of resize() for an 8-bit HxWxC array (ikc_resize_begin_u8): queues the copies and kernels and returns
        a ticket; the caller decodes its next upload, then ticket.end() returns the resized array.  `src` (if pinned) and
        the result buffer are kept alive by the ticket."""
        s = np.ascontiguousarray(src[:, :, None] if src.ndim == 2 else src)
        if s.ndim != 3 or s.dtype != np.uint8:
            raise ImageKitError(_lib.ERR_UNSUPPORTED, "resize_begin() takes an 8-bit HxW or HxWxC array")
        sh, sw, ch = s.shape
        co = ch if out_channels is None else out_channels
        check_dims(sw, sh, min(dw, 0xFFFFFFFF), min(dh, 0xFFFFFFFF))
        dst = out if out is not None else np.empty((dh, dw, co), np.uint8)
        assert dst.shape == (dh, dw, co) and dst.dtype == np.uint8 and dst.flags.c_contiguous
        h = C.c_void_p()
        _check(_lib.load().ikc_resize_begin_u8(self._h, s.ctypes.data, sw, sh, sw * ch, ch | (co << 8) if co != ch else ch, dst.ctypes.data,
                                               dw, dh, dw * co, filt, C.byref(h)))
        return Ticket(h, s, dst)

    def stats(self) -> dict:
        """Counters for a /metrics handler (ikc_get_stats)."""
        st = _lib.Stats()
        _check(_lib.load().ikc_get_stats(self._h, C.byref(st)))
        return {n: int(getattr(st, n)) for n, _ in _lib.Stats._fields_}

    def _resize_convert(self, src, dw, dh, filt, out, co):
        s = np.ascontiguousarray(src[:, :, None] if src.ndim == 2 else src)
        if s.ndim != 3 or s.dtype != np.uint8:
            raise ImageKitError(_lib.ERR_UNSUPPORTED, "channel conversion needs an 8-bit HxW or HxWxC array")
        sh, sw, ch = s.shape
        check_dims(sw, sh, min(dw, 0xFFFFFFFF), min(dh, 0xFFFFFFFF))  # before allocating the result
        dst = out if out is not None else np.empty((dh, dw, co), np.uint8)
        assert dst.shape == (dh, dw, co) and dst.dtype == np.uint8 and dst.flags.c_contiguous
        _check(_lib.load().ikc_resize_convert_u8(self._h, s.ctypes.data, sw, sh, sw * ch, ch, dst.ctypes.data, dw, dh,
                                                 dw * co, co, filt))
        return dst

    def resize_image(self, src: np.ndarray, w: int | None, h: int | None) -> np.ndarray:
        """resize_image(img, w, h) for a tight 8-bit raster through ikc_resize_image_u8."""
        squeeze = src.ndim == 2
        s = np.ascontiguousarray(src[:, :, None] if squeeze else src)
        sh, sw, ch = s.shape
        tw, th, _ = target_dims(sw, sh, w, h)
        dst = np.empty((th, tw, ch), np.uint8)
        otw, oth = C.c_uint32(), C.c_uint32()
        code = _lib.load().ikc_resize_image_u8(self._h, s.ctypes.data, sw, sh, ch, w is not None, w or 0,
                                               h is not None, h or 0, dst.ctypes.data, dst.nbytes,
                                               C.byref(otw), C.byref(oth))
        if code < 0:
            raise ImageKitError(-code, _lib.last_error())
        assert (otw.value, oth.value) == (tw, th)
        return dst[:, :, 0] if squeeze else dst

    def resize_batch(self, srcs, sizes, filt: int = _lib.FILTER_LANCZOS3, outs=None):
        """ikc_resize_batch over host arrays; sizes = [(dw, dh)]. Returns (outputs, jobs)."""
        n = len(srcs)
        jobs = (Job * n)()
        keep = []
        results = []
        for i, (s, (dw, dh)) in enumerate(zip(srcs, sizes)):
            s3 = s[:, :, None] if s.ndim == 2 else s
            if not s3.flags.c_contiguous:
                s3 = np.ascontiguousarray(s3)
            sh, sw, ch = s3.shape
            d = outs[i] if outs is not None else np.empty((dh, dw, ch), np.uint8)
            keep.append((s3, d))
            results.append(d)
            jobs[i] = Job(s3.ctypes.data, d.ctypes.data, sw, sh, dw, dh, sw * ch, dw * ch, ch, filt, 0, 0)
        rc = _lib.load().ikc_resize_batch(self._h, jobs, n)
        _check(rc)
        return results, jobs

    # ---- device-resident entry points ----------------------------------------------------------
    def resize_device(self, device_index: int, stream: int, d_src: int, sw: int, sh: int, src_pitch: int,
                      ch: int, d_dst: int, dw: int, dh: int, dst_pitch: int, filt: int = _lib.FILTER_LANCZOS3):
        _check(_lib.load().ikc_resize_u8_device(self._h, device_index, stream, d_src, sw, sh, src_pitch, ch,
                                                d_dst, dw, dh, dst_pitch, filt))

    def prepare_batch(self, device_index: int, jobs) -> "PreparedBatch":
        """jobs: list of (d_src, sw, sh, src_pitch, d_dst, dw, dh, dst_pitch, channels, filter)."""
        n = len(jobs)
        arr = (Job * n)()
        for i, (d_src, sw, sh, sp, d_dst, dw, dh, dp, ch, filt) in enumerate(jobs):
            arr[i] = Job(d_src, d_dst, sw, sh, dw, dh, sp, dp, ch, filt, 0, 0)
        h = C.c_void_p()
        rc = _lib.load().ikc_batch_prepare(self._h, device_index, arr, n, C.byref(h))
        if rc != _lib.OK and h:   # some jobs were rejected: the library still returns a batch over the others; do not leak it
            _lib.load().ikc_batch_free(h)
        _check(rc)
        return PreparedBatch(self, h, arr)


class PreparedBatch:
    def __init__(self, ctx: Context, handle, jobs):
        self._ctx = ctx
        self._h = handle
        self.jobs = jobs

    @property
    def launch_count(self) -> int:
        return _lib.load().ikc_batch_launch_count(self._h)

    def describe(self) -> str:
        buf = C.create_string_buffer(512)
        _check(_lib.load().ikc_batch_describe(self._h, buf, 512))
        return buf.value.decode()

    def launch(self, stream: int = 0) -> None:
        _check(_lib.load().ikc_batch_launch(self._h, stream))

    def free(self):
        if self._h:
            _lib.load().ikc_batch_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


_default = None
_default_lock = threading.Lock()


def default_context() -> Context:
    """Process-wide context (lazy), like the Rust wrapper's OnceLock<Context>."""
    global _default
    with _default_lock:
        if _default is None:
            _default = Context()
        return _default
