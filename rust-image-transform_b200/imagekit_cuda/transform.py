"""Host-side mirror of the reference's transform module (/root/reference/src/transform.rs).

Same three free functions, same argument meaning and error behaviour:
    decode_image(bytes)            -> (DynamicImage, Optional[ImageFormat])   transform.rs:27-43
    resize_image(img, w, h)        -> DynamicImage                            transform.rs:62-90  (GPU)
    encode_image(img, fmt, quality)-> bytes                                   transform.rs:113-150
Only resize_image is the hot path; it runs on the B200 through the C ABI (no CPU fallback).
decode/encode stay on the CPU (Pillow here stands in for the reference's image/webp/ravif crates)
and exist so that the /upload-shaped end-to-end pipeline (BASELINE config 5) can be driven from
Python with the same call sequence the Rust handlers use (src/lib.rs:281-294).
"""
from __future__ import annotations

import enum
import io

import numpy as np

from . import _lib
from .engine import ImageKitError, default_context, target_dims

DEFAULT_QUALITY = 80  # src/config.rs:31


class ImageFormat(enum.Enum):
    """src/config.rs:10-17 (lowercase variants, same spelling as the serde names)."""
    jpeg = "jpeg"
    webp = "webp"
    avif = "avif"

    def __str__(self) -> str:
        return self.value


_VARIANTS = {  # (channels, dtype) -> image::DynamicImage variant name
    (1, "uint8"): "ImageLuma8", (2, "uint8"): "ImageLumaA8", (3, "uint8"): "ImageRgb8", (4, "uint8"): "ImageRgba8",
    (1, "uint16"): "ImageLuma16", (2, "uint16"): "ImageLumaA16", (3, "uint16"): "ImageRgb16",
    (4, "uint16"): "ImageRgba16",
}


class DynamicImage:
    """image::DynamicImage: a tagged raster.  Pixels are a tight row-major HxWxC numpy array."""

    def __init__(self, pixels: np.ndarray):
        if pixels.ndim == 2:
            pixels = pixels[:, :, None]
        key = (pixels.shape[2] if pixels.ndim == 3 else 0, str(pixels.dtype))
        if pixels.ndim != 3 or key not in _VARIANTS:
            raise ImageKitError(_lib.ERR_UNSUPPORTED, f"unsupported pixel format {pixels.shape} {pixels.dtype}")
        self.pixels = np.ascontiguousarray(pixels)
        self.variant = _VARIANTS[key]

    @staticmethod
    def new_rgb8(w: int, h: int) -> "DynamicImage":
        return DynamicImage(np.zeros((h, w, 3), np.uint8))

    @staticmethod
    def new_rgba8(w: int, h: int) -> "DynamicImage":
        return DynamicImage(np.zeros((h, w, 4), np.uint8))

    @staticmethod
    def new_luma8(w: int, h: int) -> "DynamicImage":
        return DynamicImage(np.zeros((h, w, 1), np.uint8))

    def dimensions(self):
        return (self.pixels.shape[1], self.pixels.shape[0])

    def width(self) -> int:
        return self.pixels.shape[1]

    def height(self) -> int:
        return self.pixels.shape[0]

    def clone(self) -> "DynamicImage":
        return DynamicImage(self.pixels.copy())

    def _to8(self) -> np.ndarray:
        p = self.pixels
        if p.dtype == np.uint16:  # image's u16 -> u8 conversion: (x + 128) / 257
            p = ((p.astype(np.uint32) + 128) // 257).astype(np.uint8)
        return p

    def to_rgb8(self) -> np.ndarray:
        p = self._to8()
        c = p.shape[2]
        if c >= 3:
            return np.ascontiguousarray(p[:, :, :3])
        return np.repeat(p[:, :, :1], 3, axis=2)

    def to_rgba8(self) -> np.ndarray:
        p = self._to8()
        c = p.shape[2]
        if c == 4:
            return p
        if c == 3:
            return np.concatenate([p, np.full(p.shape[:2] + (1,), 255, np.uint8)], axis=2)
        alpha = p[:, :, 1:2] if c == 2 else np.full(p.shape[:2] + (1,), 255, np.uint8)
        return np.concatenate([np.repeat(p[:, :, :1], 3, axis=2), alpha], axis=2)


def decode_image(data: bytes):
    """transform.rs:27-43: guess the container from magic bytes, decode, map the format."""
    from PIL import Image, UnidentifiedImageError
    try:
        im = Image.open(io.BytesIO(data))
        fmt_name = (im.format or "").upper()
        im.load()
    except (UnidentifiedImageError, OSError, ValueError, SyntaxError) as e:
        raise ImageKitError(_lib.ERR_INVALID_ARG, str(e)) from None
    if im.mode in ("P", "CMYK", "YCbCr", "1"):
        im = im.convert("RGBA" if "transparency" in im.info else "RGB")
    if im.mode in ("I;16", "I;16L", "I;16B", "I"):
        arr = np.asarray(im).astype(np.uint16)
    elif im.mode in ("L", "LA", "RGB", "RGBA"):
        arr = np.asarray(im)
    else:
        arr = np.asarray(im.convert("RGB"))
    fmt = {"WEBP": ImageFormat.webp, "JPEG": ImageFormat.jpeg, "AVIF": ImageFormat.avif}.get(fmt_name)
    return DynamicImage(arr), fmt


def resize_image(img: DynamicImage, w: int | None, h: int | None, ctx=None,
                 encode_as: ImageFormat | None = None) -> DynamicImage:
    """transform.rs:62-90.  (None, None) returns the image untouched; otherwise the target size
    follows the reference's f32 rule, DynamicImage::resize fits it within (aspect preserved) and the
    raster is resampled with Lanczos3 -- on the GPU, through ikc_resize_u8 / ikc_resize_u16.

    encode_as (optional, not in the reference's signature): the format the caller is about to pass to
    encode_image.  The resize then stores what encode_image would convert to anyway -- to_rgb8() for
    jpeg/webp, to_rgba8() for avif (transform.rs:123,131,140) -- so the CPU conversion pass disappears
    and an RGBA source going to jpeg/webp sends 25 % fewer bytes back over PCIe.  8-bit rasters only;
    the result encodes to the same bytes as the unfused path."""
    if w is None and h is None:
        return img
    for v in (w, h):
        if v is not None and not (0 <= int(v) <= 0xFFFFFFFF):
            raise ImageKitError(_lib.ERR_INVALID_ARG, "w/h must fit in u32")
    ow, oh = img.dimensions()
    tw, th, code = target_dims(ow, oh, w, h)
    if code != _lib.DIMS_RESAMPLE:
        return img.clone()
    ctx = ctx or default_context()
    oc = None
    if encode_as is not None and img.pixels.dtype == np.uint8:
        oc = 4 if encode_as == ImageFormat.avif else 3
        if oc == img.pixels.shape[2]:
            oc = None
    out = ctx.resize(img.pixels, tw, th, _lib.FILTER_LANCZOS3, out_channels=oc)
    return DynamicImage(out)


def encode_image(img: DynamicImage, fmt: ImageFormat, quality: int) -> bytes:
    """transform.rs:113-150: jpeg/webp from to_rgb8(), avif from to_rgba8(); quality clamped 1..=100."""
    from PIL import Image
    q = max(1, min(100, int(quality)))
    out = io.BytesIO()
    try:
        if fmt == ImageFormat.jpeg:
            Image.fromarray(img.to_rgb8(), "RGB").save(out, "JPEG", quality=q)
        elif fmt == ImageFormat.webp:
            Image.fromarray(img.to_rgb8(), "RGB").save(out, "WEBP", quality=q, method=4)
        elif fmt == ImageFormat.avif:
            Image.fromarray(img.to_rgba8(), "RGBA").save(out, "AVIF", quality=q, speed=4)
        else:
            raise ImageKitError(_lib.ERR_INVALID_ARG, f"unknown format {fmt}")
    except (OSError, KeyError, ValueError) as e:
        raise ImageKitError(_lib.ERR_UNSUPPORTED, str(e)) from None
    return out.getvalue()
