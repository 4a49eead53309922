"""imagekit_cuda -- Python host side of the B200-native resize_image drop-in.

Layout mirrors the reference crate's public surface for the hot path:
  imagekit_cuda.transform   decode_image / resize_image / encode_image   (src/transform.rs)
  imagekit_cuda.ImageFormat                                              (src/config.rs:10-17)
  imagekit_cuda.engine      Context / PreparedBatch over the C ABI       (include/imagekit_cuda.h)
"""
from . import _lib
from ._lib import (FILTER_CATMULLROM, FILTER_GAUSSIAN, FILTER_LANCZOS3, FILTER_NEAREST, FILTER_TRIANGLE,
                   MODE_EXACT, MODE_FAST, MODE_FAST_F16, MODE_FAST_FP32)
from .engine import (Context, ImageKitError, PinnedArray, PreparedBatch, default_context, pass_info, pass_table,
                     target_dims)
from .transform import DEFAULT_QUALITY, DynamicImage, ImageFormat, decode_image, encode_image, resize_image

__all__ = [
    "Context", "ImageKitError", "PinnedArray", "PreparedBatch", "default_context", "target_dims", "pass_table",
    "pass_info",
    "DynamicImage", "ImageFormat", "decode_image", "encode_image", "resize_image", "DEFAULT_QUALITY",
    "FILTER_NEAREST", "FILTER_TRIANGLE", "FILTER_CATMULLROM", "FILTER_GAUSSIAN", "FILTER_LANCZOS3",
    "MODE_FAST", "MODE_EXACT", "MODE_FAST_FP32", "MODE_FAST_F16",
]
