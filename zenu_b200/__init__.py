"""zenu_b200 — B200 (sm_100a) backend for ZeNu's CNN-training hot path.

The product is libzenu_b200.so (zenu_b200/csrc, C ABI in include/); this package is its host-side
binding: `ops` mirrors the reference's operator traits, `nn` the layer / autograd / optimizer surface.
Importing the package does not need a GPU; calling any op does, and fails loudly without one.
"""
from ._lib import (ZB_F32, ZB_F64, ZB_MATH_DEFAULT, ZB_MATH_FP32, ZB_MATH_TF32, ZB_MATH_TF32X3, ZB_NCHW, ZB_NHWC,  # noqa: F401
                   ZenuB200Error)

__all__ = ["ZB_F32", "ZB_F64", "ZB_MATH_DEFAULT", "ZB_MATH_FP32", "ZB_MATH_TF32", "ZB_MATH_TF32X3", "ZB_NCHW", "ZB_NHWC", "ZenuB200Error"]
