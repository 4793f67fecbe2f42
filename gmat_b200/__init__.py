"""gmat_b200 -- B200-native pixel-transform engine behind GMAT's SwsContext / AVFilter surfaces.

Python is only the test / bench / multi-GPU harness: everything here is a thin
ctypes binding over the C ABI of ``include/gmat_b200.h`` (``libgmat_b200.so``,
hand-written sm_100a CUDA).  There is no CPU or PyTorch fallback: if the shared
library is missing or no CUDA device is usable, calls raise.

The Python-facing surface mirrors the reference's own caller
(``metrans/python/swscale.py`` + ``metrans/app/CSwscale.c`` in NVIDIA/GMAT):
``sws_getContext(..., SWS_HWACCEL_CUDA)`` -> ``sws_scale`` -> ``sws_freeContext``.
"""
from .lib import (  # noqa: F401
    GmatbImage, lib, lib_path, check, GmatbError,
    FMT, SPC, SWS, INTERP, BORDER,
)
from .image import FrameBatch, plane_layout, lcg_bytes  # noqa: F401
from .sws import SwsContext, sws_getContext, sws_scale, sws_freeContext  # noqa: F401
from .ops import (  # noqa: F401
    yuv2rgb, rgb2yuv, yuv2yuv, rgb24tobgr24, yuv2rgb_planar_f32,
    crop, flip, rotate, gaussian, median, csc_matrix_yuv2rgb, csc_matrix_rgb2yuv,
    format_nv12_to_rgbpf32, format_rgbpf32_to_nv12,
)

__all__ = [n for n in dir() if not n.startswith("_")]
