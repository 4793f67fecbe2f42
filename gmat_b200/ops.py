"""Unscaled converters and filter kernels (host-side mirror of the C ABI)."""
import ctypes as C

import numpy as np

from .image import FrameBatch
from .lib import BORDER, INTERP, SPC, check, lib


def _img(x):
    return x.image() if isinstance(x, FrameBatch) else x


def _st(stream):
    return C.c_void_p(stream or 0)


def csc_matrix_yuv2rgb(colorspace=SPC.DEFAULT):
    m = np.zeros(9, np.float32)
    lib().gmatb_csc_matrix_yuv2rgb(colorspace, m.ctypes.data_as(C.POINTER(C.c_float)))
    return m


def csc_matrix_rgb2yuv(colorspace=SPC.DEFAULT):
    m = np.zeros(9, np.float32)
    lib().gmatb_csc_matrix_rgb2yuv(colorspace, m.ctypes.data_as(C.POINTER(C.c_float)))
    return m


def yuv2rgb(src, dst, colorspace=SPC.DEFAULT, stream=None):
    s, d = _img(src), _img(dst)
    check(lib().gmatb_yuv2rgb(C.byref(s), C.byref(d), colorspace, _st(stream)), "yuv2rgb")


def yuv2rgb_planar_f32(src, dst, colorspace=SPC.DEFAULT, norm=255.0, shift=(0.0, 0.0, 0.0), stream=None):
    s, d = _img(src), _img(dst)
    sh = (C.c_float * 3)(*shift)
    check(lib().gmatb_yuv2rgb_planar_f32(C.byref(s), C.byref(d), colorspace, norm, sh, _st(stream)), "yuv2rgb_planar_f32")


def format_nv12_to_rgbpf32(src, dst, av_colorspace=2, norm=255.0, shift=None, bgr_planes=False, stream=None):
    """format_cuda NV12 -> RGBPF32 (vf_format_cuda.c:193, format_cuda_kernel.cu:583-609); av_colorspace is the
    frame's AVColorSpace (2 = unspecified -> the filter's BT.709 default)"""
    s, d = _img(src), _img(dst)
    sh = None if shift is None else (C.c_float * 3)(*shift)
    check(lib().gmatb_format_nv12_to_rgbpf32(C.byref(s), C.byref(d), av_colorspace, norm, sh, int(bgr_planes), _st(stream)),
          "format_nv12_to_rgbpf32")


def format_rgbpf32_to_nv12(src, dst, av_colorspace=2, stream=None):
    """format_cuda RGBPF32 -> NV12 (vf_format_cuda.c:198, format_cuda_kernel.cu:625-631)"""
    s, d = _img(src), _img(dst)
    check(lib().gmatb_format_rgbpf32_to_nv12(C.byref(s), C.byref(d), av_colorspace, _st(stream)), "format_rgbpf32_to_nv12")


def rgb2yuv(src, dst, colorspace=SPC.DEFAULT, stream=None):
    s, d = _img(src), _img(dst)
    check(lib().gmatb_rgb2yuv(C.byref(s), C.byref(d), colorspace, _st(stream)), "rgb2yuv")


def yuv2yuv(src, dst, stream=None):
    s, d = _img(src), _img(dst)
    check(lib().gmatb_yuv2yuv(C.byref(s), C.byref(d), _st(stream)), "yuv2yuv")


def rgb24tobgr24(src, dst, stream=None):
    s, d = _img(src), _img(dst)
    check(lib().gmatb_rgb24tobgr24(C.byref(s), C.byref(d), _st(stream)), "rgb24tobgr24")


# ---- filters: option names follow the reference's AVOption tables -------------------------
def crop(src, dst, x=-1, y=-1, stream=None):
    """crop_cuda: x/y = -1 centre the window like vf_crop_nvcv.c:149-150"""
    s, d = _img(src), _img(dst)
    if x < 0:
        x = (s.width - d.width) // 2
    if y < 0:
        y = (s.height - d.height) // 2
    check(lib().gmatb_crop(C.byref(s), C.byref(d), x, y, _st(stream)), "crop")


def flip(src, dst, code=0, stream=None):
    s, d = _img(src), _img(dst)
    check(lib().gmatb_flip(C.byref(s), C.byref(d), code, _st(stream)), "flip")


def rotate(src, dst, angle=0.0, shift_x=0.0, shift_y=0.0, interp="linear", stream=None):
    code = {"nearest": INTERP.NEAREST, "linear": INTERP.LINEAR, "cubic": INTERP.CUBIC, "area": INTERP.AREA}[interp] \
        if isinstance(interp, str) else interp
    s, d = _img(src), _img(dst)
    check(lib().gmatb_rotate(C.byref(s), C.byref(d), angle, shift_x, shift_y, code, _st(stream)), "rotate")


def gaussian(src, dst, kw=3, kh=3, sigmaX=0.0, sigmaY=0.0, border_type=BORDER.CONSTANT, stream=None):
    s, d = _img(src), _img(dst)
    check(lib().gmatb_gaussian(C.byref(s), C.byref(d), kw, kh, sigmaX, sigmaY, border_type, _st(stream)), "gaussian")


def median(src, dst, kw=3, kh=3, stream=None):
    s, d = _img(src), _img(dst)
    check(lib().gmatb_median(C.byref(s), C.byref(d), kw, kh, _st(stream)), "median")
