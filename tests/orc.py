"""ctypes binding of oracle/libgmat_oracle.so -- the CPU restatement (test infrastructure)."""
import ctypes as C
import os
import subprocess

import numpy as np

from gmat_b200 import GmatbImage

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_L = None


def orc():
    global _L
    if _L is not None:
        return _L
    so = os.path.join(ROOT, "oracle", "libgmat_oracle.so")
    src = os.path.join(ROOT, "oracle", "gmat_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "libgmat_oracle.so"], stdout=subprocess.DEVNULL)
    L = C.CDLL(so)
    IP, fp, ip, ci, cf, cd = C.POINTER(GmatbImage), C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_int, C.c_float, C.c_double
    u8 = C.c_void_p
    L.orc_matrix_yuv2rgb.argtypes = [ci, fp]
    L.orc_matrix_rgb2yuv.argtypes = [ci, fp]
    L.orc_yuv2rgb.argtypes = [IP, IP, fp]
    L.orc_yuv2rgb_planar_f32.argtypes = [IP, IP, fp, cf, fp]
    L.orc_rgb2yuv.argtypes = [IP, IP, fp]
    L.orc_yuv2yuv.argtypes = [IP, IP]
    L.orc_rgb24tobgr24.argtypes = [IP, IP]
    L.orc_filter_table.argtypes = [ci, ci, ci, cf, fp, ip]
    L.orc_resample_packed.argtypes = [u8, ci, ci, ci, u8, ci, ci, ci, ci, ci, fp, ip, fp, ip, ci, ci]
    L.orc_yuv2rgb_scale.argtypes = [IP, IP, fp, fp, ip, fp, ip, ci, ci]
    L.orc_check_norm.argtypes = [ci]
    L.orc_crop.argtypes = [IP, IP, ci, ci]
    L.orc_flip.argtypes = [IP, IP, ci]
    L.orc_rotate.argtypes = [IP, IP, cd, cd, cd, ci]
    L.orc_gaussian.argtypes = [IP, IP, ci, ci, cd, cd, ci]
    L.orc_median.argtypes = [IP, IP, ci, ci]
    L.orc_format_colorspace.argtypes = [ci]
    L.orc_format_rgbpf32_to_nv12.argtypes = [IP, IP, fp]
    _L = L
    return L


def fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def iptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def matrix_yuv2rgb(cs=0):
    m = np.zeros(9, np.float32); orc().orc_matrix_yuv2rgb(cs, fptr(m)); return m


def matrix_rgb2yuv(cs=0):
    m = np.zeros(9, np.float32); orc().orc_matrix_rgb2yuv(cs, fptr(m)); return m


def filter_table(algo, src_n, dst_n, A=0.0):
    co = np.zeros((dst_n, 4), np.float32); po = np.zeros(dst_n, np.int32)
    orc().orc_filter_table(algo, src_n, dst_n, A, fptr(co), iptr(po))
    return co, po


ALGO = {"bicubic": 0, "lanczos": 1, "bilinear": 2, "nearest": 3}


def yuv2rgb(src, dst, cs=0):
    m = matrix_yuv2rgb(cs); s, d = src.image(), dst.image()
    orc().orc_yuv2rgb(C.byref(s), C.byref(d), fptr(m))


def rgb2yuv(src, dst, cs=0):
    m = matrix_rgb2yuv(cs); s, d = src.image(), dst.image()
    orc().orc_rgb2yuv(C.byref(s), C.byref(d), fptr(m))


def yuv2rgb_scale(src, dst, tables, cs=0, ra=0, wrap=0):
    (cx, px), (cy, py) = tables
    m = matrix_yuv2rgb(cs); s, d = src.image(), dst.image()
    orc().orc_yuv2rgb_scale(C.byref(s), C.byref(d), fptr(m), fptr(cx), iptr(px), fptr(cy), iptr(py), ra, wrap)


def format_nv12_to_rgbpf32(src, dst, av_cs=2, norm=255.0, shift=None):
    """format_cuda's NV12 -> planar float: libgpuscale's planar chain with format_cuda's matrix selection"""
    m = matrix_yuv2rgb(orc().orc_format_colorspace(av_cs)); s, d = src.image(), dst.image()
    sh = None if shift is None else fptr(np.asarray(shift, np.float32))
    orc().orc_yuv2rgb_planar_f32(C.byref(s), C.byref(d), fptr(m), norm, sh)


def format_rgbpf32_to_nv12(src, dst, av_cs=2):
    m = matrix_rgb2yuv(orc().orc_format_colorspace(av_cs)); s, d = src.image(), dst.image()
    assert orc().orc_format_rgbpf32_to_nv12(C.byref(s), C.byref(d), fptr(m)) == 0
