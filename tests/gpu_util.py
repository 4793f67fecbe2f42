"""helpers shared by the -m gpu tests"""
import ctypes as C
import os

import numpy as np
import torch

import orc
from gmat_b200 import FMT, FrameBatch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
vp, ci = C.c_void_p, C.c_int


def diff(a, b):
    a = a.payload() if isinstance(a, FrameBatch) else a
    b = b.payload() if isinstance(b, FrameBatch) else b
    assert a.shape == b.shape
    return int((a != b).sum())


def assert_same(a, b, what=""):
    a = a.payload() if isinstance(a, FrameBatch) else a
    b = b.payload() if isinstance(b, FrameBatch) else b
    assert a.shape == b.shape, what
    bad = int((a != b).sum())
    if bad:
        i = int(np.nonzero(a != b)[0][0])
        raise AssertionError(f"{what}: {bad} of {a.size} bytes differ; first at {i}: {int(a[i])} vs {int(b[i])}")


def have_ref():
    return os.path.exists(os.path.join(REF, "libref_gpuscale.so"))


_o1 = None


def o1():
    """the reference's own libgpuscale kernels, compiled unmodified for sm_100a"""
    global _o1
    if _o1 is None:
        L = C.CDLL(os.path.join(REF, "libref_gpuscale.so"))
        for n in ("yuv2rgb_cuda", "rgb2yuv_cuda", "yuv2yuv_cuda"):
            f = getattr(L, n); f.restype = ci
            f.argtypes = [C.POINTER(vp), C.POINTER(ci), C.POINTER(vp), C.POINTER(ci), ci, ci, ci, ci, vp]
        L.rgb24tobgr24_cuda.argtypes = [C.POINTER(vp), C.POINTER(vp), C.POINTER(ci), C.POINTER(ci), ci, ci, vp]
        L.set_mat_yuv2rgb_cuda.argtypes = [ci]; L.set_mat_rgb2yuv_cuda.argtypes = [ci]
        L.ref_p016_to_color64.argtypes = [vp, ci, vp, ci, ci, ci, ci, vp]
        L.ref_nv12_to_bgrpf32.argtypes = [vp, ci, vp, ci, ci, ci, vp]
        _o1 = L
    return _o1


_o2 = None


def o2():
    global _o2
    if _o2 is None:
        L = C.CDLL(os.path.join(REF, "libref_o2_driver.so"))
        L.ref_o2_load.argtypes = [C.c_char_p]
        L.ref_o2_launch.argtypes = [C.c_char_p, ci, C.POINTER(vp), C.POINTER(ci), C.POINTER(ci), C.POINTER(ci),
                                    C.POINTER(ci), C.POINTER(ci), C.POINTER(vp), ci, ci, ci, ci, ci, C.c_float, ci, ci]
        assert L.ref_o2_load(os.path.join(REF, "ref_scale_cuda.cubin").encode()) == 0
        _o2 = L
    return _o2


_o3 = None


def o3():
    """the reference's format_cuda kernels (libavfilter/format_cuda_kernel.cu), compiled unmodified for sm_100a"""
    global _o3
    if _o3 is None:
        L = C.CDLL(os.path.join(REF, "libref_format_cuda.so"))
        pp, pi = C.POINTER(vp), C.POINTER(ci)
        L.nv12_to_rgbpf32.argtypes = [vp, pp, pi, pp, pi, ci, ci, ci]
        L.nv12_to_rgbpf32_shift.argtypes = [vp, pp, pi, pp, pi, ci, ci, C.c_float, C.POINTER(C.c_float), ci]
        L.nv12_to_bgrpf32_shift.argtypes = [vp, pp, pi, pp, pi, ci, ci, C.c_float, C.POINTER(C.c_float), ci]
        L.rgbpf32_to_nv12.argtypes = [vp, pp, pi, pp, pi, ci, ci, ci]
        _o3 = L
    return _o3


def o3_run(kind, src, dst, av_cs=2, norm=255.0, shift=None):
    L = o3()
    sp, ss = arrs(src.image()); dp, ds = arrs(dst.image())
    if kind == "nv12_to_rgbpf32":
        L.nv12_to_rgbpf32(None, sp, ss, dp, ds, src.w, src.h, av_cs)
    elif kind in ("nv12_to_rgbpf32_shift", "nv12_to_bgrpf32_shift"):
        getattr(L, kind)(None, sp, ss, dp, ds, src.w, src.h, norm, (C.c_float * 3)(*shift), av_cs)
    else:
        L.rgbpf32_to_nv12(None, sp, ss, dp, ds, src.w, src.h, av_cs)
    torch.cuda.synchronize()


def arrs(img):
    return (vp * 4)(*[img.data[i] for i in range(4)]), (ci * 4)(*[img.linesize[i] for i in range(4)])


def o1_run(kind, src, dst, cs=0):
    L = o1()
    L.set_mat_yuv2rgb_cuda(cs); L.set_mat_rgb2yuv_cuda(cs)
    sp, ss = arrs(src.image()); dp, ds = arrs(dst.image())
    if kind == "swap":
        L.rgb24tobgr24_cuda(sp, dp, ss, ds, src.w, src.h, None)
    else:
        getattr(L, kind)(sp, ss, dp, ds, src.w, src.h, src.fmt, dst.fmt, None)
    torch.cuda.synchronize()


def o2_packed(func, src, dst, ch, depth, param=999999.0, linear=0, integer=0, plane=0, sw=None, sh=None, dw=None, dh=None):
    """run a reference Subsample_* kernel on plane `plane` of src -> dst"""
    L = o2()
    si, di = src.image(), dst.image()
    sw = sw or src.w; sh = sh or src.h; dw = dw or dst.w; dh = dh or dst.h
    pad = lambda xs: list(xs) + [0] * (4 - len(xs))
    rc = L.ref_o2_launch(func.encode(), 1, (vp * 4)(*pad([si.data[plane]])), (ci * 4)(*pad([si.linesize[plane]])),
                         (ci * 4)(*pad([sw])), (ci * 4)(*pad([sh])), (ci * 4)(*pad([depth])), (ci * 4)(*pad([ch])),
                         (vp * 4)(*pad([di.data[plane]])), dw, dh, di.linesize[plane], sw, sh, C.c_float(param), linear, integer)
    assert rc == 0, (func, rc)
    torch.cuda.synchronize()


_o2c = None


def o2_coeffs(lanczos, fx, param=999999.0):
    """the reference's own lanczos_coeffs / bicubic_coeffs (vf_scale_cuda.cu:948-981) at the positions fx"""
    global _o2c
    if _o2c is None:
        _o2c = C.CDLL(os.path.join(REF, "libref_o2_coeffs.so"))
        _o2c.ref_o2_coeffs.argtypes = [ci, C.POINTER(C.c_float), C.c_float, C.POINTER(C.c_float), ci]
    fx = np.ascontiguousarray(fx, np.float32)
    out = np.zeros((len(fx), 4), np.float32)
    rc = _o2c.ref_o2_coeffs(int(lanczos), fx.ctypes.data_as(C.POINTER(C.c_float)), param,
                            out.ctypes.data_as(C.POINTER(C.c_float)), len(fx))
    assert rc == 0, rc
    return out


_O2_FMT = {FMT.NV12: ("nv12", 8, (1, 2)), FMT.YUV420P: ("yuv420p", 8, (1, 1, 1)),
           FMT.P010LE: ("p010le", 16, (1, 2)), FMT.P016LE: ("p016le", 16, (1, 2))}


def o2_frame(algo, src, dst, param=999999.0):
    """one frame through the reference's scale_cuda kernels exactly as scalecuda_resize does it
    (vf_scale_cuda.c:430-500): a texture per input plane, the luma launch, then the _uv launch on the chroma grid"""
    L = o2()
    name, depth, chans = _O2_FMT[src.fmt]
    assert dst.fmt == src.fmt
    si, di = src.image(), dst.image()
    npl = len(chans)
    cw, chh = (src.w + 1) // 2, (src.h + 1) // 2
    dcw, dch = (dst.w + 1) // 2, (dst.h + 1) // 2
    pad = lambda xs: list(xs) + [0] * (4 - len(xs))
    args = ((vp * 4)(*pad([si.data[i] for i in range(npl)])), (ci * 4)(*pad([si.linesize[i] for i in range(npl)])),
            (ci * 4)(*pad([src.w] + [cw] * (npl - 1))), (ci * 4)(*pad([src.h] + [chh] * (npl - 1))),
            (ci * 4)(*pad([depth] * npl)), (ci * 4)(*pad(chans)), (vp * 4)(*pad([di.data[i] for i in range(npl)])))
    fn = f"Subsample_{algo}_{name}_{name}"
    rc = L.ref_o2_launch(fn.encode(), npl, *args, dst.w, dst.h, di.linesize[0], src.w, src.h, C.c_float(param), 0, 0)
    assert rc == 0, (fn, rc)
    rc = L.ref_o2_launch((fn + "_uv").encode(), npl, *args, dcw, dch, di.linesize[1], cw, chh, C.c_float(param), 0, 0)
    assert rc == 0, (fn + "_uv", rc)
    torch.cuda.synchronize()
