"""SwsContext mirror: same names and argument meaning as the reference's public API
(libswscale/swscale.h: sws_getContext :?, sws_scale, sws_freeContext_cuda :188,
sws_setCudaStream :448) for the SWS_HWACCEL_CUDA path."""
import ctypes as C

import numpy as np

from .image import FrameBatch, ptr_arrays
from .lib import SPC, SWS, GmatbError, check, lib


class SwsContext:
    def __init__(self, srcW, srcH, srcFormat, dstW, dstH, dstFormat, flags=SWS.BICUBIC | SWS.HWACCEL_CUDA,
                 param=None, colorspace=SPC.DEFAULT):
        L = lib()
        self.srcW, self.srcH, self.srcFormat = srcW, srcH, srcFormat
        self.dstW, self.dstH, self.dstFormat = dstW, dstH, dstFormat
        self.flags = flags
        pp = None
        if param is not None:
            pp = (C.c_double * 2)(*[float(x) for x in (list(param) + [SWS.PARAM_DEFAULT] * 2)[:2]])
        self._h = L.gmatb_sws_create(srcW, srcH, srcFormat, dstW, dstH, dstFormat, flags, pp, colorspace)
        if not self._h:
            # sws_getContext returns NULL when ff_sws_init_swscale_cuda fails (utils.c:2102-2105)
            raise GmatbError(f"sws_getContext: unsupported conversion {srcW}x{srcH} fmt {srcFormat} -> "
                             f"{dstW}x{dstH} fmt {dstFormat} flags {flags:#x}")

    # sws_setCudaStream (swscale.c:1249)
    def set_stream(self, stream):
        lib().gmatb_sws_set_stream(self._h, C.c_void_p(stream or 0))

    @property
    def path(self):
        return lib().gmatb_sws_path(self._h)

    def scale(self, src, dst):
        """src/dst: GmatbImage (device pointers) or FrameBatch (whole batch)"""
        s = src.image() if isinstance(src, FrameBatch) else src
        d = dst.image() if isinstance(dst, FrameBatch) else dst
        check(lib().gmatb_sws_scale_batch(self._h, C.byref(s), C.byref(d)), "sws_scale")
        return 0

    def scale_arrays(self, src_ptrs, src_strides, dst_ptrs, dst_strides):
        """FFmpeg-style single-frame call (uint8* [4], int [4])"""
        check(lib().gmatb_sws_scale(self._h, src_ptrs, src_strides, dst_ptrs, dst_strides), "sws_scale")
        return 0

    def scale_host(self, src, dst):
        """host frames in, host frames out (H2D + convert + D2H + stream sync)"""
        s = src.image() if isinstance(src, FrameBatch) else src
        d = dst.image() if isinstance(dst, FrameBatch) else dst
        check(lib().gmatb_sws_scale_host(self._h, C.byref(s), C.byref(d)), "sws_scale_host")
        return 0

    def get_filter(self, axis):
        n = self.dstW if axis == 0 else self.dstH
        co = np.zeros((n, 4), np.float32)
        po = np.zeros(n, np.int32)
        check(lib().gmatb_sws_get_filter(self._h, axis, co.ctypes.data_as(C.POINTER(C.c_float)),
                                         po.ctypes.data_as(C.POINTER(C.c_int))), "get_filter")
        return co, po

    def free(self):
        if getattr(self, "_h", None):
            lib().gmatb_sws_free(self._h)
            self._h = None

    __del__ = free


def sws_getContext(srcW, srcH, srcFormat, dstW, dstH, dstFormat, flags, srcFilter=None, dstFilter=None, param=None):
    """libswscale/utils.c:2087.  Only the SWS_HWACCEL_CUDA path exists here."""
    if not flags & SWS.HWACCEL_CUDA:
        raise GmatbError("gmat_b200 implements only the SWS_HWACCEL_CUDA path of sws_getContext (no CPU scaler)")
    try:
        return SwsContext(srcW, srcH, srcFormat, dstW, dstH, dstFormat, flags, param)
    except GmatbError:
        return None


def sws_scale(c, srcSlice, srcStride, srcSliceY, srcSliceH, dst, dstStride):
    """libswscale/swscale.c:1204.  Returns 0 like the reference's CUDA path."""
    return c.scale_arrays(srcSlice, srcStride, dst, dstStride)


def sws_freeContext(c):
    if c is not None:
        c.free()
