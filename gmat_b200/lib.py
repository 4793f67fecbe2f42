"""ctypes binding of include/gmat_b200.h (no torch import here)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))


class GmatbError(RuntimeError):
    pass


class GmatbImage(C.Structure):
    _fields_ = [
        ("data", C.c_void_p * 4),
        ("linesize", C.c_int * 4),
        ("width", C.c_int),
        ("height", C.c_int),
        ("format", C.c_int),
        ("batch", C.c_int),
        ("batch_stride", C.c_longlong * 4),
    ]


class _NS:
    def __init__(self, **kw):
        self.__dict__.update(kw)


# enum AVPixelFormat values (libavutil/pixfmt.h of the reference)
FMT = _NS(YUV420P=0, RGB24=2, BGR24=3, NV12=23, RGBA=26, BGRA=28, RGB48LE=35, YUV420P16LE=45,
          BGR48LE=58, YUV420P10LE=62, RGBA64LE=105, BGRA64LE=107, ZRGB=118, RGB0=119, ZBGR=120, BGR0=121,
          P010LE=159, P016LE=170, RGBPF32LE=179, RGBAPF32LE=180)
SPC = _NS(DEFAULT=0, BT709=1, FCC=4, BT470BG=5, SMPTE170M=6, SMPTE240M=7, BT2020_NCL=9, BT2020_CL=10)
SWS = _NS(FAST_BILINEAR=1, BILINEAR=2, BICUBIC=4, POINT=0x10, AREA=0x20, LANCZOS=0x200,
          HWACCEL_CUDA=0x1000000, PARITY_WRAP=0x40000000, INT_CHAIN=0x20000000, MMA_CHAIN=0x10000000, TILE_KERNEL=0x08000000, PARAM_DEFAULT=123456.0)
INTERP = _NS(NEAREST=0, LINEAR=1, CUBIC=2, AREA=3)
BORDER = _NS(CONSTANT=0, REPLICATE=1, REFLECT=2, WRAP=3, REFLECT101=4)

_lib = None


def lib_path():
    return os.path.join(_HERE, "libgmat_b200.so")


def lib():
    """Load libgmat_b200.so.  Fails loudly: there is no fallback implementation."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise GmatbError(f"{p} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                         "(nvcc -gencode arch=compute_100a,code=sm_100a); gmat_b200 has no CPU fallback")
    L = C.CDLL(p, mode=C.RTLD_GLOBAL)
    IP = C.POINTER(GmatbImage)
    vp, ci, cd, cf = C.c_void_p, C.c_int, C.c_double, C.c_float
    sig = {
        "gmatb_version": (ci, []),
        "gmatb_device_count": (ci, []),
        "gmatb_last_cuda_error": (ci, []),
        "gmatb_last_cuda_error_string": (C.c_char_p, []),
        "gmatb_launch_count": (C.c_longlong, []),
        "gmatb_device_sync": (ci, []),
        "gmatb_csc_matrix_yuv2rgb": (None, [ci, C.POINTER(cf)]),
        "gmatb_csc_matrix_rgb2yuv": (None, [ci, C.POINTER(cf)]),
        "gmatb_yuv2rgb": (ci, [IP, IP, ci, vp]),
        "gmatb_rgb2yuv": (ci, [IP, IP, ci, vp]),
        "gmatb_yuv2yuv": (ci, [IP, IP, vp]),
        "gmatb_rgb24tobgr24": (ci, [IP, IP, vp]),
        "gmatb_yuv2rgb_planar_f32": (ci, [IP, IP, ci, cf, C.POINTER(cf), vp]),
        "gmatb_format_colorspace": (ci, [ci]),
        "gmatb_format_nv12_to_rgbpf32": (ci, [IP, IP, ci, cf, C.POINTER(cf), ci, vp]),
        "gmatb_format_rgbpf32_to_nv12": (ci, [IP, IP, ci, vp]),
        "gmatb_sws_create": (vp, [ci, ci, ci, ci, ci, ci, ci, C.POINTER(cd), ci]),
        "gmatb_sws_free": (None, [vp]),
        "gmatb_sws_set_stream": (None, [vp, vp]),
        "gmatb_sws_scale": (ci, [vp, C.POINTER(vp), C.POINTER(ci), C.POINTER(vp), C.POINTER(ci)]),
        "gmatb_sws_scale_batch": (ci, [vp, IP, IP]),
        "gmatb_sws_scale_host": (ci, [vp, IP, IP]),
        "gmatb_sws_get_filter": (ci, [vp, ci, C.POINTER(cf), C.POINTER(ci)]),
        "gmatb_sws_path": (ci, [vp]),
        "gmatb_crop": (ci, [IP, IP, ci, ci, vp]),
        "gmatb_flip": (ci, [IP, IP, ci, vp]),
        "gmatb_rotate": (ci, [IP, IP, cd, cd, cd, ci, vp]),
        "gmatb_gaussian": (ci, [IP, IP, ci, ci, cd, cd, ci, vp]),
        "gmatb_median": (ci, [IP, IP, ci, ci, vp]),
        # libswscale boundary symbols living in this library (include/gmat_b200_sws.h)
        "yuv2rgb_cuda": (ci, [C.POINTER(vp), C.POINTER(ci), C.POINTER(vp), C.POINTER(ci), ci, ci, ci, ci, vp]),
        "rgb2yuv_cuda": (ci, [C.POINTER(vp), C.POINTER(ci), C.POINTER(vp), C.POINTER(ci), ci, ci, ci, ci, vp]),
        "yuv2yuv_cuda": (ci, [C.POINTER(vp), C.POINTER(ci), C.POINTER(vp), C.POINTER(ci), ci, ci, ci, ci, vp]),
        "rgb24tobgr24_cuda": (None, [C.POINTER(vp), C.POINTER(vp), C.POINTER(ci), C.POINTER(ci), ci, ci, vp]),
        "rgb2rgb_init_cuda": (None, []),
        "gmatb_set_process_colorspace": (None, [ci]),
        "gmatb_get_process_colorspace": (ci, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)      # AttributeError if the .so does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(rc, what="gmat_b200 call"):
    if rc is None or rc >= 0:
        return rc
    L = lib()
    extra = ""
    if rc == -5:
        extra = f" (cuda {L.gmatb_last_cuda_error()}: {L.gmatb_last_cuda_error_string().decode()})"
    raise GmatbError(f"{what} failed with {rc}{extra}")
