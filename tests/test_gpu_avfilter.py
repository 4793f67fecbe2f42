"""-m gpu: the AVFilter surface for real.  The six filter objects of
gmat_b200/csrc/avfilter/ are linked into the reference's OWN libavfilter + libavutil (oracle/refbuild `avf`:
avfilter.c, avfiltergraph.c, buffersrc.c, buffersink.c, formats.c, framepool.c, scale_eval.c, hwcontext.c,
hwcontext_cuda.c ... compiled from /root/reference, allfilters.c unmodified with a generated filter_list.c), registered,
found by name, configured by the reference's graph parser / format negotiation, and fed CUDA frames from the reference's
frame pool (hwcontext_cuda.c:96-205: pitch aligned to the device's texture alignment, planes contiguous).
Every filter's init / config_props / filter_frame runs; the output must equal the kernel layer called directly
on the same pixels."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import gmat_b200 as g
from gmat_b200 import BORDER, FMT, SWS, FrameBatch, SwsContext
from gpu_util import REF

pytestmark = pytest.mark.gpu
LIB = os.path.join(REF, "libref_avfilter.so")
needs_lib = pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libref_avfilter.so not built")
HW = SWS.HWACCEL_CUDA


class AvfResult(C.Structure):
    _fields_ = [("out_w", C.c_int), ("out_h", C.c_int), ("out_fmt", C.c_int), ("in_pitch", C.c_int), ("out_pitch", C.c_int),
                ("frames_out", C.c_int), ("out_bytes_per_frame", C.c_longlong), ("error", C.c_char * 256)]


_L = None


def avf():
    global _L
    if _L is None:
        g.lib()                                   # libgmat_b200.so first (RTLD_GLOBAL): the filter objects call into it
        _L = C.CDLL(LIB, mode=C.RTLD_GLOBAL)
        _L.avf_list_filters.restype = C.c_char_p
        _L.avf_run.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_void_p, C.c_int, C.c_void_p, C.c_longlong, C.POINTER(AvfResult)]
    return _L


def run_graph(fmt_name, w, h, filters, frames, out_cap=None):
    """frames: uint8 array [n, frame_bytes] in av_image_copy_to_buffer layout (planes tightly packed)"""
    frames = np.ascontiguousarray(frames, np.uint8)
    n = frames.shape[0]
    out = np.zeros(out_cap or (n * max(frames.shape[1] * 4, 1 << 20)), np.uint8)
    r = AvfResult()
    rc = avf().avf_run(fmt_name.encode(), w, h, filters.encode(), frames.ctypes.data, n, out.ctypes.data, out.size, C.byref(r))
    assert rc == 0, (rc, r.error.decode())
    assert r.frames_out == n
    return out[:n * r.out_bytes_per_frame].reshape(n, r.out_bytes_per_frame), r


def batch_from(fmt, w, h, frames, dev):
    fb = FrameBatch(fmt, w, h, frames.shape[0])
    host = np.zeros(fb.frame_bytes * fb.n, np.uint8)
    for f in range(fb.n):
        off = 0
        for p in range(len(fb.planes)):
            v = fb.plane_view(host, f, p)
            v[...] = frames[f, off:off + v.size].reshape(v.shape); off += v.size
    fb.upload(host)
    return fb.to(dev)


def payload_frames(fb):
    return fb.payload().reshape(fb.n, -1)


def rand_frames(n, nbytes, seed):
    return np.random.default_rng(seed).integers(0, 256, size=(n, nbytes), dtype=np.uint8)


FMTS = [("rgb24", FMT.RGB24, 3), ("bgr24", FMT.BGR24, 3), ("bgr0", FMT.BGR0, 4), ("rgb0", FMT.RGB0, 4)]


@needs_lib
def test_filters_are_registered_in_the_reference_libavfilter(dev):
    names = avf().avf_list_filters().decode().split()
    for n in ("crop_cuda", "rotate_cuda", "flip_cuda", "smooth_cuda", "scale_cuda", "format_cuda", "buffer", "buffersink"):
        assert n in names, names


@needs_lib
@pytest.mark.parametrize("name,fmt,bpp", FMTS)
def test_each_filter_on_pool_frames_equals_the_kernel_layer(dev, name, fmt, bpp):
    w, h, n = 642, 362, 3                       # row bytes are not a multiple of the pool's pitch alignment
    frames = rand_frames(n, w * h * bpp, 7)
    src = batch_from(fmt, w, h, frames, dev)

    def direct(fn, ow=w, oh=h):
        d = FrameBatch(fmt, ow, oh, n, device=dev); fn(src, d); torch.cuda.synchronize(); return payload_frames(d)

    cases = [
        ("crop_cuda=w=320:h=180:x=10:y=21", lambda s, d: g.crop(s, d, 10, 21), 320, 180),
        ("crop_cuda=w=300:h=200", lambda s, d: g.crop(s, d, -1, -1), 300, 200),
        ("flip_cuda=code=1", lambda s, d: g.flip(s, d, 1), w, h),
        ("flip_cuda=code=-1", lambda s, d: g.flip(s, d, -1), w, h),
        ("rotate_cuda=angle=30:shift_x=-20.5:shift_y=11.25", lambda s, d: g.rotate(s, d, 30.0, -20.5, 11.25, "linear"), w, h),
        ("rotate_cuda=angle=-77.3:interp=cubic", lambda s, d: g.rotate(s, d, -77.3, 0.0, 0.0, "cubic"), w, h),
        ("smooth_cuda=type=gaussian:kw=5:kh=5:sigmaX=1.1:sigmaY=1.1:border_type=reflect101",
         lambda s, d: g.gaussian(s, d, 5, 5, 1.1, 1.1, BORDER.REFLECT101), w, h),
        ("smooth_cuda=type=median:kw=3:kh=3", lambda s, d: g.median(s, d, 3, 3), w, h),
        ("smooth_cuda=type=median:kw=5:kh=5", lambda s, d: g.median(s, d, 5, 5), w, h),
    ]
    for desc, fn, ow, oh in cases:
        out, r = run_graph(name, w, h, desc, frames)
        assert (r.out_w, r.out_h) == (ow, oh), desc
        assert r.in_pitch % 256 == 0 and r.in_pitch > w * bpp and r.out_pitch % 256 == 0, (desc, r.in_pitch, r.out_pitch)
        exp = direct(fn, ow, oh)
        assert out.shape == exp.shape and np.array_equal(out, exp), f"{name} {desc}: {int((out != exp).sum())} bytes differ"


@needs_lib
def test_bad_options_fail_at_graph_configuration(dev):
    frames = rand_frames(1, 64 * 48 * 3, 1)
    for desc in ("smooth_cuda=type=median:kw=17:kh=17", "smooth_cuda=type=gaussian:kw=4:kh=4", "crop_cuda=w=100:h=100",
                 "crop_cuda=w=0:h=10"):
        out = np.zeros(1 << 20, np.uint8); r = AvfResult()
        rc = avf().avf_run(b"rgb24", 64, 48, desc.encode(), frames.ctypes.data, 1, out.ctypes.data, out.size, C.byref(r))
        assert rc < 0 and (b"avfilter_graph_config" in r.error or b"avfilter_graph_parse_ptr" in r.error), (desc, rc, r.error)


@needs_lib
@pytest.mark.parametrize("algo,flag,param", [("bicubic", SWS.BICUBIC, None), ("lanczos", SWS.LANCZOS, None), ("bilinear", SWS.BILINEAR, None),
                                             ("nearest", SWS.POINT, None), ("bicubic:param=0.75", SWS.BICUBIC, (0.75,))])
def test_scale_cuda_on_pool_frames(dev, algo, flag, param):
    w, h, ow, oh, n = 640, 360, 400, 226, 2
    for name, fmt, nbytes in (("nv12", FMT.NV12, w * h * 3 // 2), ("yuv420p", FMT.YUV420P, w * h * 3 // 2), ("bgr0", FMT.BGR0, w * h * 4),
                              ("p010le", FMT.P010LE, w * h * 3), ("p016le", FMT.P016LE, w * h * 3)):
        frames = rand_frames(n, nbytes, 11)
        if name == "p010le":
            frames = frames.copy(); frames[:, 0::2] &= 0xC0
        out, r = run_graph(name, w, h, f"scale_cuda=w={ow}:h={oh}:interp_algo={algo}", frames)
        assert (r.out_w, r.out_h) == (ow, oh)
        src = batch_from(fmt, w, h, frames, dev)
        d = FrameBatch(fmt, ow, oh, n, device=dev)
        SwsContext(w, h, fmt, ow, oh, fmt, flag | HW, param).scale(src, d); torch.cuda.synchronize()
        exp = payload_frames(d)
        assert np.array_equal(out, exp), f"scale_cuda {name} {algo}: {int((out != exp).sum())} bytes differ"


@needs_lib
def test_scale_cuda_passthrough_and_format_change(dev):
    w, h, n = 320, 180, 2
    frames = rand_frames(n, w * h * 3 // 2, 3)
    out, r = run_graph("nv12", w, h, "scale_cuda", frames)                    # same size, same format: passthrough
    assert np.array_equal(out, frames)
    out, r = run_graph("nv12", w, h, "scale_cuda=format=yuv420p", frames)     # repack only (nearest at the same size)
    src = batch_from(FMT.NV12, w, h, frames, dev); d = FrameBatch(FMT.YUV420P, w, h, n, device=dev)
    g.yuv2yuv(src, d); torch.cuda.synchronize()
    assert np.array_equal(out, payload_frames(d))


@needs_lib
def test_format_cuda_on_pool_frames(dev):
    w, h, n = 320, 180, 2
    frames = rand_frames(n, w * h * 3 // 2, 5)
    out, r = run_graph("nv12", w, h, "format_cuda=pix_fmt=rgbpf32le", frames, out_cap=n * w * h * 12 + 4096)
    src = batch_from(FMT.NV12, w, h, frames, dev); d = FrameBatch(FMT.RGBPF32LE, w, h, n, device=dev)
    g.format_nv12_to_rgbpf32(src, d, 2); torch.cuda.synchronize()
    assert np.array_equal(out, payload_frames(d))
    # and back: planar float -> nv12 within one graph
    out2, r2 = run_graph("nv12", w, h, "format_cuda=pix_fmt=rgbpf32le,format_cuda=pix_fmt=nv12", frames)
    back = FrameBatch(FMT.NV12, w, h, n, device=dev); g.format_rgbpf32_to_nv12(d, back, 2); torch.cuda.synchronize()
    assert np.array_equal(out2, payload_frames(back))


@needs_lib
@pytest.mark.parametrize("name,fmt,bpp", [("rgb24", FMT.RGB24, 3), ("bgr0", FMT.BGR0, 4)])
def test_c4_like_chain_in_one_graph(dev, name, fmt, bpp):
    """crop -> rotate -> smooth -> (scale_cuda for the 4-byte format the reference's scale_cuda accepts) in ONE graph"""
    w, h, n = 1920, 1080, 2
    frames = rand_frames(n, w * h * bpp, 13)
    desc = "crop_cuda=w=1280:h=720:x=100:y=50,rotate_cuda=angle=30:shift_x=-94.2563:shift_y=368.2309," \
           "smooth_cuda=type=gaussian:kw=5:kh=5:sigmaX=1.1:sigmaY=1.1:border_type=reflect101,flip_cuda=code=0"
    if bpp == 4:
        desc += ",scale_cuda=w=640:h=360:interp_algo=bicubic"
    out, r = run_graph(name, w, h, desc, frames)
    src = batch_from(fmt, w, h, frames, dev)
    a = FrameBatch(fmt, 1280, 720, n, device=dev); g.crop(src, a, 100, 50)
    b = FrameBatch(fmt, 1280, 720, n, device=dev); g.rotate(a, b, 30.0, -94.2563, 368.2309, "linear")
    c = FrameBatch(fmt, 1280, 720, n, device=dev); g.gaussian(b, c, 5, 5, 1.1, 1.1, BORDER.REFLECT101)
    d = FrameBatch(fmt, 1280, 720, n, device=dev); g.flip(c, d, 0)
    if bpp == 4:
        e = FrameBatch(fmt, 640, 360, n, device=dev)
        SwsContext(1280, 720, fmt, 640, 360, fmt, SWS.BICUBIC | HW).scale(d, e); d = e
    torch.cuda.synchronize()
    assert np.array_equal(out, payload_frames(d))
