"""-m gpu: the drop-in itself.  The REFERENCE's own libswscale.so (CPU objects, built by
oracle/refbuild without its cuda/ objects) is loaded on top of libgmat_b200.so + libgmat_b200_sws.so and
driven through its public API exactly like metrans/app/CSwscale.c: sws_getContext(...,
SWS_HWACCEL_CUDA) -> sws_setCudaStream -> sws_scale -> sws_freeContext_cuda."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import gmat_b200 as g
from gmat_b200 import FMT, SWS, FrameBatch, SwsContext

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFSO = os.path.join(ROOT, "oracle", "_ref", "libref_swscale_cpu.so")
vp, ci = C.c_void_p, C.c_int


@pytest.fixture(scope="module")
def refsws():
    if not os.path.exists(REFSO):
        pytest.skip("oracle/_ref/libref_swscale_cpu.so not built")
    g.lib()
    C.CDLL(os.path.join(ROOT, "gmat_b200", "libgmat_b200_sws.so"), mode=C.RTLD_GLOBAL)
    L = C.CDLL(REFSO, mode=C.RTLD_GLOBAL)
    L.sws_getContext.restype = vp
    L.sws_getContext.argtypes = [ci, ci, ci, ci, ci, ci, ci, vp, vp, C.POINTER(C.c_double)]
    L.sws_scale.restype = ci
    L.sws_scale.argtypes = [vp, C.POINTER(vp), C.POINTER(ci), ci, ci, C.POINTER(vp), C.POINTER(ci)]
    L.sws_setCudaStream.argtypes = [vp, vp]
    L.sws_freeContext_cuda.argtypes = [vp]
    return L


def call(L, ctx, src, dst):
    si, di = src.image(), dst.image()
    sp = (vp * 4)(*[si.data[i] for i in range(4)]); ss = (ci * 4)(*[si.linesize[i] for i in range(4)])
    dp = (vp * 4)(*[di.data[i] for i in range(4)]); ds = (ci * 4)(*[di.linesize[i] for i in range(4)])
    return L.sws_scale(ctx, sp, ss, 0, src.h, dp, ds)


@pytest.mark.parametrize("sfmt,dfmt", [(FMT.NV12, FMT.RGB24), (FMT.YUV420P, FMT.BGRA), (FMT.RGB24, FMT.NV12), (FMT.NV12, FMT.YUV420P),
                                       (FMT.RGB24, FMT.BGR24), (FMT.NV12, FMT.P010LE), (FMT.NV12, FMT.RGBPF32LE)])
def test_unscaled_through_reference_sws_scale(dev, refsws, sfmt, dfmt):
    w, h = 320, 180
    src = FrameBatch(sfmt, w, h, 1, device=dev); src.fill_lcg(seed=1)
    ctx = refsws.sws_getContext(w, h, sfmt, w, h, dfmt, SWS.HWACCEL_CUDA | SWS.BICUBIC, None, None, None)
    assert ctx, "reference sws_getContext returned NULL"
    out = FrameBatch(dfmt, w, h, 1, device=dev)
    st = torch.cuda.Stream()
    refsws.sws_setCudaStream(ctx, st.cuda_stream)
    assert call(refsws, ctx, src, out) == 0            # the reference's CUDA path returns 0
    st.synchronize()
    exp = FrameBatch(dfmt, w, h, 1, device=dev)
    SwsContext(w, h, sfmt, w, h, dfmt).scale(src, exp); torch.cuda.synchronize()
    assert torch.equal(out.buf, exp.buf)
    refsws.sws_freeContext_cuda(ctx)


@pytest.mark.parametrize("flags,param", [(SWS.BICUBIC, None), (SWS.LANCZOS, None), (SWS.BILINEAR, None), (SWS.BICUBIC, (0.75, 0.0))])
def test_scaled_through_reference_sws_scale(dev, refsws, flags, param):
    sw, sh, dw, dh = 1280, 720, 640, 360
    src = FrameBatch(FMT.NV12, sw, sh, 1, device=dev); src.fill_lcg(seed=2)
    pp = (C.c_double * 2)(*param) if param else None
    ctx = refsws.sws_getContext(sw, sh, FMT.NV12, dw, dh, FMT.RGB24, SWS.HWACCEL_CUDA | flags, None, None, pp)
    assert ctx
    out = FrameBatch(FMT.RGB24, dw, dh, 1, device=dev)
    assert call(refsws, ctx, src, out) == 0
    torch.cuda.synchronize()
    exp = FrameBatch(FMT.RGB24, dw, dh, 1, device=dev)
    SwsContext(sw, sh, FMT.NV12, dw, dh, FMT.RGB24, SWS.HWACCEL_CUDA | flags, param[:1] if param else None).scale(src, exp)
    torch.cuda.synchronize()
    assert torch.equal(out.buf, exp.buf)
    refsws.sws_freeContext_cuda(ctx)


def test_unsupported_pair_makes_reference_getcontext_fail(refsws):
    assert not refsws.sws_getContext(64, 48, FMT.RGBPF32LE, 32, 24, FMT.NV12, SWS.HWACCEL_CUDA, None, None, None)


def test_reference_metrans_cswscale_caller_unmodified(dev, refsws):
    """metrans/app/CSwscale.c -- the only in-tree consumer of SWS_HWACCEL_CUDA -- compiled UNMODIFIED
    (oracle/refbuild target n1) and driven the way metrans/python/swscale.py drives it: Init -> Convert -> Delete on
    tight NV12 / planar-float buffers.  The result must equal our kernel layer called directly."""
    so = os.path.join(ROOT, "oracle", "_ref", "libref_cswscale.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libref_cswscale.so not built")
    K = C.CDLL(so)
    K.SwscaleCuda_Nv12ToRgbpf32_Init.restype = vp
    K.SwscaleCuda_Nv12ToRgbpf32_Init.argtypes = [ci, ci]
    K.SwscaleCuda_Nv12ToRgbpf32_Convert.restype = ci
    K.SwscaleCuda_Nv12ToRgbpf32_Convert.argtypes = [vp, vp, ci, vp, ci, ci, ci, vp]
    K.SwscaleCuda_Nv12ToRgbpf32_Delete.argtypes = [vp]
    w, h = 640, 360
    # tight buffers, as av_image_fill_linesizes / av_image_fill_pointers lay them out
    rng = np.random.default_rng(3)
    nv12 = torch.from_numpy(rng.integers(0, 256, w * h * 3 // 2, dtype=np.uint8)).to(dev)
    out = torch.zeros(3 * w * h, dtype=torch.float32, device=dev)
    ctx = K.SwscaleCuda_Nv12ToRgbpf32_Init(w, h)
    assert ctx
    st = torch.cuda.Stream()
    assert K.SwscaleCuda_Nv12ToRgbpf32_Convert(ctx, nv12.data_ptr(), w, out.data_ptr(), 4 * w, w, h, st.cuda_stream) == 0
    st.synchronize()
    K.SwscaleCuda_Nv12ToRgbpf32_Delete(ctx)
    # the same conversion through the kernel layer on the same tight layout
    exp = torch.zeros_like(out)
    si = g.GmatbImage(); di = g.GmatbImage()
    si.data[0] = nv12.data_ptr(); si.data[1] = nv12.data_ptr() + w * h; si.linesize[0] = w; si.linesize[1] = w
    si.width, si.height, si.format, si.batch = w, h, FMT.NV12, 1
    for p in range(3):
        di.data[p] = exp.data_ptr() + 4 * w * h * p; di.linesize[p] = 4 * w
    di.width, di.height, di.format, di.batch = w, h, FMT.RGBPF32LE, 1
    g.yuv2rgb(si, di); torch.cuda.synchronize()
    assert torch.equal(out, exp)
    v = out.cpu().numpy()
    assert v.min() >= 0.0 and v.max() <= 1.0 and len(np.unique(v)) > 200
