"""-m gpu: colour-conversion parity.  Ours (through the C ABI) vs the CPU oracle, vs the
reference's own kernels (O1) running on the same GPU, and the exhaustive 2^24 (Y,U,V) sweep."""
import ctypes as C

import numpy as np
import pytest
import torch

import gmat_b200 as g
import orc
from gmat_b200 import FMT, FrameBatch
from gpu_util import assert_same, have_ref, o1, o1_run

pytestmark = pytest.mark.gpu

SIZES = [(64, 48), (33, 17), (2, 2), (3, 3), (1, 1), (17, 33), (130, 6), (1919, 1079), (256, 8), (8, 2)]
RGB_DST = [FMT.RGB24, FMT.BGR24, FMT.RGBA, FMT.BGRA, FMT.RGBA64LE, FMT.BGRA64LE, FMT.RGB48LE, FMT.BGR48LE]


@pytest.mark.parametrize("w,h", SIZES)
@pytest.mark.parametrize("sfmt", [FMT.NV12, FMT.YUV420P])
def test_yuv2rgb_8bit_vs_oracle(dev, sfmt, w, h):
    src = FrameBatch(sfmt, w, h, 2); src.fill_lcg(seed=w * 7 + h)
    ds = src.to(dev)
    for cs in (0, 1):
        for dfmt in RGB_DST:
            ref = FrameBatch(dfmt, w, h, 2); orc.yuv2rgb(src, ref, cs)
            dd = FrameBatch(dfmt, w, h, 2, device=dev)
            g.yuv2rgb(ds, dd, cs); torch.cuda.synchronize()
            assert_same(dd, ref, f"{sfmt}->{dfmt} {w}x{h} cs{cs}")


@pytest.mark.parametrize("w,h", [(64, 48), (34, 18), (2, 2), (33, 17), (3840, 16)])
@pytest.mark.parametrize("sfmt", [FMT.P010LE, FMT.P016LE])
def test_yuv2rgb_16bit_vs_oracle(dev, sfmt, w, h):
    src = FrameBatch(sfmt, w, h, 1); src.fill_lcg(seed=w + h)
    ds = src.to(dev)
    for cs in (0, 9):
        for dfmt in (FMT.RGB48LE, FMT.BGR48LE, FMT.RGBA64LE, FMT.BGRA64LE, FMT.RGB24, FMT.BGRA):
            ref = FrameBatch(dfmt, w, h, 1); orc.yuv2rgb(src, ref, cs)
            dd = FrameBatch(dfmt, w, h, 1, device=dev)
            g.yuv2rgb(ds, dd, cs); torch.cuda.synchronize()
            assert_same(dd, ref, f"{sfmt}->{dfmt} {w}x{h} cs{cs}")


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("w,h", SIZES + [(3840, 2160)])
def test_yuv2rgb_vs_reference_kernels_live(dev, w, h):
    """bit-exact against nv122color<...> of the reference (yuv2rgb_cuda.cu:556-562), all W/H parities"""
    src = FrameBatch(FMT.NV12, w, h, 1, device=dev); src.fill_lcg(seed=11 * w + h)
    for cs in (0, 1, 7):
        for dfmt in (FMT.RGB24, FMT.BGR24, FMT.RGBA, FMT.BGRA, FMT.RGBA64LE, FMT.BGRA64LE):
            ref = FrameBatch(dfmt, w, h, 1, device=dev); o1_run("yuv2rgb_cuda", src, ref, cs)
            dd = FrameBatch(dfmt, w, h, 1, device=dev)
            g.yuv2rgb(src, dd, cs); torch.cuda.synchronize()
            assert_same(dd, ref, f"nv12->{dfmt} {w}x{h} cs{cs} vs O1")


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
def test_exhaustive_yuv_triples_vs_reference(dev):
    """every one of the 2^24 (Y,U,V) triples, one 4096x4096 NV12 frame"""
    W = H = 4096
    src = FrameBatch(FMT.NV12, W, H, 1)
    host = np.zeros(src.frame_bytes, np.uint8)
    yv = src.plane_view(host, 0, 0); uv = src.plane_view(host, 0, 1)
    by, bx = np.mgrid[0:H // 2, 0:W // 2]
    blk = by * (W // 2) + bx                       # 2^22 blocks; chroma combo = blk >> 6, luma group = blk & 63
    combo = blk >> 6
    uv[:, 0::2] = (combo & 255).astype(np.uint8)
    uv[:, 1::2] = (combo >> 8).astype(np.uint8)
    base = ((blk & 63) * 4).astype(np.uint8)
    yv[0::2, 0::2] = base; yv[0::2, 1::2] = base + 1; yv[1::2, 0::2] = base + 2; yv[1::2, 1::2] = base + 3
    ds = FrameBatch(FMT.NV12, W, H, 1, device=dev); ds.upload(host)
    for cs in (0, 1, 4, 7):
        for dfmt in (FMT.RGB24, FMT.BGRA):
            ref = FrameBatch(dfmt, W, H, 1, device=dev); o1_run("yuv2rgb_cuda", ds, ref, cs)
            dd = FrameBatch(dfmt, W, H, 1, device=dev)
            g.yuv2rgb(ds, dd, cs); torch.cuda.synchronize()
            assert torch.equal(dd.buf, ref.buf), f"cs {cs} fmt {dfmt}"


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("w,h", [(64, 48), (2, 2), (130, 6), (1920, 1080)])
def test_p016_vs_reference_template(dev, w, h):
    """P016 -> RGBA64/BGRA64: the undispatched p0162color64 template of the reference (:612-618)"""
    src = FrameBatch(FMT.P016LE, w, h, 1, device=dev); src.fill_lcg(seed=3 * w + h)
    o1().set_mat_yuv2rgb_cuda(0); torch.cuda.synchronize()
    for order, dfmt in ((0, FMT.RGBA64LE), (1, FMT.BGRA64LE)):
        ref = FrameBatch(dfmt, w, h, 1, device=dev)
        si, ri = src.image(), ref.image()
        o1().ref_p016_to_color64(si.data[0], si.linesize[0], ri.data[0], ri.linesize[0], w, h, order, None)
        torch.cuda.synchronize()
        dd = FrameBatch(dfmt, w, h, 1, device=dev)
        g.yuv2rgb(src, dd); torch.cuda.synchronize()
        assert_same(dd, ref, f"p016->{dfmt} {w}x{h} vs O1")


@pytest.mark.parametrize("w,h", [(64, 48), (33, 17), (2, 2), (1920, 1080)])
def test_planar_float_vs_oracle(dev, w, h):
    src = FrameBatch(FMT.NV12, w, h, 1); src.fill_lcg(seed=w)
    ds = src.to(dev)
    for norm, shift in ((255.0, (0.0, 0.0, 0.0)), (58.395, (123.675, 116.28, 103.53))):
        ref = FrameBatch(FMT.RGBPF32LE, w, h, 1)
        m = orc.matrix_yuv2rgb(0); sh = np.array(shift, np.float32)
        s, d = src.image(), ref.image()
        orc.orc().orc_yuv2rgb_planar_f32(C.byref(s), C.byref(d), orc.fptr(m), norm, orc.fptr(sh))
        dd = FrameBatch(FMT.RGBPF32LE, w, h, 1, device=dev)
        g.yuv2rgb_planar_f32(ds, dd, 0, norm, shift); torch.cuda.synchronize()
        assert_same(dd, ref, f"nv12->rgbpf32 {w}x{h} norm {norm}")     # floats compared bit for bit (<= 1 ulp allowed, 0 achieved)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
def test_planar_float_vs_reference_live(dev):
    w, h = 640, 360
    src = FrameBatch(FMT.NV12, w, h, 1, device=dev); src.fill_lcg(seed=8)
    ref = FrameBatch(FMT.RGBPF32LE, w, h, 1, device=dev); o1_run("yuv2rgb_cuda", src, ref, 0)
    dd = FrameBatch(FMT.RGBPF32LE, w, h, 1, device=dev)
    g.yuv2rgb(src, dd); torch.cuda.synchronize()
    assert_same(dd, ref, "nv12->rgbpf32 vs O1")


@pytest.mark.parametrize("w,h", [(64, 48), (33, 17), (2, 2), (3, 3), (1920, 1080)])
@pytest.mark.parametrize("dfmt", [FMT.NV12, FMT.YUV420P])
def test_rgb2yuv_vs_oracle(dev, dfmt, w, h):
    for sfmt in (FMT.RGB24, FMT.BGR24, FMT.RGBA, FMT.BGRA, FMT.RGBA64LE, FMT.BGRA64LE):
        src = FrameBatch(sfmt, w, h, 1); src.fill_lcg(seed=w + 13 * h)
        ref = FrameBatch(dfmt, w, h, 1); orc.rgb2yuv(src, ref, 0)
        ds = src.to(dev); dd = FrameBatch(dfmt, w, h, 1, device=dev)
        g.rgb2yuv(ds, dd, 0); torch.cuda.synchronize()
        assert_same(dd, ref, f"{sfmt}->{dfmt} {w}x{h}")


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("w,h", [(64, 48), (2, 2), (130, 6), (1920, 1080)])
def test_rgb24_to_nv12_vs_reference_live(dev, w, h):
    """RgbToYuvKernel (yuv2rgb_cuda.cu:671-702): the one rgb->yuv launcher of the reference that is sane (even sizes)"""
    src = FrameBatch(FMT.RGB24, w, h, 1, device=dev); src.fill_lcg(seed=w)
    for cs in (0, 1):
        ref = FrameBatch(FMT.NV12, w, h, 1, device=dev); o1_run("rgb2yuv_cuda", src, ref, cs)
        dd = FrameBatch(FMT.NV12, w, h, 1, device=dev)
        g.rgb2yuv(src, dd, cs); torch.cuda.synchronize()
        assert_same(dd, ref, f"rgb24->nv12 {w}x{h} cs{cs} vs O1")


YUVS = [FMT.NV12, FMT.YUV420P, FMT.P010LE, FMT.P016LE, FMT.YUV420P10LE, FMT.YUV420P16LE]


@pytest.mark.parametrize("w,h", [(64, 48), (33, 17), (2, 2), (1, 1), (1920, 1080)])
def test_yuv2yuv_vs_oracle(dev, w, h):
    for sfmt in YUVS:
        src = FrameBatch(sfmt, w, h, 1); src.fill_lcg(seed=w + h)
        ds = src.to(dev)
        for dfmt in YUVS:
            ref = FrameBatch(dfmt, w, h, 1)
            s, d = src.image(), ref.image()
            orc.orc().orc_yuv2yuv(C.byref(s), C.byref(d))
            dd = FrameBatch(dfmt, w, h, 1, device=dev)
            g.yuv2yuv(ds, dd); torch.cuda.synchronize()
            assert_same(dd, ref, f"{sfmt}->{dfmt} {w}x{h}")


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
def test_yuv2yuv_and_swap_vs_reference_live(dev):
    w, h = 640, 360
    src = FrameBatch(FMT.NV12, w, h, 1, device=dev); src.fill_lcg(seed=1)
    for dfmt in (FMT.YUV420P, FMT.P010LE, FMT.P016LE, FMT.YUV420P10LE, FMT.YUV420P16LE):
        ref = FrameBatch(dfmt, w, h, 1, device=dev); o1_run("yuv2yuv_cuda", src, ref)
        dd = FrameBatch(dfmt, w, h, 1, device=dev); g.yuv2yuv(src, dd); torch.cuda.synchronize()
        assert_same(dd, ref, f"nv12->{dfmt} vs O1")
    src = FrameBatch(FMT.YUV420P, w, h, 1, device=dev); src.fill_lcg(seed=2)
    for dfmt in (FMT.NV12, FMT.P010LE, FMT.P016LE, FMT.YUV420P10LE, FMT.YUV420P16LE):
        ref = FrameBatch(dfmt, w, h, 1, device=dev); o1_run("yuv2yuv_cuda", src, ref)
        dd = FrameBatch(dfmt, w, h, 1, device=dev); g.yuv2yuv(src, dd); torch.cuda.synchronize()
        assert_same(dd, ref, f"yuv420p->{dfmt} vs O1")
    for (w, h) in ((640, 360), (33, 17), (1, 1)):
        src = FrameBatch(FMT.RGB24, w, h, 1, device=dev); src.fill_lcg(seed=3)
        ref = FrameBatch(FMT.BGR24, w, h, 1, device=dev); o1_run("swap", src, ref)
        dd = FrameBatch(FMT.BGR24, w, h, 1, device=dev); g.rgb24tobgr24(src, dd); torch.cuda.synchronize()
        assert_same(dd, ref, f"rgb24->bgr24 {w}x{h} vs O1")


def test_unaligned_pointers_take_the_scalar_path(dev):
    """pitch / base not multiples of 16: same bytes, slower path"""
    w, h = 70, 20
    src = FrameBatch(FMT.NV12, w, h, 1, align=1); src.fill_lcg(seed=4)
    ref = FrameBatch(FMT.RGB24, w, h, 1, align=1); orc.yuv2rgb(src, ref)
    ds = src.to(dev); dd = FrameBatch(FMT.RGB24, w, h, 1, device=dev, align=1)
    g.yuv2rgb(ds, dd); torch.cuda.synchronize()
    assert_same(dd, ref, "unaligned")


def test_batch_equals_frame_by_frame(dev):
    w, h, n = 128, 64, 5
    src = FrameBatch(FMT.NV12, w, h, n, device=dev); src.fill_lcg(seed=6)
    a = FrameBatch(FMT.RGB24, w, h, n, device=dev); g.yuv2rgb(src, a)
    b = FrameBatch(FMT.RGB24, w, h, n, device=dev)
    for i in range(n):
        g.yuv2rgb(src.image(i, 1), b.image(i, 1))
    torch.cuda.synchronize()
    assert torch.equal(a.buf, b.buf)


def test_libswscale_unscaled_symbols(dev):
    """yuv2rgb_cuda(src[], srcStride[], ...) exactly as swscale_unscaled.c:1996 calls it; UV found at
    src[0] + H*pitch when the caller passes no second plane (reference convention)"""
    vp, ci = C.c_void_p, C.c_int
    w, h = 64, 48
    src = FrameBatch(FMT.NV12, w, h, 1, device=dev); src.fill_lcg(seed=10)
    ref = FrameBatch(FMT.RGB24, w, h, 1, device=dev); g.yuv2rgb(src, ref)
    dd = FrameBatch(FMT.RGB24, w, h, 1, device=dev)
    si, di = src.image(), dd.image()
    sp = (vp * 4)(si.data[0], None, None, None); ss = (ci * 4)(si.linesize[0], 0, 0, 0)
    dp = (vp * 4)(di.data[0], None, None, None); dsn = (ci * 4)(di.linesize[0], 0, 0, 0)
    assert g.lib().yuv2rgb_cuda(sp, ss, dp, dsn, w, h, FMT.NV12, FMT.RGB24, None) == 0
    torch.cuda.synchronize()
    assert torch.equal(dd.buf, ref.buf)
    assert g.lib().yuv2rgb_cuda(sp, ss, dp, dsn, w, h, FMT.RGB24, FMT.RGB24, None) == -1     # unsupported pair
