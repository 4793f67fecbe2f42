"""-m gpu: crop / flip / rotate / gaussian / median vs the CPU oracle (the specification of these
filters: CV-CUDA, which the reference calls, is a closed third-party binary) + algebraic properties."""
import ctypes as C

import numpy as np
import pytest
import torch

import gmat_b200 as g
import orc
from gmat_b200 import BORDER, FMT, FrameBatch
from gpu_util import assert_same

pytestmark = pytest.mark.gpu
FMTS = [FMT.RGB24, FMT.BGRA]
SIZES = [(64, 48), (33, 17), (130, 7), (132, 9), (1, 1), (640, 360)]


def pair(fmt, w, h, dev, n=1, seed=1):
    src = FrameBatch(fmt, w, h, n); src.fill_lcg(seed=seed)
    return src, src.to(dev)


@pytest.mark.parametrize("fmt", FMTS)
def test_crop(dev, fmt):
    for (w, h, cw, ch, x, y) in ((64, 48, 32, 24, 5, 7), (640, 360, 333, 111, 3, 1), (33, 17, 33, 17, 0, 0), (64, 48, 1, 1, 63, 47),
                                 (640, 360, 320, 180, -1, -1)):
        src, ds = pair(fmt, w, h, dev, 2)
        dd = FrameBatch(fmt, cw, ch, 2, device=dev)
        g.crop(ds, dd, x, y); torch.cuda.synchronize()
        xx = (w - cw) // 2 if x < 0 else x; yy = (h - ch) // 2 if y < 0 else y
        ref = FrameBatch(fmt, cw, ch, 2)
        s, d = src.image(), ref.image(); orc.orc().orc_crop(C.byref(s), C.byref(d), xx, yy)
        assert_same(dd, ref, f"crop {w}x{h}->{cw}x{ch}@{x},{y}")


@pytest.mark.parametrize("fmt", FMTS)
@pytest.mark.parametrize("w,h", SIZES)
def test_flip(dev, fmt, w, h):
    src, ds = pair(fmt, w, h, dev, 2)
    for code in (0, 1, -1):
        dd = FrameBatch(fmt, w, h, 2, device=dev); g.flip(ds, dd, code)
        ref = FrameBatch(fmt, w, h, 2); s, d = src.image(), ref.image(); orc.orc().orc_flip(C.byref(s), C.byref(d), code)
        torch.cuda.synchronize()
        assert_same(dd, ref, f"flip {code} {w}x{h}")
        back = FrameBatch(fmt, w, h, 2, device=dev); g.flip(dd, back, code); torch.cuda.synchronize()
        assert torch.equal(back.buf, ds.buf)            # involution


@pytest.mark.parametrize("fmt", FMTS)
@pytest.mark.parametrize("interp", ["linear", "nearest", "cubic"])
def test_rotate(dev, fmt, interp):
    code = {"nearest": 0, "linear": 1, "cubic": 2}[interp]
    for (w, h) in ((64, 48), (33, 17), (640, 360)):
        src, ds = pair(fmt, w, h, dev, 1, seed=w)
        for ang, sx, sy in ((30.0, 0.0, 0.0), (30.0, -20.5, 11.25), (-77.3, 40.0, 5.0), (180.0, w - 1.0, h - 1.0), (0.0, 0.0, 0.0), (90.0, 0.0, h - 1.0), (0.5, 0.3, 0.2), (-0.7, -0.4, 0.45)):
            dd = FrameBatch(fmt, w, h, 1, device=dev); g.rotate(ds, dd, ang, sx, sy, interp); torch.cuda.synchronize()
            ref = FrameBatch(fmt, w, h, 1); s, d = src.image(), ref.image()
            orc.orc().orc_rotate(C.byref(s), C.byref(d), ang, sx, sy, code)
            assert_same(dd, ref, f"rotate {interp} {ang} {w}x{h}")
            if ang == 0.0:
                assert torch.equal(dd.buf, ds.buf)      # identity


def test_rotate_c4_geometry(dev):
    """BASELINE C4: 4K, 30 degrees about the centre (shift = c - R^T c), linear"""
    w, h = 3840, 2160
    src = FrameBatch(FMT.RGB24, w, h, 1, device=dev); src.fill_lcg(seed=4)
    dd = FrameBatch(FMT.RGB24, w, h, 1, device=dev)
    g.rotate(src, dd, 30.0, -282.7688, 1104.6926, "linear"); torch.cuda.synchronize()
    # the centre pixel stays (almost) where it was; corners fall outside the source and are zero
    host = dd.numpy(); v = dd.plane_view(host, 0, 0)
    assert v[0, 0:3].max() == 0 and v[h - 1, (w - 1) * 3:].max() == 0
    crop_w, crop_h = 256, 64
    # oracle on a window: rotate is a pure gather, so a cropped destination equals the same rows/cols of the full one
    s = FrameBatch(FMT.RGB24, w, h, 1); s.upload(src.numpy())
    ref = FrameBatch(FMT.RGB24, w, 8, 1)
    # oracle computes full rows for y in [0,8) by rotating with the destination restricted to 8 rows
    si = s.image(); ri = ref.image(); ri.height = 8
    si2 = s.image()
    orc.orc().orc_rotate(C.byref(si2), C.byref(ri), 30.0, -282.7688, 1104.6926, 1)
    assert np.array_equal(ref.plane_view(ref.numpy(), 0, 0), v[:8])


@pytest.mark.parametrize("fmt", FMTS)
@pytest.mark.parametrize("w,h", [(64, 48), (33, 17), (640, 360), (9, 9)])
def test_gaussian(dev, fmt, w, h):
    src, ds = pair(fmt, w, h, dev, 2)
    for kw, kh, sx, sy, border in ((5, 5, 1.1, 1.1, BORDER.REFLECT101), (3, 3, 0.0, 0.0, BORDER.CONSTANT), (7, 3, 2.0, 0.0, BORDER.REPLICATE),
                                   (3, 9, 0.7, 1.9, BORDER.REFLECT), (5, 5, 1.0, 1.0, BORDER.WRAP), (1, 1, 0.0, 0.0, BORDER.CONSTANT)):
        if kw > w or kh > h:
            continue
        dd = FrameBatch(fmt, w, h, 2, device=dev); g.gaussian(ds, dd, kw, kh, sx, sy, border); torch.cuda.synchronize()
        ref = FrameBatch(fmt, w, h, 2); s, d = src.image(), ref.image()
        orc.orc().orc_gaussian(C.byref(s), C.byref(d), kw, kh, sx, sy, border)
        assert_same(dd, ref, f"gauss {kw}x{kh} s{sx},{sy} b{border} {w}x{h}")
        if kw == 1 and kh == 1:
            assert torch.equal(dd.buf, ds.buf)


@pytest.mark.parametrize("fmt", FMTS)
@pytest.mark.parametrize("w,h", [(64, 48), (33, 17), (640, 360)])
def test_median(dev, fmt, w, h):
    src, ds = pair(fmt, w, h, dev, 2)
    for kw, kh in ((3, 3), (5, 5), (1, 1), (3, 5), (7, 7)):
        dd = FrameBatch(fmt, w, h, 2, device=dev); g.median(ds, dd, kw, kh); torch.cuda.synchronize()
        ref = FrameBatch(fmt, w, h, 2); s, d = src.image(), ref.image()
        orc.orc().orc_median(C.byref(s), C.byref(d), kw, kh)
        assert_same(dd, ref, f"median {kw}x{kh} {w}x{h}")


@pytest.mark.parametrize("fmt", FMTS)
@pytest.mark.parametrize("w,h,n", [(16, 3, 1), (48, 17, 3), (256, 67, 2), (1920, 1080, 2), (3840, 130, 1)])
def test_median3_stream_kernel(dev, fmt, w, h, n):
    """the streaming 3x3 kernel (row bytes % 16 == 0): band seams, odd heights, frame edges, batches"""
    src, ds = pair(fmt, w, h, dev, n)
    dd = FrameBatch(fmt, w, h, n, device=dev); g.median(ds, dd, 3, 3); torch.cuda.synchronize()
    ref = FrameBatch(fmt, w, h, n); s, d = src.image(), ref.image()
    orc.orc().orc_median(C.byref(s), C.byref(d), 3, 3)
    assert_same(dd, ref, f"median3 stream {w}x{h}x{n}")


@pytest.mark.parametrize("fmt", FMTS)
@pytest.mark.parametrize("w,h,n", [(8, 5, 1), (16, 5, 1), (24, 9, 2), (40, 17, 3), (48, 6, 1), (256, 67, 2), (1920, 1080, 2), (3840, 130, 1), (36, 21, 1)])
def test_median5_stream_kernel(dev, fmt, w, h, n):
    """the streaming 5x5 kernel (W % 8 == 0: 8 pixels per thread, else W % 4 == 0: 4): one-thread rows (both frame edges in
    one strip), band seams, heights below the window, batches; (36, 21) takes the 4-pixel instantiation"""
    src, ds = pair(fmt, w, h, dev, n)
    dd = FrameBatch(fmt, w, h, n, device=dev); g.median(ds, dd, 5, 5); torch.cuda.synchronize()
    ref = FrameBatch(fmt, w, h, n); s, d = src.image(), ref.image()
    orc.orc().orc_median(C.byref(s), C.byref(d), 5, 5)
    assert_same(dd, ref, f"median5 stream {w}x{h}x{n}")


@pytest.mark.parametrize("fmt", FMTS)
def test_median5_stream_kernel_low_entropy(dev, fmt):
    """ties everywhere: 2-level and 4-level content (the min/max blocks must not depend on distinct samples)"""
    w, h = 64, 40
    for levels in (2, 4):
        src = FrameBatch(fmt, w, h, 2)
        src.buf[:] = np.random.default_rng(levels).integers(0, levels, src.buf.size, dtype=np.uint8) * np.uint8(255 // (levels - 1))
        ds = src.to(dev)
        dd = FrameBatch(fmt, w, h, 2, device=dev); g.median(ds, dd, 5, 5); torch.cuda.synchronize()
        ref = FrameBatch(fmt, w, h, 2); s, d = src.image(), ref.image()
        orc.orc().orc_median(C.byref(s), C.byref(d), 5, 5)
        assert_same(dd, ref, f"median5 stream levels={levels}")


def test_constant_image_is_a_fixed_point(dev):
    w, h = 1920, 1080
    src = FrameBatch(FMT.RGB24, w, h, 1, device=dev); src.buf.fill_(0)
    img = src.plane_view(np.zeros(src.frame_bytes, np.uint8), 0, 0)
    host = np.zeros(src.frame_bytes, np.uint8); src.plane_view(host, 0, 0)[...] = 137; src.upload(host)
    for fn in (lambda s, d: g.gaussian(s, d, 5, 5, 1.1, 1.1, BORDER.REFLECT101), lambda s, d: g.median(s, d, 5, 5),
               lambda s, d: g.flip(s, d, -1)):
        dd = FrameBatch(FMT.RGB24, w, h, 1, device=dev); fn(src, dd); torch.cuda.synchronize()
        assert np.array_equal(dd.payload(), src.payload())
