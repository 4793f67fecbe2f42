"""-m gpu: the any-ratio streaming kernel (scale_stream.cuh: 8-bit yuv 4:2:0 -> 8-bit packed rgb, R-B arithmetic) against
the shared-memory tile kernel it replaced on that path (scale_generic.cuh, kept behind SWS.TILE_KERNEL and itself pinned to
the reference's kernels and the CPU oracle in test_gpu_scale.py), against the CPU oracle, and -- 1080p -> 720p, the
headline parameter -- against the reference's own two-kernel pipeline (O1 + O2) live."""
import numpy as np
import pytest
import torch

import orc
from gmat_b200 import FMT, SWS, FrameBatch, SwsContext
from gpu_util import assert_same

pytestmark = pytest.mark.gpu
HW = SWS.HWACCEL_CUDA
ALGOS = [("bicubic", SWS.BICUBIC, None), ("bicubic", SWS.BICUBIC, (0.75,)), ("lanczos", SWS.LANCZOS, None),
         ("bilinear", SWS.BILINEAR, None), ("nearest", SWS.POINT, None)]       # the last two: R-A arithmetic (what the reference executes)
# (source, destination): 1.5:1, 3:1, 1:2, non-uniform, odd sizes, narrow strips, more than one warp per row, a huge reduction
SIZES = [(1920, 1080, 1280, 720), (960, 540, 320, 180), (320, 180, 640, 360), (64, 48, 40, 30), (33, 17, 50, 29), (16, 16, 7, 5),
         (100, 60, 12, 7), (62, 46, 31, 23), (250, 34, 1000, 35), (1000, 36, 250, 72), (527, 63, 333, 40), (3840, 16, 1280, 6),
         (720, 50, 24, 50), (24, 50, 720, 50), (1928, 22, 1286, 15), (18, 2, 7, 3)]


def run(dev, sfmt, dfmt, sw, sh, dw, dh, flag, param, n=2, extra=0, seed=1):
    src = FrameBatch(sfmt, sw, sh, n); src.fill_lcg(seed=seed)
    c = SwsContext(sw, sh, sfmt, dw, dh, dfmt, flag | HW | extra, param)
    ds = src.to(dev); dd = FrameBatch(dfmt, dw, dh, n, device=dev)
    dd.buf.fill_(0xA5)
    c.scale(ds, dd); torch.cuda.synchronize()
    return src, c, dd


@pytest.mark.parametrize("name,flag,param", ALGOS)
@pytest.mark.parametrize("sw,sh,dw,dh", SIZES)
def test_stream_equals_tile_kernel(dev, name, flag, param, sw, sh, dw, dh):
    for sfmt, dfmt, wrap in ((FMT.NV12, FMT.RGB24, 0), (FMT.YUV420P, FMT.BGRA, SWS.PARITY_WRAP), (FMT.NV12, FMT.BGR24, SWS.PARITY_WRAP), (FMT.YUV420P, FMT.RGBA, 0)):
        _, _, a = run(dev, sfmt, dfmt, sw, sh, dw, dh, flag | wrap, param, seed=sw + dh)
        _, _, b = run(dev, sfmt, dfmt, sw, sh, dw, dh, flag | wrap, param, extra=SWS.TILE_KERNEL, seed=sw + dh)
        if not torch.equal(a.buf, b.buf):
            assert_same(a, b, f"stream vs tile kernel {name}{param} {sfmt}->{dfmt} {sw}x{sh}->{dw}x{dh}")


@pytest.mark.parametrize("name,flag,param", ALGOS)
@pytest.mark.parametrize("sw,sh,dw,dh", [(64, 48, 40, 30), (33, 17, 50, 29), (16, 16, 7, 5), (100, 60, 12, 7), (264, 72, 176, 48), (96, 40, 300, 90)])
def test_stream_vs_oracle(dev, name, flag, param, sw, sh, dw, dh):
    for sfmt, dfmt in ((FMT.NV12, FMT.RGB24), (FMT.YUV420P, FMT.BGRA)):
        src, c, dd = run(dev, sfmt, dfmt, sw, sh, dw, dh, flag, param, n=1, seed=sw * dh)
        ref = FrameBatch(dfmt, dw, dh, 1); orc.yuv2rgb_scale(src, ref, (c.get_filter(0), c.get_filter(1)), ra=1 if name in ("bilinear", "nearest") else 0)
        assert_same(dd, ref, f"stream kernel vs oracle {name}{param} {sfmt}->{dfmt} {sw}x{sh}->{dw}x{dh}")


@pytest.mark.parametrize("sw,sh,dw,dh", [(3840, 2160, 1280, 720), (1920, 1080, 3840, 2160), (1920, 1080, 1280, 720)])
def test_stream_full_sizes(dev, sw, sh, dw, dh):
    """the three ratios VERDICT r1 names, at full size, 2 frames, headline parameter and bilinear"""
    for flag, param in ((SWS.BICUBIC, (0.75,)), (SWS.BILINEAR, None)):
        _, _, a = run(dev, FMT.NV12, FMT.RGB24, sw, sh, dw, dh, flag, param, seed=7)
        _, _, b = run(dev, FMT.NV12, FMT.RGB24, sw, sh, dw, dh, flag, param, extra=SWS.TILE_KERNEL, seed=7)
        assert torch.equal(a.buf, b.buf), f"{(a.buf != b.buf).sum().item()} bytes differ"


# ---- yuv -> yuv: every plane through the plane form of the streaming kernel (what the scale_cuda filter runs) -----------
@pytest.mark.parametrize("name,flag,param", ALGOS)
@pytest.mark.parametrize("sw,sh,dw,dh", [(1920, 1080, 1280, 720), (3840, 2160, 1920, 1080), (640, 360, 1280, 720), (64, 48, 40, 30), (66, 34, 100, 58),
                                         # exactly 2:1 (scale_plane2.cuh where the plane width is a multiple of 8; warp seams at 30 / 31 / 29 / 61 strips)
                                         (64, 48, 32, 24), (16, 4, 8, 2), (512, 132, 256, 66), (480, 12, 240, 6), (496, 12, 248, 6), (464, 72, 232, 36), (976, 8, 488, 4), (248, 12, 124, 6), (1920, 1080, 960, 540),
                                         (32, 32, 14, 10), (200, 120, 24, 14), (500, 68, 2000, 70), (2000, 72, 500, 144), (1054, 126, 666, 80),
                                         (3840, 32, 1280, 12), (1440, 100, 48, 100), (48, 100, 1440, 100), (36, 4, 14, 6)])
def test_plane_stream_equals_tile_kernel(dev, name, flag, param, sw, sh, dw, dh):
    for fmt, wrap in ((FMT.NV12, 0), (FMT.YUV420P, SWS.PARITY_WRAP), (FMT.P016LE, 0), (FMT.YUV420P16LE, SWS.PARITY_WRAP), (FMT.P010LE, 0)):
        _, _, a = run(dev, fmt, fmt, sw, sh, dw, dh, flag | wrap, param, seed=sw + dh)
        _, _, b = run(dev, fmt, fmt, sw, sh, dw, dh, flag | wrap, param, extra=SWS.TILE_KERNEL, seed=sw + dh)
        if not torch.equal(a.buf, b.buf):
            assert_same(a, b, f"plane stream vs tile kernel {name}{param} {fmt} {sw}x{sh}->{dw}x{dh}")


# ---- 4-component packed rgb (rgb0 / bgra ...) -> same format: the scale_cuda filter on rgb frames -------------------------
@pytest.mark.parametrize("name,flag,param", ALGOS)
@pytest.mark.parametrize("sw,sh,dw,dh", [(1920, 1080, 1280, 720), (1920, 1080, 960, 540), (640, 360, 1280, 720), (64, 48, 40, 30), (33, 17, 50, 29),
                                         (64, 48, 32, 24), (16, 4, 8, 2), (480, 12, 240, 6), (496, 12, 248, 6), (100, 60, 12, 7), (250, 34, 1000, 35), (18, 2, 7, 3)])
def test_rgb4_stream_equals_tile_kernel(dev, name, flag, param, sw, sh, dw, dh):
    for fmt, wrap in ((FMT.RGB0, SWS.PARITY_WRAP), (FMT.BGRA, 0), (FMT.RGB24, 0), (FMT.BGR24, SWS.PARITY_WRAP)):
        _, _, a = run(dev, fmt, fmt, sw, sh, dw, dh, flag | wrap, param, seed=sw + dh)
        _, _, b = run(dev, fmt, fmt, sw, sh, dw, dh, flag | wrap, param, extra=SWS.TILE_KERNEL, seed=sw + dh)
        if not torch.equal(a.buf, b.buf):
            assert_same(a, b, f"rgb stream vs tile kernel {name}{param} {fmt} {sw}x{sh}->{dw}x{dh}")


@pytest.mark.parametrize("name,flag,param", [("bicubic", SWS.BICUBIC, (0.75,)), ("bilinear", SWS.BILINEAR, None)])
@pytest.mark.parametrize("sw,sh,dw,dh", [(1920, 1080, 1280, 720), (1920, 1080, 960, 540), (640, 360, 1280, 720), (64, 48, 40, 30), (250, 34, 1000, 36)])
def test_rgb_to_yuv_prescale_stream_equals_tile_kernel(dev, name, flag, param, sw, sh, dw, dh):
    """rgb -> yuv with scaling = resize the rgb image, then convert (swscale_cuda.c:312-341): the resize through the plane kernels"""
    for sfmt, dfmt in ((FMT.RGB24, FMT.NV12), (FMT.BGRA, FMT.YUV420P)):
        _, _, a = run(dev, sfmt, dfmt, sw, sh, dw, dh, flag, param, seed=sw + dh)
        _, _, b = run(dev, sfmt, dfmt, sw, sh, dw, dh, flag, param, extra=SWS.TILE_KERNEL, seed=sw + dh)
        if not torch.equal(a.buf, b.buf):
            assert_same(a, b, f"rgb->yuv prescale stream vs tile {name}{param} {sfmt}->{dfmt} {sw}x{sh}->{dw}x{dh}")
