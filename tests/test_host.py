"""Host-side logic that needs no GPU: frame layouts, synthetic generator, argument checking."""
import numpy as np
import pytest

import gmat_b200 as g
from gmat_b200 import FMT, FrameBatch, plane_layout, lcg_bytes


def test_layouts_follow_the_cuda_frame_pool():
    planes, size = plane_layout(FMT.NV12, 3840, 2160)
    assert planes[0] == (0, 3840, 2160, 3840)
    assert planes[1] == (3840 * 2160, 3840, 1080, 3840)      # UV at Y + H*pitch, luma pitch
    assert size == 3840 * 2160 * 3 // 2
    planes, _ = plane_layout(FMT.RGB24, 1920, 1080)
    assert planes[0][1] == 5888 and planes[0][3] == 5760      # pitch aligned to 256
    planes, _ = plane_layout(FMT.P010LE, 7680, 4320)
    assert planes[0][1] == 15360 and planes[1][3] == 15360
    planes, _ = plane_layout(FMT.YUV420P, 33, 17)
    assert [p[3] for p in planes] == [33, 17, 17] and [p[2] for p in planes] == [17, 9, 9]


def test_algorithmic_bytes_of_the_headline_config():
    s = FrameBatch(FMT.NV12, 3840, 2160, 1); d = FrameBatch(FMT.RGB24, 1920, 1080, 1)
    assert s.payload().size + d.payload().size == 18662400          # 2.25 B / src px


def test_lcg_is_the_survey_generator():
    s = 0xC0FFEE
    exp = []
    for _ in range(70000):
        s = (s * 1664525 + 1013904223) & 0xFFFFFFFF
        exp.append(s >> 24)
    assert np.array_equal(lcg_bytes(70000, 0xC0FFEE), np.array(exp, np.uint8))


def test_ten_bit_fill_is_msb_aligned():
    b = FrameBatch(FMT.P010LE, 16, 8, 1); host = b.fill_lcg(seed=3)
    y = b.plane_view(host, 0, 0).view(np.uint16)
    assert np.all(y & 0x3F == 0) and y.max() > 1024


def test_batch_descriptor():
    b = FrameBatch(FMT.RGB24, 64, 48, 5)
    img = b.image(first=2, count=3)
    assert img.batch == 3 and img.batch_stride[0] == b.frame_bytes
    assert img.data[0] == b.ptr + 2 * b.frame_bytes


@pytest.mark.parametrize("args", [
    (0, 48, FMT.NV12, 32, 24, FMT.RGB24),           # bad size
    (64, 48, 9999, 32, 24, FMT.RGB24),              # unknown format
    (64, 48, FMT.RGB24, 64, 48, FMT.RGBA),          # unscaled rgb->rgb other than 24<->24 swap
])
def test_unsupported_contexts_return_null(args):
    with pytest.raises(g.GmatbError):
        g.SwsContext(*args)


def test_filter_argument_errors_are_negative_codes():
    L = g.lib()
    src = FrameBatch(FMT.RGB24, 64, 48, 1); dst = FrameBatch(FMT.RGB24, 32, 24, 1)
    with pytest.raises(g.GmatbError):
        g.crop(src, dst, x=40, y=0)                 # window outside the frame (vf_crop_nvcv.c:151-154)
    with pytest.raises(g.GmatbError):
        g.flip(src, dst, 0)                         # geometry mismatch
    same = FrameBatch(FMT.RGB24, 64, 48, 1)
    with pytest.raises(g.GmatbError):
        g.gaussian(src, same, kw=4, kh=3)           # even kernel
    with pytest.raises(g.GmatbError):
        g.gaussian(src, same, kw=65, kh=3)          # kernel larger than the image (vf_smooth_nvcv.c:172-175)
    with pytest.raises(g.GmatbError):
        g.rotate(src, same, interp=7)
    nv = FrameBatch(FMT.NV12, 64, 48, 1)
    with pytest.raises(g.GmatbError):
        g.rotate(nv, nv)                            # filters take packed rgb only (vf_rotate_nvcv.c:92-101)
    assert L.gmatb_launch_count() >= 0


def test_format_cuda_argument_errors_and_colourspace_map():
    """format_cuda entry points (include/gmat_b200.h): argument checks happen before any CUDA call"""
    L = g.lib()
    # GetConstants (format_cuda_kernel.cu:32-63): BT.709 is the default branch, SMPTE170M falls into it
    assert [L.gmatb_format_colorspace(c) for c in (0, 1, 2, 4, 5, 6, 7, 9, 10)] == [1, 1, 1, 4, 5, 1, 7, 9, 9]
    f = FrameBatch(FMT.RGBPF32LE, 64, 48, 1); n = FrameBatch(FMT.NV12, 64, 48, 1)
    with pytest.raises(g.GmatbError):
        g.format_rgbpf32_to_nv12(n, n)                                   # source must be planar float
    with pytest.raises(g.GmatbError):
        g.format_rgbpf32_to_nv12(f, FrameBatch(FMT.NV12, 32, 48, 1))     # geometry mismatch
    with pytest.raises(g.GmatbError):
        g.format_rgbpf32_to_nv12(FrameBatch(FMT.RGBPF32LE, 66, 48, 1), FrameBatch(FMT.NV12, 66, 48, 1))   # width % 4
    with pytest.raises(g.GmatbError):
        g.format_nv12_to_rgbpf32(f, f)                                   # source must be NV12
    with pytest.raises(g.GmatbError):
        g.median(FrameBatch(FMT.RGB24, 64, 48, 1), FrameBatch(FMT.RGB24, 64, 48, 1), 3, 49)   # window taller than the frame
