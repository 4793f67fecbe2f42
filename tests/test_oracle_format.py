"""CPU: the restatement of the format_cuda kernels (oracle/gmat_oracle.c) against golden vectors produced by
the reference's own format_cuda_kernel.cu on a B200 (tests/golden/make_golden_format.py)."""
import os

import numpy as np
import pytest

import orc
from gmat_b200 import FMT, FrameBatch

GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_format_cuda_golden.npz")


def rand_rgbpf32(w, h, n, seed, wide=False):      # same generator as tests/test_gpu_format.py (kept torch-free here)
    f = FrameBatch(FMT.RGBPF32LE, w, h, n)
    rng = np.random.default_rng(seed)
    host = f.numpy()
    for i in range(n):
        for p in range(3):
            v = f.plane_view(host, i, p).view(np.float32)
            x = rng.integers(0, 256, v.shape).astype(np.float32) / np.float32(255.0)
            if wide:
                x = rng.uniform(-0.2, 1.3, v.shape).astype(np.float32)
            v[...] = x
    return f


def test_format_colorspace_map():
    L = orc.orc()
    # GetConstants (format_cuda_kernel.cu:32-63): BT.709 is the default branch
    assert [L.orc_format_colorspace(c) for c in (0, 1, 2, 4, 5, 6, 7, 9, 10)] == [1, 1, 1, 4, 5, 1, 7, 9, 9]


@pytest.mark.skipif(not os.path.exists(GOLD), reason="golden vectors not generated yet")
@pytest.mark.parametrize("cs", [1, 2, 4, 5, 6, 7, 9])
def test_oracle_vs_reference_format_golden(cs):
    G = np.load(GOLD)
    src = FrameBatch(FMT.NV12, 64, 48, 1); src.fill_lcg(seed=1000 + cs)
    dst = FrameBatch(FMT.RGBPF32LE, 64, 48, 1)
    orc.format_nv12_to_rgbpf32(src, dst, cs)
    assert np.array_equal(dst.payload(), G[f"nv12_to_rgbpf32_cs{cs}"])
    f = rand_rgbpf32(64, 48, 1, seed=2000 + cs, wide=(cs in (5, 9)))
    d2 = FrameBatch(FMT.NV12, 64, 48, 1)
    orc.format_rgbpf32_to_nv12(f, d2, cs)
    assert np.array_equal(d2.payload(), G[f"rgbpf32_to_nv12_cs{cs}"])


@pytest.mark.skipif(not os.path.exists(GOLD), reason="golden vectors not generated yet")
def test_oracle_vs_reference_format_shift_golden():
    G = np.load(GOLD)
    src = FrameBatch(FMT.NV12, 64, 48, 1); src.fill_lcg(seed=77)
    dst = FrameBatch(FMT.RGBPF32LE, 64, 48, 1)
    orc.format_nv12_to_rgbpf32(src, dst, 1, norm=58.395, shift=(123.675, 116.28, 103.53))
    assert np.array_equal(dst.payload(), G["nv12_to_rgbpf32_shift"])
    # BGR plane order = the same values with planes 0 and 2 exchanged
    host = dst.numpy()
    r = np.ascontiguousarray(dst.plane_view(host, 0, 0)).reshape(-1)
    gg = np.ascontiguousarray(dst.plane_view(host, 0, 1)).reshape(-1)
    b = np.ascontiguousarray(dst.plane_view(host, 0, 2)).reshape(-1)
    assert np.array_equal(np.concatenate([b, gg, r]), G["nv12_to_bgrpf32_shift"])


def test_rgbpf32_to_nv12_properties():
    """known answers of the restatement: black/white/grey, the bottom-right-luma defect (:560), odd sizes refused"""
    f = FrameBatch(FMT.RGBPF32LE, 8, 4, 1)
    host = f.numpy()
    for p, val in enumerate((1.0, 0.0, 0.0)):          # pure red
        f.plane_view(host, 0, p).view(np.float32)[...] = val
    d = FrameBatch(FMT.NV12, 8, 4, 1)
    orc.format_rgbpf32_to_nv12(f, d, 2)
    y = d.plane_view(d.numpy(), 0, 0)
    # BT.709 red: Y = trunc(0.2126*219/255*255 + 16) = 62; the bottom-right pixel of each 2x2 block takes b := g = 0 too
    assert y[0, 0] == 62 and y[1, 1] == 62
    for p, val in enumerate((0.0, 0.0, 1.0)):          # pure blue: bottom-right luma loses its blue term
        f.plane_view(host, 0, p).view(np.float32)[...] = val
    orc.format_rgbpf32_to_nv12(f, d, 2)
    y = d.plane_view(d.numpy(), 0, 0)
    assert y[0, 0] == 31 and y[1, 1] == 16 and y[1, 0] == 31
    odd = FrameBatch(FMT.RGBPF32LE, 7, 4, 1)
    m = orc.matrix_rgb2yuv(1)
    import ctypes as C
    s_, d_ = odd.image(), FrameBatch(FMT.NV12, 7, 4, 1).image()
    assert orc.orc().orc_format_rgbpf32_to_nv12(C.byref(s_), C.byref(d_), orc.fptr(m)) != 0
