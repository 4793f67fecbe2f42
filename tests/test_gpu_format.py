"""format_cuda kernels: ours vs the reference's own format_cuda_kernel.cu compiled
unmodified for sm_100a (oracle O3, live on the GPU), vs the CPU restatement, vs committed golden vectors."""
import os

import numpy as np
import pytest
import torch

import gmat_b200 as g
import orc
from gmat_b200 import FMT, FrameBatch
from gpu_util import REF, assert_same, o3_run
from test_oracle_format import rand_rgbpf32

pytestmark = pytest.mark.gpu
HAVE_O3 = os.path.exists(os.path.join(REF, "libref_format_cuda.so"))
# AVColorSpace values: 1 BT709, 2 unspecified, 4 FCC, 5 BT470BG, 6 SMPTE170M, 7 SMPTE240M, 9 BT2020_NCL
SPACES = [2, 1, 5, 6, 4, 7, 9]


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.mark.skipif(not HAVE_O3, reason="oracle/_ref/libref_format_cuda.so not built")
@pytest.mark.parametrize("cs", SPACES)
@pytest.mark.parametrize("w,h", [(64, 48), (1920, 1080), (16, 2), (260, 34)])
def test_nv12_to_rgbpf32_vs_reference_live(dev, cs, w, h):
    src = FrameBatch(FMT.NV12, w, h, 1); src.fill_lcg(seed=cs * 131 + w)
    ds = src.to(dev)
    ours = FrameBatch(FMT.RGBPF32LE, w, h, 1, device=dev); ref = FrameBatch(FMT.RGBPF32LE, w, h, 1, device=dev)
    g.format_nv12_to_rgbpf32(ds, ours, cs); torch.cuda.synchronize()
    o3_run("nv12_to_rgbpf32", ds, ref, cs)
    assert_same(ours, ref, f"nv12->rgbpf32 cs={cs} {w}x{h}")
    cpu = FrameBatch(FMT.RGBPF32LE, w, h, 1)
    if w * h <= 64 * 48:
        orc.format_nv12_to_rgbpf32(src, cpu, cs)
        assert_same(ours, cpu, "vs CPU restatement")


@pytest.mark.skipif(not HAVE_O3, reason="oracle/_ref/libref_format_cuda.so not built")
@pytest.mark.parametrize("kind,bgr", [("nv12_to_rgbpf32_shift", False), ("nv12_to_bgrpf32_shift", True)])
def test_nv12_to_planar_shift_norm_vs_reference_live(dev, kind, bgr):
    w, h = 320, 180
    src = FrameBatch(FMT.NV12, w, h, 1); src.fill_lcg(seed=99)
    ds = src.to(dev)
    ours = FrameBatch(FMT.RGBPF32LE, w, h, 1, device=dev); ref = FrameBatch(FMT.RGBPF32LE, w, h, 1, device=dev)
    shift = (123.675, 116.28, 103.53)
    g.format_nv12_to_rgbpf32(ds, ours, 1, norm=58.395, shift=shift, bgr_planes=bgr); torch.cuda.synchronize()
    o3_run(kind, ds, ref, 1, norm=58.395, shift=shift)
    assert_same(ours, ref, kind)


@pytest.mark.skipif(not HAVE_O3, reason="oracle/_ref/libref_format_cuda.so not built")
@pytest.mark.parametrize("cs", SPACES)
@pytest.mark.parametrize("w,h,wide", [(64, 48, False), (1920, 1080, False), (16, 2, False), (260, 34, True), (3840, 2160, True)])
def test_rgbpf32_to_nv12_vs_reference_live(dev, cs, w, h, wide):
    src = rand_rgbpf32(w, h, 1, seed=cs * 7 + w, wide=wide)
    ds = src.to(dev)
    ours = FrameBatch(FMT.NV12, w, h, 1, device=dev); ref = FrameBatch(FMT.NV12, w, h, 1, device=dev)
    g.format_rgbpf32_to_nv12(ds, ours, cs); torch.cuda.synchronize()
    o3_run("rgbpf32_to_nv12", ds, ref, cs)
    assert_same(ours, ref, f"rgbpf32->nv12 cs={cs} {w}x{h}")
    if w * h <= 260 * 34:
        cpu = FrameBatch(FMT.NV12, w, h, 1)
        orc.format_rgbpf32_to_nv12(src, cpu, cs)
        assert_same(ours, cpu, "vs CPU restatement")


def test_format_roundtrip_and_batch(dev):
    """size-independent property at full size: NV12 -> RGBPF32 -> NV12 of a grey ramp is the identity on luma away
    from the clamps, and a batch equals its frames converted one by one"""
    w, h, n = 3840, 2160, 3
    src = FrameBatch(FMT.NV12, w, h, n); src.fill_lcg(seed=5)
    ds = src.to(dev)
    mid = FrameBatch(FMT.RGBPF32LE, w, h, n, device=dev)
    g.format_nv12_to_rgbpf32(ds, mid, 2)
    back = FrameBatch(FMT.NV12, w, h, n, device=dev)
    g.format_rgbpf32_to_nv12(mid, back, 2); torch.cuda.synchronize()
    for i in range(n):
        one_mid = FrameBatch(FMT.RGBPF32LE, w, h, 1, device=dev)
        s1 = FrameBatch(FMT.NV12, w, h, 1, device=dev); s1.buf.copy_(ds.buf[i * ds.frame_bytes:(i + 1) * ds.frame_bytes])
        g.format_nv12_to_rgbpf32(s1, one_mid, 2); torch.cuda.synchronize()
        a = mid.buf[i * mid.frame_bytes:(i + 1) * mid.frame_bytes].cpu().numpy()
        assert np.array_equal(a, one_mid.buf.cpu().numpy())
    # grey frame: U = V = 128 -> R = G = B = trunc(1.164 (Y - 16)), and back to within 1 code of Y for 16 <= Y <= 235
    grey = FrameBatch(FMT.NV12, 256, 16, 1)
    host = grey.numpy()
    grey.plane_view(host, 0, 0)[...] = np.arange(256, dtype=np.uint8)[None, :]
    grey.plane_view(host, 0, 1)[...] = 128
    dg = grey.to(dev)
    m2 = FrameBatch(FMT.RGBPF32LE, 256, 16, 1, device=dev); b2 = FrameBatch(FMT.NV12, 256, 16, 1, device=dev)
    g.format_nv12_to_rgbpf32(dg, m2, 2); g.format_rgbpf32_to_nv12(m2, b2, 2); torch.cuda.synchronize()
    y = b2.plane_view(b2.numpy(), 0, 0)[0].astype(int)
    assert np.all(np.abs(y[16:236] - np.arange(16, 236)) <= 1)


def test_format_golden_vectors(dev):
    path = os.path.join(os.path.dirname(__file__), "golden", "reference_format_cuda_golden.npz")
    if not os.path.exists(path):
        pytest.skip("golden vectors not generated yet")
    G = np.load(path)
    for cs in (2, 5):
        src = FrameBatch(FMT.NV12, 64, 48, 1); src.fill_lcg(seed=1000 + cs)
        out = FrameBatch(FMT.RGBPF32LE, 64, 48, 1, device=dev)
        g.format_nv12_to_rgbpf32(src.to(dev), out, cs); torch.cuda.synchronize()
        assert np.array_equal(out.payload(), G[f"nv12_to_rgbpf32_cs{cs}"])
        f = rand_rgbpf32(64, 48, 1, seed=2000 + cs, wide=(cs == 5))
        o2 = FrameBatch(FMT.NV12, 64, 48, 1, device=dev)
        g.format_rgbpf32_to_nv12(f.to(dev), o2, cs); torch.cuda.synchronize()
        assert np.array_equal(o2.payload(), G[f"rgbpf32_to_nv12_cs{cs}"])
