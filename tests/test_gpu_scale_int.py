"""-m gpu: the exact-integer form of the fused 2:1 kernel (scale_fused4i.cuh, opt-in: SWS.INT_CHAIN) against the float-chain
kernel (scale_fused3.cuh, itself pinned to the reference's kernels in test_gpu_scale.py) and the CPU
oracle, on content that exercises each of its three regimes: no ambiguous outputs, a few per warp
step (shared-memory queue + per-output float recomputation), many (the band continues in float)."""
import numpy as np
import pytest
import torch

import orc
from gmat_b200 import FMT, SWS, FrameBatch, SwsContext
from gpu_util import assert_same

pytestmark = pytest.mark.gpu
HW = SWS.HWACCEL_CUDA
PARAMS = [(0.75,), (0.5,), (1.0,)]
CONTENT = ["noise", "flat", "ramp", "halfflat", "leftflat", "patches", "gray", "black", "white", "lowamp", "stripes"]


def content(fb, kind, seed):
    """host bytes of `fb` (8-bit NV12 / I420) filled with `kind`"""
    rng = np.random.default_rng(seed)
    host = np.zeros(fb.frame_bytes * fb.n, np.uint8)
    for f in range(fb.n):
        for p in range(len(fb.planes)):
            v = fb.plane_view(host, f, p)
            r, c = v.shape
            noise = rng.integers(0, 256, size=(r, c), dtype=np.uint8)
            if kind == "noise":
                a = noise
            elif kind == "flat":
                a = np.full((r, c), [90, 110, 140][p % 3] + 7 * f, np.uint8)
            elif kind == "black":
                a = np.full((r, c), 16 if p == 0 else 128, np.uint8)
            elif kind == "white":
                a = np.full((r, c), 235 if p == 0 else 128, np.uint8)
            elif kind == "ramp":
                a = ((np.arange(c)[None, :] // (1 if p == 0 else 2 if len(fb.planes) == 2 else 1) + np.arange(r)[:, None] * 2) & 255).astype(np.uint8)
            elif kind == "halfflat":      # noise above, flat below: a band switches to float half way
                a = noise.copy(); a[r // 2 + 3:, :] = 77 + 20 * p
            elif kind == "leftflat":      # flat left third: some warps dense, their neighbours not
                a = noise.copy(); a[:, : c // 3] = 60 + 30 * p
            elif kind == "patches":       # small flat patches in noise: a handful of ambiguous outputs per step
                a = noise.copy()
                for _ in range(max(4, r * c // 4096)):
                    y0, x0 = int(rng.integers(0, max(1, r - 6))), int(rng.integers(0, max(1, c - 12)))
                    a[y0:y0 + int(rng.integers(2, 7)), x0:x0 + int(rng.integers(4, 13))] = int(rng.integers(0, 256))
            elif kind == "gray":
                a = noise if p == 0 else np.full((r, c), 128, np.uint8)
            elif kind == "lowamp":
                a = (100 + (noise & 3)).astype(np.uint8)
            elif kind == "stripes":       # vertical stripes of even period: N = 0 (mod 1024) on whole columns
                a = np.where((np.arange(c)[None, :] // 4) % 2 == 0, 40, 200).astype(np.uint8) * np.ones((r, 1), np.uint8)
            v[...] = a
    return host


def run_pair(dev, sfmt, dfmt, sw, sh, n, kind, param, wrap, seed=1, chain=None):
    chain = SWS.INT_CHAIN if chain is None else chain
    dw, dh = sw // 2, sh // 2
    src = FrameBatch(sfmt, sw, sh, n)
    host = content(src, kind, seed)
    src.upload(host)
    ds = src.to(dev)
    fl = SWS.BICUBIC | HW | (SWS.PARITY_WRAP if wrap else 0)
    ci = SwsContext(sw, sh, sfmt, dw, dh, dfmt, fl | chain, param)
    cf = SwsContext(sw, sh, sfmt, dw, dh, dfmt, fl, param)
    assert ci.path == 1 and cf.path == 1
    a = FrameBatch(dfmt, dw, dh, n, device=dev); b = FrameBatch(dfmt, dw, dh, n, device=dev)
    a.buf.fill_(0xA5); b.buf.fill_(0xA5)
    ci.scale(ds, a); cf.scale(ds, b); torch.cuda.synchronize()
    return src, ci, a, b


@pytest.mark.parametrize("kind", CONTENT)
@pytest.mark.parametrize("param", PARAMS)
@pytest.mark.parametrize("sw,sh", [(64, 48), (16, 4), (8, 2), (512, 130), (240, 12), (248, 140), (488, 66), (1920, 1080)])
def test_int_equals_float_chain(dev, kind, param, sw, sh):
    for sfmt, dfmt, wrap in ((FMT.NV12, FMT.RGB24, False), (FMT.YUV420P, FMT.BGRA, True), (FMT.NV12, FMT.BGR24, True), (FMT.YUV420P, FMT.RGBA, False)):
        _, _, a, b = run_pair(dev, sfmt, dfmt, sw, sh, 2, kind, param, wrap, seed=sw + sh)
        if not torch.equal(a.buf, b.buf):
            assert_same(a, b, f"int vs float chain {kind} {param} {sfmt}->{dfmt} wrap={wrap} {sw}x{sh}")


@pytest.mark.parametrize("kind", ["noise", "patches", "halfflat", "stripes", "lowamp"])
@pytest.mark.parametrize("param", PARAMS)
def test_int_vs_oracle(dev, kind, param):
    """the integer kernel against the CPU restatement of the reference's float chain"""
    for sfmt, dfmt in ((FMT.NV12, FMT.RGB24), (FMT.YUV420P, FMT.BGRA)):
        src, ci, a, _ = run_pair(dev, sfmt, dfmt, 264, 72, 2, kind, param, False, seed=5)
        ref = FrameBatch(dfmt, 132, 36, 2)
        orc.yuv2rgb_scale(src, ref, (ci.get_filter(0), ci.get_filter(1)))
        assert_same(a, ref, f"int kernel vs oracle {kind} {param} {sfmt}->{dfmt}")


@pytest.mark.parametrize("kind", ["noise", "patches", "halfflat", "leftflat", "flat", "stripes"])
def test_int_headline_4k(dev, kind):
    """BASELINE C2 at full size, 3 frames"""
    _, _, a, b = run_pair(dev, FMT.NV12, FMT.RGB24, 3840, 2160, 3, kind, (0.75,), False, seed=11)
    assert torch.equal(a.buf, b.buf), f"4K {kind}: {(a.buf != b.buf).sum().item()} bytes differ"


def test_int_selected_only_for_dyadic_weights(dev):
    """other parameters keep the float chain (INT_CHAIN is then a no-op)"""
    for param in ((0.6,), (0.3,), None):
        _, _, a, b = run_pair(dev, FMT.NV12, FMT.RGB24, 512, 130, 1, "noise", param, False)
        assert torch.equal(a.buf, b.buf)


# ---- the tensor-pipe form (scale_fused5m.cuh, SWS.MMA_CHAIN): NV12 sources, horizontal pass on IMMA ----------------------
@pytest.mark.parametrize("kind", CONTENT)
@pytest.mark.parametrize("param", PARAMS)
@pytest.mark.parametrize("sw,sh", [(64, 48), (16, 4), (8, 2), (512, 130), (240, 12), (248, 140), (488, 66), (1920, 1080), (3848, 34)])
def test_mma_equals_float_chain(dev, kind, param, sw, sh):
    for dfmt, wrap in ((FMT.RGB24, False), (FMT.BGRA, True), (FMT.BGR24, True), (FMT.RGBA, False)):
        _, _, a, b = run_pair(dev, FMT.NV12, dfmt, sw, sh, 2, kind, param, wrap, seed=sw + sh, chain=SWS.MMA_CHAIN)
        if not torch.equal(a.buf, b.buf):
            assert_same(a, b, f"mma vs float chain {kind} {param} NV12->{dfmt} wrap={wrap} {sw}x{sh}")


@pytest.mark.parametrize("kind", ["noise", "patches", "halfflat", "stripes", "lowamp"])
@pytest.mark.parametrize("param", PARAMS)
def test_mma_vs_oracle(dev, kind, param):
    """the tensor-pipe kernel against the CPU restatement of the reference's float chain"""
    for dfmt in (FMT.RGB24, FMT.BGRA):
        src, ci, a, _ = run_pair(dev, FMT.NV12, dfmt, 264, 72, 2, kind, param, False, seed=5, chain=SWS.MMA_CHAIN)
        ref = FrameBatch(dfmt, 132, 36, 2)
        orc.yuv2rgb_scale(src, ref, (ci.get_filter(0), ci.get_filter(1)))
        assert_same(a, ref, f"mma kernel vs oracle {kind} {param} NV12->{dfmt}")


@pytest.mark.parametrize("kind", ["noise", "patches", "halfflat", "leftflat", "flat", "stripes"])
def test_mma_headline_4k(dev, kind):
    """BASELINE C2 at full size, 3 frames"""
    _, _, a, b = run_pair(dev, FMT.NV12, FMT.RGB24, 3840, 2160, 3, kind, (0.75,), False, seed=11, chain=SWS.MMA_CHAIN)
    assert torch.equal(a.buf, b.buf), f"4K {kind}: {(a.buf != b.buf).sum().item()} bytes differ"
