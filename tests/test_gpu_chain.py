"""-m gpu: BASELINE configs[3] and [4] end to end at full size.
C4: 4K rgb24 rotate(30 deg about the centre, linear) -> gaussian 5x5 (sigma 1.1, reflect101) -> bicubic scale to 1080p,
    each stage materialising its frame (a filtergraph), against the CPU oracle chain on the whole frame.
C5: a mixed 1080p / 4K / 8K NV12 -> RGB24 half-size batch, sharded by cost like bench.py --workload c5: every frame
    equals the frame converted alone, whatever rank / order / batch it ran in, and matches the oracle on a crop."""
import ctypes as C

import numpy as np
import pytest
import torch

import gmat_b200 as g
import orc
from gmat_b200 import BORDER, FMT, SWS, FrameBatch, SwsContext
from gmat_b200.dist import shard_by_cost

pytestmark = pytest.mark.gpu
HW = SWS.HWACCEL_CUDA
C4_ANGLE, C4_SHIFT = 30.0, (-282.7688, 1104.6926)        # shift = c - R^T c for 3840x2160


def video_like(w, h, seed):
    """gradients + an 8x8 checker + noise (SURVEY 8d's second synthetic set), rgb24"""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    base = np.stack([(x * 255 // w), (y * 255 // h), ((x + y) * 255 // (w + h))], -1).astype(np.int32)
    chk = ((((x // 8) + (y // 8)) & 1) * 40)[..., None]
    return np.clip(base + chk + rng.integers(-12, 13, size=(h, w, 3)), 0, 255).astype(np.uint8)


@pytest.mark.parametrize("param", [None, (0.75,)])
def test_c4_chain_4k_vs_oracle_chain(dev, param):
    w, h, dw, dh = 3840, 2160, 1920, 1080
    img = video_like(w, h, 4)
    src = FrameBatch(FMT.RGB24, w, h, 1)
    host = np.zeros(src.frame_bytes, np.uint8); src.plane_view(host, 0, 0)[...] = img.reshape(h, w * 3); src.upload(host)
    a = src.to(dev)
    b = FrameBatch(FMT.RGB24, w, h, 1, device=dev); c = FrameBatch(FMT.RGB24, w, h, 1, device=dev)
    d = FrameBatch(FMT.RGB24, dw, dh, 1, device=dev)
    sc = SwsContext(w, h, FMT.RGB24, dw, dh, FMT.RGB24, SWS.BICUBIC | HW, param)
    g.rotate(a, b, C4_ANGLE, C4_SHIFT[0], C4_SHIFT[1], "linear")
    g.gaussian(b, c, 5, 5, 1.1, 1.1, BORDER.REFLECT101)
    sc.scale(c, d); torch.cuda.synchronize()
    # the oracle chain, stage by stage, on the host
    rb = FrameBatch(FMT.RGB24, w, h, 1); rc = FrameBatch(FMT.RGB24, w, h, 1); rd = FrameBatch(FMT.RGB24, dw, dh, 1)
    s_, d_ = src.image(), rb.image(); orc.orc().orc_rotate(C.byref(s_), C.byref(d_), C4_ANGLE, C4_SHIFT[0], C4_SHIFT[1], 1)
    s_, d_ = rb.image(), rc.image(); orc.orc().orc_gaussian(C.byref(s_), C.byref(d_), 5, 5, 1.1, 1.1, BORDER.REFLECT101)
    (cx, px), (cy, py) = sc.get_filter(0), sc.get_filter(1)
    s_, d_ = rc.image(), rd.image()
    orc.orc().orc_resample_packed(s_.data[0], s_.linesize[0], w, h, d_.data[0], d_.linesize[0], dw, dh, 3, 0,
                                  orc.fptr(cx), orc.iptr(px), orc.fptr(cy), orc.iptr(py), 0, 0)
    for name, got, ref in (("rotate", b, rb), ("gaussian", c, rc), ("scale", d, rd)):
        ga, ra = got.payload(), ref.payload()
        assert np.array_equal(ga, ra), f"C4 stage {name}: {int((ga != ra).sum())} of {ga.size} bytes differ"


def test_c4_chain_batch_is_frame_independent(dev):
    """a batch of 4 frames through the chain in one launch per stage == each frame alone"""
    w, h, dw, dh, n = 3840, 2160, 1920, 1080, 4
    a = FrameBatch(FMT.RGB24, w, h, n, device=dev); a.fill_lcg(seed=44)
    b = FrameBatch(FMT.RGB24, w, h, n, device=dev); c = FrameBatch(FMT.RGB24, w, h, n, device=dev)
    d = FrameBatch(FMT.RGB24, dw, dh, n, device=dev)
    sc = SwsContext(w, h, FMT.RGB24, dw, dh, FMT.RGB24, SWS.BICUBIC | HW)

    def chain(x, y, z, o):
        g.rotate(x, y, C4_ANGLE, C4_SHIFT[0], C4_SHIFT[1], "linear")
        g.gaussian(y, z, 5, 5, 1.1, 1.1, BORDER.REFLECT101)
        sc.scale(z, o)
    chain(a, b, c, d); torch.cuda.synchronize()
    one_in = FrameBatch(FMT.RGB24, w, h, 1, device=dev); t1 = FrameBatch(FMT.RGB24, w, h, 1, device=dev)
    t2 = FrameBatch(FMT.RGB24, w, h, 1, device=dev); one_out = FrameBatch(FMT.RGB24, dw, dh, 1, device=dev)
    for f in (0, 3):
        one_in.buf.copy_(a.buf[f * a.frame_bytes:(f + 1) * a.frame_bytes])
        chain(one_in, t1, t2, one_out); torch.cuda.synchronize()
        assert torch.equal(one_out.buf, d.buf[f * d.frame_bytes:(f + 1) * d.frame_bytes])


def test_c5_mixed_sizes_sharded_by_cost(dev):
    sizes = [(1920, 1080)] * 3 + [(3840, 2160)] * 3 + [(7680, 4320)] * 2
    costs = [w * h for (w, h) in sizes]
    parts = shard_by_cost(costs, 8)
    assert sorted(sum(parts, [])) == list(range(len(sizes)))
    ctxs = {wh: SwsContext(wh[0], wh[1], FMT.NV12, wh[0] // 2, wh[1] // 2, FMT.RGB24, SWS.BICUBIC | HW, (0.75,)) for wh in set(sizes)}
    frames = []
    for i, (w, h) in enumerate(sizes):
        s = FrameBatch(FMT.NV12, w, h, 1, device=dev); s.fill_lcg(seed=500 + i)
        frames.append(s)
    # alone
    alone = []
    for s in frames:
        d = FrameBatch(FMT.RGB24, s.w // 2, s.h // 2, 1, device=dev); ctxs[(s.w, s.h)].scale(s, d); alone.append(d)
    torch.cuda.synchronize()
    # as "ranks" would run them: each rank's list sorted by size, same-size frames batched into one launch
    for part in parts:
        by_size = {}
        for i in part:
            by_size.setdefault(sizes[i], []).append(i)
        for wh, idx in sorted(by_size.items()):
            sb = FrameBatch(FMT.NV12, wh[0], wh[1], len(idx), device=dev)
            for k, i in enumerate(idx):
                sb.buf[k * sb.frame_bytes:(k + 1) * sb.frame_bytes] = frames[i].buf
            db = FrameBatch(FMT.RGB24, wh[0] // 2, wh[1] // 2, len(idx), device=dev)
            ctxs[wh].scale(sb, db); torch.cuda.synchronize()
            for k, i in enumerate(idx):
                assert torch.equal(db.buf[k * db.frame_bytes:(k + 1) * db.frame_bytes], alone[i].buf), (wh, i)
    # oracle on the top-left crop of one frame of every size
    for s, d in ((frames[0], alone[0]), (frames[3], alone[3]), (frames[6], alone[6])):
        cw, ch = 264, 72
        crop = FrameBatch(FMT.NV12, cw, ch, 1)
        hc = np.zeros(crop.frame_bytes, np.uint8); sh_ = s.numpy()
        crop.plane_view(hc, 0, 0)[...] = s.plane_view(sh_, 0, 0)[:ch, :cw]
        crop.plane_view(hc, 0, 1)[...] = s.plane_view(sh_, 0, 1)[:ch // 2, :cw]
        crop.upload(hc)
        cc = SwsContext(cw, ch, FMT.NV12, cw // 2, ch // 2, FMT.RGB24, SWS.BICUBIC | HW, (0.75,))
        ref = FrameBatch(FMT.RGB24, cw // 2, ch // 2, 1); orc.yuv2rgb_scale(crop, ref, (cc.get_filter(0), cc.get_filter(1)))
        got = d.plane_view(d.numpy(), 0, 0)[:ch // 2 - 1, :(cw // 2 - 1) * 3]
        exp = ref.plane_view(ref.numpy(), 0, 0)[:ch // 2 - 1, :(cw // 2 - 1) * 3]
        assert np.array_equal(got, exp), (s.w, s.h)
