"""-m gpu: scaling parity.  The fused 2:1 kernel and the generic kernel against the CPU oracle
(small sizes), against each other (full BASELINE sizes), and against the reference's scale_cuda
kernels (O2) running on the same GPU."""
import ctypes as C
import zlib

import numpy as np
import pytest
import torch

import gmat_b200 as g
import orc
from gmat_b200 import FMT, SWS, FrameBatch, SwsContext
from gpu_util import assert_same, have_ref, o1_run, o2_packed

pytestmark = pytest.mark.gpu
HW = SWS.HWACCEL_CUDA
ALGOS = [("bicubic", SWS.BICUBIC, None), ("bicubic", SWS.BICUBIC, (0.75,)), ("bicubic", SWS.BICUBIC, (0.5,)),
         ("lanczos", SWS.LANCZOS, None)]


def tables(c):
    return c.get_filter(0), c.get_filter(1)


@pytest.mark.parametrize("name,flag,param", ALGOS)
@pytest.mark.parametrize("sw,sh,dw,dh", [(64, 48, 32, 24), (16, 4, 8, 2), (8, 2, 4, 1), (512, 130, 256, 65), (264, 64, 132, 32),
                                         (240, 12, 120, 6), (248, 12, 124, 6), (232, 70, 116, 35), (488, 6, 244, 3)])   # 30 / 31 / 29 / 61 strips: warp seams of the overlapped layout
@pytest.mark.parametrize("sfmt,dfmt", [(FMT.NV12, FMT.RGB24), (FMT.YUV420P, FMT.BGRA), (FMT.P010LE, FMT.RGB48LE), (FMT.P016LE, FMT.BGRA64LE)])
def test_fused_2to1_vs_oracle(dev, name, flag, param, sw, sh, dw, dh, sfmt, dfmt):
    src = FrameBatch(sfmt, sw, sh, 2); src.fill_lcg(seed=sw + dh)
    c = SwsContext(sw, sh, sfmt, dw, dh, dfmt, flag | HW, param)
    assert c.path == 1, "expected the fused 2:1 kernel"
    ds = src.to(dev); dd = FrameBatch(dfmt, dw, dh, 2, device=dev)
    c.scale(ds, dd); torch.cuda.synchronize()
    ref = FrameBatch(dfmt, dw, dh, 2); orc.yuv2rgb_scale(src, ref, tables(c))
    assert_same(dd, ref, f"{name}{param} {sfmt}->{dfmt} {sw}x{sh}->{dw}x{dh}")


@pytest.mark.parametrize("name,flag,param", ALGOS + [("bilinear", SWS.BILINEAR, None), ("nearest", SWS.POINT, None)])
@pytest.mark.parametrize("sw,sh,dw,dh", [(64, 48, 40, 30), (33, 17, 50, 29), (16, 16, 7, 5), (3, 3, 9, 9), (100, 60, 12, 7), (62, 46, 31, 23)])
def test_generic_any_ratio_vs_oracle(dev, name, flag, param, sw, sh, dw, dh):
    for sfmt, dfmt in ((FMT.NV12, FMT.RGB24), (FMT.YUV420P, FMT.RGBA), (FMT.P016LE, FMT.RGB48LE)):
        src = FrameBatch(sfmt, sw, sh, 1); src.fill_lcg(seed=sw * dh)
        c = SwsContext(sw, sh, sfmt, dw, dh, dfmt, flag | HW, param)
        ds = src.to(dev); dd = FrameBatch(dfmt, dw, dh, 1, device=dev)
        c.scale(ds, dd); torch.cuda.synchronize()
        ra = 1 if name in ("bilinear", "nearest") else 0
        ref = FrameBatch(dfmt, dw, dh, 1); orc.yuv2rgb_scale(src, ref, tables(c), ra=ra)
        assert_same(dd, ref, f"{name}{param} {sfmt}->{dfmt} {sw}x{sh}->{dw}x{dh} path {c.path}")


@pytest.mark.parametrize("sw,sh,dw,dh", [(64, 48, 32, 24), (16, 4, 8, 2), (8, 2, 4, 1), (512, 130, 256, 65), (264, 64, 132, 32)])
@pytest.mark.parametrize("sfmt,dfmt", [(FMT.NV12, FMT.RGB24), (FMT.YUV420P, FMT.BGRA), (FMT.NV12, FMT.BGR24), (FMT.YUV420P, FMT.RGBA)])
def test_fused_bilinear_2to1_vs_oracle(dev, sw, sh, dw, dh, sfmt, dfmt):
    """SWS_BILINEAR (what the reference always runs): integer fast path vs the R-A restatement"""
    src = FrameBatch(sfmt, sw, sh, 2); src.fill_lcg(seed=sw + dh)
    c = SwsContext(sw, sh, sfmt, dw, dh, dfmt, SWS.BILINEAR | HW)
    assert c.path == 1
    ds = src.to(dev); dd = FrameBatch(dfmt, dw, dh, 2, device=dev)
    c.scale(ds, dd); torch.cuda.synchronize()
    ref = FrameBatch(dfmt, dw, dh, 2); orc.yuv2rgb_scale(src, ref, tables(c), ra=1)
    assert_same(dd, ref, f"bilinear {sfmt}->{dfmt} {sw}x{sh}->{dw}x{dh}")


def test_fused_bilinear_4k_equals_generic(dev):
    sw, sh, dw, dh = 3840, 2160, 1920, 1080
    src = FrameBatch(FMT.NV12, sw, sh, 1, device=dev); host = src.fill_lcg(seed=78)
    c = SwsContext(sw, sh, FMT.NV12, dw, dh, FMT.RGB24, SWS.BILINEAR | HW)
    a = FrameBatch(FMT.RGB24, dw, dh, 1, device=dev); c.scale(src, a)
    odd = FrameBatch(FMT.NV12, sw, sh, 1, device=dev, align=1)
    odd.planes = [(0, sw + 3, sh, sw), ((sw + 3) * sh, sw + 3, sh // 2, sw)]
    odd.frame_bytes = (sw + 3) * (sh + sh // 2)
    odd.buf = torch.zeros(odd.frame_bytes, dtype=torch.uint8, device=dev)
    h2 = np.zeros(odd.frame_bytes, np.uint8)
    for p in range(2):
        odd.plane_view(h2, 0, p)[...] = src.plane_view(host, 0, p)
    odd.upload(h2)
    b = FrameBatch(FMT.RGB24, dw, dh, 1, device=dev); c.scale(odd, b); torch.cuda.synchronize()
    assert torch.equal(a.buf, b.buf)


@pytest.mark.parametrize("name,flag,param", ALGOS)
@pytest.mark.parametrize("sw,sh,dw,dh", [(64, 48, 32, 24), (16, 4, 8, 2), (8, 2, 4, 1), (512, 130, 256, 65)])
def test_fused_rgb24_2to1_vs_oracle(dev, name, flag, param, sw, sh, dw, dh):
    """rgb24 -> rgb24 at 2:1 runs the fused kernel without its colour conversion"""
    for fmt in (FMT.RGB24, FMT.BGR24):
        src = FrameBatch(fmt, sw, sh, 2); src.fill_lcg(seed=sw + dh)
        c = SwsContext(sw, sh, fmt, dw, dh, fmt, flag | HW, param)
        assert c.path == 1
        ds = src.to(dev); dd = FrameBatch(fmt, dw, dh, 2, device=dev)
        c.scale(ds, dd); torch.cuda.synchronize()
        (cx, px), (cy, py) = tables(c)
        ref = FrameBatch(fmt, dw, dh, 2)
        for f in range(2):
            s, d = src.image(f, 1), ref.image(f, 1)
            orc.orc().orc_resample_packed(s.data[0], s.linesize[0], sw, sh, d.data[0], d.linesize[0], dw, dh, 3, 0,
                                          orc.fptr(cx), orc.iptr(px), orc.fptr(cy), orc.iptr(py), 0, 0)
        assert_same(dd, ref, f"rgb {name}{param} {sw}x{sh}->{dw}x{dh}")


def test_fused_rgb24_4k_equals_generic(dev):
    sw, sh, dw, dh = 3840, 2160, 1920, 1080
    src = FrameBatch(FMT.RGB24, sw, sh, 1, device=dev); host = src.fill_lcg(seed=79)
    c = SwsContext(sw, sh, FMT.RGB24, dw, dh, FMT.RGB24, SWS.BICUBIC | HW, (0.75,))
    a = FrameBatch(FMT.RGB24, dw, dh, 1, device=dev); c.scale(src, a)
    odd = FrameBatch(FMT.RGB24, sw, sh, 1, device=dev, align=1)
    odd.planes = [(0, sw * 3 + 5, sh, sw * 3)]
    odd.frame_bytes = (sw * 3 + 5) * sh
    odd.buf = torch.zeros(odd.frame_bytes, dtype=torch.uint8, device=dev)
    h2 = np.zeros(odd.frame_bytes, np.uint8)
    odd.plane_view(h2, 0, 0)[...] = src.plane_view(host, 0, 0)
    odd.upload(h2)
    b = FrameBatch(FMT.RGB24, dw, dh, 1, device=dev); c.scale(odd, b); torch.cuda.synchronize()
    assert torch.equal(a.buf, b.buf)


def test_device_filter_tables_vs_cpu_restatement(dev):
    """bicubic / bilinear / nearest tables are bit-exact on the CPU; Lanczos (GPU __sinf) within 2e-6"""
    for (s, d) in ((3840, 1920), (1920, 1281), (33, 50), (100, 12)):
        for name, flag, param, A in (("bicubic", SWS.BICUBIC, None, 0.0), ("bicubic", SWS.BICUBIC, (0.75,), -0.75),
                                     ("bilinear", SWS.BILINEAR, None, 0.0), ("nearest", SWS.POINT, None, 0.0)):
            c = SwsContext(s, 64, FMT.NV12, d, 32 if d != s else 64, FMT.RGB24, flag | HW, param)
            co, po = c.get_filter(0)
            ro, rp = orc.filter_table(orc.ALGO[name], s, d, A)
            assert np.array_equal(co.view(np.uint32), ro.view(np.uint32)), (name, s, d)
            assert np.array_equal(po, rp)
        c = SwsContext(s, 64, FMT.NV12, d, 32, FMT.RGB24, SWS.LANCZOS | HW)
        co, po = c.get_filter(0)
        ro, rp = orc.filter_table(orc.ALGO["lanczos"], s, d)
        assert np.array_equal(po, rp) and np.abs(co - ro).max() < 2e-6


@pytest.mark.parametrize("flag,param", [(SWS.BICUBIC, None), (SWS.BICUBIC, (0.75,)), (SWS.LANCZOS, None)])
def test_headline_4k_fused_equals_generic_and_oracle_sample(dev, flag, param):
    """BASELINE C2 at full size: the fused kernel must equal the generic kernel bit for bit (the generic
    kernel is forced by mis-aligning the source by one row of padding), plus an oracle check on a crop."""
    sw, sh, dw, dh = 3840, 2160, 1920, 1080
    src = FrameBatch(FMT.NV12, sw, sh, 1, device=dev); host = src.fill_lcg(seed=77)
    c = SwsContext(sw, sh, FMT.NV12, dw, dh, FMT.RGB24, flag | HW, param)
    assert c.path == 1
    a = FrameBatch(FMT.RGB24, dw, dh, 1, device=dev); c.scale(src, a)
    # same frame with a pitch that is not a multiple of 8 -> generic kernel
    odd = FrameBatch(FMT.NV12, sw, sh, 1, device=dev, align=1)
    odd.planes = [(0, sw + 3, sh, sw), ((sw + 3) * sh, sw + 3, sh // 2, sw)]
    odd.frame_bytes = (sw + 3) * (sh + sh // 2)
    odd.buf = torch.zeros(odd.frame_bytes, dtype=torch.uint8, device=dev)
    h2 = np.zeros(odd.frame_bytes, np.uint8)
    for p in range(2):
        odd.plane_view(h2, 0, p)[...] = src.plane_view(host, 0, p)
    odd.upload(h2)
    b = FrameBatch(FMT.RGB24, dw, dh, 1, device=dev); c.scale(odd, b); torch.cuda.synchronize()
    assert torch.equal(a.buf, b.buf)
    # oracle on the top-left 256x128 source crop (output rows/cols not touching the crop's far edges)
    cw, ch = 256, 128
    crop = FrameBatch(FMT.NV12, cw, ch, 1)
    hc = np.zeros(crop.frame_bytes, np.uint8)
    crop.plane_view(hc, 0, 0)[...] = src.plane_view(host, 0, 0)[:ch, :cw]
    crop.plane_view(hc, 0, 1)[...] = src.plane_view(host, 0, 1)[:ch // 2, :cw]
    crop.upload(hc)
    cc = SwsContext(cw, ch, FMT.NV12, cw // 2, ch // 2, FMT.RGB24, flag | HW, param)
    ref = FrameBatch(FMT.RGB24, cw // 2, ch // 2, 1); orc.yuv2rgb_scale(crop, ref, tables(cc))
    got = a.plane_view(a.numpy(), 0, 0)[:ch // 2 - 1, :(cw // 2 - 1) * 3]
    exp = ref.plane_view(ref.numpy(), 0, 0)[:ch // 2 - 1, :(cw // 2 - 1) * 3]
    assert np.array_equal(got, exp)


def test_c3_8k_p010_lanczos_fused_equals_generic(dev):
    """BASELINE C3: 8K P010 -> 4K RGB48, Lanczos"""
    sw, sh, dw, dh = 7680, 4320, 3840, 2160
    src = FrameBatch(FMT.P010LE, sw, sh, 1, device=dev); src.fill_lcg(seed=5)
    c = SwsContext(sw, sh, FMT.P010LE, dw, dh, FMT.RGB48LE, SWS.LANCZOS | HW)
    assert c.path == 1
    a = FrameBatch(FMT.RGB48LE, dw, dh, 1, device=dev); c.scale(src, a)
    odd = FrameBatch(FMT.P010LE, sw, sh, 1, device=dev, align=1)
    odd.planes = [(0, 2 * sw + 2, sh, 2 * sw), ((2 * sw + 2) * sh, 2 * sw + 2, sh // 2, 2 * sw)]
    odd.frame_bytes = (2 * sw + 2) * (sh + sh // 2)
    odd.buf = torch.zeros(odd.frame_bytes, dtype=torch.uint8, device=dev)
    for p in range(2):
        off, pitch, rows, rb = odd.planes[p]
        so, sp, _, _ = src.planes[p]
        dst2 = odd.buf[off:off + pitch * rows].view(rows, pitch)[:, :rb]
        dst2.copy_(src.buf[so:so + sp * rows].view(rows, sp)[:, :rb])
    b = FrameBatch(FMT.RGB48LE, dw, dh, 1, device=dev); c.scale(odd, b); torch.cuda.synchronize()
    assert torch.equal(a.buf, b.buf)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("algo,flag,param,oparam", [("Bicubic", SWS.BICUBIC, None, 999999.0), ("Bicubic", SWS.BICUBIC, (0.75,), 0.75),
                                                    ("Lanczos", SWS.LANCZOS, None, 999999.0)])
@pytest.mark.parametrize("sw,sh,dw,dh", [(64, 48, 32, 24), (64, 48, 40, 30), (33, 17, 50, 29), (1920, 1080, 960, 540), (1280, 720, 1920, 1080)])
def test_packed_resample_vs_reference_scale_cuda_live(dev, algo, flag, param, oparam, sw, sh, dw, dh):
    """rgb0 -> rgb0 through our context (rgb->rgb scaling) vs Subsample_<algo>_rgb0_rgb0 of the reference,
    with GMATB_SWS_PARITY_WRAP reproducing its missing upper clamp: bit-exact, Lanczos included."""
    src = FrameBatch(FMT.RGBA, sw, sh, 1, device=dev); src.fill_lcg(seed=sw + dw)
    ref = FrameBatch(FMT.RGBA, dw, dh, 1, device=dev)
    o2_packed(f"Subsample_{algo}_rgb0_rgb0", src, ref, 4, 8, oparam)
    c = SwsContext(sw, sh, FMT.RGBA, dw, dh, FMT.RGBA, flag | HW | SWS.PARITY_WRAP, param)
    dd = FrameBatch(FMT.RGBA, dw, dh, 1, device=dev)
    c.scale(src, dd); torch.cuda.synchronize()
    assert_same(dd, ref, f"{algo} rgb0 {sw}x{sh}->{dw}x{dh} vs O2")
    # production mode saturates instead of wrapping: only pixels where the reference overflowed may differ
    c2 = SwsContext(sw, sh, FMT.RGBA, dw, dh, FMT.RGBA, flag | HW, param)
    d2 = FrameBatch(FMT.RGBA, dw, dh, 1, device=dev); c2.scale(src, d2); torch.cuda.synchronize()
    a, b = d2.payload(), dd.payload()
    assert np.all((a == b) | (a == 255))


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("algo,flag,param,oparam", [("Bicubic", SWS.BICUBIC, None, 999999.0), ("Bicubic", SWS.BICUBIC, (0.75,), 0.75),
                                                    ("Bicubic", SWS.BICUBIC, (0.5,), 0.5), ("Lanczos", SWS.LANCZOS, None, 999999.0)])
@pytest.mark.parametrize("sw,sh", [(1920, 1080), (3840, 2160)])
@pytest.mark.parametrize("extra", [0, SWS.INT_CHAIN])
def test_fused_pipeline_vs_reference_two_kernel_pipeline_live(dev, algo, flag, param, oparam, sw, sh, extra):
    """The reference's unfused pipeline rebuilt from its own kernels: O1 nv12->rgba at source size, then
    O2 Subsample_*_rgb0_rgb0.  Our fused kernel (rgba output) must reproduce its r,g,b bytes exactly -- including
    the BASELINE C2 headline itself (param0 = 0.75 at 3840x2160 -> 1920x1080), on both forms of the kernel."""
    dw, dh = sw // 2, sh // 2
    src = FrameBatch(FMT.NV12, sw, sh, 1, device=dev); src.fill_lcg(seed=99)
    mid = FrameBatch(FMT.RGBA, sw, sh, 1, device=dev); o1_run("yuv2rgb_cuda", src, mid, 0)
    ref = FrameBatch(FMT.RGBA, dw, dh, 1, device=dev); o2_packed(f"Subsample_{algo}_rgb0_rgb0", mid, ref, 4, 8, oparam)
    c = SwsContext(sw, sh, FMT.NV12, dw, dh, FMT.RGBA, flag | HW | SWS.PARITY_WRAP | extra, param)
    assert c.path == 1
    dd = FrameBatch(FMT.RGBA, dw, dh, 1, device=dev); c.scale(src, dd); torch.cuda.synchronize()
    assert_same(dd, ref, f"fused vs O1+O2 {algo} {param} {sw}x{sh}")
    # the headline output format: same bytes without the alpha channel, production clamp where the reference wrapped
    c3 = SwsContext(sw, sh, FMT.NV12, dw, dh, FMT.RGB24, flag | HW | extra, param)
    d3 = FrameBatch(FMT.RGB24, dw, dh, 1, device=dev); c3.scale(src, d3); torch.cuda.synchronize()
    a = d3.plane_view(d3.numpy(), 0, 0).reshape(dh, dw, 3)
    b = ref.plane_view(ref.numpy(), 0, 0).reshape(dh, dw, 4)[:, :, :3]
    assert np.all((a == b) | (a == 255)) and (a != b).mean() < 0.2


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
def test_plane_scaling_vs_reference_nv12_kernels_live(dev):
    """yuv->yuv scaling: Y and interleaved UV planes vs Subsample_Bicubic_nv12_nv12[_uv]"""
    sw, sh, dw, dh = 640, 360, 400, 226
    src = FrameBatch(FMT.NV12, sw, sh, 1, device=dev); src.fill_lcg(seed=12)
    ref = FrameBatch(FMT.NV12, dw, dh, 1, device=dev)
    o2_packed("Subsample_Bicubic_nv12_nv12", src, ref, 1, 8, 0.75, plane=0)
    # the _uv kernel reads texture slot 1: give it as plane list [Y, UV]
    import gpu_util
    L = gpu_util.o2(); si, ri = src.image(), ref.image()
    vp, ci = C.c_void_p, C.c_int
    rc = L.ref_o2_launch(b"Subsample_Bicubic_nv12_nv12_uv", 2, (vp * 4)(si.data[0], si.data[1], None, None),
                         (ci * 4)(si.linesize[0], si.linesize[1], 0, 0), (ci * 4)(sw, sw // 2, 0, 0), (ci * 4)(sh, sh // 2, 0, 0),
                         (ci * 4)(8, 8, 0, 0), (ci * 4)(1, 2, 0, 0), (vp * 4)(ri.data[0], ri.data[1], None, None),
                         dw // 2, dh // 2, ri.linesize[1], sw // 2, sh // 2, C.c_float(0.75), 0, 0)
    assert rc == 0
    torch.cuda.synchronize()
    c = SwsContext(sw, sh, FMT.NV12, dw, dh, FMT.NV12, SWS.BICUBIC | HW | SWS.PARITY_WRAP, (0.75,))
    dd = FrameBatch(FMT.NV12, dw, dh, 1, device=dev); c.scale(src, dd); torch.cuda.synchronize()
    assert_same(dd, ref, "nv12->nv12 bicubic vs O2")


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("fmt", [FMT.NV12, FMT.YUV420P, FMT.P010LE, FMT.P016LE])
@pytest.mark.parametrize("algo,flag,param,oparam", [("Bicubic", SWS.BICUBIC, None, 999999.0), ("Bicubic", SWS.BICUBIC, (0.75,), 0.75),
                                                    ("Lanczos", SWS.LANCZOS, None, 999999.0)])
@pytest.mark.parametrize("sw,sh,dw,dh", [(640, 360, 400, 226), (256, 144, 128, 72), (136, 68, 200, 100), (1920, 1080, 1280, 720)])   # planes 512-byte aligned: textures
def test_yuv_plane_scaling_vs_reference_scale_cuda_live(dev, fmt, algo, flag, param, oparam, sw, sh, dw, dh):
    """yuv -> same yuv format at another size = what the scale_cuda filter does: every plane of NV12 / YUV420P / P010 /
    P016 against Subsample_{Bicubic,Lanczos}_<fmt>_<fmt>[_uv] of the reference, launched like scalecuda_resize
    (vf_scale_cuda.c:430-500); 16-bit planes included (they were only checked on constant frames before)."""
    from gpu_util import o2_frame
    src = FrameBatch(fmt, sw, sh, 1, device=dev); src.fill_lcg(seed=sw + dh)
    ref = FrameBatch(fmt, dw, dh, 1, device=dev); ref.buf.fill_(0)
    o2_frame(algo, src, ref, oparam)
    c = SwsContext(sw, sh, fmt, dw, dh, fmt, flag | HW | SWS.PARITY_WRAP, param)
    dd = FrameBatch(fmt, dw, dh, 1, device=dev); dd.buf.fill_(0)
    c.scale(src, dd); torch.cuda.synchronize()
    assert_same(dd, ref, f"{fmt} {algo}{param} {sw}x{sh}->{dw}x{dh} vs O2")


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built")
def test_device_filter_tables_vs_reference_coeff_functions_live(dev):
    """our filter_table_kernel against the reference's own lanczos_coeffs / bicubic_coeffs (vf_scale_cuda.cu:948-981,
    #included in place by oracle/ref_o2_coeffs.cu) at the same fractional positions: bit-exact, Lanczos' fast-math
    sines included"""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_golden_lanczos import frac_positions
    from gpu_util import o2_coeffs
    for (s_, d_) in ((3840, 1920), (7680, 3840), (1920, 1281), (33, 50), (100, 12), (1080, 720), (64, 40)):
        fx, pos = frac_positions(s_, d_)
        for flag, param, lz, op in ((SWS.LANCZOS, None, 1, 999999.0), (SWS.BICUBIC, None, 0, 999999.0), (SWS.BICUBIC, (0.75,), 0, 0.75),
                                    (SWS.BICUBIC, (0.6,), 0, 0.6)):
            c = SwsContext(s_, 64, FMT.NV12, d_, 32, FMT.RGB24, flag | HW, param)
            co, po = c.get_filter(0)
            ref = o2_coeffs(lz, fx, op)
            assert np.array_equal(po, pos), (s_, d_)
            assert np.array_equal(co.view(np.uint32), ref.view(np.uint32)), (s_, d_, flag, param, int((co != ref).sum()))


@pytest.mark.parametrize("fmt", [FMT.YUV420P, FMT.NV12, FMT.P010LE, FMT.P016LE, FMT.RGB0, FMT.BGR0])
@pytest.mark.parametrize("flag", [SWS.BICUBIC, SWS.LANCZOS, SWS.POINT, SWS.BILINEAR])
def test_scale_cuda_filter_format_set(dev, fmt, flag):
    """every (format, algorithm) the scale_cuda glue (csrc/avfilter/vf_scale_cuda.c) can ask for: the context is
    created, runs, and a constant frame stays constant (each plane keeps its value: all four filters sum to 1)"""
    sw, sh, dw, dh = 256, 144, 160, 90
    c = SwsContext(sw, sh, fmt, dw, dh, fmt, flag | HW)
    src = FrameBatch(fmt, sw, sh, 2, device=dev); src.buf.fill_(0x40)       # every byte 0x40: 8-bit 64, 16-bit 0x4040
    dd = FrameBatch(fmt, dw, dh, 2, device=dev); dd.buf.fill_(0xEE)
    c.scale(src, dd); torch.cuda.synchronize()
    out = dd.payload()
    if fmt in (FMT.P010LE, FMT.P016LE):
        v = out.view(np.uint16)
        assert v.min() >= 0x403F and v.max() <= 0x4040, (int(v.min()), int(v.max()))
    else:
        assert out.min() >= 0x3F and out.max() <= 0x40, (int(out.min()), int(out.max()))


def test_rgb2yuv_scaled_is_resize_then_convert(dev):
    """swscale_cuda.c:312-341 ordering: resize the rgb source to dst size, then rgb2yuv"""
    sw, sh, dw, dh = 320, 200, 128, 96
    src = FrameBatch(FMT.RGB24, sw, sh, 1, device=dev); src.fill_lcg(seed=1)
    c = SwsContext(sw, sh, FMT.RGB24, dw, dh, FMT.NV12, SWS.BICUBIC | HW)
    dd = FrameBatch(FMT.NV12, dw, dh, 1, device=dev); c.scale(src, dd)
    c1 = SwsContext(sw, sh, FMT.RGB24, dw, dh, FMT.RGB24, SWS.BICUBIC | HW)
    mid = FrameBatch(FMT.RGB24, dw, dh, 1, device=dev); c1.scale(src, mid)
    ref = FrameBatch(FMT.NV12, dw, dh, 1, device=dev); g.rgb2yuv(mid, ref); torch.cuda.synchronize()
    assert torch.equal(dd.buf, ref.buf)


def test_constant_frames_and_linearity_properties(dev):
    """size-independent properties at full 4K: a constant frame maps to a constant (the value the chain
    gives a constant row), and permuting frames in a batch permutes the outputs"""
    sw, sh, dw, dh = 3840, 2160, 1920, 1080
    c = SwsContext(sw, sh, FMT.NV12, dw, dh, FMT.RGB24, SWS.BICUBIC | HW, (0.75,))
    src = FrameBatch(FMT.NV12, sw, sh, 3, device=dev)
    src.buf.fill_(0)
    fb = src.frame_bytes
    src.buf[0:fb] = 128; src.buf[fb:2 * fb] = 200
    src.buf[2 * fb:3 * fb].random_(0, 256)
    out = FrameBatch(FMT.RGB24, dw, dh, 3, device=dev); c.scale(src, out); torch.cuda.synchronize()
    host = out.numpy()
    for f in (0, 1):
        v = out.plane_view(host, f, 0)
        assert len(np.unique(v[:, 0::3])) == 1 and len(np.unique(v[:, 1::3])) == 1 and len(np.unique(v[:, 2::3])) == 1
    perm = FrameBatch(FMT.NV12, sw, sh, 3, device=dev)
    perm.buf[0:fb] = src.buf[2 * fb:3 * fb]; perm.buf[fb:2 * fb] = src.buf[0:fb]; perm.buf[2 * fb:] = src.buf[fb:2 * fb]
    out2 = FrameBatch(FMT.RGB24, dw, dh, 3, device=dev); c.scale(perm, out2); torch.cuda.synchronize()
    ofb = out.frame_bytes
    assert torch.equal(out2.buf[0:ofb], out.buf[2 * ofb:3 * ofb]) and torch.equal(out2.buf[ofb:2 * ofb], out.buf[0:ofb])


def test_scale_host_roundtrip_equals_device_path(dev):
    sw, sh, dw, dh, n = 640, 360, 320, 180, 3
    hs = FrameBatch(FMT.NV12, sw, sh, n, pinned=True); host = hs.fill_lcg(seed=3)
    hd = FrameBatch(FMT.RGB24, dw, dh, n, pinned=True)
    c = SwsContext(sw, sh, FMT.NV12, dw, dh, FMT.RGB24, SWS.BICUBIC | HW)
    c.scale_host(hs, hd)
    ds = FrameBatch(FMT.NV12, sw, sh, n, device=dev); ds.upload(host)
    dd = FrameBatch(FMT.RGB24, dw, dh, n, device=dev); c.scale(ds, dd); torch.cuda.synchronize()
    assert np.array_equal(hd.payload(), dd.payload())


def test_ffmpeg_style_single_frame_call(dev):
    vp, ci = C.c_void_p, C.c_int
    sw, sh, dw, dh = 256, 128, 128, 64
    src = FrameBatch(FMT.NV12, sw, sh, 1, device=dev); src.fill_lcg(seed=2)
    c = g.sws_getContext(sw, sh, FMT.NV12, dw, dh, FMT.RGB24, SWS.BICUBIC | HW)
    a = FrameBatch(FMT.RGB24, dw, dh, 1, device=dev); c.scale(src, a)
    b = FrameBatch(FMT.RGB24, dw, dh, 1, device=dev)
    si, bi = src.image(), b.image()
    assert g.sws_scale(c, (vp * 4)(si.data[0], si.data[1], None, None), (ci * 4)(si.linesize[0], si.linesize[1], 0, 0), 0, sh,
                       (vp * 4)(bi.data[0], None, None, None), (ci * 4)(bi.linesize[0], 0, 0, 0)) == 0
    torch.cuda.synchronize()
    assert torch.equal(a.buf, b.buf)
    g.sws_freeContext(c)
