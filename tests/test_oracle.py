"""CPU tests of the oracle (oracle/gmat_oracle.c): known-answer tests from SURVEY 8a and the
golden vectors produced by the REFERENCE's own CUDA kernels on a B200
(tests/golden/make_golden.py -> tests/golden/reference_gpu_golden.npz)."""
import os
import re

import numpy as np
import pytest

import orc
from gmat_b200 import FMT, FrameBatch

# SURVEY 8a KAT: IEEE-754 bit patterns of the 9 floats the reference uploads
KAT_Y2R = {
    0: "3f950a85 00000000 3fd0f4a3 3f950a85 becd296b bf54df0d 3f950a85 40040ce8 00000000",
    1: "3f950a85 00000000 3feab5bd 3f950a85 be5f5a23 bf0b8a1d 3f950a85 400a47c4 00000000",
    4: "3f950a85 00000000 3fd0a854 3f950a85 bec5d883 bf5431b2 3f950a85 4004a586 00000000",
    7: "3f950a85 00000000 3feae386 3f950a85 be871a9c bf0e1290 3f950a85 40081314 00000000",
    9: "3f959f90 00000000 3fdca26f 3f959f90 be44f7c5 bf2af9ba 3f959f90 400cc029 00000000",
}
KAT_R2Y = {
    0: "3e8379bf 3f010ea0 3dc882e1 be14646e be91a9a6 3edbdbdc 3edbdbdc beb81ab5 bd8f049c",
    1: "3e3af7cb 3f1d3e37 3d7dfb1d bdc9847b bea97abd 3edbdbdc 3edbdbdc bec7b2f5 bd214738",
    4: "3e83ea51 3f01b77a 3dc179cc be14382d be91bfc5 3edbdbdc 3edbdbdc beb94f40 bd8a326d",
    7: "3e3a70b6 3f1a1efc 3d990591 bdcc34cb bea8cea8 3edbdbdc 3edbdbdc bec395c7 bd4230a6",
    9: "3e6620f3 3f147bf8 3d4fca56 bdf4a2b7 be9dd82d 3edb00db 3edb00db bec963a2 bd0ce9bc",
}


def bits(m):
    return " ".join("%08x" % x for x in m.view(np.uint32))


@pytest.mark.parametrize("cs", sorted(KAT_Y2R))
def test_matrix_kat(cs):
    assert bits(orc.matrix_yuv2rgb(cs)) == KAT_Y2R[cs]
    assert bits(orc.matrix_rgb2yuv(cs)) == KAT_R2Y[cs]
    # colourspaces 5 and 6 (BT470BG, SMPTE170M) are the default matrix
    if cs == 0:
        for alias in (5, 6, 2, 3):
            assert bits(orc.matrix_yuv2rgb(alias)) == KAT_Y2R[0]


def test_matrix_matches_reference_upload(golden):
    """what the reference's set_mat_*_cuda actually put into constant memory on the B200"""
    for cs in (0, 1, 4, 5, 6, 7, 9):
        assert bits(orc.matrix_yuv2rgb(cs)) == bits(golden[f"mat_y2r_{cs}"])
        assert bits(orc.matrix_rgb2yuv(cs)) == bits(golden[f"mat_r2y_{cs}"])


def test_product_matrix_equals_oracle():
    import gmat_b200 as g
    for cs in (0, 1, 4, 5, 6, 7, 9, 10):
        assert bits(g.csc_matrix_yuv2rgb(cs)) == bits(orc.matrix_yuv2rgb(cs))
        assert bits(g.csc_matrix_rgb2yuv(cs)) == bits(orc.matrix_rgb2yuv(cs))


def test_worked_example():
    """(Y,U,V) = (81,90,240), default matrix"""
    src = FrameBatch(FMT.NV12, 2, 2, 1)
    host = np.zeros(src.frame_bytes, np.uint8)
    src.plane_view(host, 0, 0)[...] = 81
    src.plane_view(host, 0, 1)[0, 0] = 90; src.plane_view(host, 0, 1)[0, 1] = 240
    src.upload(host)
    dst = FrameBatch(FMT.RGB24, 2, 2, 1)
    orc.yuv2rgb(src, dst)
    m = orc.matrix_yuv2rgb(0).astype(np.float64)
    fy, fu, fv = 65.0, -38.0, 112.0
    exp = [int(min(max(m[3 * i] * fy + m[3 * i + 1] * fu + m[3 * i + 2] * fv, 0), 255)) for i in range(3)]
    assert list(dst.payload()[:3]) == exp == [255, 0, 0] or list(dst.payload()[:3]) == exp


def test_norm_two_term_is_exact():
    """FFMA(j,khi,RN(j*klo)) == RN(j/max) for every sample value (gmat_b200/csrc/resample_core.cuh)"""
    assert orc.orc().orc_check_norm(0) == 0
    assert orc.orc().orc_check_norm(1) == 0


def test_norm_geometric_form_is_exact():
    """sat(FFMA(hi, c2, hi)), hi = j*c1 exactly, == RN(j/max) for every sample value: the form the
    fused kernel uses (resample_core.cuh quant_norm2).  hi*(1+c2) has <= 48 significant bits, so the
    float64 product-sum is exact and its cast to float32 is the FFMA's single rounding."""
    for maxv, c1, c2 in ((255, 65793.0 / 2 ** 24, 2.0 ** -24), (65535, 2.0 ** -16, 2.0 ** -16 + 2.0 ** -32)):
        j = np.arange(maxv + 1, dtype=np.float64)
        assert np.float32(c1) == c1 and np.float32(c2) == c2
        hi = j * c1
        assert np.all(hi.astype(np.float32).astype(np.float64) == hi)           # representable: FFMA(m, c1, -2^23 c1) is exact
        m = 8388608.0 + j
        assert np.all((m * c1 - 8388608.0 * c1) == hi)
        p = (hi * c2 + hi).astype(np.float32)
        ref = j.astype(np.float32) / np.float32(maxv)                           # IEEE division: correctly rounded
        assert np.array_equal(p, ref)


def test_norm_denormal_domain_form_is_exact_and_scaled():
    """quant_norm2d (resample_core.cuh), the all-packed form of the headline kernel: j read as the denormal float j*2^-149,
    hi' = j*kd exactly (kd = c1*2^(149-shift)), p' = FFMA(hi', c2, hi') == RN(j/max) * 2^-shift for every sample value, and a
    4-tap chain on the scaled samples times factor*2^shift gives the bytes of the unscaled chain."""
    rng = np.random.default_rng(5)
    for maxv, c1, c2, shift in ((255, 65793.0 / 2 ** 24, 2.0 ** -24, 14), (65535, 2.0 ** -16, 2.0 ** -16 + 2.0 ** -32, 6)):
        kd = np.float32(c1 * 2.0 ** (149 - shift))
        assert np.isfinite(kd) and float(kd) == c1 * 2.0 ** (149 - shift)
        j = np.arange(maxv + 1, dtype=np.float64)
        d = j * 2.0 ** -149                                                       # exact in float64
        hi = d * float(kd)
        assert np.all(hi == j * c1 * 2.0 ** -shift) and np.all(hi.astype(np.float32).astype(np.float64) == hi)
        p = (hi * c2 + hi).astype(np.float32)                                     # <= 48 significant bits: exact in float64
        ref = j.astype(np.float32) / np.float32(maxv)
        assert np.array_equal(p.astype(np.float64) * 2.0 ** shift, ref.astype(np.float64))
        # linearity of the chains under the power-of-two scale (float32 FMA emulated in float64: products of two float32 are
        # exact, the sum is rounded once to float32 -- double rounding cannot occur within 53 bits here because the operands
        # share the scale): same integer result for random 4-tap windows and the dyadic / Lanczos-like weights
        for w in (np.array([-0.09375, 0.59375, 0.59375, -0.09375]), np.array([-0.0703125, 0.5703125, 0.5703125, -0.0703125]),
                  np.array([-0.064453125, 0.564453125, 0.564453125, -0.064453125])):
            jj = rng.integers(0, maxv + 1, (4000, 4))
            ps = p[jj].astype(np.float64); pu = ref[jj].astype(np.float64)

            def chain(x):
                t = np.float32(w[1] * x[:, 1]).astype(np.float64)
                for k in (0, 2, 3):
                    t = np.float32(w[k] * x[:, k] + t).astype(np.float64)
                return t
            hs, hu = chain(ps), chain(pu)
            assert np.array_equal(hs * 2.0 ** shift, hu)
            os_ = np.trunc(np.float32(hs * (maxv * 2.0 ** shift)).astype(np.float64))
            ou = np.trunc(np.float32(hu * maxv).astype(np.float64))
            assert np.array_equal(os_, ou)


def test_median5_stream_blocks_are_in_sync_and_exact(tmp_path):
    """tools/gen_median5_stream.py verifies its min/max blocks (0-1 principle on sorted inputs + random bytes + the whole
    25-sample pair construction) before writing them; the committed median5_nets.inc must be what it generates"""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "median5_nets.inc"
    subprocess.run([sys.executable, os.path.join(root, "tools", "gen_median5_stream.py"), str(out)], check=True, capture_output=True)
    assert out.read_text() == open(os.path.join(root, "gmat_b200", "csrc", "median5_nets.inc")).read()


def test_bicubic_coefficients_closed_form():
    # exact 2:1: fx = 0.5 for every output; A = -0.75 gives dyadic weights, default A = 0 a 2x2 box
    co, po = orc.filter_table(orc.ALGO["bicubic"], 3840, 1920, -0.75)
    assert np.all(co == np.array([-0.09375, 0.59375, 0.59375, -0.09375], np.float32))
    assert np.all(po == 2 * np.arange(1920) - 1)
    co, _ = orc.filter_table(orc.ALGO["bicubic"], 3840, 1920, 0.0)
    assert np.all(co == np.array([0, 0.5, 0.5, 0], np.float32))
    # partition of unity to 1 ulp at arbitrary ratios
    co, _ = orc.filter_table(orc.ALGO["bicubic"], 1920, 1281, -0.5)
    assert np.abs(co.sum(1) - 1).max() < 3e-7


def _case(name):
    m = re.match(r"o1_(\w+?)_(\w+?)_(\d+)x(\d+)(?:_cs(\d))?$", name)
    return m


FMTS = {"nv12": FMT.NV12, "yuv420p": FMT.YUV420P, "p010": FMT.P010LE, "p016": FMT.P016LE, "yuv420p10": FMT.YUV420P10LE,
        "yuv420p16": FMT.YUV420P16LE, "rgb24": FMT.RGB24, "bgr24": FMT.BGR24, "rgba": FMT.RGBA, "bgra": FMT.BGRA,
        "rgba64": FMT.RGBA64LE, "bgra64": FMT.BGRA64LE}
SEEDS = {"nv12_rgb": lambda w, h: 1234 + w * 131 + h}


def test_oracle_vs_reference_csc_golden(golden):
    """every O1 golden vector: the reference's libgpuscale kernels vs the CPU restatement"""
    import ctypes as C
    n = 0
    for name in golden.files:
        m = _case(name)
        if not m or name.endswith("_crc"):
            continue
        sname, dname, w, h, cs = m.group(1), m.group(2), int(m.group(3)), int(m.group(4)), int(m.group(5) or 0)
        if dname == "rgbpf32":
            src = FrameBatch(FMT.NV12, w, h, 1); src.fill_lcg(seed=1234 + w * 131 + h)
            dst = FrameBatch(FMT.RGBPF32LE, w, h, 1)
            mm = orc.matrix_yuv2rgb(0); sh = np.zeros(3, np.float32)
            s, d = src.image(), dst.image()
            orc.orc().orc_yuv2rgb_planar_f32(C.byref(s), C.byref(d), orc.fptr(mm), 255.0, orc.fptr(sh))
            assert np.array_equal(dst.payload(), golden[name]), name
            n += 1
            continue
        sfmt, dfmt = FMTS[sname], FMTS[dname]
        src = FrameBatch(sfmt, w, h, 1)
        dst = FrameBatch(dfmt, w, h, 1)
        yuv_s = sname in ("nv12", "yuv420p", "p010", "p016")
        yuv_d = dname in ("nv12", "yuv420p", "p010", "p016", "yuv420p10", "yuv420p16")
        if yuv_s and not yuv_d:
            src.fill_lcg(seed=(99 + w) if sname == "p016" else (1234 + w * 131 + h))
            orc.yuv2rgb(src, dst, cs)
        elif not yuv_s and yuv_d:
            src.fill_lcg(seed=4321 + w)
            orc.rgb2yuv(src, dst, cs)
        elif yuv_s and yuv_d:
            src.fill_lcg(seed=(555 + w) if sname == "nv12" else (777 + w))
            s, d = src.image(), dst.image()
            orc.orc().orc_yuv2yuv(C.byref(s), C.byref(d))
        else:
            src.fill_lcg(seed=888 + w)
            s, d = src.image(), dst.image()
            orc.orc().orc_rgb24tobgr24(C.byref(s), C.byref(d))
        got, exp = dst.payload(), golden[name]
        assert got.shape == exp.shape, name
        if (sname, dname) in (("nv12", "yuv420p"), ("yuv420p", "nv12")):
            # reference defect: the 8-bit (de)interleave launchers round their grid DOWN
            # (yuv2yuv_cuda.cu:293,304 "(height + 3) / 4 / 2" blocks of 8 rows, "(width + 31) / 32 / 2" of 64
            # columns), so trailing rows/columns are never written.  Compare what the reference covers.
            rows = min(h, ((h + 3) // 4 // 2) * 8); cols = min(w, ((w + 31) // 32 // 2) * 64)
            assert cols == w, name
            if rows < h:
                keep = np.zeros(0, bool)
                for p_, (off, pitch, prow, rb) in enumerate(dst.planes):
                    lim = rows if p_ == 0 else rows // 2
                    keep = np.concatenate([keep, (np.arange(prow)[:, None] < lim).repeat(rb, 1).reshape(-1)])
                assert not exp[~keep].any(), name          # untouched (still zero) in the reference's output
                got, exp = got[keep], exp[keep]
        bad = int((got != exp).sum())
        assert bad == 0, f"{name}: {bad} of {got.size} bytes differ from the reference kernel's output"
        n += 1
    assert n >= 100


def _o2_cases(golden):
    for name in golden.files:
        m = re.match(r"o2_(Bicubic|Lanczos|Nearest)_(def|[\d.]+)_(rgb0|y8|y16)_(\d+)x(\d+)_(\d+)x(\d+)$", name)
        if m:
            yield name, m.group(1), m.group(2), m.group(3), tuple(int(x) for x in m.groups()[3:])


def test_oracle_vs_reference_resample_golden(golden):
    """O2 golden vectors (the reference's scale_cuda kernels) vs the CPU restatement of R-B.
    Bicubic / nearest coefficient tables are computed on the CPU (bit-exact restatement); Lanczos tables need the
    GPU's fast-math __sinf, so the committed tables of the reference's own lanczos_coeffs
    (tests/golden/reference_lanczos_tables.npz, made by make_golden_lanczos.py on a B200) are used: the resample
    arithmetic of the oracle is pinned for Lanczos -- C3's filter -- as well."""
    import ctypes as C
    L = orc.orc()
    lz = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_lanczos_tables.npz"))
    n = nl = 0
    for name, algo, pn, kind, (sw, sh, dw, dh) in _o2_cases(golden):
        A = 0.0 if pn == "def" else -float(pn)
        if algo == "Lanczos":
            cx, px = np.ascontiguousarray(lz[f"lanczos_{sw}_{dw}"]), np.ascontiguousarray(lz[f"pos_{sw}_{dw}"])
            cy, py = np.ascontiguousarray(lz[f"lanczos_{sh}_{dh}"]), np.ascontiguousarray(lz[f"pos_{sh}_{dh}"])
            nl += 1
        else:
            a = orc.ALGO["bicubic"] if algo == "Bicubic" else orc.ALGO["nearest"]
            cx, px = orc.filter_table(a, sw, dw, A)
            cy, py = orc.filter_table(a, sh, dh, A)
        ra = 1 if algo == "Nearest" else 0
        if kind == "rgb0":
            src = FrameBatch(FMT.RGBA, sw, sh, 1); src.fill_lcg(seed=31 + sw + dw)
            dst = FrameBatch(FMT.RGBA, dw, dh, 1)
            s, d = src.image(), dst.image()
            L.orc_resample_packed(s.data[0], s.linesize[0], sw, sh, d.data[0], d.linesize[0], dw, dh, 4, 0,
                                  orc.fptr(cx), orc.iptr(px), orc.fptr(cy), orc.iptr(py), ra, 1)
            got = dst.payload()
        else:
            b16 = kind == "y16"
            fmt = FMT.YUV420P16LE if b16 else FMT.YUV420P
            src = FrameBatch(fmt, sw, sh, 1); src.fill_lcg(seed=(91 if b16 else 77) + sw + dw)
            dst = FrameBatch(fmt, dw, dh, 1)
            s, d = src.image(), dst.image()
            L.orc_resample_packed(s.data[0], s.linesize[0], sw, sh, d.data[0], d.linesize[0], dw, dh, 1, int(b16),
                                  orc.fptr(cx), orc.iptr(px), orc.fptr(cy), orc.iptr(py), ra, 1)
            got = np.ascontiguousarray(dst.plane_view(dst.numpy(), 0, 0)).reshape(-1)
        exp = golden[name]
        bad = int((got != exp).sum())
        assert bad == 0, f"{name}: {bad} of {got.size} bytes differ from the reference scale_cuda output"
        n += 1
    assert n >= 40 and nl >= 12


def test_reference_tables_vs_cpu_restatement():
    """the committed reference tables: positions and the bicubic rows are reproduced bit for bit on the CPU (which
    validates how the fractional positions fed to the reference's functions were formed); the CPU's libm Lanczos
    stays within 2e-6 of the reference's fast-math one"""
    lz = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_lanczos_tables.npz"))
    n = 0
    for k in lz.files:
        if not k.startswith("lanczos_"):
            continue
        sn, dn = (int(v) for v in k.split("_")[1:])
        co, po = orc.filter_table(orc.ALGO["bicubic"], sn, dn, -0.75)
        assert np.array_equal(po, lz[f"pos_{sn}_{dn}"])
        assert np.array_equal(co.view(np.uint32), lz[f"bicubic075_{sn}_{dn}"].view(np.uint32)), k
        cl, _ = orc.filter_table(orc.ALGO["lanczos"], sn, dn)
        assert np.abs(cl - lz[k]).max() < 2e-6, k
        n += 1
    assert n >= 20


# ---- exactness arguments the kernels lean on, checked in IEEE binary32 on the CPU -------------------------------
def test_two_tap_half_weights_identity():
    """scale_fused3.cuh, TAPS2: with weights exactly {0,.5,.5,0} the reference chain
         h = FFMA(.5, p2, FMUL(.5, p1));  v = FFMA(.5, h_bottom, FMUL(.5, h_top));  out = trunc(FMUL(v, 255))
       equals  out = trunc(FMUL(RN(RN(p1+p2) + RN(p3+p4)), 255/4))  (scaling by 2^-k commutes with rounding).
       p = RN(j/255); all 2^16 horizontal pairs exhaustively, 2^21 random 2x2 blocks, 8- and 16-bit factors."""
    f = np.float32
    half = f(0.5)
    for mx, fac in ((255, f(255.0)), (65535, f(65535.0))):
        j = np.arange(256, dtype=np.float32) if mx == 255 else np.linspace(0, 65535, 256).round().astype(np.float32)
        p = (j / f(mx)).astype(np.float32)
        p1, p2 = np.meshgrid(p, p)
        h_ref = (half * p2 + half * p1).astype(np.float32)          # both products exact -> one rounding, like the FFMA
        h_new = ((p1 + p2).astype(np.float32) * half).astype(np.float32)
        assert np.array_equal(h_ref, h_new)
        rng = np.random.default_rng(7)
        q = p[rng.integers(0, 256, (4, 1 << 21))]
        ht = (half * q[1] + half * q[0]).astype(np.float32); hb = (half * q[3] + half * q[2]).astype(np.float32)
        v = (half * hb + half * ht).astype(np.float32)
        ref = np.trunc((v * fac).astype(np.float32))
        s = ((q[0] + q[1]).astype(np.float32) + (q[2] + q[3]).astype(np.float32)).astype(np.float32)
        new = np.trunc((s * (fac * f(0.25))).astype(np.float32))
        assert np.array_equal(ref, new)
        assert float(fac * f(0.25)) * 4 == float(fac)               # factor / 4 is exact


def test_integer_mean_through_float_floor():
    """csc.cu rgb2yuv_tile8: (a+b+c+d)/4 in integers == RZ(0.25*(a+b+c+d) + 2^23) - 2^23 for every block sum"""
    s = np.arange(0, 4 * 255 + 1, dtype=np.float32)
    q = (s * np.float32(0.25)).astype(np.float32)                    # exact: multiples of 0.25
    m = np.float32(8388608.0)
    assert np.array_equal(np.floor(q + np.float64(m)).astype(np.float32) - m, (s.astype(np.int64) // 4).astype(np.float32))


def test_planar_float_table_is_the_division():
    """csc.cu yuv2rgb_planar_f32_kernel: the per-block table holds exactly the 3 x 256 possible quotients"""
    for norm, shift in ((255.0, 0.0), (58.395, 123.675), (1.0, -3.5)):
        c = np.arange(256, dtype=np.float32)
        tab = ((c - np.float32(shift)).astype(np.float32) / np.float32(norm)).astype(np.float32)
        for v in (0, 17, 128, 255):
            assert tab[v] == np.float32(np.float32(np.float32(v) - np.float32(shift)) / np.float32(norm))


def test_median3_column_sort_identity():
    """median3_stream.cuh: median of a 3x3 window = med3(max of column minima, med3 of column medians, min of column
    maxima), and u16 lanes holding 257*byte order like the bytes"""
    rng = np.random.default_rng(11)
    w = rng.integers(0, 256, (200000, 3, 3)).astype(np.int64)        # [window, row, column]
    w[:5000] = rng.integers(0, 3, (5000, 3, 3))                      # many ties
    s = np.sort(w, axis=1)                                           # sort every column
    lo, mid, hi = s[:, 0, :], s[:, 1, :], s[:, 2, :]
    med3 = lambda a, b, c: np.maximum(np.maximum(np.minimum(a, b), np.minimum(b, c)), np.minimum(a, c))
    got = med3(lo.max(axis=1), med3(mid[:, 0], mid[:, 1], mid[:, 2]), hi.min(axis=1))
    assert np.array_equal(got, np.median(w.reshape(-1, 9), axis=1).astype(np.int64))
    b = np.arange(256)
    assert np.all(np.diff(b * 257) > 0) and np.all((b * 257) & 0xFF == b) and (255 * 257) < 65536


# ---- the exactness argument behind the integer form of the fused 2:1 kernel (scale_fused4i.cuh) -------------
def _fma32(a, b, c):
    """float32 fma on arrays: the product of two float32 is exact in float64; the float64 sum is then rounded to
    float32 (double rounding can differ from a true fma only on exact float32 ties of the float64 sum, ~2^-29 of
    the cases, and by one float32 ulp -- far inside the margin asserted below)"""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def _chain4(w, p0, p1, p2, p3):
    t = (w[1] * p1).astype(np.float32)
    t = _fma32(np.float32(w[0]), p0, t); t = _fma32(np.float32(w[2]), p2, t); t = _fma32(np.float32(w[3]), p3, t)
    return t


@pytest.mark.parametrize("a,b,s", [(-3, 19, 5), (-1, 9, 4), (-1, 5, 3)])
def test_integer_form_of_the_dyadic_2to1_chain(a, b, s):
    """R-B bicubic at exactly 2:1 with param0 = 0.75 / 0.5 / 1.0 has weights (a, b, b, a) / 2^s.  On the quantised
    bytes j the exact value of 255 v is N / 2^(2s) with N = sum W[y] W[x] j[y][x]; the float chain the reference
    runs (vf_scale_cuda.cu:1040-1074: p = RN(j/255), two 4-tap FMA chains, trunc(255 v)) stays within 2.5e-4 of it,
    so it truncates to N >> 2s whenever N is not a multiple of 2^(2s) -- what the integer kernel stores -- and
    can land on either side only when it is (those outputs are recomputed with the float chain)."""
    rng = np.random.default_rng(1234 + s)
    n = 400_000
    w = np.array([a, b, b, a], np.float32) / np.float32(1 << s)
    wi = np.array([a, b, b, a], np.int64)
    mode = rng.integers(0, 4, n)
    base = rng.integers(0, 256, n)
    amp = np.array([256, 16, 2, 1])[mode]
    j = (base[:, None, None] * (mode != 0)[:, None, None] + rng.integers(0, 1 << 30, (n, 4, 4)) % amp[:, None, None])
    sat = rng.integers(0, 8, (n, 4, 4))
    j = np.where((mode == 0)[:, None, None] & (sat == 0), 0, np.where((mode == 0)[:, None, None] & (sat == 1), 255, j))
    j = np.clip(j, 0, 255)
    p = (j.astype(np.float64) / 255.0).astype(np.float32)          # RN(j/255)
    h = [_chain4(w, p[:, y, 0], p[:, y, 1], p[:, y, 2], p[:, y, 3]) for y in range(4)]
    v = _chain4(w, h[0], h[1], h[2], h[3])
    f = (v * np.float32(255.0)).astype(np.float32)
    got = np.where(f < 0, 0, np.trunc(f)).astype(np.int64)
    N = np.einsum("y,x,nyx->n", wi, wi, j.astype(np.int64))
    sh = 2 * s
    E = np.where(N < 0, 0, N >> sh)
    err = np.abs(f.astype(np.float64) - N / float(1 << sh)).max()
    assert err < 2.5e-4 < 1.0 / (1 << sh) / 2 or sh > 10, err
    unamb = (N & ((1 << sh) - 1)) != 0
    assert np.array_equal(got[unamb], E[unamb])
    amb = ~unamb & (N > 0)
    assert amb.sum() > 1000 and np.all((got[amb] == E[amb]) | (got[amb] == E[amb] - 1))
    assert (got[amb] != E[amb]).any(), "the ambiguous outputs do need the float chain"
