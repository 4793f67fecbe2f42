#!/usr/bin/env python3
"""Generates tests/golden/reference_lanczos_tables.npz ON A B200 (gpurun): the Lanczos coefficient tables of the
reference's own lanczos_coeffs (libavfilter/vf_scale_cuda.cu:948-968, fast-math __sinf), evaluated by
oracle/_ref/libref_o2_coeffs.so (the reference file #included in place) at the fractional positions of every
(src, dst) axis the O2 golden vectors of reference_gpu_golden.npz use, plus the BASELINE C3 axes.  The CPU oracle's
Lanczos cannot reproduce __sinf; these tables pin its resample arithmetic to the reference's outputs instead
(tests/test_oracle.py::test_oracle_vs_reference_resample_golden).

    gpurun -- 'python tests/golden/make_golden_lanczos.py gpurun_out/reference_lanczos_tables.npz'
"""
import ctypes as C
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def frac_positions(src_n, dst_n):
    """fx of every output of one axis exactly as vf_scale_cuda.cu computes it in binary32:
    xi = fma(o + 0.5, src/dst, -0.5); fx = xi - floor(xi)   (resample_core.cuh)"""
    scale = np.float32(src_n) / np.float32(dst_n)
    o = np.arange(dst_n, dtype=np.float32) + np.float32(0.5)
    xi = (o.astype(np.float64) * np.float64(scale) - 0.5).astype(np.float32)      # exact product, one rounding: an fma
    return (xi - np.floor(xi)).astype(np.float32), np.floor(xi).astype(np.int32) - 1


def ref_coeffs(lanczos, x, param=999999.0):
    L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_o2_coeffs.so"))
    L.ref_o2_coeffs.argtypes = [C.c_int, C.POINTER(C.c_float), C.c_float, C.POINTER(C.c_float), C.c_int]
    out = np.zeros((len(x), 4), np.float32)
    x = np.ascontiguousarray(x, np.float32)
    rc = L.ref_o2_coeffs(int(lanczos), x.ctypes.data_as(C.POINTER(C.c_float)), param,
                         out.ctypes.data_as(C.POINTER(C.c_float)), len(x))
    assert rc == 0, rc
    return out


def axes():
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_gpu_golden.npz"))
    s = set()
    for name in g.files:
        m = re.match(r"o2_Lanczos_def_\w+?_(\d+)x(\d+)_(\d+)x(\d+)$", name)
        if m:
            sw, sh, dw, dh = (int(v) for v in m.groups())
            s.add((sw, dw)); s.add((sh, dh))
    s |= {(7680, 3840), (4320, 2160), (3840, 1920), (2160, 1080), (1920, 1280), (1080, 720), (264, 132), (72, 36)}
    return sorted(s)


if __name__ == "__main__":
    out = {}
    for (sn, dn) in axes():
        fx, pos = frac_positions(sn, dn)
        out[f"lanczos_{sn}_{dn}"] = ref_coeffs(1, fx)
        out[f"pos_{sn}_{dn}"] = pos
        # the bicubic tables are reproducible on the CPU; kept as a cross-check of frac_positions() itself
        out[f"bicubic075_{sn}_{dn}"] = ref_coeffs(0, fx, 0.75)
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "reference_lanczos_tables.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays")
