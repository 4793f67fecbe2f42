#!/usr/bin/env python3
"""Generate the golden vectors that pin oracle/gmat_oracle.c.

Runs the REFERENCE's own CUDA kernels, compiled unmodified for sm_100a into
oracle/_ref/ (oracle/refbuild/Makefile):
  O1  libswscale/cuda/{yuv2rgb,yuv2yuv,rgb2rgb}_cuda*.cu   -> libref_gpuscale.so
  O2  libavfilter/vf_scale_cuda.cu (Subsample_* kernels)    -> ref_scale_cuda.cubin
on seeded LCG inputs and stores inputs' seeds + full outputs (small frames) or
CRC32s (large frames) in tests/golden/*.npz.  Needs a GPU:

    gpurun -- python tests/golden/make_golden.py gpurun_out/golden
    cp gpurun_out/golden/*.npz tests/golden/

The committed .npz files are what tests/test_oracle.py checks the CPU restatement
against (no GPU and no /root/reference needed at test time).
"""
import ctypes as C
import os
import sys
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gmat_b200 import FMT, FrameBatch  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
vp, ci = C.c_void_p, C.c_int


def load_o1():
    L = C.CDLL(os.path.join(REF, "libref_gpuscale.so"))
    for n in ("yuv2rgb_cuda", "rgb2yuv_cuda", "yuv2yuv_cuda"):
        f = getattr(L, n)
        f.restype = ci
        f.argtypes = [C.POINTER(vp), C.POINTER(ci), C.POINTER(vp), C.POINTER(ci), ci, ci, ci, ci, vp]
    L.rgb24tobgr24_cuda.argtypes = [C.POINTER(vp), C.POINTER(vp), C.POINTER(ci), C.POINTER(ci), ci, ci, vp]
    L.set_mat_yuv2rgb_cuda.argtypes = [ci]
    L.set_mat_rgb2yuv_cuda.argtypes = [ci]
    L.ref_p016_to_color64.argtypes = [vp, ci, vp, ci, ci, ci, ci, vp]
    L.ref_get_mat.argtypes = [ci, C.POINTER(C.c_float)]
    return L


def load_o2():
    L = C.CDLL(os.path.join(REF, "libref_o2_driver.so"))
    L.ref_o2_load.argtypes = [C.c_char_p]
    L.ref_o2_launch.argtypes = [C.c_char_p, ci, C.POINTER(vp), C.POINTER(ci), C.POINTER(ci), C.POINTER(ci),
                                C.POINTER(ci), C.POINTER(ci), C.POINTER(vp), ci, ci, ci, ci, ci, C.c_float, ci, ci]
    rc = L.ref_o2_load(os.path.join(REF, "ref_scale_cuda.cubin").encode())
    assert rc == 0, rc
    return L


def arrs(img):
    return (vp * 4)(*[img.data[i] for i in range(4)]), (ci * 4)(*[img.linesize[i] for i in range(4)])


def o1_convert(L, kind, src, dst, w, h):
    sp, ss = arrs(src.image())
    dp, ds = arrs(dst.image())
    if kind == "swap":
        L.rgb24tobgr24_cuda(sp, dp, ss, ds, w, h, None)
    else:
        getattr(L, kind)(sp, ss, dp, ds, w, h, src.fmt, dst.fmt, None)
    torch.cuda.synchronize()


def o2_scale(L2, func, src_planes, dst_ptrs, dw, dh, dpitch, sw, sh, param, linear=0, integer=0):
    """src_planes: list of (ptr, pitch, w, h, depth, channels)"""
    n = len(src_planes)
    pad = lambda xs: list(xs) + [0] * (4 - len(xs))
    rc = L2.ref_o2_launch(func.encode(), n,
                          (vp * 4)(*pad([p[0] for p in src_planes])), (ci * 4)(*pad([p[1] for p in src_planes])),
                          (ci * 4)(*pad([p[2] for p in src_planes])), (ci * 4)(*pad([p[3] for p in src_planes])),
                          (ci * 4)(*pad([p[4] for p in src_planes])), (ci * 4)(*pad([p[5] for p in src_planes])),
                          (vp * 4)(*pad(dst_ptrs)), dw, dh, dpitch, sw, sh, C.c_float(param), linear, integer)
    assert rc == 0, (func, rc)


SIZES = [(64, 48), (33, 17), (2, 2), (3, 3), (130, 6), (1, 1), (17, 33)]
PARAM_DEFAULT = 999999.0


def main(out):
    os.makedirs(out, exist_ok=True)
    dev = torch.device("cuda:0")
    L = load_o1()
    L2 = load_o2()
    G = {}
    # ---- matrices the reference uploads (SURVEY 8a KAT) -------------------------------
    for cs in (0, 1, 4, 5, 6, 7, 9):
        L.set_mat_yuv2rgb_cuda(cs); L.set_mat_rgb2yuv_cuda(cs)
        torch.cuda.synchronize()
        a = np.zeros(9, np.float32); b = np.zeros(9, np.float32)
        L.ref_get_mat(0, a.ctypes.data_as(C.POINTER(C.c_float))); L.ref_get_mat(1, b.ctypes.data_as(C.POINTER(C.c_float)))
        G[f"mat_y2r_{cs}"] = a; G[f"mat_r2y_{cs}"] = b
    # ---- O1 colour conversion -----------------------------------------------------------
    for cs in (0, 1):
        L.set_mat_yuv2rgb_cuda(cs); L.set_mat_rgb2yuv_cuda(cs)
        for (w, h) in SIZES:
            for sfmt, sname in ((FMT.NV12, "nv12"),):
                src = FrameBatch(sfmt, w, h, 1, device=dev)
                src.fill_lcg(seed=1234 + w * 131 + h)
                for dfmt, dname in ((FMT.RGB24, "rgb24"), (FMT.BGR24, "bgr24"), (FMT.RGBA, "rgba"), (FMT.BGRA, "bgra"),
                                    (FMT.RGBA64LE, "rgba64"), (FMT.BGRA64LE, "bgra64")):
                    dst = FrameBatch(dfmt, w, h, 1, device=dev)
                    o1_convert(L, "yuv2rgb_cuda", src, dst, w, h)
                    G[f"o1_{sname}_{dname}_{w}x{h}_cs{cs}"] = dst.payload()
        # rgb24 -> nv12 (even sizes only in the reference)
        for (w, h) in ((64, 48), (2, 2), (130, 6)):
            src = FrameBatch(FMT.RGB24, w, h, 1, device=dev)
            src.fill_lcg(seed=4321 + w)
            dst = FrameBatch(FMT.NV12, w, h, 1, device=dev)
            o1_convert(L, "rgb2yuv_cuda", src, dst, w, h)
            G[f"o1_rgb24_nv12_{w}x{h}_cs{cs}"] = dst.payload()
    L.set_mat_yuv2rgb_cuda(0); L.set_mat_rgb2yuv_cuda(0)
    # NV12 -> planar float (even sizes; the planar launcher skips odd edges)
    for (w, h) in ((64, 48), (2, 2), (130, 6)):
        src = FrameBatch(FMT.NV12, w, h, 1, device=dev); src.fill_lcg(seed=1234 + w * 131 + h)
        dst = FrameBatch(FMT.RGBPF32LE, w, h, 1, device=dev)
        o1_convert(L, "yuv2rgb_cuda", src, dst, w, h)
        G[f"o1_nv12_rgbpf32_{w}x{h}"] = dst.payload()
    # P016 -> RGBA64 / BGRA64 (undispatched template, even sizes)
    for (w, h) in ((64, 48), (2, 2), (130, 6)):
        src = FrameBatch(FMT.P016LE, w, h, 1, device=dev)
        src.fill_lcg(seed=99 + w)
        for order, dfmt, dname in ((0, FMT.RGBA64LE, "rgba64"), (1, FMT.BGRA64LE, "bgra64")):
            dst = FrameBatch(dfmt, w, h, 1, device=dev)
            g = src.image()
            L.ref_p016_to_color64(g.data[0], g.linesize[0], dst.image().data[0], dst.image().linesize[0], w, h, order, None)
            torch.cuda.synchronize()
            G[f"o1_p016_{dname}_{w}x{h}"] = dst.payload()
    # yuv2yuv + swap
    for (w, h) in ((64, 48), (34, 18)):
        src = FrameBatch(FMT.NV12, w, h, 1, device=dev); src.fill_lcg(seed=555 + w)
        for dfmt, dname in ((FMT.YUV420P, "yuv420p"), (FMT.P010LE, "p010"), (FMT.P016LE, "p016"),
                            (FMT.YUV420P10LE, "yuv420p10"), (FMT.YUV420P16LE, "yuv420p16")):
            dst = FrameBatch(dfmt, w, h, 1, device=dev)
            o1_convert(L, "yuv2yuv_cuda", src, dst, w, h)
            G[f"o1_nv12_{dname}_{w}x{h}"] = dst.payload()
        src = FrameBatch(FMT.YUV420P, w, h, 1, device=dev); src.fill_lcg(seed=777 + w)
        for dfmt, dname in ((FMT.NV12, "nv12"), (FMT.P010LE, "p010"), (FMT.YUV420P16LE, "yuv420p16")):
            dst = FrameBatch(dfmt, w, h, 1, device=dev)
            o1_convert(L, "yuv2yuv_cuda", src, dst, w, h)
            G[f"o1_yuv420p_{dname}_{w}x{h}"] = dst.payload()
        src = FrameBatch(FMT.RGB24, w, h, 1, device=dev); src.fill_lcg(seed=888 + w)
        dst = FrameBatch(FMT.BGR24, w, h, 1, device=dev)
        o1_convert(L, "swap", src, dst, w, h)
        G[f"o1_rgb24_bgr24_{w}x{h}"] = dst.payload()
    # one big frame: CRC only
    src = FrameBatch(FMT.NV12, 1920, 1080, 1, device=dev); src.fill_lcg(seed=2024)
    dst = FrameBatch(FMT.RGB24, 1920, 1080, 1, device=dev)
    o1_convert(L, "yuv2rgb_cuda", src, dst, 1920, 1080)
    G["o1_nv12_rgb24_1920x1080_crc"] = np.array([zlib.crc32(dst.payload().tobytes())], np.uint32)

    # ---- O2 resample: rgb0 (4 x u8), single-plane u8 (yuv420p luma), u16 (yuv420p16 luma) ----
    cases = [(64, 48, 32, 24), (64, 48, 40, 30), (33, 17, 50, 29), (16, 16, 7, 5), (130, 66, 65, 33), (8, 8, 8, 3)]
    for algo in ("Bicubic", "Lanczos"):
        for param in (PARAM_DEFAULT, 0.75, 0.5):
            if algo == "Lanczos" and param != PARAM_DEFAULT:
                continue
            pn = "def" if param == PARAM_DEFAULT else str(param)
            for (sw, sh, dw, dh) in cases:
                src = FrameBatch(FMT.RGBA, sw, sh, 1, device=dev); src.fill_lcg(seed=31 + sw + dw)
                dst = FrameBatch(FMT.RGBA, dw, dh, 1, device=dev)
                si, di = src.image(), dst.image()
                o2_scale(L2, f"Subsample_{algo}_rgb0_rgb0", [(si.data[0], si.linesize[0], sw, sh, 8, 4)],
                         [di.data[0]], dw, dh, di.linesize[0], sw, sh, param)
                G[f"o2_{algo}_{pn}_rgb0_{sw}x{sh}_{dw}x{dh}"] = dst.payload()
                # single 8-bit plane through the yuv420p kernel (luma launch only)
                src1 = FrameBatch(FMT.YUV420P, sw, sh, 1, device=dev); src1.fill_lcg(seed=77 + sw + dw)
                dst1 = FrameBatch(FMT.YUV420P, dw, dh, 1, device=dev)
                s1, d1 = src1.image(), dst1.image()
                cw, chh = (sw + 1) // 2, (sh + 1) // 2
                planes = [(s1.data[0], s1.linesize[0], sw, sh, 8, 1)]     # the luma launch only reads texture 0
                o2_scale(L2, f"Subsample_{algo}_yuv420p_yuv420p", planes, [d1.data[0], d1.data[1], d1.data[2]],
                         dw, dh, d1.linesize[0], sw, sh, param)
                host = dst1.numpy()
                G[f"o2_{algo}_{pn}_y8_{sw}x{sh}_{dw}x{dh}"] = np.ascontiguousarray(dst1.plane_view(host, 0, 0)).reshape(-1)
                # 16-bit plane: luma launch of the p016le kernel (plane 0 = 1 x u16, plane 1 = 2 x u16)
                src2 = FrameBatch(FMT.P016LE, sw, sh, 1, device=dev); src2.fill_lcg(seed=91 + sw + dw)
                dst2 = FrameBatch(FMT.P016LE, dw, dh, 1, device=dev)
                s2, d2 = src2.image(), dst2.image()
                planes = [(s2.data[0], s2.linesize[0], sw, sh, 16, 1)]
                o2_scale(L2, f"Subsample_{algo}_p016le_p016le", planes, [d2.data[0], d2.data[1]],
                         dw, dh, d2.linesize[0], sw, sh, param)
                host = dst2.numpy()
                G[f"o2_{algo}_{pn}_y16_{sw}x{sh}_{dw}x{dh}"] = np.ascontiguousarray(dst2.plane_view(host, 0, 0)).reshape(-1)
    # nearest (integer path)
    for (sw, sh, dw, dh) in cases:
        src = FrameBatch(FMT.RGBA, sw, sh, 1, device=dev); src.fill_lcg(seed=31 + sw + dw)
        dst = FrameBatch(FMT.RGBA, dw, dh, 1, device=dev)
        si, di = src.image(), dst.image()
        o2_scale(L2, "Subsample_Nearest_rgb0_rgb0", [(si.data[0], si.linesize[0], sw, sh, 8, 4)],
                 [di.data[0]], dw, dh, di.linesize[0], sw, sh, PARAM_DEFAULT, 0, 1)
        G[f"o2_Nearest_def_rgb0_{sw}x{sh}_{dw}x{dh}"] = dst.payload()
    np.savez_compressed(os.path.join(out, "reference_gpu_golden.npz"), **G)
    print("wrote", len(G), "golden arrays")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
