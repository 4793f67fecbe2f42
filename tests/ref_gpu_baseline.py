#!/usr/bin/env python3
"""Reference-GPU baseline ("reported, same box, same run"): the reference's OWN CUDA kernels,
compiled unmodified for sm_100a into oracle/_ref (O1 = libswscale/cuda/*.cu, O2 = libavfilter/vf_scale_cuda.cu),
timed on this box beside our kernels on the same frames.

  * O1 `yuv2rgb_cuda` NV12 -> RGB24 at 4K, alone (one launch per frame, as the reference issues it)
  * O1 NV12 -> RGBA at 4K followed by O2 `Subsample_Bicubic_rgb0_rgb0` 4K -> 1080p: the closest in-tree
    proxy for the reference's unfused CSC + resize pipeline (its real resize is closed CV-CUDA)
  * ours: the unscaled converter and the fused kernel, per frame (same launch granularity) and batched

This is test infrastructure (it executes oracle/_ref); it is not a pytest module and not part of bench.py.
Usage on the GPU box:  python tests/ref_gpu_baseline.py > gpurun_out/ref_gpu_baseline.json
"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
import ctypes as C

import gpu_util
import gmat_b200 as g
from gmat_b200 import FMT, SWS, FrameBatch, SwsContext
from gpu_util import ci, vp

HW = SWS.HWACCEL_CUDA
dev = torch.device("cuda:0")
N = 16
SW, SH, DW, DH = 3840, 2160, 1920, 1080
PX = SW * SH


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def frames(fmt, w, h):
    return [FrameBatch(fmt, w, h, 1, device=dev) for _ in range(N)]


def main():
    src = frames(FMT.NV12, SW, SH)
    for i, s in enumerate(src):
        s.fill_lcg(seed=0xC0FFEE + i)
    rgb24 = frames(FMT.RGB24, SW, SH)
    rgba = frames(FMT.RGBA, SW, SH)
    out = frames(FMT.RGBA, DW, DH)
    res = {"frames": N, "src": "3840x2160 NV12", "unit": "Gpx/s (source pixels)"}

    L1 = gpu_util.o1()
    L1.set_mat_yuv2rgb_cuda(0)

    def o1_all(dsts):
        for s, d in zip(src, dsts):
            sp, ss = gpu_util.arrs(s.image()); dp, ds = gpu_util.arrs(d.image())
            assert L1.yuv2rgb_cuda(sp, ss, dp, ds, s.w, s.h, int(s.fmt), int(d.fmt), None) == 0

    ms = timeit(lambda: o1_all(rgb24))
    res["ref_o1_nv12_to_rgb24_4k"] = {"ms_per_frame": ms / N, "gpx_s": N * PX / ms / 1e6}

    L2 = gpu_util.o2()
    pad = lambda xs: list(xs) + [0] * (4 - len(xs))

    def o2_all():
        for m, d in zip(rgba, out):
            si, di = m.image(), d.image()
            rc = L2.ref_o2_launch(b"Subsample_Bicubic_rgb0_rgb0", 1, (vp * 4)(*pad([si.data[0]])), (ci * 4)(*pad([si.linesize[0]])),
                                  (ci * 4)(*pad([SW])), (ci * 4)(*pad([SH])), (ci * 4)(*pad([8])), (ci * 4)(*pad([4])),
                                  (vp * 4)(*pad([di.data[0]])), DW, DH, di.linesize[0], SW, SH, C.c_float(0.75), 0, 0)
            assert rc == 0

    def ref_pipeline():
        o1_all(rgba); o2_all()

    ms = timeit(ref_pipeline)
    res["ref_o1_plus_o2_nv12_to_rgba_1080p_bicubic"] = {"ms_per_frame": ms / N, "gpx_s": N * PX / ms / 1e6}
    ms = timeit(o2_all)
    res["ref_o2_bicubic_rgb0_4k_to_1080p_alone"] = {"ms_per_frame": ms / N, "gpx_s": N * PX / ms / 1e6}

    # ---- ours, same frames --------------------------------------------------------------
    def ours_csc():
        for s, d in zip(src, rgb24):
            g.yuv2rgb(s, d)

    ms = timeit(ours_csc)
    res["ours_nv12_to_rgb24_4k_per_frame_launch"] = {"ms_per_frame": ms / N, "gpx_s": N * PX / ms / 1e6}
    ctx = SwsContext(SW, SH, FMT.NV12, DW, DH, FMT.RGBA, SWS.BICUBIC | HW, (0.75,))

    def ours_fused():
        for s, d in zip(src, out):
            ctx.scale(s, d)

    ms = timeit(ours_fused)
    res["ours_fused_nv12_to_rgba_1080p_bicubic_per_frame_launch"] = {"ms_per_frame": ms / N, "gpx_s": N * PX / ms / 1e6}
    bs = FrameBatch(FMT.NV12, SW, SH, 64, device=dev); bs.buf.random_(0, 256)
    bd = FrameBatch(FMT.RGBA, DW, DH, 64, device=dev)
    ms = timeit(lambda: ctx.scale(bs, bd))
    res["ours_fused_nv12_to_rgba_1080p_bicubic_batch64"] = {"ms_per_frame": ms / 64, "gpx_s": 64 * PX / ms / 1e6}
    if "--brief" in sys.argv:
        res = {"unit": "Gpx/s (source pixels), 3840x2160 NV12, one launch per frame as the reference issues it",
               "o1_nv12_to_rgb24_4k": res["ref_o1_nv12_to_rgb24_4k"]["gpx_s"],
               "o1_plus_o2_nv12_to_1080p_bicubic": res["ref_o1_plus_o2_nv12_to_rgba_1080p_bicubic"]["gpx_s"],
               "o2_bicubic_rgb0_4k_to_1080p_alone": res["ref_o2_bicubic_rgb0_4k_to_1080p_alone"]["gpx_s"],
               "ours_nv12_to_rgb24_4k_per_frame_launch": res["ours_nv12_to_rgb24_4k_per_frame_launch"]["gpx_s"],
               "ours_fused_per_frame_launch": res["ours_fused_nv12_to_rgba_1080p_bicubic_per_frame_launch"]["gpx_s"],
               "ours_fused_batch64": res["ours_fused_nv12_to_rgba_1080p_bicubic_batch64"]["gpx_s"],
               "what": "the reference's own kernels (libswscale/cuda/yuv2rgb_cuda.cu, libavfilter/vf_scale_cuda.cu) compiled "
                       "unmodified for sm_100a; O1+O2 is the closest in-tree proxy of its unfused CSC -> resize pipeline"}
        print(json.dumps(res))
    else:
        print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
