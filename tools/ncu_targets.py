#!/usr/bin/env python3
"""a few launches of the kernels we want ncu captures of (run under ncu with -k filters)"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import gmat_b200 as g
from gmat_b200 import FMT, SWS, BORDER, INTERP, FrameBatch, SwsContext
dev = torch.device("cuda:0"); B = int(os.environ.get("NCU_B", "64"))
HW = SWS.HWACCEL_CUDA
src = FrameBatch(FMT.NV12, 3840, 2160, B, device=dev); src.buf.random_(0, 256)
d1080 = FrameBatch(FMT.RGB24, 1920, 1080, B, device=dev)
d720 = FrameBatch(FMT.RGB24, 1280, 720, B, device=dev)
which = sys.argv[1:] or ["fused", "bilinear", "generic", "rotate", "gauss", "median", "rgb2yuv", "fliph"]
if "fused" in which:
    SwsContext(3840, 2160, FMT.NV12, 1920, 1080, FMT.RGB24, SWS.BICUBIC | HW, (0.75,)).scale(src, d1080)
if "fused_mma" in which:
    SwsContext(3840, 2160, FMT.NV12, 1920, 1080, FMT.RGB24, SWS.BICUBIC | SWS.MMA_CHAIN | HW, (0.75,)).scale(src, d1080)
if "fused_default" in which:
    SwsContext(3840, 2160, FMT.NV12, 1920, 1080, FMT.RGB24, SWS.BICUBIC | HW).scale(src, d1080)
if "c3" in which:
    s8 = FrameBatch(FMT.P010LE, 7680, 4320, 16, device=dev); s8.buf.random_(0, 256); s8.buf[0::2] &= 0xC0
    d8 = FrameBatch(FMT.RGB48LE, 3840, 2160, 16, device=dev)
    SwsContext(7680, 4320, FMT.P010LE, 3840, 2160, FMT.RGB48LE, SWS.LANCZOS | HW).scale(s8, d8)
if "bilinear" in which:
    SwsContext(3840, 2160, FMT.NV12, 1920, 1080, FMT.RGB24, SWS.BILINEAR | HW).scale(src, d1080)
if "stream" in which:
    s1080 = FrameBatch(FMT.NV12, 1920, 1080, B, device=dev); s1080.buf.random_(0, 256)
    SwsContext(1920, 1080, FMT.NV12, 1280, 720, FMT.RGB24, SWS.BICUBIC | HW, (0.75,)).scale(s1080, d720)
if "plane" in which:
    dn = FrameBatch(FMT.NV12, 1920, 1080, B, device=dev)
    SwsContext(3840, 2160, FMT.NV12, 1920, 1080, FMT.NV12, SWS.BICUBIC | HW).scale(src, dn)
if "generic" in which:
    SwsContext(3840, 2160, FMT.NV12, 1280, 720, FMT.RGB24, SWS.BICUBIC | HW).scale(src, d720)
a = FrameBatch(FMT.RGB24, 3840, 2160, B, device=dev); a.buf.random_(0, 256)
b = FrameBatch(FMT.RGB24, 3840, 2160, B, device=dev)
if "rotate" in which:
    g.rotate(a, b, 30.0, -282.7688, 1104.6926, INTERP.LINEAR)      # BASELINE C4: about the centre
if "gauss" in which:
    g.gaussian(a, b, 5, 5, 1.1, 1.1, BORDER.REFLECT101)
if "median" in which:
    g.median(a, b, 3, 3)
if "median5" in which:
    g.median(a, b, 5, 5)
if "rgb2yuv" in which:
    back = FrameBatch(FMT.NV12, 3840, 2160, B, device=dev); g.rgb2yuv(a, back)
if "fliph" in which:
    g.flip(a, b, 1)
torch.cuda.synchronize()
