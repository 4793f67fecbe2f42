#!/usr/bin/env python3
"""filter kernels at 4K rgb24 (CUDA events, % of the measured HBM copy peak; algorithmic bytes = read + write once)"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import gmat_b200 as g
from gmat_b200 import FMT, BORDER, FrameBatch
dev = torch.device("cuda:0"); B = int(os.environ.get("PERF_B", "32")); PEAK = 6552.0
which = sys.argv[1:] or ["rotate", "gauss", "median3", "median5", "crop", "flip"]
for fmt, bpp in ((FMT.RGB24, 3), (FMT.BGRA, 4)):
    a = FrameBatch(fmt, 3840, 2160, B, device=dev); a.buf.random_(0, 256)
    b = FrameBatch(fmt, 3840, 2160, B, device=dev)
    ops = {"rotate": lambda: g.rotate(a, b, 30.0, -282.7688, 1104.6926, "linear"),
           "rotate7": lambda: g.rotate(a, b, 7.0, 130.0, -220.0, "linear"),
           "gauss": lambda: g.gaussian(a, b, 5, 5, 1.1, 1.1, BORDER.REFLECT101),
           "median3": lambda: g.median(a, b, 3, 3), "median5": lambda: g.median(a, b, 5, 5),
           "flip": lambda: g.flip(a, b, 1)}
    for name in which:
        if name not in ops: continue
        f = ops[name]
        for _ in range(3): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        px = B * 3840 * 2160
        print(f"{name:8s} bpp{bpp}: {ms:.3f} ms  {px/ms/1e6:7.1f} Gpx/s  {2*bpp*px/ms/1e6:7.1f} GB/s  {2*bpp*px/ms/1e6/PEAK*100:5.1f}%", flush=True)
