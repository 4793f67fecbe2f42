#!/usr/bin/env python3
"""unscaled colour conversions at 4K, 64 frames resident (CUDA events; % of the measured HBM copy peak)"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import gmat_b200 as g
from gmat_b200 import FMT, FrameBatch
dev = torch.device("cuda:0"); B = 64; PEAK = 6552.0

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def report(name, ms, nbytes):
    print(f"{name:34s} {ms:7.3f} ms {B*3840*2160/ms/1e6:8.1f} Gpx/s {nbytes/ms/1e6:7.0f} GB/s {nbytes/ms/1e6/PEAK*100:6.1f}%", flush=True)

src = FrameBatch(FMT.NV12, 3840, 2160, B, device=dev); src.buf.random_(0, 256)
i420 = FrameBatch(FMT.YUV420P, 3840, 2160, B, device=dev); i420.buf.random_(0, 256)
full = FrameBatch(FMT.RGB24, 3840, 2160, B, device=dev)
fa = FrameBatch(FMT.RGBA, 3840, 2160, B, device=dev)
back = FrameBatch(FMT.NV12, 3840, 2160, B, device=dev)
report("NV12 -> RGB24", timeit(lambda: g.yuv2rgb(src, full)), B * (12441600 + 24883200))
report("YUV420P -> RGB24", timeit(lambda: g.yuv2rgb(i420, full)), B * (12441600 + 24883200))
report("NV12 -> RGBA", timeit(lambda: g.yuv2rgb(src, fa)), B * (12441600 + 33177600))
report("RGB24 -> NV12", timeit(lambda: g.rgb2yuv(full, back)), B * (12441600 + 24883200))
a = FrameBatch(FMT.RGB24, 3840, 2160, 32, device=dev); a.buf.random_(0, 256)
cr = FrameBatch(FMT.RGB24, 1920, 1080, 32, device=dev)
ms = timeit(lambda: g.crop(a, cr, -1, -1))
print(f"crop 4K -> 1080p centre rgb24        {ms:7.3f} ms {32*2*6220800/ms/1e6:7.0f} GB/s {32*2*6220800/ms/1e6/PEAK*100:6.1f}%", flush=True)
ms = timeit(lambda: g.crop(a, cr, 961, 3))
print(f"crop 4K -> 1080p at (961, 3) rgb24   {ms:7.3f} ms {32*2*6220800/ms/1e6:7.0f} GB/s {32*2*6220800/ms/1e6/PEAK*100:6.1f}%", flush=True)
