#!/usr/bin/env python3
"""kernel-level A/B timings on a GPU box (CUDA events, frames resident, inputs > L2)"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gmat_b200 as g
from gmat_b200 import FMT, SWS, BORDER, FrameBatch, SwsContext
dev = torch.device("cuda:0")
PEAK = 6552.0

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def report(name, ms, px, nbytes):
    print(f"{name:58s} {ms:8.3f} ms  {px/ms/1e6:8.1f} Gpx/s  {nbytes/ms/1e6:7.0f} GB/s  {nbytes/ms/1e6/PEAK*100:5.1f}% of measured peak", flush=True)

B = 64
src = FrameBatch(FMT.NV12, 3840, 2160, B, device=dev); src.buf.random_(0, 256)
dst = FrameBatch(FMT.RGB24, 1920, 1080, B, device=dev)
for label, flag, param in (("bicubic A=-0.75", SWS.BICUBIC, (0.75,)), ("bicubic default (A=0)", SWS.BICUBIC, None), ("lanczos", SWS.LANCZOS, None),
                           ("BILINEAR (reference's actual algo)", SWS.BILINEAR, None)):
    c = SwsContext(3840, 2160, FMT.NV12, 1920, 1080, FMT.RGB24, flag | SWS.HWACCEL_CUDA, param)
    ms = timeit(lambda: c.scale(src, dst))
    report(f"C2 4K NV12->1080p RGB24 {label}", ms, B * 3840 * 2160, B * 18662400)
g720 = FrameBatch(FMT.RGB24, 1280, 720, B, device=dev)
cg = SwsContext(3840, 2160, FMT.NV12, 1280, 720, FMT.RGB24, SWS.BICUBIC | SWS.HWACCEL_CUDA)
report("generic: 4K NV12->720p RGB24 bicubic (3:1)", timeit(lambda: cg.scale(src, g720), 3), B * 3840 * 2160, B * (12441600 + 2764800))
cg2 = SwsContext(3840, 2160, FMT.NV12, 1280, 720, FMT.RGBA, SWS.LANCZOS | SWS.HWACCEL_CUDA)
g720a = FrameBatch(FMT.RGBA, 1280, 720, B, device=dev)
report("generic: 4K NV12->720p RGBA lanczos (3:1)", timeit(lambda: cg2.scale(src, g720a), 3), B * 3840 * 2160, B * (12441600 + 3686400))
del g720, g720a
s1080 = FrameBatch(FMT.NV12, 1920, 1080, B, device=dev); s1080.buf.random_(0, 256)
g7 = FrameBatch(FMT.RGB24, 1280, 720, B, device=dev)
cg3 = SwsContext(1920, 1080, FMT.NV12, 1280, 720, FMT.RGB24, SWS.BICUBIC | SWS.HWACCEL_CUDA)
report("generic: 1080p NV12->720p RGB24 bicubic (1.5:1)", timeit(lambda: cg3.scale(s1080, g7), 3), B * 1920 * 1080, B * (3110400 + 2764800))
g4k = FrameBatch(FMT.RGB24, 3840, 2160, 16, device=dev)
cg4 = SwsContext(1920, 1080, FMT.NV12, 3840, 2160, FMT.RGB24, SWS.BICUBIC | SWS.HWACCEL_CUDA)
s16 = FrameBatch(FMT.NV12, 1920, 1080, 16, device=dev); s16.buf.random_(0, 256)
report("generic: 1080p NV12->4K RGB24 bicubic (1:2 up)", timeit(lambda: cg4.scale(s16, g4k), 3), 16 * 1920 * 1080, 16 * (3110400 + 24883200))
n7 = FrameBatch(FMT.NV12, 1280, 720, B, device=dev)
cg5 = SwsContext(1920, 1080, FMT.NV12, 1280, 720, FMT.NV12, SWS.BICUBIC | SWS.HWACCEL_CUDA)
report("generic: 1080p NV12->720p NV12 bicubic (planes)", timeit(lambda: cg5.scale(s1080, n7), 3), B * 1920 * 1080, B * (3110400 + 1382400))
del g7, g4k, s1080, s16, n7
d4 = FrameBatch(FMT.RGBA, 1920, 1080, B, device=dev)
c = SwsContext(3840, 2160, FMT.NV12, 1920, 1080, FMT.RGBA, SWS.BICUBIC | SWS.HWACCEL_CUDA, (0.75,))
report("C2 -> RGBA bicubic A=-0.75", timeit(lambda: c.scale(src, d4)), B * 3840 * 2160, B * (12441600 + 8294400))
full = FrameBatch(FMT.RGB24, 3840, 2160, B, device=dev)
report("unscaled 4K NV12->RGB24", timeit(lambda: g.yuv2rgb(src, full)), B * 3840 * 2160, B * (12441600 + 24883200))
fa = FrameBatch(FMT.RGBA, 3840, 2160, B, device=dev)
report("unscaled 4K NV12->RGBA", timeit(lambda: g.yuv2rgb(src, fa)), B * 3840 * 2160, B * (12441600 + 33177600))
back = FrameBatch(FMT.NV12, 3840, 2160, B, device=dev)
report("unscaled 4K RGB24->NV12", timeit(lambda: g.rgb2yuv(full, back)), B * 3840 * 2160, B * (12441600 + 24883200))
i420 = FrameBatch(FMT.YUV420P, 3840, 2160, B, device=dev)
report("unscaled 4K NV12->YUV420P", timeit(lambda: g.yuv2yuv(src, i420)), B * 3840 * 2160, B * 2 * 12441600)
bgr = FrameBatch(FMT.BGR24, 3840, 2160, B, device=dev)
report("unscaled 4K RGB24->BGR24", timeit(lambda: g.rgb24tobgr24(full, bgr)), B * 3840 * 2160, B * 2 * 24883200)
del full, fa, back, i420, bgr, d4
B3 = 16
s3 = FrameBatch(FMT.P010LE, 7680, 4320, B3, device=dev); s3.buf.random_(0, 256); s3.buf[0::2] &= 0xC0
d3 = FrameBatch(FMT.RGB48LE, 3840, 2160, B3, device=dev)
c3 = SwsContext(7680, 4320, FMT.P010LE, 3840, 2160, FMT.RGB48LE, SWS.LANCZOS | SWS.HWACCEL_CUDA)
report("C3 8K P010->4K RGB48 lanczos", timeit(lambda: c3.scale(s3, d3), 5), B3 * 7680 * 4320, B3 * 149299200)
del s3, d3
# filters, 4K rgb24 (C4 pieces)
Bf = 32
a = FrameBatch(FMT.RGB24, 3840, 2160, Bf, device=dev); a.buf.random_(0, 256)
b = FrameBatch(FMT.RGB24, 3840, 2160, Bf, device=dev)
fb = Bf * 2 * 24883200
report("rotate 30deg linear 4K rgb24", timeit(lambda: g.rotate(a, b, 30.0, -282.7688, 1104.6926, "linear"), 5), Bf * 3840 * 2160, fb)
report("gaussian 5x5 s1.1 reflect101 4K rgb24", timeit(lambda: g.gaussian(a, b, 5, 5, 1.1, 1.1, BORDER.REFLECT101), 5), Bf * 3840 * 2160, fb)
report("median 5x5 4K rgb24", timeit(lambda: g.median(a, b, 5, 5), 3), Bf * 3840 * 2160, fb)
report("median 3x3 4K rgb24", timeit(lambda: g.median(a, b, 3, 3), 3), Bf * 3840 * 2160, fb)
report("flip horizontal 4K rgb24", timeit(lambda: g.flip(a, b, 1), 5), Bf * 3840 * 2160, fb)
report("flip vertical 4K rgb24", timeit(lambda: g.flip(a, b, 0), 5), Bf * 3840 * 2160, fb)
cr = FrameBatch(FMT.RGB24, 1920, 1080, Bf, device=dev)
report("crop 4K->1080p centre rgb24", timeit(lambda: g.crop(a, cr, -1, -1), 5), Bf * 1920 * 1080, Bf * 2 * 6220800)
sc = SwsContext(3840, 2160, FMT.RGB24, 1920, 1080, FMT.RGB24, SWS.BICUBIC | SWS.HWACCEL_CUDA)
report("scale rgb24 4K->1080p bicubic (fused rgb path)", timeit(lambda: sc.scale(a, cr), 3), Bf * 3840 * 2160, Bf * (24883200 + 6220800))
# format_cuda kernels
fsrc = FrameBatch(FMT.NV12, 3840, 2160, 16, device=dev); fsrc.buf.random_(0, 256)
fpl = FrameBatch(FMT.RGBPF32LE, 3840, 2160, 16, device=dev)
report("format_cuda 4K NV12->RGBPF32", timeit(lambda: g.format_nv12_to_rgbpf32(fsrc, fpl, 2)), 16 * 3840 * 2160, 16 * (12441600 + 3 * 4 * 3840 * 2160))
fback = FrameBatch(FMT.NV12, 3840, 2160, 16, device=dev)
report("format_cuda 4K RGBPF32->NV12", timeit(lambda: g.format_rgbpf32_to_nv12(fpl, fback, 2)), 16 * 3840 * 2160, 16 * (12441600 + 3 * 4 * 3840 * 2160))
