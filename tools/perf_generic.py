#!/usr/bin/env python3
"""any-ratio yuv -> rgb scaling: the streaming kernel (scale_stream.cuh) vs the shared-memory tile kernel (SWS.TILE_KERNEL);
source Gpx/s and % of the measured HBM copy peak (algorithmic bytes: 1.5 B per source pixel + 3 B per destination pixel)"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from gmat_b200 import FMT, SWS, FrameBatch, SwsContext
dev = torch.device("cuda:0"); PEAK = 6552.0
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (sw, sh, dw, dh, B) in ((1920, 1080, 1280, 720, 64), (3840, 2160, 1280, 720, 32), (1920, 1080, 3840, 2160, 16), (3840, 2160, 2560, 1440, 32)):
    src = FrameBatch(FMT.NV12, sw, sh, B, device=dev); src.buf.random_(0, 256)
    dst = FrameBatch(FMT.RGB24, dw, dh, B, device=dev)
    alg = B * (sw * sh * 1.5 + dw * dh * 3)
    for name, fl, par in (("bicubic .75", SWS.BICUBIC, (0.75,)), ("bilinear", SWS.BILINEAR, None)):
        for kname, extra in (("stream", 0), ("tile", SWS.TILE_KERNEL)):
            c = SwsContext(sw, sh, FMT.NV12, dw, dh, FMT.RGB24, fl | SWS.HWACCEL_CUDA | extra, par)
            ms = timeit(lambda: c.scale(src, dst))
            print(f"{sw}x{sh}->{dw}x{dh} {name:12s} {kname:6s}: {ms:.3f} ms {B*sw*sh/ms/1e6:7.1f} Gpx/s(src) {alg/ms/1e6:7.1f} GB/s {alg/ms/1e6/PEAK*100:5.1f}%", flush=True)
for (sw, sh, dw, dh, B) in ((3840, 2160, 1920, 1080, 32), (1920, 1080, 1280, 720, 64), (1920, 1080, 3840, 2160, 16)):
    src = FrameBatch(FMT.NV12, sw, sh, B, device=dev); src.buf.random_(0, 256)
    dst = FrameBatch(FMT.NV12, dw, dh, B, device=dev)
    alg = B * (sw * sh * 1.5 + dw * dh * 1.5)
    for name, fl, par in (("bicubic", SWS.BICUBIC, None), ("bilinear", SWS.BILINEAR, None)):
        for kname, extra in (("stream", 0), ("tile", SWS.TILE_KERNEL)):
            c = SwsContext(sw, sh, FMT.NV12, dw, dh, FMT.NV12, fl | SWS.HWACCEL_CUDA | extra, par)
            ms = timeit(lambda: c.scale(src, dst))
            print(f"nv12->nv12 {sw}x{sh}->{dw}x{dh} {name:9s} {kname:6s}: {ms:.3f} ms {B*sw*sh/ms/1e6:7.1f} Gpx/s(src) {alg/ms/1e6:7.1f} GB/s {alg/ms/1e6/PEAK*100:5.1f}%", flush=True)
for (sw, sh, dw, dh, B) in ((3840, 2160, 1920, 1080, 32), (1920, 1080, 1280, 720, 64)):
    src = FrameBatch(FMT.P010LE, sw, sh, B, device=dev); src.buf.random_(0, 256)
    dst = FrameBatch(FMT.P010LE, dw, dh, B, device=dev)
    alg = B * (sw * sh * 3.0 + dw * dh * 3.0)
    for kname, extra in (("stream", 0), ("tile", SWS.TILE_KERNEL)):
        c = SwsContext(sw, sh, FMT.P010LE, dw, dh, FMT.P010LE, SWS.BICUBIC | SWS.HWACCEL_CUDA | extra)
        ms = timeit(lambda: c.scale(src, dst))
        print(f"p010->p010 {sw}x{sh}->{dw}x{dh} bicubic   {kname:6s}: {ms:.3f} ms {B*sw*sh/ms/1e6:7.1f} Gpx/s(src) {alg/ms/1e6:7.1f} GB/s {alg/ms/1e6/PEAK*100:5.1f}%", flush=True)
for (sw, sh, dw, dh, B) in ((3840, 2160, 1920, 1080, 16), (1920, 1080, 1280, 720, 32)):
    src = FrameBatch(FMT.RGB0, sw, sh, B, device=dev); src.buf.random_(0, 256)
    dst = FrameBatch(FMT.RGB0, dw, dh, B, device=dev)
    alg = B * (sw * sh * 4.0 + dw * dh * 4.0)
    for kname, extra in (("stream", 0), ("tile", SWS.TILE_KERNEL)):
        c = SwsContext(sw, sh, FMT.RGB0, dw, dh, FMT.RGB0, SWS.BICUBIC | SWS.HWACCEL_CUDA | extra)
        ms = timeit(lambda: c.scale(src, dst))
        print(f"rgb0->rgb0 {sw}x{sh}->{dw}x{dh} bicubic   {kname:6s}: {ms:.3f} ms {B*sw*sh/ms/1e6:7.1f} Gpx/s(src) {alg/ms/1e6:7.1f} GB/s {alg/ms/1e6/PEAK*100:5.1f}%", flush=True)
for (sw, sh, dw, dh, B) in ((1920, 1080, 1280, 720, 32), (3840, 2160, 1280, 720, 16)):
    src = FrameBatch(FMT.RGB24, sw, sh, B, device=dev); src.buf.random_(0, 256)
    dst = FrameBatch(FMT.RGB24, dw, dh, B, device=dev)
    alg = B * (sw * sh * 3.0 + dw * dh * 3.0)
    for kname, extra in (("stream", 0), ("tile", SWS.TILE_KERNEL)):
        c = SwsContext(sw, sh, FMT.RGB24, dw, dh, FMT.RGB24, SWS.BICUBIC | SWS.HWACCEL_CUDA | extra)
        ms = timeit(lambda: c.scale(src, dst))
        print(f"rgb24->rgb24 {sw}x{sh}->{dw}x{dh} bicubic   {kname:6s}: {ms:.3f} ms {B*sw*sh/ms/1e6:7.1f} Gpx/s(src) {alg/ms/1e6:7.1f} GB/s {alg/ms/1e6/PEAK*100:5.1f}%", flush=True)
