#!/usr/bin/env python3
"""C3 (8K P010 -> 4K RGB48, Lanczos) and C2 timings in one process (CUDA events, frames resident)"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from gmat_b200 import FMT, SWS, FrameBatch, SwsContext
dev = torch.device("cuda:0")
def run(label, sf, sw, sh, df, dw, dh, flag, param, B, bpf):
    src = FrameBatch(sf, sw, sh, B, device=dev); src.buf.random_(0, 256)
    dst = FrameBatch(df, dw, dh, B, device=dev)
    c = SwsContext(sw, sh, sf, dw, dh, df, flag | SWS.HWACCEL_CUDA, param)
    for _ in range(3): c.scale(src, dst)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): c.scale(src, dst)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{label}: {ms:.3f} ms {B*sw*sh/ms/1e6:.1f} Gpx/s {B*bpf/ms/1e6/6552.0*100:.1f}%", flush=True)
run("C3 8K P010->4K RGB48 lanczos B=16", FMT.P010LE, 7680, 4320, FMT.RGB48LE, 3840, 2160, SWS.LANCZOS, None, 16, 149299200)
run("C3 B=8", FMT.P010LE, 7680, 4320, FMT.RGB48LE, 3840, 2160, SWS.LANCZOS, None, 8, 149299200)
run("C2 A=-0.75 B=64", FMT.NV12, 3840, 2160, FMT.RGB24, 1920, 1080, SWS.BICUBIC, (0.75,), 64, 18662400)
run("C2 I420->BGRA A=-0.75 B=64", FMT.YUV420P, 3840, 2160, FMT.BGRA, 1920, 1080, SWS.BICUBIC, (0.75,), 64, 12441600 + 8294400)
