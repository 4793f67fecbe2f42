#!/usr/bin/env python3
"""C2 A/B: exact-integer fused kernel vs the float-chain kernel, noise and flat content (CUDA events)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from gmat_b200 import FMT, SWS, FrameBatch, SwsContext
dev = torch.device("cuda:0"); B = int(os.environ.get("PERF_B", "64"))
src = FrameBatch(FMT.NV12, 3840, 2160, B, device=dev)
dst = FrameBatch(FMT.RGB24, 1920, 1080, B, device=dev)
def fill(kind):
    if kind == "noise": src.buf.random_(0, 256)
    elif kind == "flat": src.buf.fill_(100)
    elif kind == "half":
        src.buf.random_(0, 256)
        v = src.buf.view(B, -1); v[:, v.shape[1] // 2:] = 100     # second half of each frame's bytes (lower luma rows + all chroma)
def run(label, flags, param):
    c = SwsContext(3840, 2160, FMT.NV12, 1920, 1080, FMT.RGB24, flags | SWS.HWACCEL_CUDA, param)
    for _ in range(3): c.scale(src, dst)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): c.scale(src, dst)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{label}: {ms:.3f} ms {B*3840*2160/ms/1e6:.1f} Gpx/s {B*18662400/ms/1e6/6552.0*100:.1f}%", flush=True)
for kind in ("noise", "flat", "half"):
    fill(kind)
    for p in ((0.75,), (0.5,)):
        run(f"{kind} mma   param {p}", SWS.BICUBIC | SWS.MMA_CHAIN, p)
        if os.environ.get("PERF_INT"): run(f"{kind} int   param {p}", SWS.BICUBIC | SWS.INT_CHAIN, p)
        run(f"{kind} float param {p}", SWS.BICUBIC, p)
