#!/usr/bin/env python3
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from gmat_b200 import FMT, SWS, FrameBatch, SwsContext
dev = torch.device("cuda:0"); B = int(os.environ.get("PERF_B", "64"))
src = FrameBatch(FMT.NV12, 3840, 2160, B, device=dev); src.buf.random_(0, 256)
dst = FrameBatch(FMT.RGB24, 1920, 1080, B, device=dev)
for label, flag, param in (("A=-0.75", SWS.BICUBIC, (0.75,)), ("A=0", SWS.BICUBIC, None), ("bilinear", SWS.BILINEAR, None)):
    c = SwsContext(3840, 2160, FMT.NV12, 1920, 1080, FMT.RGB24, flag | SWS.HWACCEL_CUDA, param)
    for _ in range(3): c.scale(src, dst)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): c.scale(src, dst)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"MINB={os.environ.get('GMATB_FUSED_MINB','-')} {label}: {ms:.3f} ms {B*3840*2160/ms/1e6:.1f} Gpx/s {B*18662400/ms/1e6/6552.0*100:.1f}%", flush=True)
