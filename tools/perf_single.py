#!/usr/bin/env python3
"""single-frame launches (the granularity of an FFmpeg sws_scale call): 16 distinct 4K frames, one launch each"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from gmat_b200 import FMT, SWS, FrameBatch, SwsContext
dev = torch.device("cuda:0"); N = 16
src = [FrameBatch(FMT.NV12, 3840, 2160, 1, device=dev) for _ in range(N)]
for s in src: s.buf.random_(0, 256)
dst = [FrameBatch(FMT.RGB24, 1920, 1080, 1, device=dev) for _ in range(N)]
c = SwsContext(3840, 2160, FMT.NV12, 1920, 1080, FMT.RGB24, SWS.BICUBIC | SWS.HWACCEL_CUDA, (0.75,))
def run():
    for s, d in zip(src, dst): c.scale(s, d)
for _ in range(3): run()
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(f"MINBAND={os.environ.get('GMATB_FUSED_MINBAND','-')} single-frame launches: {best/N*1000:.1f} us/frame {N*3840*2160/best/1e6:.1f} Gpx/s", flush=True)
