#!/usr/bin/env python3
"""First-light check on a GPU box: a few conversions against the CPU oracle."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gmat_b200 as g
from gmat_b200 import FMT, SWS, FrameBatch, SwsContext
import orc

dev = torch.device("cuda:0")
def cmp(name, a, b):
    a = a.payload(); b = b.payload()
    n = int((a != b).sum())
    print(f"{name}: {'OK' if n == 0 else 'MISMATCH'} ({n} of {a.size} bytes differ, max |d| {int(np.abs(a.astype(int)-b.astype(int)).max())})")
    return n

for (w, h) in ((64, 48), (33, 17), (1920, 1080)):
    src = FrameBatch(FMT.NV12, w, h, 1); host = src.fill_lcg(seed=5)
    for df in (FMT.RGB24, FMT.BGRA, FMT.RGBA64LE):
        ref = FrameBatch(df, w, h, 1); orc.yuv2rgb(src, ref)
        ds = src.to(dev); dd = FrameBatch(df, w, h, 1, device=dev)
        g.yuv2rgb(ds, dd); torch.cuda.synchronize()
        cmp(f"nv12->{df} {w}x{h}", dd, ref)

for algo, flag, param in (("bicubic", SWS.BICUBIC, None), ("bicubic", SWS.BICUBIC, (0.75,)), ("lanczos", SWS.LANCZOS, None)):
    for (sw, sh, dw, dh) in ((64, 48, 32, 24), (512, 256, 256, 128), (64, 48, 40, 30), (33, 17, 50, 29)):
        src = FrameBatch(FMT.NV12, sw, sh, 2); src.fill_lcg(seed=9)
        ds = src.to(dev); dd = FrameBatch(FMT.RGB24, dw, dh, 2, device=dev)
        c = SwsContext(sw, sh, FMT.NV12, dw, dh, FMT.RGB24, flag | SWS.HWACCEL_CUDA, param)
        c.scale(ds, dd); torch.cuda.synchronize()
        tabs = (c.get_filter(0), c.get_filter(1))
        ref = FrameBatch(FMT.RGB24, dw, dh, 2); orc.yuv2rgb_scale(src, ref, tabs)
        cmp(f"nv12->rgb24 {algo} {param} {sw}x{sh}->{dw}x{dh} path {c.path}", dd, ref)

# timing of the headline configuration
B = 32
src = FrameBatch(FMT.NV12, 3840, 2160, B, device=dev); src.buf.random_(0, 256)
dst = FrameBatch(FMT.RGB24, 1920, 1080, B, device=dev)
for param in (None, (0.75,)):
    c = SwsContext(3840, 2160, FMT.NV12, 1920, 1080, FMT.RGB24, SWS.BICUBIC | SWS.HWACCEL_CUDA, param)
    for _ in range(3): c.scale(src, dst)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): c.scale(src, dst)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gpx = B * 3840 * 2160 / ms / 1e6
    print(f"C2 param={param}: {ms:.3f} ms per {B} frames, {gpx:.1f} Gpx/s, {gpx*2.25:.0f} GB/s algorithmic, path {c.path}")
s2 = FrameBatch(FMT.NV12, 3840, 2160, B, device=dev); s2.buf.copy_(src.buf)
d2 = FrameBatch(FMT.RGB24, 3840, 2160, B, device=dev)
for _ in range(3): g.yuv2rgb(s2, d2)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): g.yuv2rgb(s2, d2)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"unscaled NV12->RGB24 4K: {ms:.3f} ms per {B} frames, {B*3840*2160/ms/1e6:.1f} Gpx/s, {B*3840*2160*4.5/ms/1e6:.0f} GB/s")
