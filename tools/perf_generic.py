import os, sys, torch
sys.path.insert(0, '/root/repo')
import gmat_b200 as g
from gmat_b200 import FMT, SWS, FrameBatch, SwsContext
dev = torch.device("cuda:0")
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
B = 32
src = FrameBatch(FMT.NV12, 3840, 2160, B, device=dev); src.buf.random_(0, 256)
d720 = FrameBatch(FMT.RGB24, 1280, 720, B, device=dev)
c = SwsContext(3840, 2160, FMT.NV12, 1280, 720, FMT.RGB24, SWS.BICUBIC | SWS.HWACCEL_CUDA)
ms = timeit(lambda: c.scale(src, d720)); print(f"SMEM={os.environ.get('GMATB_GEN_SMEM')} 4K->720p: {B*3840*2160/ms/1e6:.1f} Gpx/s")
s1080 = FrameBatch(FMT.NV12, 1920, 1080, B, device=dev); s1080.buf.random_(0, 256)
c2 = SwsContext(1920, 1080, FMT.NV12, 1280, 720, FMT.RGB24, SWS.BICUBIC | SWS.HWACCEL_CUDA)
ms = timeit(lambda: c2.scale(s1080, d720)); print(f"SMEM={os.environ.get('GMATB_GEN_SMEM')} 1080p->720p: {B*1920*1080/ms/1e6:.1f} Gpx/s")
