"""Frame containers: FFmpeg-style planes + strides, on the host (numpy) or on a GPU (torch).

Layouts follow the reference's CUDA frame pool (libavutil/hwcontext_cuda.c:143-199):
one allocation per frame, planes contiguous, pitch aligned (256 B here); NV12 / P010
keep UV at Y + H*pitch with the luma pitch, which is what the reference's kernels
assume (yuv2rgb_cuda.cu:226).
"""
import ctypes as C

import numpy as np

from .lib import FMT, GmatbImage

_YUV_SEMI = (FMT.NV12, FMT.P010LE, FMT.P016LE)
_YUV_PLANAR = (FMT.YUV420P, FMT.YUV420P10LE, FMT.YUV420P16LE)
_B16 = (FMT.P010LE, FMT.P016LE, FMT.YUV420P10LE, FMT.YUV420P16LE, FMT.RGB48LE, FMT.BGR48LE, FMT.RGBA64LE, FMT.BGRA64LE)
_RGB3 = (FMT.RGB24, FMT.BGR24, FMT.RGB48LE, FMT.BGR48LE)
_F32 = (FMT.RGBPF32LE, FMT.RGBAPF32LE)


def _align(v, a):
    return (v + a - 1) // a * a


def plane_layout(fmt, w, h, align=256):
    """-> (list of (offset, pitch, rows, row_bytes)), frame_bytes"""
    bs = 2 if fmt in _B16 else 1
    planes = []
    off = 0
    if fmt in _YUV_SEMI or fmt in _YUV_PLANAR:
        cw, ch = (w + 1) // 2, (h + 1) // 2
        pitch = _align(w * bs, align)
        planes.append((off, pitch, h, w * bs)); off += pitch * h
        if fmt in _YUV_SEMI:
            # same pitch as luma so that UV == Y + H*pitch (reference convention)
            planes.append((off, pitch, ch, cw * 2 * bs)); off += pitch * ch
        else:
            cp = _align(cw * bs, align)
            planes.append((off, cp, ch, cw * bs)); off += cp * ch
            planes.append((off, cp, ch, cw * bs)); off += cp * ch
    elif fmt in _F32:
        n = 3 if fmt == FMT.RGBPF32LE else 4
        pitch = _align(w * 4, align)
        for _ in range(n):
            planes.append((off, pitch, h, w * 4)); off += pitch * h
    else:
        ch = 3 if fmt in _RGB3 else 4
        pitch = _align(w * ch * bs, align)
        planes.append((off, pitch, h, w * ch * bs)); off += pitch * h
    return planes, _align(off, 256)


def lcg_bytes(n, seed):
    """SURVEY 8d synthetic generator: s = s*1664525 + 1013904223 (mod 2^32), byte = s >> 24."""
    a, c, M = 1664525, 1013904223, 1 << 32
    B = 1 << 16
    ai = np.empty(B, np.uint64); ci = np.empty(B, np.uint64)
    x, y = 1, 0
    for i in range(B):
        x = (x * a) % M; y = (y * a + c) % M
        ai[i] = x; ci[i] = y
    out = np.empty(n, np.uint8)
    s = seed & 0xFFFFFFFF
    pos = 0
    while pos < n:
        m = min(B, n - pos)
        blk = (ai[:m] * np.uint64(s) + ci[:m]) & np.uint64(M - 1)
        out[pos:pos + m] = (blk >> np.uint64(24)).astype(np.uint8)
        s = int(blk[m - 1]); pos += m
    return out


class FrameBatch:
    """`n` frames of one format in ONE buffer (numpy on the host, torch uint8 on a device)."""

    def __init__(self, fmt, w, h, n=1, device=None, pinned=False, align=256, buffer=None):
        self.fmt, self.w, self.h, self.n = fmt, w, h, n
        self.planes, self.frame_bytes = plane_layout(fmt, w, h, align)
        self.device = device
        total = self.frame_bytes * n
        if buffer is not None:
            self.buf = buffer
        elif device is None and not pinned:
            self.buf = np.zeros(total, np.uint8)
        else:
            import torch
            if device is None:
                self.buf = torch.zeros(total, dtype=torch.uint8, pin_memory=True)
            else:
                self.buf = torch.zeros(total, dtype=torch.uint8, device=device)
        self.is_torch = not isinstance(self.buf, np.ndarray)

    # ---- raw access ---------------------------------------------------------------------
    @property
    def ptr(self):
        return self.buf.data_ptr() if self.is_torch else self.buf.ctypes.data

    def numpy(self):
        """host view (copy for device tensors)"""
        if not self.is_torch:
            return self.buf
        return self.buf.cpu().numpy()

    def image(self, first=0, count=None):
        g = GmatbImage()
        count = self.n - first if count is None else count
        for i, (off, pitch, _rows, _rb) in enumerate(self.planes):
            g.data[i] = self.ptr + first * self.frame_bytes + off
            g.linesize[i] = pitch
            g.batch_stride[i] = self.frame_bytes
        g.width, g.height, g.format, g.batch = self.w, self.h, self.fmt, count
        return g

    def plane_view(self, host, frame, p):
        """2-D uint8 view [rows, row_bytes] of plane p of `frame` inside a host array"""
        off, pitch, rows, rb = self.planes[p]
        base = frame * self.frame_bytes + off
        return np.lib.stride_tricks.as_strided(host[base:], shape=(rows, rb), strides=(pitch, 1), writeable=True)

    def payload(self, host=None):
        """concatenated visible bytes of all planes of all frames (for exact comparisons / CRCs)"""
        host = self.numpy() if host is None else host
        parts = []
        for f in range(self.n):
            for p in range(len(self.planes)):
                parts.append(np.ascontiguousarray(self.plane_view(host, f, p)).reshape(-1))
        return np.concatenate(parts)

    def fill_lcg(self, seed=0xC0FFEE, ten_bit=None):
        """deterministic synthetic frames; seed + frame index per frame"""
        host = np.zeros(self.frame_bytes * self.n, np.uint8)
        for f in range(self.n):
            for p in range(len(self.planes)):
                v = self.plane_view(host, f, p)
                raw = lcg_bytes(v.size, seed + 7919 * f + 104729 * p).reshape(v.shape)
                if self.fmt in (FMT.P010LE, FMT.YUV420P10LE) if ten_bit is None else ten_bit:
                    # 10-bit samples are MSB-aligned: low 6 bits of every little-endian word are zero
                    raw = raw.copy(); raw[:, 0::2] &= 0xC0
                v[...] = raw
        self.upload(host)
        return host

    def upload(self, host):
        if self.is_torch:
            import torch
            self.buf.copy_(torch.from_numpy(host))
        else:
            self.buf[...] = host

    def to(self, device):
        import torch
        o = FrameBatch(self.fmt, self.w, self.h, self.n, device=device)
        o.planes, o.frame_bytes = self.planes, self.frame_bytes
        src = self.buf if self.is_torch else torch.from_numpy(self.buf)
        o.buf = src.to(device)
        o.is_torch = True
        return o


def ptr_arrays(img):
    """GmatbImage -> (uint8*[4], int[4]) like FFmpeg's data / linesize arrays"""
    p = (C.c_void_p * 4)(*[img.data[i] for i in range(4)])
    s = (C.c_int * 4)(*[img.linesize[i] for i in range(4)])
    return p, s
