#!/usr/bin/env python3
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c2_default|c3|csc|c4|c5]

A "step" is `launches_per_step` passes of the hot path over one batch of synthetic frames (one launch converts the
whole batch; the repeat count is chosen so that the K timed steps last about a second: config.launches_per_step).
At N=1 the workload
is BASELINE config[1] ("C2"): 3840x2160 NV12 -> 1920x1080 RGB24, bicubic (R-B; headline
param0 = 0.75 so that all 16 taps are live; the default-parameter result is reported
beside it).  `value` is whole-job source Gpixels/s with the frames resident in HBM; `e2e` is the
same metric through the public host-buffer call (gmatb_sws_scale_host: H2D + kernels + D2H every
step); `roofline` relates the kernel to the measured HBM copy peak; `cpu_baseline` times the
reference's own CPU libswscale on this box's cores.  N > 1: one process per GPU (torchrun),
frames sharded by batch index, no data-path collective (weak scaling); time = max over ranks.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (src fmt name, sw, sh, dst fmt name, dw, dh, sws flag name, param0, algorithmic bytes per frame)
    "c2":         ("NV12", 3840, 2160, "RGB24", 1920, 1080, "BICUBIC", 0.75),
    "c2_default": ("NV12", 3840, 2160, "RGB24", 1920, 1080, "BICUBIC", None),
    "c3":         ("P010LE", 7680, 4320, "RGB48LE", 3840, 2160, "LANCZOS", None),
    "csc":        ("NV12", 3840, 2160, "RGB24", 3840, 2160, "BICUBIC", None),
    "c1":         ("NV12", 1920, 1080, "RGB24", 1920, 1080, "BICUBIC", None),     # BASELINE configs[0]
}
# BASELINE configs[3] / configs[4] (chain and mixed sizes) are measured by run_c4 / run_c5 below
METRIC = "Gpixels/s 4K NV12->RGB24+bicubic->1080p; %HBM roofline; 1/2/4/8 GPU"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons / power DURING the timed region, read through NVML (nvidia_ml_py) every 20 ms; falls
    back to one `nvidia-smi -lms` child when NVML cannot be loaded.  (Round 1 polled nvidia-smi only: with eight ranks
    starting it at once it had not produced a sample by the time the 11 ms timed region was over.)"""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw"
    BITS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4),
            ("hw_power_brake_slowdown", 0x80))

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag = index, False
        self.sm, self.reasons, self.power, self.sm_max, self.source, self.proc = [], set(), [], None, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.physical_index(index))
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.source = "nvml"
        except Exception:
            self.nv = None

    @staticmethod
    def physical_index(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v for v in vis.split(",") if v.strip() != ""]
            if index < len(ids) and ids[index].strip().isdigit():
                return int(ids[index])
        return index

    def run(self):
        if self.nv is not None:
            while not self.stop_flag:
                try:
                    self.sm.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                    r = int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                    self.reasons |= {n for n, bit in self.BITS if r & bit}
                    self.power.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                except Exception:
                    pass
                time.sleep(0.02)
            return
        try:
            p = subprocess.Popen(["nvidia-smi", f"--id={self.physical_index(self.index)}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                  "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        self.proc, self.source = p, "nvidia-smi"
        names = [n for n, _ in self.BITS]
        for line in p.stdout:
            if self.stop_flag:
                break
            f = [x.strip() for x in line.split(",")]
            if f and f[0].replace(".", "").isdigit():
                self.sm.append(float(f[0])); self.sm_max = float(f[1])
                self.reasons |= {names[i] for i in range(4) if len(f) > 2 + i and f[2 + i].lower().startswith("active")}
                if len(f) > 6 and f[6].replace(".", "").isdigit():
                    self.power.append(float(f[6]))
        p.kill()

    def stop(self):
        self.stop_flag = True
        if self.proc:
            self.proc.kill()

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["no clock samples (NVML and nvidia-smi unavailable)"]}
        return {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "power_w_max": max(self.power) if self.power else None, "samples": len(self.sm), "source": self.source}


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


# --------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own CPU libswscale (oracle/_ref), all host threads
# --------------------------------------------------------------------------------------------
def cpu_reference(workload, budget_s=12.0, threads=None, with_sliced=True):
    """The reference's own CPU libswscale (oracle/_ref/libref_swscale_cpu.so: C only, no x86 asm) on this box.
    Two figures: (i) frame-parallel -- one single-threaded context per host thread, each converting its own
    frames: the fair throughput number, and `value`; (ii) `sliced`: ONE context with threads = nproc through
    sws_scale_frame, libswscale's own slice threading (swscale.c:1130-1198).
    CPU libswscale's param[0], param[1] are the B and C of the BC-spline family (libswscale/utils.c:476-497): the R-B
    bicubic with A = -0.75 that the GPU side runs is B = 0, C = 0.75 there, which is what is passed."""
    from gmat_b200 import FMT, SWS, FrameBatch
    sname, sw, sh, dname, dw, dh, flag, param0 = WORKLOADS[workload]
    ref = os.path.join(ROOT, "oracle", "_ref", "libref_swscale_cpu.so")
    stubs = os.path.join(ROOT, "oracle", "_ref", "libref_cuda_stubs.so")
    if not (os.path.exists(ref) and os.path.exists(stubs)):
        return None
    C.CDLL(stubs, mode=C.RTLD_GLOBAL)
    L = C.CDLL(ref, mode=C.RTLD_GLOBAL)
    vp, ci = C.c_void_p, C.c_int
    L.sws_getContext.restype = vp
    L.sws_getContext.argtypes = [ci, ci, ci, ci, ci, ci, ci, vp, vp, C.POINTER(C.c_double)]
    L.sws_scale.restype = ci
    L.sws_scale.argtypes = [vp, C.POINTER(vp), C.POINTER(ci), ci, ci, C.POINTER(vp), C.POINTER(ci)]
    L.sws_freeContext.argtypes = [vp]
    sfmt, dfmt = getattr(FMT, sname), getattr(FMT, dname)
    flags = getattr(SWS, flag)
    if flag == "LANCZOS":
        prm = (4.0, SWS.PARAM_DEFAULT)                       # C3 "Lanczos-4" 
    elif param0 is not None:
        prm = (0.0, float(param0))                           # B = 0, C = -A
    else:
        prm = (SWS.PARAM_DEFAULT, SWS.PARAM_DEFAULT)
    pp = (C.c_double * 2)(*prm)
    nthreads = threads or len(os.sched_getaffinity(0)) or 1
    src = FrameBatch(sfmt, sw, sh, 1); src.fill_lcg(seed=0xC0FFEE)

    def worker(nframes, out):
        ctx = L.sws_getContext(sw, sh, sfmt, dw, dh, dfmt, flags, None, None, pp)
        dst = FrameBatch(dfmt, dw, dh, 1)
        si, di = src.image(), dst.image()
        sp = (vp * 4)(*[si.data[i] for i in range(4)]); ss = (ci * 4)(*[si.linesize[i] for i in range(4)])
        dp = (vp * 4)(*[di.data[i] for i in range(4)]); ds = (ci * 4)(*[di.linesize[i] for i in range(4)])
        for _ in range(nframes):
            L.sws_scale(ctx, sp, ss, 0, sh, dp, ds)
        L.sws_freeContext(ctx)
        out.append(nframes)

    t0 = time.time(); worker(1, []); one = time.time() - t0          # calibration (also warms the tables)
    per_thread = max(1, int(budget_s / max(one, 1e-4) / 1.0))
    per_thread = min(per_thread, 64)
    done = []
    ths = [threading.Thread(target=worker, args=(per_thread, done)) for _ in range(nthreads)]
    t0 = time.time()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.time() - t0
    frames = sum(done)
    res = {"value": frames * sw * sh / dt / 1e9, "unit": "Gpx/s", "cores": nthreads, "kind": "reference", "cpu": cpu_model(),
           "sample": f"{frames} frames of {sw}x{sh} {sname}->{dw}x{dh} {dname} through the reference's libswscale "
                     f"(C only, no x86 asm; oracle/_ref/libref_swscale_cpu.so), {nthreads} frame-parallel single-thread contexts, "
                     f"param (B, C) = {prm if prm[0] != SWS.PARAM_DEFAULT else 'default'}, {dt:.1f} s",
           "seconds": dt, "frames": frames}
    if with_sliced:
        try:
            res["sliced"] = cpu_reference_sliced(L, src, sfmt, dfmt, sw, sh, dw, dh, flags, prm, nthreads, min(4.0, budget_s / 2))
        except Exception as e:                                # reported-only figure: never fail the bench line over it
            res["sliced"] = {"value": None, "error": repr(e)[:200]}
    return res


class _AVFrameHead(C.Structure):
    """the leading fields of AVFrame (libavutil/frame.h:325-400, libavutil 57): all this harness has to fill"""
    _fields_ = [("data", C.c_void_p * 8), ("linesize", C.c_int * 8), ("extended_data", C.c_void_p),
                ("width", C.c_int), ("height", C.c_int), ("nb_samples", C.c_int), ("format", C.c_int)]


def cpu_reference_sliced(L, src, sfmt, dfmt, sw, sh, dw, dh, flags, prm, nthreads, budget_s):
    """one SwsContext with threads = nproc, driven through sws_scale_frame (the only threaded entry point)"""
    vp, ci = C.c_void_p, C.c_int
    L.sws_alloc_context.restype = vp
    L.av_opt_set_int.argtypes = [vp, C.c_char_p, C.c_int64, ci]
    L.av_opt_set_double.argtypes = [vp, C.c_char_p, C.c_double, ci]
    L.sws_init_context.argtypes = [vp, vp, vp]
    L.av_frame_alloc.restype = vp
    L.av_frame_get_buffer.argtypes = [vp, ci]
    L.av_frame_free.argtypes = [C.POINTER(vp)]
    L.sws_scale_frame.argtypes = [vp, vp, vp]
    ctx = L.sws_alloc_context()
    for k, v in (("srcw", sw), ("srch", sh), ("src_format", sfmt), ("dstw", dw), ("dsth", dh), ("dst_format", dfmt),
                 ("sws_flags", flags), ("threads", nthreads)):
        assert L.av_opt_set_int(ctx, k.encode(), int(v), 0) == 0, k
    from gmat_b200 import SWS
    if prm[0] != SWS.PARAM_DEFAULT:
        L.av_opt_set_double(ctx, b"param0", prm[0], 0)
    if prm[1] != SWS.PARAM_DEFAULT:
        L.av_opt_set_double(ctx, b"param1", prm[1], 0)
    assert L.sws_init_context(ctx, None, None) >= 0
    fs, fd = vp(L.av_frame_alloc()), vp(L.av_frame_alloc())
    hs, hd = _AVFrameHead.from_address(fs.value), _AVFrameHead.from_address(fd.value)
    hs.width, hs.height, hs.format = sw, sh, int(sfmt)
    hd.width, hd.height, hd.format = dw, dh, int(dfmt)
    assert L.av_frame_get_buffer(fs, 0) == 0 and L.av_frame_get_buffer(fd, 0) == 0
    si = src.image()
    for p, (off, pitch, rows, rb) in enumerate(src.planes):      # copy the synthetic frame into the AVFrame's planes
        for r in range(rows):
            C.memmove(hs.data[p] + r * hs.linesize[p], si.data[p] + r * pitch, rb)
    assert L.sws_scale_frame(ctx, fd, fs) >= 0                   # warm-up
    n, t0 = 0, time.time()
    while time.time() - t0 < budget_s:
        L.sws_scale_frame(ctx, fd, fs); n += 1
    dt = time.time() - t0
    L.av_frame_free(C.byref(fs)); L.av_frame_free(C.byref(fd)); L.sws_freeContext(ctx)
    return {"value": n * sw * sh / dt / 1e9, "unit": "Gpx/s", "threads": nthreads, "frames": n, "seconds": dt,
            "how": "one context, threads = nproc, sws_scale_frame (libswscale slice threading)"}


def reference_gpu_baseline(timeout_s=180):
    """SURVEY 8d "reference-GPU baseline (reported, same box, same run)": the reference's own CUDA kernels compiled
    unmodified for sm_100a (oracle O1 / O2), timed by tests/ref_gpu_baseline.py in a child process after our own
    measurements are over (the oracle never shares a process with a timed region of ours)."""
    script = os.path.join(ROOT, "tests", "ref_gpu_baseline.py")
    if not (os.path.exists(script) and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_gpuscale.so"))):
        return None
    try:
        out = subprocess.run([sys.executable, script, "--brief"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=timeout_s)
        return json.loads(out.stdout)
    except Exception as e:
        return {"error": repr(e)[:200]}


def c4_workload(dev, B):
    """BASELINE configs[3]: 4K rgb24 rotate(30 deg about the centre, linear) -> gaussian 5x5 (sigma 1.1, reflect101) -> bicubic
    scale to 1080p, every stage materialising its frame like a filtergraph: 130 636 800 algorithmic bytes per frame"""
    import gmat_b200 as g
    from gmat_b200 import BORDER, FMT, SWS, FrameBatch, SwsContext
    a = FrameBatch(FMT.RGB24, 3840, 2160, B, device=dev); a.buf.random_(0, 256)
    b = FrameBatch(FMT.RGB24, 3840, 2160, B, device=dev)
    c = FrameBatch(FMT.RGB24, 3840, 2160, B, device=dev)
    d = FrameBatch(FMT.RGB24, 1920, 1080, B, device=dev)
    sc = SwsContext(3840, 2160, FMT.RGB24, 1920, 1080, FMT.RGB24, SWS.BICUBIC | SWS.HWACCEL_CUDA)

    def step():
        g.rotate(a, b, 30.0, -282.7688, 1104.6926, "linear")
        g.gaussian(b, c, 5, 5, 1.1, 1.1, BORDER.REFLECT101)
        sc.scale(c, d)
    return step, B * 3840 * 2160, B * 130636800, \
        f"C4: 4K rgb24 rotate(30deg, linear) -> gaussian 5x5 sigma 1.1 reflect101 -> bicubic scale to 1080p, each stage materialised, {B} frames per GPU"


def c5_workload(dev, n):
    """BASELINE configs[4]: equal thirds (by count) of 1080p / 4K / 8K NV12 -> RGB24 at half size, bicubic R-B param0 = 0.75"""
    from gmat_b200 import FMT, SWS, FrameBatch, SwsContext
    items = []
    for (w, h) in ((1920, 1080), (3840, 2160), (7680, 4320)):
        s_ = FrameBatch(FMT.NV12, w, h, n, device=dev); s_.buf.random_(0, 256)
        d_ = FrameBatch(FMT.RGB24, w // 2, h // 2, n, device=dev)
        items.append((SwsContext(w, h, FMT.NV12, w // 2, h // 2, FMT.RGB24, SWS.BICUBIC | SWS.HWACCEL_CUDA, (0.75,)), s_, d_))

    def step():
        for ctx, s_, d_ in items:
            ctx.scale(s_, d_)
    px = sum(s_.n * s_.w * s_.h for _, s_, _ in items)
    return step, px, int(px * 2.25), f"C5: mixed 1080p/4K/8K NV12 -> RGB24 at half size, bicubic R-B param0=0.75, {n}+{n}+{n} frames per GPU"


class Timer:
    """CUDA-event timing of `steps` calls bracketed by barrier + synchronize on both sides, max over ranks"""

    def __init__(self, torch, dist, world, dev):
        self.torch, self.dist, self.world, self.dev = torch, dist, world, dev

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world > 1:
            t = self.torch.tensor([v], device=self.dev, dtype=self.torch.float64)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            return t.item()
        return v

    def run(self, fn, steps, warmup=3):
        for _ in range(warmup):
            fn()
        self.barrier()
        e0, e1 = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1)) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + ["c4", "c5"])
    ap.add_argument("--batch", type=int, default=64, help="frames per GPU per launch")
    ap.add_argument("--repeat", type=int, default=0, help="launches per step (0: chosen so that the timed region lasts ~1 s)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / reference_gpu legs")
    ap.add_argument("--no-secondary", action="store_true", help="headline only")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        # the reference's own CPU implementation of the path, all host threads; rank 0 only
        if rank != 0:
            return 0
        wl = args.workload if args.workload in WORKLOADS else "c2"
        sname, sw, sh, dname, dw, dh, flag, param0 = WORKLOADS[wl]
        wl_desc = workload_desc(wl)
        r = cpu_reference(wl, budget_s=max(2.0, 1.5 * args.steps))
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libref_swscale_cpu.so not built"}))
            return 0
        cb = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "cpu")}
        cb["sliced"] = r.get("sliced")
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "Gpx/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] / max(args.steps, 1), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": shared_config(wl, args.batch),
                "note": "reference CPU libswscale (its CUDA path needs the closed CV-CUDA library)",
                "cpu_baseline": cb,
                "e2e": {"value": r["value"], "unit": "Gpx/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    import gmat_b200 as g
    from gmat_b200 import FMT, SWS, FrameBatch, SwsContext
    from gmat_b200.dist import bind_rank_to_cores, init, pipelined_scatter_compute_gather
    local = int(os.environ.get("LOCAL_RANK", "0"))
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    cores = bind_rank_to_cores(local, local_world) if world > 1 else None      # before any pinned allocation
    rank, world, local = init()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: gmat_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    L = g.lib()
    T = Timer(torch, dist, world, dev)
    peak, peak_src = measured_peak()
    B = args.batch

    # ---- the workload of this run ---------------------------------------------------------------------
    if args.workload in ("c4", "c5"):
        if args.workload == "c4":
            fn, px, alg_bytes_launch, wl_desc = c4_workload(dev, max(1, 256 // max(world, 8)))
        else:
            fn, px, alg_bytes_launch, wl_desc = c5_workload(dev, max(1, B // 3))
        ctx = src = dst = None
        kernel_path = "filters" if args.workload == "c4" else "fused_csc_scale2"
    else:
        sname, sw, sh, dname, dw, dh, flag, param0 = WORKLOADS[args.workload]
        wl_desc = workload_desc(args.workload)
        sfmt, dfmt = getattr(FMT, sname), getattr(FMT, dname)
        flags = getattr(SWS, flag) | SWS.HWACCEL_CUDA
        ctx = SwsContext(sw, sh, sfmt, dw, dh, dfmt, flags, None if param0 is None else (param0,))
        src = FrameBatch(sfmt, sw, sh, B, device=dev)
        # synthetic frames: LCG bytes, one distinct frame per slot built on the device from a host seed frame
        seed = FrameBatch(sfmt, sw, sh, 1); hseed = seed.fill_lcg(seed=0xC0FFEE + rank)
        t = torch.from_numpy(hseed).to(dev)
        for i in range(B):
            src.buf[i * src.frame_bytes:(i + 1) * src.frame_bytes] = torch.roll(t, shifts=i * 4099) if i else t
        if sname == "P010LE":
            src.buf[0::2] &= 0xC0
        dst = FrameBatch(dfmt, dw, dh, B, device=dev)
        alg_frame = sum(p[2] * p[3] for p in src.planes) + sum(p[2] * p[3] for p in dst.planes)
        px, alg_bytes_launch = B * sw * sh, B * alg_frame
        kernel_path = {0: "unscaled", 1: "fused_csc_scale2", 2: "generic"}[ctx.path]

        def fn():
            ctx.scale(src, dst)

    # ---- launches per step: the K timed steps should last about a second -----------------------------------
    one = T.run(fn, 3, warmup=max(args.warmup, 3))                       # ms per launch, max over ranks
    R = args.repeat if args.repeat > 0 else int(min(512, max(1, round(1000.0 / (args.steps * max(one, 1e-3))))))

    def step():
        for _ in range(R):
            fn()

    for _ in range(max(args.warmup, 3)):
        step()
    T.barrier()
    sampler = ClockSampler(local); sampler.start(); time.sleep(0.1)
    l0 = L.gmatb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    T.barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    T.barrier()
    launches = L.gmatb_launch_count() - l0
    ms = T.max_over_ranks(e0.elapsed_time(e1))
    time.sleep(0.05); sampler.stop()
    ms_step = ms / args.steps
    ms_launch = ms_step / R
    value = world * px / (ms_launch * 1e-3) / 1e9
    achieved = alg_bytes_launch / (ms_launch * 1e-3) / 1e9

    # ---- secondary measurements (reported beside the headline, every N) --------------------------------------
    secondary = []
    if not args.no_secondary and args.workload == "c2":
        def sec(label, f, spx, salg, steps=max(5, args.steps // 2)):
            m = T.run(f, steps)
            secondary.append({"workload": label, "value": world * spx / (m * 1e-3) / 1e9, "unit": "Gpx/s", "ms_per_launch": m,
                              "achieved_gbs_per_gpu": salg / (m * 1e-3) / 1e9, "frac": salg / (m * 1e-3) / 1e9 / peak})
        for label, fl, par in (("C2 frames, default bicubic parameter (A=0)", SWS.BICUBIC, None),
                               ("C2 frames, SWS_BILINEAR (what the reference really runs for every flag, R-A arithmetic)", SWS.BILINEAR, None),
                               ("C2 frames, exact-integer form of the headline kernel (SWS.INT_CHAIN, same bytes)", SWS.BICUBIC | SWS.INT_CHAIN, (0.75,)),
                               ("C2 frames, tensor-pipe form of the headline kernel (SWS.MMA_CHAIN: horizontal pass on IMMA u8 x s8, same bytes)", SWS.BICUBIC | SWS.MMA_CHAIN, (0.75,))):
            c2 = SwsContext(sw, sh, sfmt, dw, dh, dfmt, fl | SWS.HWACCEL_CUDA, par)
            sec(label, lambda c2=c2: c2.scale(src, dst), px, alg_bytes_launch)
        # other ratios than 2:1 (the any-ratio streaming kernel, scale_stream.cuh), 32 frames per GPU
        for label, (aw, ah, bw, bh), afmt, bfmt, fl, par, obpp in (
                ("1080p NV12 -> 720p RGB24, bicubic R-B param0=0.75 (any-ratio streaming kernel)", (1920, 1080, 1280, 720), FMT.NV12, FMT.RGB24, SWS.BICUBIC, (0.75,), 3.0),
                ("4K NV12 -> 720p RGB24, bicubic R-B param0=0.75 (any-ratio streaming kernel)", (3840, 2160, 1280, 720), FMT.NV12, FMT.RGB24, SWS.BICUBIC, (0.75,), 3.0),
                ("1080p NV12 -> 4K RGB24, bicubic R-B param0=0.75 (any-ratio streaming kernel)", (1920, 1080, 3840, 2160), FMT.NV12, FMT.RGB24, SWS.BICUBIC, (0.75,), 3.0),
                ("4K NV12 -> 1080p NV12, bicubic R-B default parameter (exact-2:1 plane kernel: the scale_cuda filter's path)", (3840, 2160, 1920, 1080), FMT.NV12, FMT.NV12, SWS.BICUBIC, None, 1.5)):
            nb_ = 16 if bw > aw else 32
            sa = FrameBatch(afmt, aw, ah, nb_, device=dev); sa.buf.random_(0, 256)
            sb = FrameBatch(bfmt, bw, bh, nb_, device=dev)
            cs_ = SwsContext(aw, ah, afmt, bw, bh, bfmt, fl | SWS.HWACCEL_CUDA, par)
            sec(label, lambda cs_=cs_, sa=sa, sb=sb: cs_.scale(sa, sb), nb_ * aw * ah, nb_ * (aw * ah * 1.5 + bw * bh * obpp), steps=5)
            del sa, sb, cs_
        torch.cuda.empty_cache()
        f4, p4, a4, d4 = c4_workload(dev, max(1, 256 // max(world, 8)))      # BASELINE configs[3]: batch 256 over 8 GPUs
        sec(d4, f4, p4, a4, steps=5)
        del f4
        torch.cuda.empty_cache()
        f5, p5, a5, d5 = c5_workload(dev, 8)                                  # BASELINE configs[4]
        sec(d5, f5, p5, a5, steps=5)
        del f5
        torch.cuda.empty_cache()

    # ---- e2e (1): host buffers in, host buffers out, through the public call (PCIe both ways) --------------
    e2e = None
    e2e_nvlink = None
    if ctx is not None:
        Be = min(B, 16)
        hs = FrameBatch(sfmt, sw, sh, Be, pinned=True); hs.buf.copy_(src.buf[:Be * src.frame_bytes].cpu())
        hd = FrameBatch(dfmt, dw, dh, Be, pinned=True)
        for _ in range(2):
            ctx.scale_host(hs, hd)
        T.barrier()
        t0 = time.perf_counter()
        esteps = max(3, args.steps // 4)
        for _ in range(esteps):
            ctx.scale_host(hs, hd)          # synchronises its streams before returning
        torch.cuda.synchronize()
        et = T.max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": world * Be * sw * sh * esteps / et / 1e9, "unit": "Gpx/s",
               "h2d_bytes_per_step": Be * sum(p[1] * p[2] for p in hs.planes), "d2h_bytes_per_step": Be * sum(p[1] * p[2] for p in hd.planes),
               "frames_per_step": Be, "steps": esteps, "path": "pinned host -> PCIe -> kernels -> PCIe -> pinned host (gmatb_sws_scale_host)",
               "cores_bound": cores}
        # ---- e2e (2), N > 1: the batch lives on GPU 0 and is scattered / gathered over NVLink (SURVEY 8e report (2)) ----
        if world > 1:
            Bn = 16                                                           # frames per rank
            total = world * Bn
            full = None
            if rank == 0:
                full = torch.empty(total * src.frame_bytes, dtype=torch.uint8, device=dev)
                for i in range(total):
                    full[i * src.frame_bytes:(i + 1) * src.frame_bytes] = src.buf[(i % B) * src.frame_bytes:((i % B) + 1) * src.frame_bytes]

            def compute(tin, tout, n):
                ctx.scale(FrameBatch(sfmt, sw, sh, n, buffer=tin), FrameBatch(dfmt, dw, dh, n, buffer=tout))

            def nv_step():
                return pipelined_scatter_compute_gather(full, src.frame_bytes, dst.frame_bytes, total, compute, chunk=8, root=0, device=dev)
            nv_step(); nv_step()
            T.barrier()
            t0 = time.perf_counter()
            nsteps = 5
            for _ in range(nsteps):
                nv_step()
            torch.cuda.synchronize()
            nt = T.max_over_ranks(time.perf_counter() - t0)
            e2e_nvlink = {"value": total * sw * sh * nsteps / nt / 1e9, "unit": "Gpx/s", "frames_per_step": total, "steps": nsteps,
                          "scatter_bytes_per_step": (total - Bn) * src.frame_bytes, "gather_bytes_per_step": (total - Bn) * dst.frame_bytes,
                          "path": "batch resident on GPU 0 -> NCCL send/recv over NVLink in chunks of 8 frames, overlapped with the kernels -> "
                                  "results gathered on GPU 0 (gmat_b200.dist.pipelined_scatter_compute_gather)"}
    if world > 1:
        dist.barrier()                      # every rank reaches this point; rank 0 alone reports
        dist.destroy_process_group()
    if rank != 0:
        return 0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp)).get(args.workload)
        if tj:      # ncu dram bytes per frame of the dominant kernel, scaled to this launch's frame count
            traffic = tj["per_frame"] * B
    line = {"metric": METRIC, "value": value, "unit": "Gpx/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u16 in/out, f32 arithmetic" if args.workload == "c3" else "u8 in/out, f32 arithmetic", "data": "synthetic",
            "config": shared_config(args.workload, B) if ctx is not None else {"workload": wl_desc},
            "run": {"launches_per_step": R, "ms_per_launch": ms_launch, "timed_region_s": ms * 1e-3,
                    "parallelism": f"frame-sharded dp{world}", "kernel_path": kernel_path,
                    "output_gpx_s": value * (dw * dh) / (sw * sh) if ctx is not None else None},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes_launch,
                         "frac_of_8000_nominal": achieved / 8000.0},
            "gpu_launches": int(launches), "clocks": sampler.summary()}
    if e2e:
        line["e2e"] = e2e
    if e2e_nvlink:
        line["e2e_nvlink"] = e2e_nvlink
    if secondary:
        line["secondary"] = secondary
    if world == 1 and not args.no_cpu and ctx is not None:
        cb = cpu_reference(args.workload, budget_s=10.0)
        if cb is None:
            cb = {"value": None, "unit": "Gpx/s", "cores": 0, "kind": "port", "sample": "oracle/_ref not built", "cpu": cpu_model()}
        line["cpu_baseline"] = {k: cb.get(k) for k in ("value", "unit", "cores", "kind", "sample", "cpu", "sliced")}
        if args.workload == "c2":
            line["reference_gpu"] = reference_gpu_baseline()
    print(json.dumps(line))
    return 0


def shared_config(wl, batch):
    """the `config` object both arms print (the driver compares them for equality); run-specific details go to `run`"""
    from gmat_b200.image import plane_layout
    from gmat_b200 import FMT
    sname, sw, sh, dname, dw, dh, flag, param0 = WORKLOADS[wl]
    sb = plane_layout(getattr(FMT, sname), sw, sh)[1]; db = plane_layout(getattr(FMT, dname), dw, dh)[1]
    return {"workload": workload_desc(wl), "batch_per_gpu": batch,
            "frames_resident": f"inputs larger than L2 ({batch * sb / 1e6:.0f} MB read + {batch * db / 1e6:.0f} MB written per launch per GPU)"}


def workload_desc(wl):
    sname, sw, sh, dname, dw, dh, flag, param0 = WORKLOADS[wl]
    return (f"{wl.upper()}: {sw}x{sh} {sname} -> {dw}x{dh} {dname}, "
            f"{flag.lower()} R-B" + (f" param0={param0}" if param0 is not None else " default param"))


if __name__ == "__main__":
    sys.exit(main())
