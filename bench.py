#!/usr/bin/env python3
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c2_default|c3|csc]

A "step" is one pass of the hot path over one batch of synthetic frames.  At N=1 the workload
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
    """nvidia-smi clocks / throttle reasons DURING the timed region"""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        try:
            p = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                 stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        self.proc = p
        for line in p.stdout:
            if self.stop_flag:
                break
            self.samples.append([x.strip() for x in line.split(",")])
        p.kill()

    def stop(self):
        self.stop_flag = True
        if getattr(self, "proc", None):
            self.proc.kill()

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 2 + i and s[2 + i].lower().startswith("active") for s in self.samples)]
        pw = [float(s[6]) for s in self.samples if len(s) > 6 and s[6].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "power_w_max": max(pw) if pw else None, "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own CPU libswscale (oracle/_ref), all host threads
# --------------------------------------------------------------------------------------------
def cpu_reference(workload, budget_s=12.0, threads=None):
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
    # CPU flags: C1/C2 SWS_BICUBIC, C3 SWS_LANCZOS with param0 = 4
    flags = getattr(SWS, flag)
    pp = (C.c_double * 2)(4.0 if flag == "LANCZOS" else (param0 if param0 is not None else SWS.PARAM_DEFAULT), SWS.PARAM_DEFAULT)
    nthreads = threads or os.cpu_count() or 1
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
    return {"value": frames * sw * sh / dt / 1e9, "unit": "Gpx/s", "cores": nthreads, "kind": "reference",
            "sample": f"{frames} frames of {sw}x{sh} {sname}->{dw}x{dh} {dname} through the reference's libswscale "
                      f"(C only, no x86 asm; oracle/_ref/libref_swscale_cpu.so), {nthreads} frame-parallel threads, {dt:.1f} s",
            "seconds": dt, "frames": frames}


def run_extra(args):
    """BASELINE configs[3] (C4: rotate 30 deg -> gaussian 5x5 -> scale to 1080p on 4K rgb24, every filter
    materialising its frame like a filtergraph) and configs[4] (C5: equal thirds of 1080p / 4K / 8K NV12 ->
    RGB24 at half size), frames sharded by batch index over the ranks.  Secondary benchmarks: same JSON shape."""
    import torch
    import torch.distributed as dist
    import gmat_b200 as g
    from gmat_b200 import BORDER, FMT, SWS, FrameBatch, SwsContext
    from gmat_b200.dist import init
    rank, world, local = init()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    peak, peak_src = measured_peak()
    if args.workload == "c4":
        B = max(1, 256 // 8 if world == 1 else 256 // world)          # batch 256 over 8 GPUs = 32 per GPU
        a = FrameBatch(FMT.RGB24, 3840, 2160, B, device=dev); a.buf.random_(0, 256)
        b = FrameBatch(FMT.RGB24, 3840, 2160, B, device=dev)
        c = FrameBatch(FMT.RGB24, 3840, 2160, B, device=dev)
        d = FrameBatch(FMT.RGB24, 1920, 1080, B, device=dev)
        sc = SwsContext(3840, 2160, FMT.RGB24, 1920, 1080, FMT.RGB24, SWS.BICUBIC | SWS.HWACCEL_CUDA)

        def step():
            g.rotate(a, b, 30.0, -282.7688, 1104.6926, "linear")
            g.gaussian(b, c, 5, 5, 1.1, 1.1, BORDER.REFLECT101)
            sc.scale(c, d)
        px = B * 3840 * 2160
        alg = B * 130636800
        desc = "C4: 4K rgb24 rotate(30deg, linear) -> gaussian 5x5 sigma 1.1 reflect101 -> bicubic scale to 1080p, each stage materialised"
    else:
        n = max(1, args.batch // 3)
        sizes = ((1920, 1080), (3840, 2160), (7680, 4320))
        items = []
        for (w, h) in sizes:
            s_ = FrameBatch(FMT.NV12, w, h, n if w < 7680 else max(1, n // 2), device=dev); s_.buf.random_(0, 256)
            d_ = FrameBatch(FMT.RGB24, w // 2, h // 2, s_.n, device=dev)
            items.append((SwsContext(w, h, FMT.NV12, w // 2, h // 2, FMT.RGB24, SWS.BICUBIC | SWS.HWACCEL_CUDA, (0.75,)), s_, d_))

        def step():
            for ctx, s_, d_ in items:
                ctx.scale(s_, d_)
        px = sum(s_.n * s_.w * s_.h for _, s_, _ in items)
        alg = int(px * 2.25)
        desc = "C5: mixed 1080p/4K/8K NV12 -> RGB24 at half size, bicubic R-B param0=0.75, " + "+".join(str(s_.n) for _, s_, _ in items) + " frames per GPU"
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = g.lib().gmatb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = g.lib().gmatb_launch_count() - l0
    if world > 1:
        tt = torch.tensor([ms], device=dev); dist.all_reduce(tt, op=dist.ReduceOp.MAX); ms = tt.item()
        dist.barrier(); dist.destroy_process_group()
    if rank != 0:
        return 0
    ms_step = ms / args.steps
    print(json.dumps({"metric": METRIC, "value": world * px / (ms_step * 1e-3) / 1e9, "unit": "Gpx/s", "n_gpus": world, "steps": args.steps,
                      "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                      "dtype": "u8 in/out, f32 arithmetic", "data": "synthetic", "config": {"workload": desc},
                      "roofline": {"bound": "hbm", "achieved": alg / (ms_step * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                   "frac": alg / (ms_step * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src},
                      "gpu_launches": int(launches)}))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + ["c4", "c5"])
    ap.add_argument("--batch", type=int, default=64, help="frames per GPU per step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.workload in ("c4", "c5"):
        return run_extra(args)
    sname, sw, sh, dname, dw, dh, flag, param0 = WORKLOADS[args.workload]
    wl_desc = (f"{args.workload.upper()}: {sw}x{sh} {sname} -> {dw}x{dh} {dname}, "
               f"{flag.lower()} R-B" + (f" param0={param0}" if param0 is not None else " default param"))

    if args.impl == "reference":
        # the reference's own CPU implementation of the path, all host threads; rank 0 only
        if rank != 0:
            return 0
        r = cpu_reference(args.workload, budget_s=max(2.0, 1.5 * args.steps))
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libref_swscale_cpu.so not built"}))
            return 0
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "Gpx/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] / max(args.steps, 1), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": wl_desc, "note": "reference CPU libswscale (its CUDA path needs the closed CV-CUDA library)"},
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": "Gpx/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    import gmat_b200 as g
    from gmat_b200 import FMT, SWS, FrameBatch, SwsContext
    from gmat_b200.dist import init
    rank, world, local = init()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: gmat_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    L = g.lib()
    sfmt, dfmt = getattr(FMT, sname), getattr(FMT, dname)
    B = args.batch
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
    alg_bytes = sum(p[2] * p[3] for p in src.planes) + sum(p[2] * p[3] for p in dst.planes)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        ctx.scale(src, dst)
    barrier()
    sampler = ClockSampler(local); sampler.start(); time.sleep(0.25)
    l0 = L.gmatb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        ctx.scale(src, dst)
    e1.record()
    barrier()
    launches = L.gmatb_launch_count() - l0
    ms = e0.elapsed_time(e1)
    time.sleep(0.15); sampler.stop()
    if world > 1:
        tt = torch.tensor([ms], device=dev); dist.all_reduce(tt, op=dist.ReduceOp.MAX); ms = tt.item()
    ms_step = ms / args.steps
    value = world * B * sw * sh / (ms_step * 1e-3) / 1e9
    peak, peak_src = measured_peak()
    # ---- secondary measurements on the same frames (reported beside the headline) ----------------
    #  * default-parameter bicubic (A = 0: at exactly 2:1 the outer taps vanish)
    #  * SWS_BILINEAR: what the reference really executes for EVERY flag (swscale_cuda.c:305 bug)
    secondary = None
    if args.workload == "c2":
        secondary = []
        for label, fl, par in (("default bicubic parameter (A=0)", SWS.BICUBIC, None),
                               ("SWS_BILINEAR (the reference's actual resize, R-A arithmetic)", SWS.BILINEAR, None)):
            c2 = SwsContext(sw, sh, sfmt, dw, dh, dfmt, fl | SWS.HWACCEL_CUDA, par)
            for _ in range(3):
                c2.scale(src, dst)
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(args.steps):
                c2.scale(src, dst)
            a1.record(); torch.cuda.synchronize()
            m2 = a0.elapsed_time(a1) / args.steps
            secondary.append({"workload": "same frames, " + label, "value": B * sw * sh / (m2 * 1e-3) / 1e9, "unit": "Gpx/s (this rank)",
                              "achieved_gbs": B * alg_bytes / (m2 * 1e-3) / 1e9, "frac": B * alg_bytes / (m2 * 1e-3) / 1e9 / peak})
    # ---- e2e: host buffers in, host buffers out, through the public call --------------------------
    Be = min(B, 16)
    hs = FrameBatch(sfmt, sw, sh, Be, pinned=True); hs.buf.copy_(src.buf[:Be * src.frame_bytes].cpu())
    hd = FrameBatch(dfmt, dw, dh, Be, pinned=True)
    for _ in range(2):
        ctx.scale_host(hs, hd)
    barrier()
    t0 = time.perf_counter()
    esteps = max(3, args.steps // 4)
    for _ in range(esteps):
        ctx.scale_host(hs, hd)          # synchronises its stream before returning
    torch.cuda.synchronize()
    et = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([et], device=dev); dist.all_reduce(tt, op=dist.ReduceOp.MAX); et = tt.item()
    e2e = {"value": world * Be * sw * sh * esteps / et / 1e9, "unit": "Gpx/s",
           "h2d_bytes_per_step": Be * sum(p[1] * p[2] for p in hs.planes), "d2h_bytes_per_step": Be * sum(p[1] * p[2] for p in hd.planes),
           "frames_per_step": Be, "steps": esteps}
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
            "dtype": "u8 in/out, f32 arithmetic" if sname != "P010LE" else "u16 in/out, f32 arithmetic", "data": "synthetic",
            "config": {"workload": wl_desc, "batch_per_gpu": B, "frames_resident": "inputs larger than L2 "
                       f"({B * src.frame_bytes / 1e6:.0f} MB read + {B * dst.frame_bytes / 1e6:.0f} MB written per step per GPU)",
                       "parallelism": f"frame-sharded dp{world}", "kernel_path": {0: "unscaled", 1: "fused_csc_scale2", 2: "generic"}[ctx.path],
                       "output_gpx_s": value * dw * dh / (sw * sh)},
            "roofline": {"bound": "hbm", "achieved": B * alg_bytes / (ms_step * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": B * alg_bytes / (ms_step * 1e-3) / 1e9 / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": B * alg_bytes, "frac_of_8000_nominal": B * alg_bytes / (ms_step * 1e-3) / 1e9 / 8000.0},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": sampler.summary()}
    if secondary:
        line["secondary"] = secondary
    if world == 1 and not args.no_cpu:
        cb = cpu_reference(args.workload, budget_s=10.0)
        if cb is None:
            cb = {"value": None, "unit": "Gpx/s", "cores": 0, "kind": "port", "sample": "oracle/_ref not built"}
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
