#!/usr/bin/env python3
"""Golden vectors for the format_cuda kernels.

Runs the REFERENCE's own libavfilter/format_cuda_kernel.cu, compiled unmodified for sm_100a into
oracle/_ref/libref_format_cuda.so (oracle O3, oracle/refbuild/Makefile target o3), on seeded inputs:

    gpurun -- python tests/golden/make_golden_format.py gpurun_out/golden
    cp gpurun_out/golden/reference_format_cuda_golden.npz tests/golden/

tests/test_oracle_format.py checks the CPU restatement (oracle/gmat_oracle.c) against the committed
file without a GPU; tests/test_gpu_format.py checks our kernels against it on the GPU.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gmat_b200 import FMT, FrameBatch  # noqa: E402
from gpu_util import o3_run  # noqa: E402
from test_oracle_format import rand_rgbpf32  # noqa: E402


def main(outdir):
    dev = torch.device("cuda:0")
    out = {}
    for cs in (1, 2, 4, 5, 6, 7, 9):
        src = FrameBatch(FMT.NV12, 64, 48, 1); src.fill_lcg(seed=1000 + cs)
        dst = FrameBatch(FMT.RGBPF32LE, 64, 48, 1, device=dev)
        o3_run("nv12_to_rgbpf32", src.to(dev), dst, cs)
        out[f"nv12_to_rgbpf32_cs{cs}"] = dst.payload()
        f = rand_rgbpf32(64, 48, 1, seed=2000 + cs, wide=(cs in (5, 9)))
        d2 = FrameBatch(FMT.NV12, 64, 48, 1, device=dev)
        o3_run("rgbpf32_to_nv12", f.to(dev), d2, cs)
        out[f"rgbpf32_to_nv12_cs{cs}"] = d2.payload()
    src = FrameBatch(FMT.NV12, 64, 48, 1); src.fill_lcg(seed=77)
    for kind in ("nv12_to_rgbpf32_shift", "nv12_to_bgrpf32_shift"):
        dst = FrameBatch(FMT.RGBPF32LE, 64, 48, 1, device=dev)
        o3_run(kind, src.to(dev), dst, 1, norm=58.395, shift=(123.675, 116.28, 103.53))
        out[kind] = dst.payload()
    os.makedirs(outdir, exist_ok=True)
    np.savez_compressed(os.path.join(outdir, "reference_format_cuda_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
