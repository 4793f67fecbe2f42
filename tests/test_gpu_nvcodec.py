"""-m gpu: SURVEY 8f N3 -- metrans' NvCodec helpers.  Ours (C ABI of include/gmat_b200_nvcodec.h and the C++ drop-in
signatures of libgmat_b200_nvcodec.so) against the reference's Resize.cu / ColorSpace.cu compiled unmodified for
sm_100a (oracle O4, oracle/_ref/libref_nvcodec.so), live on the same GPU: bit-exact for every launcher."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import gmat_b200 as g
from gpu_util import REF, ROOT

pytestmark = pytest.mark.gpu
O4 = os.path.join(REF, "libref_nvcodec.so")
needs = pytest.mark.skipif(not os.path.exists(O4), reason="oracle/_ref/libref_nvcodec.so not built")
vp, ci = C.c_void_p, C.c_int
CONV_SIG = "PhiS_iiiiP11CUstream_st"          # (uint8_t*, int, uint8_t*, int, int, int, int, cudaStream_t)
CONVF_SIG = "PhiPfiiiiP11CUstream_st"         # float* destination
KINDS = [("Nv12ToBgra32", 0, 8, "p4"), ("Nv12ToRgba32", 1, 8, "p4"), ("Nv12ToBgra64", 2, 8, "p8"), ("P016ToBgra32", 3, 16, "p4"),
         ("P016ToBgra64", 4, 16, "p8"), ("Nv12ToBgrPlanar", 5, 8, "pl1"), ("Nv12ToRgbPlanar", 6, 8, "pl1"), ("P016ToBgrPlanar", 7, 16, "pl1"),
         ("Nv12ToBgrFloatPlanar", 8, 8, "pl4"), ("Nv12ToRgbFloatPlanar", 9, 8, "pl4"), ("P016ToBgrFloatPlanar", 10, 16, "pl4")]


def mangled(name, fl=False):
    return f"_Z{len(name)}{name}{CONVF_SIG if fl else CONV_SIG}"


def libs():
    ref = C.CDLL(O4)                                            # RTLD_LOCAL: same C++ names as our drop-in library
    ours = g.lib()
    cxx = C.CDLL(os.path.join(ROOT, "gmat_b200", "libgmat_b200_nvcodec.so"))
    ours.gmatb_nvcodec_convert.argtypes = [ci, vp, ci, vp, ci, ci, ci, ci, vp]
    ours.gmatb_nvcodec_scale_nv12_bicubic.argtypes = [vp, ci, ci, ci, vp, ci, ci, ci, vp]
    return ref, ours, cxx


def yuv_frame(dev, w, h, bits, pitch, seed):
    t = torch.zeros(pitch * (h * 3 // 2) + 64, dtype=torch.uint8, device=dev)
    gen = torch.Generator(device=dev); gen.manual_seed(seed)
    t.random_(0, 256, generator=gen)
    return t


@needs
@pytest.mark.parametrize("name,kind,bits,layout", KINDS)
@pytest.mark.parametrize("w,h", [(1920, 1080), (642, 362), (64, 48), (130, 34)])
@pytest.mark.parametrize("matrix", [1, 6, 9, 2])
def test_colorspace_launchers_vs_reference_live(dev, name, kind, bits, layout, w, h, matrix):
    ref, ours, cxx = libs()
    sb = bits // 8
    spitch = (w * sb + 255) // 256 * 256
    src = yuv_frame(dev, w, h, bits, spitch, w + kind)
    bpp = {"p4": 4, "p8": 8, "pl1": 1, "pl4": 4}[layout]
    planar = layout.startswith("pl")
    dpitch = (w * bpp + 255) // 256 * 256
    nbytes = dpitch * h * (3 if planar else 1)
    outs = []
    fl = layout == "pl4"
    for which in ("ref", "abi", "cxx"):
        d = torch.full((nbytes + 64,), 0x5A, dtype=torch.uint8, device=dev)
        if which == "ref":
            f = getattr(ref, mangled(name, fl)); f.restype = None; f.argtypes = [vp, ci, vp, ci, ci, ci, ci, vp]
            f(src.data_ptr(), spitch, d.data_ptr(), dpitch, w, h, matrix, None)
        elif which == "abi":
            assert ours.gmatb_nvcodec_convert(kind, src.data_ptr(), spitch, d.data_ptr(), dpitch, w, h, matrix, None) == 0
        else:
            f = getattr(cxx, mangled(name, fl)); f.restype = None; f.argtypes = [vp, ci, vp, ci, ci, ci, ci, vp]
            f(src.data_ptr(), spitch, d.data_ptr(), dpitch, w, h, matrix, None)
        torch.cuda.synchronize()
        outs.append(d.cpu().numpy())
    assert np.array_equal(outs[0], outs[1]), f"{name} {w}x{h} m{matrix}: {int((outs[0] != outs[1]).sum())} bytes differ (C ABI)"
    assert np.array_equal(outs[0], outs[2]), f"{name} {w}x{h} m{matrix}: C++ drop-in differs"
    assert (outs[0] != 0x5A).any()


@needs
@pytest.mark.parametrize("w,h", [(1920, 1080), (642, 362), (64, 48)])
@pytest.mark.parametrize("matrix", [1, 6, 9])
def test_bgra64_to_p016_vs_reference_live(dev, w, h, matrix):
    ref, ours, cxx = libs()
    spitch = (w * 8 + 255) // 256 * 256
    dpitch = (w * 2 + 255) // 256 * 256
    src = torch.zeros(spitch * h + 64, dtype=torch.uint8, device=dev); src.random_(0, 256)
    outs = []
    for which in ("ref", "abi"):
        d = torch.full((dpitch * (h * 3 // 2) + 64,), 0x5A, dtype=torch.uint8, device=dev)
        if which == "ref":
            f = getattr(ref, mangled("Bgra64ToP016")); f.restype = None; f.argtypes = [vp, ci, vp, ci, ci, ci, ci, vp]
            f(src.data_ptr(), spitch, d.data_ptr(), dpitch, w, h, matrix, None)
        else:
            assert ours.gmatb_nvcodec_convert(11, src.data_ptr(), spitch, d.data_ptr(), dpitch, w, h, matrix, None) == 0
        torch.cuda.synchronize(); outs.append(d.cpu().numpy())
    assert np.array_equal(outs[0], outs[1]), f"{int((outs[0] != outs[1]).sum())} bytes differ"


@needs
@pytest.mark.parametrize("sw,sh,dw,dh", [(1920, 1080, 1280, 720), (1280, 720, 1920, 1080), (640, 360, 320, 180), (3840, 2160, 1920, 1080),
                                         (64, 48, 100, 70), (642, 362, 322, 182)])
def test_scale_nv12_bicubic_vs_reference_live(dev, sw, sh, dw, dh):
    ref, ours, cxx = libs()
    spitch = (sw + 255) // 256 * 256
    dpitch = (dw + 255) // 256 * 256
    # one extra row after the chroma plane: the reference reads (weight 0) one row past it (Resize.cu:101-123 with fy = H/2 - 2)
    src = torch.zeros(spitch * (sh * 3 // 2 + 2), dtype=torch.uint8, device=dev); src.random_(0, 256)
    name = "_Z17ScaleNv12_BicubicPhiiiS_iii"
    outs = []
    for which in ("ref", "abi", "cxx"):
        d = torch.full((dpitch * (dh * 3 // 2) + 64,), 0x5A, dtype=torch.uint8, device=dev)
        if which == "abi":
            assert ours.gmatb_nvcodec_scale_nv12_bicubic(src.data_ptr(), spitch, sw, sh, d.data_ptr(), dpitch, dw, dh, None) == 0
        else:
            f = getattr(ref if which == "ref" else cxx, name); f.restype = None; f.argtypes = [vp, ci, ci, ci, vp, ci, ci, ci]
            f(src.data_ptr(), spitch, sw, sh, d.data_ptr(), dpitch, dw, dh)
        torch.cuda.synchronize(); outs.append(d.cpu().numpy())
    assert np.array_equal(outs[0], outs[1]), f"bicubic {sw}x{sh}->{dw}x{dh}: {int((outs[0] != outs[1]).sum())} bytes differ"
    assert np.array_equal(outs[0], outs[2])
