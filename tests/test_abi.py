"""The drop-in boundary without a GPU: the libraries load and export every symbol the headers declare."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NINE = ["ff_sws_init_swscale_cuda", "ff_swscale_cuda", "ff_sws_free_swscale_cuda", "ff_yuv2rgb_init_tables_cuda",
        "yuv2rgb_cuda", "rgb2yuv_cuda", "yuv2yuv_cuda", "rgb24tobgr24_cuda", "rgb2rgb_init_cuda"]


def declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b((?:gmatb_|ff_|yuv2|rgb2)\w+)\s*\(", txt)))


def exported(so):
    out = subprocess.check_output(["nm", "-D", "--defined-only", so]).decode()
    return {l.split()[-1] for l in out.splitlines() if " T " in l}


def test_library_exports_everything_the_header_declares():
    import gmat_b200
    L = gmat_b200.lib()          # raises if missing; binds every prototype
    syms = exported(gmat_b200.lib_path())
    for name in declared("gmat_b200.h"):
        assert name in syms, name
    assert L.gmatb_version() == 0x000100


def test_nvcodec_header_and_cxx_dropin_symbols():
    """include/gmat_b200_nvcodec.h: the C ABI is in libgmat_b200.so, the reference's own C++ signatures
    (metrans/include/NvCodec/NvCommon.h:232-255, Itanium-mangled) in libgmat_b200_nvcodec.so"""
    import gmat_b200
    syms = exported(gmat_b200.lib_path())
    for name in declared("gmat_b200_nvcodec.h"):
        assert name in syms, name
    cxx = exported(os.path.join(ROOT, "gmat_b200", "libgmat_b200_nvcodec.so"))
    conv = ["Nv12ToBgra32", "Nv12ToRgba32", "Nv12ToBgra64", "P016ToBgra32", "P016ToBgra64", "Nv12ToBgrPlanar", "Nv12ToRgbPlanar",
            "P016ToBgrPlanar", "Bgra64ToP016"]
    for n in conv:
        assert f"_Z{len(n)}{n}PhiS_iiiiP11CUstream_st" in cxx, n
    for n in ("Nv12ToBgrFloatPlanar", "Nv12ToRgbFloatPlanar", "P016ToBgrFloatPlanar"):
        assert f"_Z{len(n)}{n}PhiPfiiiiP11CUstream_st" in cxx, n
    assert "_Z17ScaleNv12_BicubicPhiiiS_iii" in cxx
    ref = os.path.join(ROOT, "oracle", "_ref", "libref_nvcodec.so")
    if os.path.exists(ref):        # the reference's own objects export the same names (minus the texture-filter scalers)
        r = exported(ref)
        assert {s for s in cxx if s.startswith("_Z")} <= r


def test_avfilter_harness_registers_the_six_filters():
    """oracle/_ref/libref_avfilter.so = the reference's libavfilter + libavutil with our filter objects in filter_list.c"""
    so = os.path.join(ROOT, "oracle", "_ref", "libref_avfilter.so")
    if not os.path.exists(so):
        pytest.skip("oracle/refbuild avf not built")
    out = subprocess.check_output(["nm", "-D", "--defined-only", so]).decode()
    for n in ("crop", "rotate", "flip", "smooth", "scale", "format"):
        assert f"ff_vf_{n}_cuda" in out, n
    for n in ("avfilter_graph_config", "av_buffersrc_add_frame", "av_hwframe_transfer_data", "avf_run"):
        assert n in out, n


def test_nine_libswscale_symbols_are_provided():
    import gmat_b200
    a = exported(gmat_b200.lib_path())
    shim = os.path.join(ROOT, "gmat_b200", "libgmat_b200_sws.so")
    assert os.path.exists(shim), "libgmat_b200_sws.so not built (needs the reference headers at build time)"
    b = exported(shim)
    for s in NINE:
        assert s in (a | b), s
    for s in declared("gmat_b200_sws.h"):
        assert s in (a | b), s


def test_reference_libswscale_links_against_us():
    """The reference's own CPU libswscale.so (built by oracle/refbuild) leaves exactly the nine
    symbols undefined; loading it after our two libraries must resolve all of them (RTLD_NOW)."""
    ref = os.path.join(ROOT, "oracle", "_ref", "libref_swscale_cpu.so")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref not built (make -C oracle ref)")
    und = subprocess.check_output(["nm", "-D", "--undefined-only", ref]).decode()
    und = sorted(l.split()[-1] for l in und.splitlines() if "cuda" in l)
    assert und == sorted(NINE)
    import gmat_b200
    gmat_b200.lib()
    C.CDLL(os.path.join(ROOT, "gmat_b200", "libgmat_b200_sws.so"), mode=C.RTLD_GLOBAL)
    L = C.CDLL(ref, mode=C.RTLD_GLOBAL)           # ctypes always adds RTLD_NOW
    assert L.sws_getContext and L.sws_scale and L.sws_setCudaStream


def test_no_cpu_fallback_without_device():
    """On a box without a GPU, device work must fail loudly (never silently run elsewhere)."""
    import gmat_b200 as g
    L = g.lib()
    if L.gmatb_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(g.GmatbError):
        g.SwsContext(64, 48, g.FMT.NV12, 32, 24, g.FMT.RGB24)      # needs device tables -> NULL context
    c = g.SwsContext(64, 48, g.FMT.NV12, 64, 48, g.FMT.RGB24)      # unscaled context is host-only state
    src = g.FrameBatch(g.FMT.NV12, 64, 48, 1); dst = g.FrameBatch(g.FMT.RGB24, 64, 48, 1)
    with pytest.raises(g.GmatbError):
        c.scale(src, dst)
    assert g.sws_getContext(64, 48, g.FMT.NV12, 32, 24, g.FMT.RGB24, g.SWS.BICUBIC | g.SWS.HWACCEL_CUDA) is None
    with pytest.raises(g.GmatbError):
        g.sws_getContext(64, 48, g.FMT.NV12, 32, 24, g.FMT.RGB24, g.SWS.BICUBIC)   # no CPU scaler here


def test_product_does_not_reference_the_oracle():
    """the oracle is test infrastructure: nothing under gmat_b200/ may import, link or call it"""
    for dp, _, files in os.walk(os.path.join(ROOT, "gmat_b200")):
        if "build" in dp:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "gmat_oracle" not in txt or f.endswith((".cu", ".cuh")) and "oracle/gmat_oracle.c" in txt, (dp, f)
                assert "libgmat_oracle" not in txt and "import orc" not in txt, (dp, f)


def test_avfilter_glue_compiles_against_reference_headers(tmp_path):
    """vf_{crop,rotate,flip,smooth,format,scale}_cuda.c are plain C translation units for ffmpeg-gpu's libavfilter;
    they must compile against the reference's own headers and export `const AVFilter ff_vf_<name>`."""
    ref = "/root/reference/ffmpeg-gpu"
    inc = os.path.join(ROOT, "oracle", "_ref", "include")
    if not (os.path.isdir(ref) and os.path.exists(os.path.join(inc, "config.h"))):
        pytest.skip("reference tree / generated config.h not available here")
    d = os.path.join(ROOT, "gmat_b200", "csrc", "avfilter")
    for name in ("crop", "rotate", "flip", "smooth", "format", "scale"):
        obj = str(tmp_path / f"vf_{name}_cuda.o")
        subprocess.check_call(["gcc", "-c", "-std=c11", "-O1", "-fPIC", "-Wall", "-Werror=implicit-function-declaration",
                               "-DHAVE_AV_CONFIG_H", f"-I{inc}", f"-I{ref}", "-I/usr/local/cuda/include",
                               f"-I{os.path.join(ROOT, 'include')}", os.path.join(d, f"vf_{name}_cuda.c"), "-o", obj])
        syms = subprocess.check_output(["nm", obj]).decode()
        assert f"ff_vf_{name}_cuda" in syms
        if name == "scale":
            assert "U gmatb_sws_create" in syms and "U gmatb_sws_scale" in syms and "U ff_scale_eval_dimensions" in syms
        if name == "format":
            assert "U gmatb_format_nv12_to_rgbpf32" in syms and "U gmatb_format_rgbpf32_to_nv12" in syms
        for fn in ("gmatb_crop", "gmatb_rotate", "gmatb_flip", "gmatb_gaussian"):
            if fn.split("_")[1][:4] in name or (name == "smooth" and fn == "gmatb_gaussian"):
                assert f"U {fn}" in syms
