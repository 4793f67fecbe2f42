#!/usr/bin/env python3
"""Write the three headers FFmpeg's `configure` would generate (config.h,
config_components.h, libavutil/avconfig.h, libavutil/ffversion.h) for a
C-only, x86-asm-free, swscale+avutil-only build of the reference tree.

BUILD SCAFFOLDING for this sandbox (the shim of gmat_b200/csrc/Makefile and the reference builds of oracle/refbuild);
inside a real ffmpeg-gpu tree the headers come from its own configure.  This is *our* recipe (the reference's own build
system is never run): every HAVE_/CONFIG_/ARCH_ token that the reference's
libavutil / libswscale / compat sources mention is defined to 0, then a short
whitelist describing this container (glibc, pthreads, gcc) is set to 1.

usage: gen_ffmpeg_config.py <reference ffmpeg-gpu dir> <output include dir> [avfilter]

With `avfilter` the headers describe the AVFilter harness build instead (oracle/refbuild `avf`): libavfilter's
tokens are scanned too and CONFIG_AVFILTER / CONFIG_CUDA / CONFIG_FFNVCODEC are on (hwcontext_cuda.c is compiled
against oracle/ffnv_shim), everything else as above.
"""
import os
import re
import sys

ONES = """
ARCH_GENERIC
HAVE_THREADS HAVE_PTHREADS HAVE_PTHREAD_CANCEL
HAVE_ATOMICS_NATIVE HAVE_ATOMIC_COMPARE_EXCHANGE HAVE_SYNC_VAL_COMPARE_AND_SWAP
HAVE_FAST_64BIT HAVE_FAST_CLZ HAVE_FAST_CMOV HAVE_FAST_UNALIGNED HAVE_LOCAL_ALIGNED
HAVE_SIMD_ALIGN_16 HAVE_SIMD_ALIGN_32 HAVE_SIMD_ALIGN_64
HAVE_ATTRIBUTE_MAY_ALIAS HAVE_ATTRIBUTE_PACKED HAVE_PRAGMA_DEPRECATED HAVE_INLINE_ASM_LABELS
HAVE_ATANF HAVE_ATAN2F HAVE_CBRT HAVE_CBRTF HAVE_COPYSIGN HAVE_COSF HAVE_ERF HAVE_EXP2 HAVE_EXP2F
HAVE_EXPF HAVE_HYPOT HAVE_ISFINITE HAVE_ISINF HAVE_ISNAN HAVE_LDEXPF HAVE_LLRINT HAVE_LLRINTF
HAVE_LOG2 HAVE_LOG2F HAVE_LOG10F HAVE_LRINT HAVE_LRINTF HAVE_POWF HAVE_RINT HAVE_ROUND HAVE_ROUNDF
HAVE_SINF HAVE_TRUNC HAVE_TRUNCF
HAVE_UNISTD_H HAVE_SYS_TIME_H HAVE_SYS_RESOURCE_H HAVE_SYS_PARAM_H HAVE_DIRENT_H HAVE_LINUX_PERF_EVENT_H
HAVE_GETTIMEOFDAY HAVE_CLOCK_GETTIME HAVE_NANOSLEEP HAVE_USLEEP HAVE_SCHED_GETAFFINITY HAVE_SYSCONF
HAVE_POSIX_MEMALIGN HAVE_MEMALIGN HAVE_ALIGNED_MALLOC_NOT HAVE_MMAP HAVE_MKSTEMP HAVE_ISATTY HAVE_STRERROR_R
HAVE_GMTIME_R HAVE_LOCALTIME_R HAVE_FCNTL HAVE_LSTAT HAVE_ACCESS HAVE_GETRUSAGE HAVE_GETENV
HAVE_MALLOC_H HAVE_STDATOMIC HAVE_ARPA_INET_H HAVE_STRUCT_RUSAGE_RU_MAXRSS HAVE_RDTSC_NOT
HAVE_SECTION_DATA_REL_RO HAVE_VALGRIND_VALGRIND_H_NOT HAVE_SYMVER_NOT
CONFIG_SWSCALE CONFIG_AVUTIL CONFIG_SHARED CONFIG_PIC CONFIG_SAFE_BITSTREAM_READER
CONFIG_FAST_UNALIGNED CONFIG_CVCUDA CONFIG_CUDART CONFIG_RUNTIME_CPUDETECT CONFIG_SWSCALE_ALPHA
""".split()

STRINGS = {
    "FFMPEG_CONFIGURATION": '"gmat-b200 oracle recipe: C-only, no x86 asm, swscale+avutil"',
    "FFMPEG_LICENSE": '"LGPL version 2.1 or later"',
    "CONFIG_THIS_YEAR": "2023",
    "FFMPEG_DATADIR": '"/usr/local/share/ffmpeg"',
    "AVCONV_DATADIR": '"/usr/local/share/ffmpeg"',
    "CC_IDENT": '"gcc"',
    "OS_NAME": "linux",
    "av_restrict": "restrict",
    "EXTERN_PREFIX": '""',
    "EXTERN_ASM": "",
    "BUILDSUF": '""',
    "SLIBSUF": '".so"',
    "HAVE_MMX2": "HAVE_MMXEXT",
    "SWS_MAX_FILTER_SIZE": "256",
}

TOKEN = re.compile(r"\b((?:HAVE|CONFIG|ARCH)_[A-Z0-9_]+)\b")


def scan(root, subdirs):
    toks = set()
    for sub in subdirs:
        for dirpath, _, files in os.walk(os.path.join(root, sub)):
            for f in files:
                if f.endswith((".c", ".h")):
                    try:
                        with open(os.path.join(dirpath, f), errors="ignore") as fh:
                            toks.update(TOKEN.findall(fh.read()))
                    except OSError:
                        pass
    return toks


# cpu-extension macros are built by token pasting (libavutil/cpu_internal.h:27,
# HAVE_ ## ext ## suffix), so they never appear literally in the sources
EXTS = """AESNI AMD3DNOW AMD3DNOWEXT AVX AVX2 AVX512 AVX512ICL FMA3 FMA4 MMX MMXEXT SSE SSE2 SSE3
SSE4 SSE42 SSSE3 XOP CPUNOP I686 ARMV5TE ARMV6 ARMV6T2 ARMV8 NEON VFP VFPV3 SETEND ALTIVEC DCBZL
LDBRX POWER8 PPC4XX VSX MIPSFPU MIPS32R2 MIPS32R5 MIPS64R2 MIPS32R6 MIPS64R6 MIPSDSP MIPSDSPR2 MSA
LOONGSON2 LOONGSON3 MMI LSX LASX RVV""".split()


def main():
    ref, out = sys.argv[1], sys.argv[2]
    avf = len(sys.argv) > 3 and sys.argv[3] == "avfilter"
    toks = scan(ref, ["libavutil", "libswscale", "compat"] + (["libavfilter"] if avf else []))
    for e in EXTS:
        for suf in ("", "_EXTERNAL", "_INLINE"):
            toks.add(f"HAVE_{e}{suf}")
    # tokens that are macro *names* used with defined()/ifdef semantics in the
    # sources must stay undefined when off
    skip = {"HAVE_AV_CONFIG_H", "CONFIG_H", "HAVE_MMX2"}
    ones = {t for t in ONES if not t.endswith("_NOT")}
    if avf:
        ones |= {"CONFIG_AVFILTER", "CONFIG_CUDA", "CONFIG_FFNVCODEC"}
        ones -= {"CONFIG_CVCUDA"}          # the nvcv filters / libgpuscale glue are what we replace; not compiled here
    os.makedirs(os.path.join(out, "libavutil"), exist_ok=True)
    with open(os.path.join(out, "config.h"), "w") as f:
        f.write("/* generated by tools/gen_ffmpeg_config.py -- not by the reference's configure */\n")
        f.write("#ifndef FFMPEG_CONFIG_H\n#define FFMPEG_CONFIG_H\n")
        for k, v in STRINGS.items():
            f.write(f"#define {k} {v}\n")
        for t in sorted(toks | ones):
            if t in skip or t in STRINGS:
                continue
            f.write(f"#define {t} {1 if t in ones else 0}\n")
        f.write("#endif\n")
    with open(os.path.join(out, "config_components.h"), "w") as f:
        f.write("/* all optional components off */\n#ifndef FFMPEG_CONFIG_COMPONENTS_H\n#define FFMPEG_CONFIG_COMPONENTS_H\n#endif\n")
    with open(os.path.join(out, "libavutil", "avconfig.h"), "w") as f:
        f.write("#ifndef AVUTIL_AVCONFIG_H\n#define AVUTIL_AVCONFIG_H\n"
                "#define AV_HAVE_BIGENDIAN 0\n#define AV_HAVE_FAST_UNALIGNED 1\n#endif\n")
    with open(os.path.join(out, "libavutil", "ffversion.h"), "w") as f:
        f.write("#ifndef AVUTIL_FFVERSION_H\n#define AVUTIL_FFVERSION_H\n"
                '#define FFMPEG_VERSION "gmat-ref-5.1.2-oracle"\n#endif\n')


if __name__ == "__main__":
    main()
