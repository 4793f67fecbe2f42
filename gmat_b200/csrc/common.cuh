// common.cuh -- device helpers shared by all gmat_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/gmat_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "gmat_b200 kernels are written for sm_100a (packed f32x2 arithmetic)"
#endif

namespace gmatb {

// ---------------------------------------------------------------------------
// Packed fp32 pairs (Blackwell FFMA2/FADD2/FMUL2).  Each lane is an IEEE op, so
// results are bit-identical to the scalar FFMA/FADD/FMUL the reference's
// kernels execute; the packed form only halves the issue slots.
// ---------------------------------------------------------------------------
typedef unsigned long long f2;   // two floats in a 64-bit register pair

__device__ __forceinline__ f2 pk(float lo, float hi) {
    f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ f2 bc(float v) { return pk(v, v); }
__device__ __forceinline__ void upk(f2 v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ void upki(f2 v, int &lo, int &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
    f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
    f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
    f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
// round-toward-zero multiply.  With b = 2^-149 the product of a small float is a
// denormal whose BIT PATTERN is trunc(|a|) (sign in bit 31): float->int
// truncation on the FMA pipe instead of the quarter-rate F2I unit.
__device__ __forceinline__ f2 mul2_rz(f2 a, f2 b) {
    f2 r; asm("mul.rz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
__device__ __forceinline__ float mul_rz(float a, float b) {
    float r; asm("mul.rz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r;
}
__device__ __forceinline__ float mul_sat(float a, float b) {
    float r; asm("mul.rn.sat.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r;
}

#define GMATB_TWO_M149 1.401298464324817e-45f   /* 2^-149, smallest denormal */
#define GMATB_TWO_P149 7.136238463529799e+44f   /* 2^149                     */

// byte k of `w` -> float(2^23 + byte)  (PRMT into 0x4B0000xx)
template <int K> __device__ __forceinline__ float byte_magic(uint32_t w) {
    uint32_t r; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(0x4B000000u), "r"(0x7440u | K));
    return __uint_as_float(r);
}
// 16-bit half k (0/1) of `w` -> float(2^23 + half)
template <int K> __device__ __forceinline__ float half_magic(uint32_t w) {
    uint32_t r; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(0x4B000000u),
                    "r"(K == 0 ? 0x7410u : 0x7432u));
    return __uint_as_float(r);
}
#define GMATB_MAGIC 8388608.0f   /* 2^23 */

// saturating pack: returns sat_u8(b0) | sat_u8(b1)<<8 | (upper & 0xffff)<<16  (SASS I2IP)
__device__ __forceinline__ uint32_t pack2_u8(int b0, int b1, uint32_t upper) {
    uint32_t r; asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(b1), "r"(b0), "r"(upper));
    return r;
}
__device__ __forceinline__ uint32_t pack4_u8(int b0, int b1, int b2, int b3) {
    return pack2_u8(b0, b1, pack2_u8(b2, b3, 0u));
}
// saturating pack of two s32 into u16x2
__device__ __forceinline__ uint32_t pack2_u16(int lo, int hi) {
    uint32_t r; asm("cvt.pack.sat.u16.s32 %0, %1, %2;" : "=r"(r) : "r"(hi), "r"(lo));
    return r;
}

// streaming (read-once / write-once) global accesses: keep L1 clean, evict-first in L2
__device__ __forceinline__ uint4 ldg128(const void *p) { return __ldcs(reinterpret_cast<const uint4 *>(p)); }
__device__ __forceinline__ uint2 ldg64(const void *p)  { return __ldcs(reinterpret_cast<const uint2 *>(p)); }
__device__ __forceinline__ uint32_t ldg32(const void *p) { return __ldcs(reinterpret_cast<const uint32_t *>(p)); }
__device__ __forceinline__ void stg128(void *p, uint4 v) { __stcs(reinterpret_cast<uint4 *>(p), v); }
__device__ __forceinline__ void stg64(void *p, uint2 v)  { __stcs(reinterpret_cast<uint2 *>(p), v); }
__device__ __forceinline__ void stg32(void *p, uint32_t v) { __stcs(reinterpret_cast<uint32_t *>(p), v); }

// ---------------------------------------------------------------------------
// Kernel-argument descriptors (passed by value -> constant bank, no global
// __constant__ state: two contexts with different matrices cannot race, unlike
// the reference's process-global matYuv2Rgb, yuv2rgb_cuda.cu:16-17,830).
// ---------------------------------------------------------------------------
struct Plane {
    uint8_t  *p;
    int       pitch;
    long long bstride;   // batch stride
};
struct Img {
    Plane pl[4];
    int w, h;
};
struct Mat9 { float m[9]; };

}  // namespace gmatb

// host-side helpers (host.cu)
namespace gmatb {
int  set_cuda_error(cudaError_t e);
void count_launch(int n = 1);
void gmatb_log(const char *msg);
bool to_img(const GmatbImage *g, Img *out, int nplanes);
int  fmt_planes(int fmt);
}
