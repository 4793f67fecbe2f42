// csc_core.cuh -- the colour-space arithmetic, bit-exact with the reference's
// libgpuscale kernels (libswscale/cuda/yuv2rgb_cuda.cu:72-106, :653-669) as
// nvcc 12.9 compiles them for sm_100a (SASS read from oracle/_ref/libref_gpuscale.so):
//
//   yuv->rgb, per channel c with matrix row (mA, mB, mC):
//       t1 = FMUL(fu, mB); t2 = FMUL(fv, mC)          (once per 2x2 block)
//       x  = FFMA(fy, mA, t1); r = FADD(x, t2)
//       r  = r < 0 ? 0 : min(r, max)  ->  F2I.U32.TRUNC
//   rgb->yuv:  Y = trunc(FADD(FFMA(b, m2, FFMA(r, m0, FMUL(g, m1))), low))
//              U/V the same shape on the integer 2x2 mean, + mid.
//
// Here the same IEEE operations run two pixels at a time on FFMA2/FADD2/FMUL2,
// int->float goes through PRMT magic numbers and float->int truncation through a
// round-toward-zero multiply by 2^-149 (the denormal's bit pattern IS the
// integer), with the clamp folded into the saturating byte pack (I2IP).  Each
// lane performs exactly the reference's roundings, so results are identical.
#pragma once
#include "common.cuh"

namespace gmatb {

enum { L_NV12 = 0, L_I420 = 1, L_RGB3 = 2 };   // source layout: semi-planar yuv / planar yuv / packed 3-byte rgb (fused scaler only)
enum { D_RGB24 = 0, D_BGR24, D_RGBA, D_BGRA, D_RGB48, D_BGR48, D_RGBA64, D_BGRA64, D_COUNT };

__host__ __device__ constexpr int dst_bpp(int d) {
    return d <= D_BGR24 ? 3 : d <= D_BGRA ? 4 : d <= D_BGR48 ? 6 : 8;
}
__host__ __device__ constexpr bool dst_is16(int d) { return d >= D_RGB48; }
__host__ __device__ constexpr bool dst_swap(int d) { return d == D_BGR24 || d == D_BGRA || d == D_BGR48 || d == D_BGRA64; }
__host__ __device__ constexpr bool dst_alpha(int d) { return d == D_RGBA || d == D_BGRA || d == D_RGBA64 || d == D_BGRA64; }

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) {
    uint32_t r; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(s)); return r;
}

// Products of one chroma sample with the matrix (shared by the 2x2 luma block).
struct ChromaTerms {
    float t1r, t1g, t1b;   // m[1]*fu, m[4]*fu, m[7]*fu
    float t2r, t2g, t2b;   // m[2]*fv, m[5]*fv, m[8]*fv   (FADD form only)
    float fv;              // (FMA form only)
};

// SPARSE: m[1] == 0 and m[8] == 0 (true for every matrix the reference builds,
// yuv2rgb_cuda.cu:820-824).  FFMA(fy,m0,+-0) == FMUL(fy,m0) and x + (+-0) == x,
// so skipping the zero terms cannot change a result bit (up to the sign of zero,
// which the truncation discards).
//
// The reference's kernels come in TWO roundings of the same expression, because nvcc
// contracts them differently (SASS of oracle/_ref/libref_gpuscale.so):
//   FADD form  r = FADD(FFMA(fy,mA,FMUL(fu,mB)), FMUL(fv,mC))   yuv2rgb_odd_kernel / yuv02rgb_odd_kernel:
//                                                                NV12/I420 -> packed 8/16-bit rgb
//   FMA  form  r = FFMA(fv,mC, FFMA(fy,mA,FMUL(fu,mB)))          yuv2rgb_kernel / yuv2rgb_planar_kernel:
//                                                                P010/P016 sources and planar-float output
// FMAFORM selects which one a kernel reproduces.
template <bool SPARSE, bool FMAFORM = false>
__device__ __forceinline__ ChromaTerms chroma_terms(float fu, float fv, const Mat9 &M) {
    ChromaTerms t;
    float a, b;
    t.fv = fv;
    if (FMAFORM) {
        t.t1g = __fmul_rn(fu, M.m[4]); t.t1b = __fmul_rn(fu, M.m[7]);
        t.t1r = SPARSE ? 0.f : __fmul_rn(fu, M.m[1]);
        t.t2r = t.t2g = t.t2b = 0.f;
        return t;
    }
    upk(mul2(pk(fu, fv), pk(M.m[4], M.m[5])), a, b); t.t1g = a; t.t2g = b;
    upk(mul2(pk(fu, fv), pk(M.m[7], M.m[2])), a, b); t.t1b = a; t.t2r = b;
    if (SPARSE) { t.t1r = 0.f; t.t2b = 0.f; }
    else {
        upk(mul2(pk(fu, fv), pk(M.m[1], M.m[8])), a, b); t.t1r = a; t.t2b = b;
    }
    return t;
}

// Two horizontally adjacent pixels (same chroma): fy2 = (y0-low, y1-low).
// Outputs are the UNCLAMPED float results of the reference chain.
template <bool SPARSE, bool FMAFORM = false>
__device__ __forceinline__ void csc_pair_f(f2 fy2, const ChromaTerms &t, const Mat9 &M,
                                           f2 &r, f2 &g, f2 &b) {
    // NB: ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even though both
    // carry .rn (it does not for the scalar forms), and it also sees through an fma
    // with a literal zero addend.  A run-time zero is opaque to it: in the SPARSE case
    // the addend is m[1] itself (== 0.0f, checked on the host), so this is
    // FFMA2(fy, m0, 0) = RN(fy*m0), exactly the reference's FFMA(fy, m0, +-0).
    f2 xr = fma2(fy2, bc(M.m[0]), bc(SPARSE ? M.m[1] : t.t1r));
    f2 xg = fma2(fy2, bc(M.m[3]), bc(t.t1g));
    f2 xb = fma2(fy2, bc(M.m[6]), bc(t.t1b));
    if (FMAFORM) {
        r = fma2(bc(t.fv), bc(M.m[2]), xr);
        g = fma2(bc(t.fv), bc(M.m[5]), xg);
        b = SPARSE ? xb : fma2(bc(t.fv), bc(M.m[8]), xb);     // FFMA(fv, 0, x) == x
    } else {
        r = add2(xr, bc(t.t2r));
        g = add2(xg, bc(t.t2g));
        b = SPARSE ? xb : add2(xb, bc(t.t2b));
    }
}

// ... and truncated to integers (sign-magnitude for negatives: any negative value
// is a large negative s32, which the saturating packs clamp to 0).
template <bool SPARSE, bool FMAFORM = false>
__device__ __forceinline__ void csc_pair_i(f2 fy2, const ChromaTerms &t, const Mat9 &M,
                                           int &r0, int &r1, int &g0, int &g1, int &b0, int &b1) {
    f2 r, g, b;
    csc_pair_f<SPARSE, FMAFORM>(fy2, t, M, r, g, b);
    const f2 z = bc(GMATB_TWO_M149);
    upki(mul2_rz(r, z), r0, r1);
    upki(mul2_rz(g, z), g0, g1);
    upki(mul2_rz(b, z), b0, b1);
}

__device__ __forceinline__ int clamp_i(int v, int hi) { return min(max(v, 0), hi); }

// rgb -> one luma / chroma sample; inputs are exact small integers as floats.
__device__ __forceinline__ float rgb2y_f(float r, float g, float b, const Mat9 &M, int row, float off) {
    float t = __fmul_rn(g, M.m[row * 3 + 1]);
    t = __fmaf_rn(r, M.m[row * 3 + 0], t);
    t = __fmaf_rn(b, M.m[row * 3 + 2], t);
    return __fadd_rn(t, off);
}
__device__ __forceinline__ int trunc_i(float v) { return __float_as_int(mul_rz(v, GMATB_TWO_M149)); }

}  // namespace gmatb
