// resample_core.cuh -- the pinned resample definition "R-B" (SURVEY 8c): the
// arithmetic of the reference's own in-tree CUDA scaler, libavfilter/vf_scale_cuda.cu
// (Subsample_Bicubic :1040-1074, bicubic_coeffs :970-981, lanczos_coeffs :948-968,
// apply_coeffs :983-992), in the operation order nvcc 12.9 emits for sm_100a (SASS of
// oracle/_ref/ref_scale_cuda.cubin):
//
//   hscale = div.rn((float)src_w, (float)dst_w)
//   xi = FFMA(xo + 0.5f, hscale, -0.5f);  px = floor(xi);  fx = xi - px
//   taps at px-1 .. px+2, coordinates clamped to the image (texture clamp mode),
//   samples p = texel / 255 (u8) or / 65535 (u16), as the texture unit's
//   normalised-float read returns them
//   row:  t = FMUL(w1,p1); t = FFMA(w0,p0,t); t = FFMA(w2,p2,t); t = FFMA(w3,p3,t)
//   the same chain vertically over the 4 row results, then FMUL(t, 255|65535) and
//   F2I.U32.TRUNC (negative -> 0; values above the maximum are not clamped by the
//   reference -> they wrap when stored; we saturate unless GMATB_SWS_PARITY_WRAP).
//
// The chain is separable exactly as written: the 4 horizontal results of a source
// row depend only on (row, xo), so computing them once per source row and reusing
// them for every output row that taps the row performs the reference's operations
// verbatim -- bit-identical output at a fraction of the work.
#pragma once
#include "common.cuh"

namespace gmatb {

enum { RS_BICUBIC = 0, RS_LANCZOS = 1, RS_BILINEAR = 2, RS_NEAREST = 3 };

// bicubic_coeffs (vf_scale_cuda.cu:970-981) in SASS order.  A = 0 when the
// parameter is the default, else -param.
__device__ __forceinline__ float4 bicubic_coeffs_rb(float x, float A) {
    const float x1  = __fadd_rn(x, 1.0f);
    const float omx = __fadd_rn(-x, 1.0f);
    const float Ap2 = __fadd_rn(A, 2.0f), Ap3 = __fadd_rn(A, 3.0f);
    const float A5 = __fmul_rn(A, 5.0f), A4 = __fmul_rn(A, 4.0f);
    float4 r;
    float s = __fmaf_rn(x1, A, -A5);            // A*(x+1) - 5A
    s = __fmul_rn(x1, s);
    s = __fmaf_rn(A, 8.0f, s);                  // ... + 8A
    r.x = __fmaf_rn(x1, s, -A4);                // ...*(x+1) - 4A
    float t = __fmaf_rn(x, Ap2, -Ap3);          // (A+2)x - (A+3)
    t = __fmul_rn(x, t);
    r.y = __fmaf_rn(x, t, 1.0f);
    float u = __fmaf_rn(Ap2, omx, -Ap3);        // (A+2)(1-x) - (A+3)
    u = __fmul_rn(omx, u);
    r.z = __fmaf_rn(omx, u, 1.0f);
    r.w = __fadd_rn(__fadd_rn(__fadd_rn(-r.x, 1.0f), -r.y), -r.z);   // 1 - x - y - z
    return r;
}

// lanczos_coeffs (vf_scale_cuda.cu:948-968): a = 2, fast-math sines, normalised.
__device__ __forceinline__ float lanczos_tap_rb(float t) {
    if (t == 0.0f) return 1.0f;
    const float num = __fmul_rn(__sinf(t), __sinf(__fmul_rn(t, 0.5f)));   // t/2.0f == t*0.5f exactly
    const float den = __fmul_rn(__fmul_rn(t, t), 0.5f);                   // FMUL.D2
    return __fdiv_rn(num, den);
}
__device__ __forceinline__ float4 lanczos_coeffs_rb(float x) {
    const float pi = 3.141592654f;
    float4 r;
    r.x = lanczos_tap_rb(__fmul_rn(pi, __fadd_rn(x, 1.0f)));
    r.y = lanczos_tap_rb(__fmul_rn(pi, x));
    r.z = lanczos_tap_rb(__fmul_rn(pi, __fadd_rn(x, -1.0f)));
    r.w = lanczos_tap_rb(__fmul_rn(pi, __fadd_rn(x, -2.0f)));
    const float sum = __fadd_rn(__fadd_rn(__fadd_rn(r.x, r.y), r.z), r.w);
    r.x = __fdiv_rn(r.x, sum); r.y = __fdiv_rn(r.y, sum); r.z = __fdiv_rn(r.z, sum); r.w = __fdiv_rn(r.w, sum);
    return r;
}

// one axis of the filter bank: coeffs[o] (4 taps) and pos[o] = px - 1
static __global__ void filter_table_kernel(int algo, int src_n, int dst_n, float A, float4 *coeffs, int *pos) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= dst_n) return;
    const float scale = __fdiv_rn((float)src_n, (float)dst_n);
    float4 c; int p;
    if (algo == RS_BICUBIC || algo == RS_LANCZOS) {
        const float xi = __fmaf_rn(__fadd_rn((float)o, 0.5f), scale, -0.5f);
        const float pf = floorf(xi);
        const float f = __fadd_rn(xi, -pf);
        c = algo == RS_BICUBIC ? bicubic_coeffs_rb(f, A) : lanczos_coeffs_rb(f);
        p = (int)pf - 1;
    } else if (algo == RS_BILINEAR) {
        // R-A: half-pixel-centre bilinear, clamp, fp32:
        // sx = (o+.5)*s - .5; taps at floor(sx), floor(sx)+1 with weights (1-f, f)
        const float xi = __fmaf_rn(__fadd_rn((float)o, 0.5f), scale, -0.5f);
        const float pf = floorf(xi);
        const float f = __fadd_rn(xi, -pf);
        c = make_float4(0.0f, __fadd_rn(1.0f, -f), f, 0.0f);
        p = (int)pf - 1;
    } else {   // nearest: Subsample_Nearest (vf_scale_cuda.cu:998-1007): texel floor((o+.5)*s)
        const float xi = __fmul_rn(__fadd_rn((float)o, 0.5f), scale);
        c = make_float4(0.0f, 1.0f, 0.0f, 0.0f);
        p = (int)floorf(xi) - 1;
    }
    coeffs[o] = c;
    pos[o] = p;
}

// Quantise-and-normalise one CSC result: r (float, unclamped) -> the sample the
// reference's resize stage reads back from its u8/u16 intermediate image through the
// texture unit, p = RN(clamp(trunc(r), 0, max) / max).
//   m  = FADD.RZ(r, 2^23)            = 2^23 + trunc(r)    (r >= 0; r < 0 gives m < 2^23)
//   hi = FFMA(m, c1, -2^23*c1)       = j*c1 EXACTLY, j = m - 2^23:
//          8 bit: c1 = 0x010101 * 2^-24, j*0x010101 < 2^24 is representable
//         16 bit: c1 = 2^-16
//   p  = sat(FFMA(hi, c2, hi))       c2 = 2^-24 (8 bit) or 2^-16 + 2^-32 (16 bit)
// because 1/255 = 0x010101*2^-24 * (1 + 2^-24 + 2^-48 + ...) and 1/65535 =
// 2^-16 * (1 + 2^-16 + 2^-32 + ...): the neglected tail is 2^-48 relative, far inside
// the distance of any j/max from a rounding boundary (>= 2^-33 / 2^-41 relative), so
// the single rounding of the last FFMA gives the correctly rounded quotient for every
// j in range (exhaustive check in tests/test_oracle.py).  RN(j * RN(1/255)) alone is
// wrong for about half of all j.  The saturation implements both clamps (j < 0 -> 0,
// j > max -> 1).  The generic kernel keeps the older two-term form sat(FFMA(j, khi,
// RN(j*klo))), khi + klo = 1/max to 48 bits (same exhaustive check).
struct NormK { float khi, klo, c1, c0, c2; float kd; int jmax; };

__device__ __forceinline__ f2 add2_rz(f2 a, f2 b) {
    f2 r; asm("add.rz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
__device__ __forceinline__ float fma_sat(float a, float b, float c) {
    float r; asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;
}
// two values at a time (packed where the ISA allows; .sat exists only on scalar ops)
__device__ __forceinline__ f2 quant_norm2(f2 r, const NormK &k) {
    const f2 m = add2_rz(r, bc(GMATB_MAGIC));
    const f2 hi = fma2(m, bc(k.c1), bc(k.c0));
    float h0, h1;
    upk(hi, h0, h1);
    return pk(fma_sat(h0, k.c2, h0), fma_sat(h1, k.c2, h1));
}

// The same sample, scaled by 2^-NORM_SHIFT (14 for 8-bit, 6 for 16-bit sources), with every FP32 operation PACKED and the clamp
// on the integer pipe in ONE instruction (the headline kernel's form, scale_fused3.cuh):
//   d  = FMUL2.RZ(r, 2^-149)         bit pattern = trunc(r), sign-magnitude: as s32, negative for r < 0
//   j  = VIMNMX.RELU(d, max)         = max(min(d, max), 0): both clamps; still the denormal float j * 2^-149
//   hi = FMUL2(j, kd)                kd = c1 * 2^(149 - NORM_SHIFT): j*c1 * 2^-NORM_SHIFT EXACTLY (j*0x010101 < 2^24)
//   p' = FFMA2(hi, c2, hi)           = RN(j/max) * 2^-NORM_SHIFT: scaling by a power of two commutes with the rounding
// The horizontal and vertical chains are linear in the samples with constant weights, so every intermediate is the
// reference's value times 2^-NORM_SHIFT exactly (no underflow: the host refuses weights below 2^-40), and the final
// multiply uses factor * 2^NORM_SHIFT.  Why: FFMA.SAT exists only as a scalar instruction, and a stream that alternates
// scalar and packed FP32 instructions retires ~81-89 % of the lane-operations of a homogeneous one
// (profiles/r1c_probe2_operand_forms.txt: "alt 4 FFMA2 + 4 FFMA" 3.24 units/clk/SM vs 4.0) -- the 19 % the fma pipe of
// the round-1 kernel was idle while throttling.
__device__ __forceinline__ f2 quant_norm2d(f2 r, const NormK &k) {
    int j0, j1;
    upki(mul2_rz(r, bc(GMATB_TWO_M149)), j0, j1);
    j0 = __vimin_s32_relu(j0, k.jmax); j1 = __vimin_s32_relu(j1, k.jmax);
    f2 jj; asm("mov.b64 %0, {%1, %2};" : "=l"(jj) : "r"(j0), "r"(j1));
    const f2 hi = mul2(jj, bc(k.kd));
    return fma2(hi, bc(k.c2), hi);
}
constexpr int norm_shift(int bits) { return bits == 8 ? 14 : 6; }

}  // namespace gmatb
