// scale_fused.cuh -- building blocks of the headline kernel (scale_fused3.cuh): YUV 4:2:0 -> packed
// RGB colour conversion FUSED with an exact 2:1 four-tap resample (4K NV12 -> 1080p RGB24, 8K P010 ->
// 4K RGB48): row-pair loads, unpacking, the packed CSC + quantise + normalise column, the 4-tap pass.
//
// The reference runs two kernels with a full-resolution RGB image written to and read
// back from HBM in between (swscale_cuda.c:342-371: yuv2rgb_cuda -> cvcudaResizeSubmit;
// 68.4 MB of traffic per 4K frame).  Here the source is read once and only the
// destination is written (18.7 MB): 2.25 bytes per source pixel.
//
// Work decomposition -- "register-streaming column strips":
//   * a thread owns a strip of 8 source columns (= 4 output columns) and walks down the
//     frame one chroma row (= 2 luma rows) per iteration; a warp therefore reads 256
//     contiguous luma bytes per row (64-bit coalesced loads), loads for the next
//     iteration are issued before the current one is consumed;
//   * per iteration the 8x2 pixels are converted with the reference's CSC chain packed
//     as (top,bottom) f32x2 pairs (the two rows share their chroma), quantised to the
//     u8/u16 value the reference's intermediate image would hold and normalised the way
//     its texture fetch does;
//   * the horizontal 4-tap pass runs packed over the row pair; the two halo columns come
//     from the neighbouring lanes by warp shuffle (strips overlap by one lane per side);
//   * the vertical pass is a running accumulation in the reference's operand order
//     (row 2yo first, then 2yo-1, 2yo+1, 2yo+2): only 12 partial sums + 12 saved
//     horizontal results are carried between iterations, no shared memory, no barriers;
//   * output rows leave as 12/16/24/32-byte pieces per thread, contiguous across the warp.
// The frame is cut into horizontal bands (blockIdx.y) and frames of a batch sit on
// blockIdx.z, so one launch covers a whole batch with tens of thousands of warps.
#pragma once
#include "csc_core.cuh"
#include "resample_core.cuh"

namespace gmatb {

template <int SBITS> struct RawRow;   // raw loaded words of one row pair
template <> struct RawRow<8>  { uint2 yt, yb, c0; };   // I420: c0.x = 4 U bytes, c0.y = 4 V bytes
template <> struct RawRow<16> { uint4 yt, yb, c0; };   // I420: c0.xy = 4 U, c0.zw = 4 V

// packed 3-byte rgb source (rgb24 -> rgb24 scaling: no colour conversion, same resample)
struct RawRowRGB { uint2 t[3], b[3]; };
// byte B (0..23) of a 24-byte row -> magic float
template <int B> __device__ __forceinline__ float rgb_byte_magic(const uint2 (&w)[3]) {
    const uint32_t word = (B & 4) ? w[B >> 3].y : w[B >> 3].x;
    return byte_magic<B & 3>(word);
}
// normalised sample pair of integer-valued inputs that are already in range: p = RN(j/max), no clamp needed
__device__ __forceinline__ f2 norm2_inrange(float mt, float mb, const NormK &k) {
    const f2 hi = fma2(pk(mt, mb), bc(k.c1), bc(k.c0));
    return fma2(hi, bc(k.c2), hi);
}
__device__ __forceinline__ float norm_inrange(float m, const NormK &k) {      // one value: m = 2^23 + j
    const float hi = __fmaf_rn(m, k.c1, k.c0);
    return __fmaf_rn(hi, k.c2, hi);
}
template <int C> __device__ __forceinline__ void rgb_column(const RawRowRGB &R, const NormK &nk, f2 (&out)[3]) {
    out[0] = norm2_inrange(rgb_byte_magic<3 * C>(R.t), rgb_byte_magic<3 * C>(R.b), nk);
    out[1] = norm2_inrange(rgb_byte_magic<3 * C + 1>(R.t), rgb_byte_magic<3 * C + 1>(R.b), nk);
    out[2] = norm2_inrange(rgb_byte_magic<3 * C + 2>(R.t), rgb_byte_magic<3 * C + 2>(R.b), nk);
}

// unpack the raw words into magic floats: ym[row][col], um/vm[chroma sample]
template <int L>
__device__ __forceinline__ void fused_unpack(const RawRow<8> &R, float (&yt)[8], float (&yb)[8], float (&um)[4], float (&vm)[4]) {
    yt[0] = byte_magic<0>(R.yt.x); yt[1] = byte_magic<1>(R.yt.x); yt[2] = byte_magic<2>(R.yt.x); yt[3] = byte_magic<3>(R.yt.x);
    yt[4] = byte_magic<0>(R.yt.y); yt[5] = byte_magic<1>(R.yt.y); yt[6] = byte_magic<2>(R.yt.y); yt[7] = byte_magic<3>(R.yt.y);
    yb[0] = byte_magic<0>(R.yb.x); yb[1] = byte_magic<1>(R.yb.x); yb[2] = byte_magic<2>(R.yb.x); yb[3] = byte_magic<3>(R.yb.x);
    yb[4] = byte_magic<0>(R.yb.y); yb[5] = byte_magic<1>(R.yb.y); yb[6] = byte_magic<2>(R.yb.y); yb[7] = byte_magic<3>(R.yb.y);
    if (L == L_NV12) {
        um[0] = byte_magic<0>(R.c0.x); vm[0] = byte_magic<1>(R.c0.x); um[1] = byte_magic<2>(R.c0.x); vm[1] = byte_magic<3>(R.c0.x);
        um[2] = byte_magic<0>(R.c0.y); vm[2] = byte_magic<1>(R.c0.y); um[3] = byte_magic<2>(R.c0.y); vm[3] = byte_magic<3>(R.c0.y);
    } else {
        um[0] = byte_magic<0>(R.c0.x); um[1] = byte_magic<1>(R.c0.x); um[2] = byte_magic<2>(R.c0.x); um[3] = byte_magic<3>(R.c0.x);
        vm[0] = byte_magic<0>(R.c0.y); vm[1] = byte_magic<1>(R.c0.y); vm[2] = byte_magic<2>(R.c0.y); vm[3] = byte_magic<3>(R.c0.y);
    }
}
template <int L>
__device__ __forceinline__ void fused_unpack(const RawRow<16> &R, float (&yt)[8], float (&yb)[8], float (&um)[4], float (&vm)[4]) {
    yt[0] = half_magic<0>(R.yt.x); yt[1] = half_magic<1>(R.yt.x); yt[2] = half_magic<0>(R.yt.y); yt[3] = half_magic<1>(R.yt.y);
    yt[4] = half_magic<0>(R.yt.z); yt[5] = half_magic<1>(R.yt.z); yt[6] = half_magic<0>(R.yt.w); yt[7] = half_magic<1>(R.yt.w);
    yb[0] = half_magic<0>(R.yb.x); yb[1] = half_magic<1>(R.yb.x); yb[2] = half_magic<0>(R.yb.y); yb[3] = half_magic<1>(R.yb.y);
    yb[4] = half_magic<0>(R.yb.z); yb[5] = half_magic<1>(R.yb.z); yb[6] = half_magic<0>(R.yb.w); yb[7] = half_magic<1>(R.yb.w);
    if (L == L_NV12) {
        um[0] = half_magic<0>(R.c0.x); vm[0] = half_magic<1>(R.c0.x); um[1] = half_magic<0>(R.c0.y); vm[1] = half_magic<1>(R.c0.y);
        um[2] = half_magic<0>(R.c0.z); vm[2] = half_magic<1>(R.c0.z); um[3] = half_magic<0>(R.c0.w); vm[3] = half_magic<1>(R.c0.w);
    } else {
        um[0] = half_magic<0>(R.c0.x); um[1] = half_magic<1>(R.c0.x); um[2] = half_magic<0>(R.c0.y); um[3] = half_magic<1>(R.c0.y);
        vm[0] = half_magic<0>(R.c0.z); vm[1] = half_magic<1>(R.c0.z); vm[2] = half_magic<0>(R.c0.w); vm[3] = half_magic<1>(R.c0.w);
    }
}

__device__ __forceinline__ f2 shfl_up2(f2 v) { return __shfl_up_sync(0xffffffffu, v, 1); }
__device__ __forceinline__ f2 shfl_dn2(f2 v) { return __shfl_down_sync(0xffffffffu, v, 1); }

template <bool TAPS2>
__device__ __forceinline__ f2 hpass(const float (&w)[4], f2 p0, f2 p1, f2 p2, f2 p3) {
    f2 t = mul2(bc(w[1]), p1);
    if (!TAPS2) t = fma2(bc(w[0]), p0, t);
    t = fma2(bc(w[2]), p2, t);
    if (!TAPS2) t = fma2(bc(w[3]), p3, t);
    return t;
}

// The kernel itself is fused_csc_scale2_v3_kernel in scale_fused3.cuh (v1, which converted a halo column
// per warp and selected it into lanes 0/31, is kept in experiments/scale_fused_v1.cuh).

}  // namespace gmatb
