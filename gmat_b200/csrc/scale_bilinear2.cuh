// scale_bilinear2.cuh -- fused yuv 4:2:0 (8-bit) -> packed rgb + exact 2:1 BILINEAR downscale.
//
// This is what the reference actually executes for every scaled yuv->rgb conversion: its
// ff_swscale_cuda passes c->cspace (always 0) to map_resize_algo, so the resize is always
// NVCV_INTERP_LINEAR whatever SWS_* flag was given (swscale_cuda.c:69-74,305; SURVEY finding 8a).
// Arithmetic = "R-A" (CV-CUDA is closed, so our restatement is the spec): CSC at source
// resolution with the reference's chain, quantise to u8, then half-pixel-centre bilinear in fp32
// on the integer samples, round-to-nearest-even, saturate.  At exactly 2:1 both weights are 0.5
// on both axes, every product and sum is exact, and the result is RNE((a+b+c+d)/4) of the 2x2
// block -- so after the (bit-exact, packed-fp32) colour conversion everything is integer work:
//   float -> int : FMUL2.RZ by 2^-149 (the denormal's bit pattern is the integer)
//   clamp + pack : I2IP.U8.S32.SAT puts the 4 samples of a 2x2 block into one register
//   sum          : IDP4A against 0x01010101, accumulating INTO 0x4B000000: the result is the bit
//                  pattern of the float 2^23 + s
//   RNE(s/4)     : FFMA(2^23 + s, 0.25, 2^23 - 2^21) = 2^23 + s/4 rounded to an integer by the
//                  FMA's own round-to-nearest-even (ulp(2^23) = 1); the low byte is the pixel
// A thread owns an 8x2 source tile = 4 destination pixels; there is no halo, no shared memory and
// no dependency between rows, so the kernel is a pure streaming kernel at ~12 instructions per
// source pixel (the exact 4-tap kernels need ~35).
#pragma once
#include "csc_core.cuh"

namespace gmatb {

template <int L> __device__ __forceinline__ void bl2_load(const Img &s, long long fz, int x0, int y0,
                                                          float (&ym)[2][8], float (&um)[4], float (&vm)[4]) {
    const uint8_t *py = s.pl[0].p + fz * s.pl[0].bstride;
#pragma unroll
    for (int r = 0; r < 2; r++) {
        uint2 w = ldg64(py + (size_t)(y0 + r) * s.pl[0].pitch + x0);
        ym[r][0] = byte_magic<0>(w.x); ym[r][1] = byte_magic<1>(w.x); ym[r][2] = byte_magic<2>(w.x); ym[r][3] = byte_magic<3>(w.x);
        ym[r][4] = byte_magic<0>(w.y); ym[r][5] = byte_magic<1>(w.y); ym[r][6] = byte_magic<2>(w.y); ym[r][7] = byte_magic<3>(w.y);
    }
    if (L == L_NV12) {
        uint2 w = ldg64(s.pl[1].p + fz * s.pl[1].bstride + (size_t)(y0 >> 1) * s.pl[1].pitch + x0);
        um[0] = byte_magic<0>(w.x); vm[0] = byte_magic<1>(w.x); um[1] = byte_magic<2>(w.x); vm[1] = byte_magic<3>(w.x);
        um[2] = byte_magic<0>(w.y); vm[2] = byte_magic<1>(w.y); um[3] = byte_magic<2>(w.y); vm[3] = byte_magic<3>(w.y);
    } else {
        uint32_t wu = ldg32(s.pl[1].p + fz * s.pl[1].bstride + (size_t)(y0 >> 1) * s.pl[1].pitch + (x0 >> 1));
        uint32_t wv = ldg32(s.pl[2].p + fz * s.pl[2].bstride + (size_t)(y0 >> 1) * s.pl[2].pitch + (x0 >> 1));
        um[0] = byte_magic<0>(wu); um[1] = byte_magic<1>(wu); um[2] = byte_magic<2>(wu); um[3] = byte_magic<3>(wu);
        vm[0] = byte_magic<0>(wv); vm[1] = byte_magic<1>(wv); vm[2] = byte_magic<2>(wv); vm[3] = byte_magic<3>(wv);
    }
}

// low bytes of four registers -> one word (3 PRMT)
__device__ __forceinline__ uint32_t lowbytes4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    return prmt(prmt(a, b, 0x0040u), prmt(c, d, 0x0040u), 0x5410u);
}

// DST: D_RGB24 / D_BGR24 / D_RGBA / D_BGRA.  Source width % 8 == 0, height even, 16-byte aligned planes.
template <int L, int DST, int MINB>
__global__ void __launch_bounds__(256, MINB) fused_csc_bilinear2_kernel(Img src, Img dst, Mat9 M) {
    const int x0 = (blockIdx.x * 32 + threadIdx.x) * 8;
    const int y0 = (blockIdx.y * 8 + threadIdx.y) * 2;
    if (x0 >= src.w || y0 >= src.h) return;
    const long long fz = blockIdx.z;
    float ym[2][8], um[4], vm[4];
    bl2_load<L>(src, fz, x0, y0, ym, um, vm);
    constexpr float YB = -(GMATB_MAGIC + 16.f), CB = -(GMATB_MAGIC + 128.f);
    const f2 z = bc(GMATB_TWO_M149);
    uint32_t o[4][3];      // bit patterns 0x4B0000xx: the pixel value is the low byte
#pragma unroll
    for (int j = 0; j < 4; j++) {
        float fu, fv;
        upk(add2(pk(um[j], vm[j]), bc(CB)), fu, fv);
        const ChromaTerms t = chroma_terms<true>(fu, fv, M);
        int q[2][3][2];        // [column][channel][top/bottom]
#pragma unroll
        for (int cidx = 0; cidx < 2; cidx++) {
            f2 r, g, b;
            csc_pair_f<true>(add2(pk(ym[0][2 * j + cidx], ym[1][2 * j + cidx]), bc(YB)), t, M, r, g, b);
            upki(mul2_rz(r, z), q[cidx][0][0], q[cidx][0][1]);
            upki(mul2_rz(g, z), q[cidx][1][0], q[cidx][1][1]);
            upki(mul2_rz(b, z), q[cidx][2][0], q[cidx][2][1]);
        }
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const unsigned w = pack4_u8(q[0][c][0], q[0][c][1], q[1][c][0], q[1][c][1]);   // clamped 2x2 block
            const float f = __uint_as_float(__dp4a(w, 0x01010101u, 0x4B000000u));            // 2^23 + (a+b+c+d)
            o[j][c] = __float_as_uint(__fmaf_rn(f, 0.25f, 6291456.0f));                     // 2^23 + RNE(sum/4)
        }
    }
    constexpr bool SW = dst_swap(DST);
    uint8_t *pd = dst.pl[0].p + fz * dst.pl[0].bstride + (size_t)(y0 >> 1) * dst.pl[0].pitch + (size_t)(x0 >> 1) * dst_bpp(DST);
#define CH(i, c) o[i][SW ? 2 - (c) : (c)]
    if (DST == D_RGB24 || DST == D_BGR24) {
        stg32(pd,     lowbytes4(CH(0, 0), CH(0, 1), CH(0, 2), CH(1, 0)));
        stg32(pd + 4, lowbytes4(CH(1, 1), CH(1, 2), CH(2, 0), CH(2, 1)));
        stg32(pd + 8, lowbytes4(CH(2, 2), CH(3, 0), CH(3, 1), CH(3, 2)));
    } else {
        stg128(pd, make_uint4(lowbytes4(CH(0, 0), CH(0, 1), CH(0, 2), 255u), lowbytes4(CH(1, 0), CH(1, 1), CH(1, 2), 255u),
                              lowbytes4(CH(2, 0), CH(2, 1), CH(2, 2), 255u), lowbytes4(CH(3, 0), CH(3, 1), CH(3, 2), 255u)));
    }
#undef CH
}

}  // namespace gmatb
