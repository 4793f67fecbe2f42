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
// There is no halo, no shared memory and no dependency between rows: a pure streaming kernel at ~12
// instructions per source pixel (the exact 4-tap kernel needs ~22).  DST: D_RGB24 / D_BGR24 / D_RGBA /
// D_BGRA; source width % 8 == 0, height even, 16-byte aligned planes.
#pragma once
#include "csc_core.cuh"
#include "scale_fused.cuh"

namespace gmatb {

// low bytes of four registers -> one word (3 PRMT)
__device__ __forceinline__ uint32_t lowbytes4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    return prmt(prmt(a, b, 0x0040u), prmt(c, d, 0x0040u), 0x5410u);
}

// A lane owns an 8-column strip and walks down a band of row pairs (8x2 source pixels = 4 destination
// pixels per step), so the index / pointer set-up is paid once per band, and the loads of row pair k+2
// are issued while pair k is converted (two ping-pong register buffers; DRAM latency under load is about
// two steps of a warp's stream).  Measured on B200, C2 x 64 frames: one thread per 8x2 tile 1672 Gpx/s,
// this form 2092 Gpx/s = 71.8 % of the measured HBM copy peak.
struct alignas(8) Bl2Params {
    Img src, dst;
    float cm45[2], cm72[2];     // (m4, m5), (m7, m2): 8-byte aligned pairs FMUL2 takes from uniform registers
    float m0, m1, m3, m6;       // m1 == 0.0f at run time (csc_core.cuh)
    int band;                   // row pairs (= output rows) per CTA
};

template <int L> struct Bl2Row { uint2 yt, yb, c0; };   // I420: c0.x = 4 U bytes, c0.y = 4 V bytes

template <int L, int DST>
__device__ __forceinline__ void bl2_convert_store(const Bl2Row<L> &R, const Bl2Params &P, uint8_t *pd) {
    constexpr float YB = -(GMATB_MAGIC + 16.f), CB = -(GMATB_MAGIC + 128.f);
    const f2 z = bc(GMATB_TWO_M149);
    const f2 k45 = *reinterpret_cast<const f2 *>(P.cm45), k72 = *reinterpret_cast<const f2 *>(P.cm72);
    float yt[8], yb[8], um[4], vm[4];
    RawRow<8> rr; rr.yt = R.yt; rr.yb = R.yb; rr.c0 = R.c0;
    fused_unpack<L>(rr, yt, yb, um, vm);
    uint32_t o[4][3];      // bit patterns 0x4B0000xx: the pixel value is the low byte
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const f2 uv = add2(pk(um[j], vm[j]), bc(CB));
        float t1g, t2g, t1b, t2r;
        upk(mul2(uv, k45), t1g, t2g);
        upk(mul2(uv, k72), t1b, t2r);
        int q[2][3][2];        // [column][channel][top/bottom]
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const f2 fy2 = add2(pk(yt[2 * j + h], yb[2 * j + h]), bc(YB));
            const f2 r = add2(fma2(fy2, bc(P.m0), bc(P.m1)), bc(t2r));
            const f2 g = add2(fma2(fy2, bc(P.m3), bc(t1g)), bc(t2g));
            const f2 b = fma2(fy2, bc(P.m6), bc(t1b));
            upki(mul2_rz(r, z), q[h][0][0], q[h][0][1]);
            upki(mul2_rz(g, z), q[h][1][0], q[h][1][1]);
            upki(mul2_rz(b, z), q[h][2][0], q[h][2][1]);
        }
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const unsigned w = pack4_u8(q[0][c][0], q[0][c][1], q[1][c][0], q[1][c][1]);   // clamped 2x2 block
            const float f = __uint_as_float(__dp4a(w, 0x01010101u, 0x4B000000u));            // 2^23 + (a+b+c+d)
            o[j][c] = __float_as_uint(__fmaf_rn(f, 0.25f, 6291456.0f));                     // 2^23 + RNE(sum/4)
        }
    }
    constexpr bool SW = dst_swap(DST);
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

template <int L, int DST, int MINB>
__global__ void __launch_bounds__(32, MINB) fused_csc_bilinear2_stream_kernel(const Bl2Params P) {
    const int strip = blockIdx.x * 32 + threadIdx.x;
    if (strip * 8 >= P.src.w) return;
    const long long fz = blockIdx.z;
    const int HC = P.src.h >> 1;
    const int k0 = blockIdx.y * P.band, k1 = min(k0 + P.band, HC);
    const uint8_t *py = P.src.pl[0].p + fz * P.src.pl[0].bstride + (size_t)strip * 8;
    const uint8_t *pu, *pv = py;
    if (L == L_NV12) pu = P.src.pl[1].p + fz * P.src.pl[1].bstride + (size_t)strip * 8;
    else {
        pu = P.src.pl[1].p + fz * P.src.pl[1].bstride + (size_t)strip * 4;
        pv = P.src.pl[2].p + fz * P.src.pl[2].bstride + (size_t)strip * 4;
    }
    const unsigned pitch_y = P.src.pl[0].pitch, pitch_c = P.src.pl[1].pitch, pitch_c2 = P.src.pl[2].pitch, pitch_d = P.dst.pl[0].pitch;
    uint8_t *pd = P.dst.pl[0].p + fz * P.dst.pl[0].bstride + (size_t)k0 * pitch_d + (size_t)strip * (4 * dst_bpp(DST));
    // warp-uniform offsets of the next row pair to load
    unsigned ot = (unsigned)(2 * k0) * pitch_y, oc = (unsigned)k0 * pitch_c, oc2 = (unsigned)k0 * pitch_c2;
    auto load = [&](Bl2Row<L> &R) {
        R.yt = ldg64(py + ot); R.yb = ldg64(py + ot + pitch_y);
        if (L == L_NV12) R.c0 = ldg64(pu + oc);
        else { R.c0.x = ldg32(pu + oc); R.c0.y = ldg32(pv + oc2); }
        ot += 2 * pitch_y; oc += pitch_c; oc2 += pitch_c2;
    };
    Bl2Row<L> A, B;
    load(A);
    if (k0 + 1 < k1) load(B);
    int k = k0;
    for (; k + 3 < k1; k += 2) {         // pairs k+2 and k+3 exist
        { const Bl2Row<L> now = A; load(A); bl2_convert_store<L, DST>(now, P, pd); pd += pitch_d; }
        { const Bl2Row<L> now = B; load(B); bl2_convert_store<L, DST>(now, P, pd); pd += pitch_d; }
    }
#pragma unroll 1
    for (; k < k1; k++) {                // the last (up to three) pairs
        const Bl2Row<L> now = A;
        A = B;
        if (k + 2 < k1) load(B);
        bl2_convert_store<L, DST>(now, P, pd); pd += pitch_d;
    }
}

}  // namespace gmatb
