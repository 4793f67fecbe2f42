// scale_plane2.cuh -- exact 2:1 four-tap resample of ONE 8- or 16-bit PLANE with CH = 1 or 2 interleaved components (or 4 of
// 8 bits: rgb0 / bgr0 / rgba / bgra):
// the Y / U / V planes of yuv420p(16) and the Y / UV planes of nv12 / p010 / p016 in yuv -> yuv scaling at half size (4K -> 1080p,
// 1080p -> 540p: the step of an ABR ladder; the scale_cuda filter's path and sws_scale's yuv -> yuv branch,
// swscale_cuda.c:372-476).  The register-streaming layout of scale_fused3.cuh without the colour conversion:
//   * one warp per CTA, a lane owns a strip of 8 source pixels (4 outputs) and walks down a band one row pair per
//     step; loads two steps ahead; strips overlap by one lane per side, so halo columns are plain SHFL results;
//   * samples p = RN(j / max) as packed (top, bottom) pairs; the horizontal chain once per source row, the
//     vertical one as a running accumulation in the reference's operand order (resample_core.cuh): bit-identical
//     to the tile and streaming kernels and to the reference's Subsample_* kernels;
//   * a lane stores its 4 output pixels (4 .. 16 bytes) per output row, contiguous across the warp.
// Nothing but the source bytes is read and nothing but the destination written: 1.25 B per source sample, and
// ~5.5 FP32 lane-operations -- unlike the colour-converting kernels this one can be HBM-bound.
#pragma once
#include "scale_fused.cuh"

namespace gmatb {

struct Plane2Params {
    Plane src, dst;
    int W, H, dstW, dstH;          // source plane size in pixels; destination = half
    float wx[4], wy[4];
    NormK nk;
    int band, wrap;
};

template <int CH, int SBITS, int RA>
__global__ void __launch_bounds__(32, CH == 4 ? 12 : 16) plane_scale2_kernel(const __grid_constant__ Plane2Params P) {
    constexpr int OWN = 30;
    constexpr int BP = CH * SBITS / 8;                            // bytes per pixel: 1, 2, 2, 4
    constexpr int NWD = 2 * BP;                                   // words of 8 pixels
    constexpr int SMAX = SBITS == 8 ? 255 : 65535;
    const int lane = threadIdx.x;
    const int nstrips = P.W >> 3;
    const int strip = blockIdx.x * OWN + lane - 1;
    const bool owner = lane >= 1 && lane <= 30 && strip < nstrips;
    const int sl = min(max(strip, 0), nstrips - 1);
    const bool lrep = strip < 0, rrep = strip >= nstrips;          // out-of-frame provider strips: replicate the edge column
    const bool edge = blockIdx.x == 0 || (int)(blockIdx.x + 1) * OWN >= nstrips;       // warp-uniform
    const long long fz = blockIdx.z;
    const int H = P.H;
    const int yo_begin = blockIdx.y * P.band, yo_end = min(yo_begin + P.band, P.dstH);
    const uint8_t *ps = P.src.p + fz * P.src.bstride + (size_t)sl * (8 * BP);
    const unsigned pitch_s = P.src.pitch, pitch_d = P.dst.pitch;
    // pair k finishes output row k-1 and starts row k: pairs yo_begin-1 .. yo_end; the first two only prime the accumulators
    const int kfirst = yo_begin - 1, klast = yo_end, kstore = kfirst + 2;
    uint8_t *pd = P.dst.p + fz * P.dst.bstride + ((long long)kfirst - 1) * (long long)pitch_d + (long long)(owner ? strip : 0) * (4 * BP);

    struct Rows { uint32_t t[NWD], b[NWD]; };
    auto load_row = [&](const uint8_t *q, uint32_t (&w)[NWD]) {
        if (NWD == 2) { const uint2 a = ldg64(q); w[0] = a.x; w[1] = a.y; }
        else {
#pragma unroll
            for (int i = 0; i < NWD / 4; i++) {
                const uint4 a = ldg128(q + 16 * i);
                w[4 * i] = a.x; w[4 * i + 1] = a.y; w[(4 * i + 2) % NWD] = a.z; w[(4 * i + 3) % NWD] = a.w;
            }
        }
        if (edge) {                              // out-of-frame provider strips: pixel 7 := pixel 0 (left), pixel 0 := pixel 7 (right)
            if (BP == 1) {
                if (lrep) w[1] = prmt(w[1], w[0], 0x4210u);
                if (rrep) w[0] = prmt(w[0], w[1], 0x3217u);
            } else if (BP == 2) {                // a pixel is half a word
                if (lrep) w[NWD - 1] = prmt(w[NWD - 1], w[0], 0x5410u);
                if (rrep) w[0] = prmt(w[0], w[NWD - 1], 0x3276u);
            } else {                             // a pixel is a word
                if (lrep) w[NWD - 1] = w[0];
                if (rrep) w[0] = w[NWD - 1];
            }
        }
    };
    auto load_pair = [&](int k, Rows &R) {       // rows 2k, 2k+1 clamped to the frame
        const unsigned rt = (unsigned)min(max(2 * k, 0), H - 1), rb = (unsigned)min(max(2 * k + 1, 0), H - 1);
        load_row(ps + rt * pitch_s, R.t); load_row(ps + rb * pitch_s, R.b);
    };

    float acc[4][CH], hb_prev[4][CH];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int c = 0; c < CH; c++) { acc[i][c] = 0.f; hb_prev[i][c] = 0.f; }

    // (top, bottom) magic floats -> the samples: RN(j / max) (R-B) or j (R-A: SWS_BILINEAR / SWS_POINT)
    auto smp = [&](float mt, float mb) -> f2 { return RA ? add2(pk(mt, mb), bc(-GMATB_MAGIC)) : norm2_inrange(mt, mb, P.nk); };
    auto step = [&](const Rows &now, bool store) {
        // the 8 columns as (top, bottom) sample pairs
        f2 C[8][CH];
#pragma unroll
        for (int x = 0; x < 8; x++) {
            if (BP == 1) {
                const uint32_t a = now.t[x >> 2], b = now.b[x >> 2];
                C[x][0] = (x & 3) == 0 ? smp(byte_magic<0>(a), byte_magic<0>(b)) : (x & 3) == 1 ? smp(byte_magic<1>(a), byte_magic<1>(b))
                        : (x & 3) == 2 ? smp(byte_magic<2>(a), byte_magic<2>(b)) : smp(byte_magic<3>(a), byte_magic<3>(b));
            } else if (SBITS == 8 && CH == 2) {  // 2 components of 8 bits: a word is two pixels
                const uint32_t a = now.t[(x >> 1) % NWD], b = now.b[(x >> 1) % NWD];
                if (x & 1) { C[x][0] = smp(byte_magic<2>(a), byte_magic<2>(b)); C[x][CH - 1] = smp(byte_magic<3>(a), byte_magic<3>(b)); }
                else       { C[x][0] = smp(byte_magic<0>(a), byte_magic<0>(b)); C[x][CH - 1] = smp(byte_magic<1>(a), byte_magic<1>(b)); }
            } else if (SBITS == 8 && CH == 4) {  // 4 components of 8 bits: a word is one pixel
                const uint32_t a = now.t[x % NWD], b = now.b[x % NWD];
                C[x][0] = smp(byte_magic<0>(a), byte_magic<0>(b)); C[x][1 % CH] = smp(byte_magic<1>(a), byte_magic<1>(b));
                C[x][2 % CH] = smp(byte_magic<2>(a), byte_magic<2>(b)); C[x][3 % CH] = smp(byte_magic<3>(a), byte_magic<3>(b));
            } else if (CH == 1) {                // 1 component of 16 bits: a word is two pixels
                const uint32_t a = now.t[(x >> 1) % NWD], b = now.b[(x >> 1) % NWD];
                C[x][0] = (x & 1) ? smp(half_magic<1>(a), half_magic<1>(b)) : smp(half_magic<0>(a), half_magic<0>(b));
            } else {                             // 2 components of 16 bits: a word is one pixel
                const uint32_t a = now.t[x % NWD], b = now.b[x % NWD];
                C[x][0] = smp(half_magic<0>(a), half_magic<0>(b)); C[x][CH - 1] = smp(half_magic<1>(a), half_magic<1>(b));
            }
        }
        f2 PL[CH], PR[CH];
#pragma unroll
        for (int c = 0; c < CH; c++) { PL[c] = shfl_up2(C[7][c]); PR[c] = shfl_dn2(C[0][c]); }
        float ht[4][CH], hb[4][CH];
#pragma unroll
        for (int xo = 0; xo < 4; xo++)
#pragma unroll
            for (int c = 0; c < CH; c++) {
                const f2 p0 = xo == 0 ? PL[c] : C[2 * xo - 1][c];
                const f2 p3 = xo == 3 ? PR[c] : C[2 * xo + 2][c];
                upk(hpass<false>(P.wx, p0, C[2 * xo][c], C[2 * xo + 1][c], p3), ht[xo][c], hb[xo][c]);
            }
        if (store) {
            int o[4][CH];
#pragma unroll
            for (int xo = 0; xo < 4; xo++)
#pragma unroll
                for (int c = 0; c < CH; c++) {
                    const float v = __fmaf_rn(P.wy[3], ht[xo][c], acc[xo][c]);
                    if (RA) o[xo][c] = __float_as_int(__fadd_rn(v, 12582912.0f)) - 0x4B400000;      // rint; the pack saturates
                    else {
                        o[xo][c] = trunc_i(__fmul_rn(v, (float)SMAX));
                        if (P.wrap) o[xo][c] = max(o[xo][c], 0) & SMAX;
                    }
                }
            if (BP == 1) stg32(pd, pack4_u8(o[0][0], o[1][0], o[2][0], o[3][0]));
            else if (SBITS == 8 && CH == 4) stg128(pd, make_uint4(pack4_u8(o[0][0], o[0][1 % CH], o[0][2 % CH], o[0][3 % CH]), pack4_u8(o[1][0], o[1][1 % CH], o[1][2 % CH], o[1][3 % CH]),
                                                                  pack4_u8(o[2][0], o[2][1 % CH], o[2][2 % CH], o[2][3 % CH]), pack4_u8(o[3][0], o[3][1 % CH], o[3][2 % CH], o[3][3 % CH])));
            else if (SBITS == 8) stg64(pd, make_uint2(pack4_u8(o[0][0], o[0][CH - 1], o[1][0], o[1][CH - 1]), pack4_u8(o[2][0], o[2][CH - 1], o[3][0], o[3][CH - 1])));
            else if (CH == 1) stg64(pd, make_uint2(pack2_u16(o[0][0], o[1][0]), pack2_u16(o[2][0], o[3][0])));
            else stg128(pd, make_uint4(pack2_u16(o[0][0], o[0][CH - 1]), pack2_u16(o[1][0], o[1][CH - 1]), pack2_u16(o[2][0], o[2][CH - 1]), pack2_u16(o[3][0], o[3][CH - 1])));
        }
#pragma unroll
        for (int xo = 0; xo < 4; xo++)
#pragma unroll
            for (int c = 0; c < CH; c++) {
                float t = __fmul_rn(P.wy[1], ht[xo][c]);
                t = __fmaf_rn(P.wy[0], hb_prev[xo][c], t);
                t = __fmaf_rn(P.wy[2], hb[xo][c], t);
                acc[xo][c] = t;
                hb_prev[xo][c] = hb[xo][c];
            }
    };

    // A holds pair k, B pair k+1; each step refills its own buffer with the pair two steps ahead
    Rows A, B;
    int k = kfirst;
    load_pair(k, A);
    load_pair(k + 1, B);
#pragma unroll 1
    for (;;) {
        {
            const Rows now = A;
            if (k + 2 <= klast) load_pair(k + 2, A);
            step(now, owner && k >= kstore);
            pd += pitch_d;
            if (++k > klast) break;
        }
        {
            const Rows now = B;
            if (k + 2 <= klast) load_pair(k + 2, B);
            step(now, owner && k >= kstore);
            pd += pitch_d;
            if (++k > klast) break;
        }
    }
}

}  // namespace gmatb
