// scale_fused5m.cuh -- the headline kernel with the horizontal pass on the TENSOR pipe: 8-bit NV12 -> packed
// RGB colour conversion FUSED with the exact 2:1 four-tap resample of scale_fused4i.cuh (dyadic weights
// (WA, WB, WB, WA) / 2^WS on both axes: R-B bicubic with param0 = 0.75 / 0.5 / 1.0), bit for bit the results of
// scale_fused3.cuh.
//
// Why.  The float chain (scale_fused3.cuh) is fma-pipe bound: 27 FP32 lane-operations per source pixel, 16.5 of
// them in quantise/normalise and the two 4-tap passes.  The exact-integer restatement (scale_fused4i.cuh: the
// exactness argument is there) turns the passes into integer dot products, but IDP4A / IMAD retire on the same
// fma pipe at half rate, so it gained nothing (0.93-0.97 x).  IMMA (mma.sync m16n8k32 u8 x s8 -> s32) retires on
// the tensor pipe, beside the fma and alu pipes (tools/probe_imma.cu: 2 IMMA + 16 IDP4A + 16 PRMT per warp take
// the time of the 16 + 16 alone).  The horizontal pass is a banded matrix product
//     T[row][xo] = sum_k J[row][2 xo - 1 + k] W[k]
// of the quantised intermediate image J (bytes, the values the reference's first kernel would store) with a
// 32 x 8 band matrix of s8 weights, and the first half of the vertical pass rides on the accumulator operand.
//
// Layout.  One warp per CTA walks down a band of the frame one row PAIR per step, as v3 / v4i do.  It covers
// 8 "lines" of 32 source columns (lane = 4 g + t: g = line, t = position in the line).  The M = 16 rows of the
// A fragment are (line g, top row) and (line g, bottom row); K = 32 source columns; N = 8 output columns:
//     A regs of a lane: a0/a1 = 4 quantised pixels (one channel) of the top/bottom row at columns 4t..4t+3 of a
//     16-column block, a2/a3 = the same of the next block -- i.e. exactly the bytes the lane's own colour
//     conversion produces: a lane converts two 4-column chunks, 16 columns apart, of both rows (6 LDG.32).
//     Line g consists of blocks b0 = [s-4, s+12) and b1 = [s+12, s+28), s = 32 g (+ the warp's origin); outputs
//     0..7 of the line come from MMA(b0 | b1) and outputs 8..15 from MMA(b1 | b2), b2 = b0 of line g+1 = the a0/a1
//     registers of lane + 4 (one SHFL per register).  The band matrix B[k][n] = W[k - 2n - 3] is the same for
//     both (the window starts 4 columns left of the first output's centre pair), a per-lane constant pair.
//     Line 7 has no right neighbour: its second MMA is not stored, the warp owns 7 * 16 + 8 = 120 output columns
//     = the 240 source columns of a v3 warp (so a band can be handed to fused3_band with the same blockIdx.x).
//   * per (channel, MMA): D = A x B (c0,c1 = T of the top row = e, c2,c3 = T of the bottom row = o) and
//     D' = A x (WA B) + (P, P, 0, 0): c0',c1' = N[k-1] = P[k-1] + WA e[k], the finished integer of output row k-1,
//     c2',c3' = WA o[k].  Then P[k] = WB (e[k] + o[k]) + WA o[k-1]: one IADD and one IMAD per output.
//   * output = sat(N >> 2WS); outputs with N = 0 (mod 2^2WS), N != 0, are recomputed with the float chain from
//     a shared-memory ring of the quantised bytes of the last four steps (few per step), or the band continues in
//     the float loop (many: flat content) -- as scale_fused4i.cuh.
//   * a lane owns output columns 2t, 2t+1 of each MMA: 6 bytes of rgb24 (one 32-bit and one 16-bit store whose
//     order depends on the parity of t) or 8 bytes of rgba.
// Frame edges: chunk -1 / chunk W/4 are loaded from the edge chunk and byte-replicated in the raw words (only
// source columns -1 and W ever meet a non-zero weight).
#pragma once
#include "scale_fused4i.cuh"

namespace gmatb {

#define GMATB_F5M_DENSE 8    /* more ambiguous outputs than this in one warp step: the band continues in float */

// D = A (16 x 32, u8, row) x B (32 x 8, s8, col) + C   (SASS IMMA.16832.U8.S8)
__device__ __forceinline__ void mma_u8s8(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1,
                                         int c0, int c1, int c2, int c3) {
    asm("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%10, %11, %12, %13};"
        : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "r"(c0), "r"(c1), "r"(c2), "r"(c3));
}
__device__ __forceinline__ void stg16(void *p, uint32_t v) { __stcs(reinterpret_cast<unsigned short *>(p), (unsigned short)v); }

// raw words of a lane's two 4-column chunks of one row pair (NV12: c = U0 V0 U1 V1)
struct Row5 { uint32_t yt[2], yb[2], c[2]; };

// 2 x (4 x 2) pixels -> the quantised intermediate image as planar bytes: J?[c][b] = channel c of chunk b
__device__ __forceinline__ void produce5m(const Row5 &R, const Fused3Params &P, uint32_t (&Jt)[3][2], uint32_t (&Jb)[3][2]) {
    constexpr float CB = -(GMATB_MAGIC + 128.f), YB = -(GMATB_MAGIC + 16.f);
    const f2 k45 = *reinterpret_cast<const f2 *>(P.cm45), k72 = *reinterpret_cast<const f2 *>(P.cm72);
    const f2 z = bc(GMATB_TWO_M149);
#pragma unroll
    for (int b = 0; b < 2; b++) {
        const float yt[4] = {byte_magic<0>(R.yt[b]), byte_magic<1>(R.yt[b]), byte_magic<2>(R.yt[b]), byte_magic<3>(R.yt[b])};
        const float yb[4] = {byte_magic<0>(R.yb[b]), byte_magic<1>(R.yb[b]), byte_magic<2>(R.yb[b]), byte_magic<3>(R.yb[b])};
        const float um[2] = {byte_magic<0>(R.c[b]), byte_magic<2>(R.c[b])}, vm[2] = {byte_magic<1>(R.c[b]), byte_magic<3>(R.c[b])};
        int it[4][3], ib[4][3];
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const f2 uv = add2(pk(um[j], vm[j]), bc(CB));
            float t1g, t2g, t1b, t2r;
            upk(mul2(uv, k45), t1g, t2g);
            upk(mul2(uv, k72), t1b, t2r);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int col = 2 * j + h;
                const f2 fy2 = add2(pk(yt[col], yb[col]), bc(YB));
                f2 xr = fma2(fy2, bc(P.m0), bc(P.m1));      // m1 is a run-time 0.0f (csc_core.cuh)
                f2 xg = fma2(fy2, bc(P.m3), bc(t1g));
                const f2 xb = fma2(fy2, bc(P.m6), bc(t1b));
                xr = add2(xr, bc(t2r)); xg = add2(xg, bc(t2g));
                upki(mul2_rz(xr, z), it[col][0], ib[col][0]);
                upki(mul2_rz(xg, z), it[col][1], ib[col][1]);
                upki(mul2_rz(xb, z), it[col][2], ib[col][2]);
            }
        }
#pragma unroll
        for (int c = 0; c < 3; c++) {
            Jt[c][b] = pack4_u8(it[0][c], it[1][c], it[2][c], it[3][c]);
            Jb[c][b] = pack4_u8(ib[0][c], ib[1][c], ib[2][c], ib[3][c]);
        }
    }
}

// ring word of (channel c, row r: 0 top / 1 bottom, chunk cw = 0..63 of the warp's strip) inside one ring slot
__device__ __forceinline__ int f5m_ring_word(int c, int r, int cw) {
    return (4 * (cw >> 3) + (cw & 3)) * 12 + c * 4 + ((cw >> 2) & 1) * 2 + r;
}

// One ambiguous output -- channel c of output column xw (0..119, relative to the warp) of output row k-1, found
// while pair k is processed -- recomputed by ONE lane with the float chain (the operations of scale_fused3.cuh /
// resample_core.cuh) from the quantised bytes in the ring (rows b[k-2], t[k-1], b[k-1], t[k]).
template <bool WRAP>
__device__ __noinline__ int fused5m_fix(const Fused3Params &P, const uint32_t *ring, int k, int xw, int c) {
    const int x0 = 2 * xw + 3;                  // first tap, in columns of the warp's strip (strip column 0 = source column 240 bx - 4)
    const int cw = x0 >> 2, sh = (x0 & 3) * 8;
    float h[4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const uint32_t *slot = ring + ((k - 2 + ((r + 1) >> 1)) & 3) * (32 * 12);
        const int rb = (r & 1) ? 0 : 1;
        const uint32_t w = __funnelshift_r(slot[f5m_ring_word(c, rb, cw)], slot[f5m_ring_word(c, rb, cw + 1)], sh);
        const float p0 = norm_inrange(byte_magic<0>(w), P.nk), p1 = norm_inrange(byte_magic<1>(w), P.nk);
        const float p2 = norm_inrange(byte_magic<2>(w), P.nk), p3 = norm_inrange(byte_magic<3>(w), P.nk);
        float a = __fmul_rn(P.wx[1], p1);
        a = __fmaf_rn(P.wx[0], p0, a); a = __fmaf_rn(P.wx[2], p2, a); a = __fmaf_rn(P.wx[3], p3, a);
        h[r] = a;
    }
    float v = __fmul_rn(P.wy[1], h[1]);
    v = __fmaf_rn(P.wy[0], h[0], v); v = __fmaf_rn(P.wy[2], h[2], v); v = __fmaf_rn(P.wy[3], h[3], v);
    const int o = trunc_i(__fmul_rn(v, P.factor));
    return WRAP ? (max(o, 0) & 0xFF) : min(max(o, 0), 255);
}

// 2^e as a float constant expression (normal range)
__host__ __device__ constexpr float f5m_pow2(int e) { return e == 0 ? 1.0f : e > 0 ? 2.0f * f5m_pow2(e - 1) : 0.5f * f5m_pow2(e + 1); }

template <int DST, bool WRAP, int WA, int WB, int WS>
__device__ __forceinline__ void fused5m_band(const Fused3Params &P, uint32_t *ring) {
    static_assert(WB <= 127 && WA >= -128, "weights must fit s8");
    constexpr int SH = 2 * WS;
    constexpr int BPP = dst_bpp(DST);
    constexpr bool SW = dst_swap(DST);
    // The vertical pass runs on the fma pipe in packed fp32 on EXACT integers.  The IMMA accumulator starts at BIAS, so
    // every horizontal result D = T + BIAS is a positive integer, whose bit pattern is the denormal float D * 2^-149:
    // no conversion instruction.  An FMA with the weight scaled by 2^120 takes it into the normal range; everything
    // downstream is (integer) * 2^-29 with |integer| < 2^24: exact in every operation.
    constexpr int BIAS = 2048;                         // > -min T = 2 |WA| 255
    constexpr float S29 = f5m_pow2(-29), K120 = f5m_pow2(120);
    constexpr float KWA = (float)WA * K120, KWB = (float)WB * K120;
    static_assert(2 * (-WA) * 255 < BIAS, "bias too small");
    const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    const long long fz = blockIdx.z;
    const int yo_begin = blockIdx.y * P.band;
    const int yo_end = min(yo_begin + P.band, P.dstH);
    const int H = P.src.h, HC = H >> 1;
    const int nchunks = P.src.w >> 2;
    // the lane's chunks: columns 4 q0 .. 4 q0 + 3 and the same 16 columns further right
    const int q0 = (int)blockIdx.x * 60 + 8 * g - 1 + t, q1 = q0 + 4;
    const bool edge = blockIdx.x == 0 || (int)blockIdx.x * 60 + 62 >= nchunks;       // warp-uniform: some chunk lies outside the frame
    const int rep0 = q0 < 0 ? 1 : q0 >= nchunks ? 2 : 0, rep1 = q1 >= nchunks ? 2 : 0;
    const int a0c = min(max(q0, 0), nchunks - 1), a1c = min(q1, nchunks - 1);

    const uint8_t *py0 = P.src.pl[0].p + fz * P.src.pl[0].bstride + (size_t)a0c * 4;
    const uint8_t *py1 = P.src.pl[0].p + fz * P.src.pl[0].bstride + (size_t)a1c * 4;
    const uint8_t *pc0 = P.src.pl[1].p + fz * P.src.pl[1].bstride + (size_t)a0c * 4;
    const uint8_t *pc1 = P.src.pl[1].p + fz * P.src.pl[1].bstride + (size_t)a1c * 4;
    const unsigned pitch_y = P.src.pl[0].pitch, pitch_c = P.src.pl[1].pitch;

    // band matrix fragments: B[k][n] = W[k - 2n - 3], this lane holds n = g, k = 4t.. and 16 + 4t..
    uint32_t bp0 = 0, bp1 = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int i0 = 4 * t + j - 2 * g - 3, i1 = i0 + 16;
        const int w0 = (i0 == 0 || i0 == 3) ? WA : (i0 == 1 || i0 == 2) ? WB : 0;
        const int w1 = (i1 == 0 || i1 == 3) ? WA : (i1 == 1 || i1 == 2) ? WB : 0;
        bp0 |= (uint32_t)(w0 & 0xFF) << (8 * j); bp1 |= (uint32_t)(w1 & 0xFF) << (8 * j);
    }

    // output ownership: piece m = columns xw_m, xw_m + 1 (relative to the warp) of MMA m
    const int xw0 = 16 * g + 2 * t, xw1 = xw0 + 8;
    const int xg = (int)blockIdx.x * 120;
    const bool own0 = xg + xw0 < P.dstW, own1 = xw1 < 120 && xg + xw1 < P.dstW;
    // pair k finishes output row k-1 and starts row k: pairs yo_begin-1 .. yo_end are consumed, the first two only
    // prime the accumulators; the store pointers address row k-1 while pair k is processed
    const int kfirst = yo_begin - 1, klast = yo_end, kstore = kfirst + 2;
    const unsigned pitch_d = P.dst.pl[0].pitch;
    uint8_t *pd = P.dst.pl[0].p + fz * P.dst.pl[0].bstride + ((long long)kfirst - 1) * (long long)pitch_d
                + (long long)(own0 ? xg + xw0 : 0) * BPP;
    // rgb24: the 6 bytes of a piece leave as a word and a half-word; which comes first depends on the alignment (t)
    const bool odd = t & 1;
    uint8_t *pd32 = pd + (odd ? 2 : 0), *pd16 = pd + (odd ? 0 : 4);
    const uint32_t sel32 = odd ? 0x5432u : 0x3210u, sel16 = odd ? 0x3210u : 0x7654u;

    unsigned ot, ob, oc;
    auto seek = [&](int kk) {
        ot = (unsigned)min(max(2 * kk, 0), H - 1) * pitch_y; ob = (unsigned)min(max(2 * kk + 1, 0), H - 1) * pitch_y;
        oc = (unsigned)min(max(kk, 0), HC - 1) * pitch_c;
    };
    auto load_here = [&](Row5 &R) {
        if (!edge) {                             // the second chunk is 16 bytes further right in every plane
            const uint8_t *at = py0 + ot, *ab = py0 + ob, *ac = pc0 + oc;
            R.yt[0] = ldg32(at); R.yt[1] = ldg32(at + 16);
            R.yb[0] = ldg32(ab); R.yb[1] = ldg32(ab + 16);
            R.c[0] = ldg32(ac); R.c[1] = ldg32(ac + 16);
        } else {
            R.yt[0] = ldg32(py0 + ot); R.yt[1] = ldg32(py1 + ot);
            R.yb[0] = ldg32(py0 + ob); R.yb[1] = ldg32(py1 + ob);
            R.c[0] = ldg32(pc0 + oc); R.c[1] = ldg32(pc1 + oc);
        }
    };
    // out-of-frame chunks replicate the frame's edge column; applied when a pair is consumed (two steps after its loads)
    auto replicate = [&](Row5 &R) {
        if (rep0 == 1) { R.yt[0] = prmt(R.yt[0], 0, 0x0000u); R.yb[0] = prmt(R.yb[0], 0, 0x0000u); R.c[0] = prmt(R.c[0], 0, 0x1010u); }
        if (rep0 == 2) { R.yt[0] = prmt(R.yt[0], 0, 0x3333u); R.yb[0] = prmt(R.yb[0], 0, 0x3333u); R.c[0] = prmt(R.c[0], 0, 0x3232u); }
        if (rep1 == 2) { R.yt[1] = prmt(R.yt[1], 0, 0x3333u); R.yb[1] = prmt(R.yb[1], 0, 0x3333u); R.c[1] = prmt(R.c[1], 0, 0x3232u); }
    };
    const unsigned sy = 2 * pitch_y;
    auto load_next = [&](Row5 &R, int kk) {      // loads pair kk >= 1 (the offsets are on it), then moves them to pair kk+1
        load_here(R);
        if (kk + 1 < HC) { ot += sy; ob += sy; oc += pitch_c; }
        else ot = ob;                            // pair HC: rows H-1, H-1, chroma row HC-1
    };

    // vertical state, scaled by 2^-29, one packed pair (output columns 2t, 2t+1) per (MMA, channel):
    //   Pp = P[k] - BIAS WA            P[k] = WB (e[k] + o[k]) + WA o[k-1]
    //   Ap = WA o[k] - 2 BIAS WB - BIAS WA
    // so that, with the biased horizontal results E = e + BIAS, O = o + BIAS (denormals, times 2^120 inside the FMA):
    //   N[k-1]  = WA E[k] + Pp[k-1]
    //   Pp[k]   = WB (E[k] + O[k]) + Ap[k-1]
    //   Ap[k]   = WA O[k] - (2 BIAS WA + 2 BIAS WB)
    constexpr float APC = (float)(-2 * BIAS * WA - 2 * BIAS * WB) * S29;
    f2 Pp[2][3], Ap[2][3];
    uint32_t N[2][3][2];
#pragma unroll
    for (int m = 0; m < 2; m++)
#pragma unroll
        for (int c = 0; c < 3; c++) { Pp[m][c] = 0ull; Ap[m][c] = 0ull; }

    int alpha_i = 0;
    if (dst_alpha(DST)) {   // the chain applied to the reference's constant intermediate alpha (scale_fused3.cuh)
        float ah = __fmul_rn(P.wx[1], 1.0f);
        ah = __fmaf_rn(P.wx[0], 1.0f, ah); ah = __fmaf_rn(P.wx[2], 1.0f, ah); ah = __fmaf_rn(P.wx[3], 1.0f, ah);
        float av = __fmul_rn(P.wy[1], ah);
        av = __fmaf_rn(P.wy[0], ah, av); av = __fmaf_rn(P.wy[2], ah, av); av = __fmaf_rn(P.wy[3], ah, av);
        alpha_i = trunc_i(__fmul_rn(av, P.factor));
    }
    int bias = BIAS;
    asm volatile("" : "+r"(bias));               // one register quad for every accumulator operand, not a literal per MMA

    // one row pair; returns the smallest |N| << (32 - SH) of the lane: 0 <=> some N = 0 (mod 2^SH)
    auto step = [&](Row5 &cur, int k, bool st0, bool st1, uint32_t *slot) -> uint32_t {
        Row5 now = cur;
        if (k + 2 <= klast) load_next(cur, k + 2);
        if (edge) replicate(now);
        uint32_t Jt[3][2], Jb[3][2];
        produce5m(now, P, Jt, Jb);
        uint4 *s4 = reinterpret_cast<uint4 *>(slot);          // the A operands of the first MMA of each channel, as they are
#pragma unroll
        for (int c = 0; c < 3; c++) s4[c] = make_uint4(Jt[c][0], Jb[c][0], Jt[c][1], Jb[c][1]);
        uint32_t umin = 0xFFFFFFFFu;
        int o[2][3][2];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const uint32_t ht = __shfl_down_sync(0xffffffffu, Jt[c][0], 4), hb = __shfl_down_sync(0xffffffffu, Jb[c][0], 4);
#pragma unroll
            for (int m = 0; m < 2; m++) {
                // second MMA: K order (b2 | b1) with the band matrix halves swapped, so that b1 keeps its registers
                const uint32_t a0 = m ? ht : Jt[c][0], a1 = m ? hb : Jb[c][0], a2 = Jt[c][1], a3 = Jb[c][1];
                int D[4];
                mma_u8s8(D, a0, a1, a2, a3, m ? bp1 : bp0, m ? bp0 : bp1, bias, bias, bias, bias);
                f2 E, O;                                     // (T + BIAS) 2^-149 of the top / bottom row, columns 2t, 2t+1
                asm("mov.b64 %0, {%1, %2};" : "=l"(E) : "r"(D[0]), "r"(D[1]));
                asm("mov.b64 %0, {%1, %2};" : "=l"(O) : "r"(D[2]), "r"(D[3]));
                const f2 n2 = fma2(E, bc(KWA), Pp[m][c]);                    // N[k-1] 2^-29
                Pp[m][c] = fma2(add2(E, O), bc(KWB), Ap[m][c]);
                Ap[m][c] = fma2(O, bc(KWA), bc(APC));
                int q0i, q1i, y0, y1;
                upki(mul2_rz(n2, bc(f5m_pow2(29 - SH - 126) * f5m_pow2(-23))), q0i, q1i);     // trunc(N / 2^SH) as denormal bits
                upki(mul2(n2, bc(f5m_pow2(29 - 126) * f5m_pow2(-23))), y0, y1);               // |N| (+ sign bit) as denormal bits
                N[m][c][0] = (uint32_t)y0; N[m][c][1] = (uint32_t)y1;
                umin = min(umin, min((uint32_t)y0 << (32 - SH), (uint32_t)y1 << (32 - SH)));
                o[m][c][0] = q0i; o[m][c][1] = q1i;
                if (WRAP) { o[m][c][0] = max(q0i, 0) & 0xFF; o[m][c][1] = max(q1i, 0) & 0xFF; }
            }
        }
#define CH(m, i, c) o[m][SW ? 2 - (c) : (c)][i]
#pragma unroll
        for (int m = 0; m < 2; m++) {
            if (m ? st1 : st0) {
                if (BPP == 3) {
                    const uint32_t lo = pack4_u8(CH(m, 0, 0), CH(m, 0, 1), CH(m, 0, 2), CH(m, 1, 0));
                    const uint32_t hi = pack2_u8(CH(m, 1, 1), CH(m, 1, 2), 0u);
                    stg32(pd32 + m * 24, prmt(lo, hi, sel32));
                    stg16(pd16 + m * 24, prmt(lo, hi, sel16));
                } else {
                    stg64(pd + m * 32, make_uint2(pack4_u8(CH(m, 0, 0), CH(m, 0, 1), CH(m, 0, 2), alpha_i),
                                                  pack4_u8(CH(m, 1, 0), CH(m, 1, 1), CH(m, 1, 2), alpha_i)));
                }
            }
        }
#undef CH
        return umin;
    };

    // The outputs of row k-1 with N = 0 (mod 2^SH), N > 0: recompute each with the float chain (few: every lane that has
    // some recomputes its own, one per pass), or report a dense step (many).  Entered by the whole warp after the
    // stores of the step.
    auto ambiguous = [&](bool st0, bool st1, int k) -> bool {
        uint32_t mask = 0;
#pragma unroll
        for (int m = 0; m < 2; m++)
#pragma unroll
            for (int c = 0; c < 3; c++)
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    const uint32_t n = N[m][c][i];            // sign-magnitude: negative values fail the first test
                    const bool f = (m ? st1 : st0) && (int)n > 0 && (n & (0u - n)) >= (1u << SH);      // lowest set bit >= 2^SH
                    mask |= (uint32_t)f << (m * 6 + c * 2 + i);
                }
        const int total = __reduce_add_sync(0xffffffffu, __popc(mask));
        if (total == 0) return false;
        if (total > GMATB_F5M_DENSE) return true;
        __syncwarp();                       // this step's ring stores and output stores are visible to the whole warp
        uint8_t *orow = P.dst.pl[0].p + fz * P.dst.pl[0].bstride + (long long)(k - 1) * (long long)pitch_d + (long long)xg * BPP;
        while (mask) {
            const int b = __ffs(mask) - 1;
            mask &= mask - 1;
            const int m = b / 6, c = (b - 6 * m) >> 1, i = b & 1;
            const int xw = 16 * g + 8 * m + 2 * t + i;
            orow[xw * BPP + (SW ? 2 - c : c)] = (uint8_t)fused5m_fix<WRAP>(P, ring, k, xw, c);
        }
        __syncwarp();
        return false;
    };

    Row5 A, B;
    int k = kfirst;
    seek(k); load_here(A);
    seek(k + 1); load_here(B);
    seek(k + 2);
    bool dense = false;
    uint32_t *lane_ring = ring + lane * 12;
#pragma unroll 1
    for (;;) {
        {
            const bool st0 = own0 && k >= kstore, st1 = own1 && k >= kstore;
            const uint32_t um = step(A, k, st0, st1, lane_ring + (k & 3) * (32 * 12));
            if (__any_sync(0xffffffffu, st0 && um == 0u) && ambiguous(st0, st1, k)) { dense = true; break; }
            if (BPP == 3) { pd32 += pitch_d; pd16 += pitch_d; } else pd += pitch_d;
            if (++k > klast) break;
        }
        {
            const bool st0 = own0 && k >= kstore, st1 = own1 && k >= kstore;
            const uint32_t um = step(B, k, st0, st1, lane_ring + (k & 3) * (32 * 12));
            if (__any_sync(0xffffffffu, st0 && um == 0u) && ambiguous(st0, st1, k)) { dense = true; break; }
            if (BPP == 3) { pd32 += pitch_d; pd16 += pitch_d; } else pd += pitch_d;
            if (++k > klast) break;
        }
    }
    // a dense step at pair k: output rows k-1 .. yo_end-1 are (re)done by the float loop (its edge-strip form is
    // correct for every warp; a v3 warp owns the same 240 source columns)
    if (dense) fused3_band<L_NV12, 8, DST, false, WRAP, true>(P, blockIdx.x, fz, k - 1, yo_end);
}

template <int DST, bool WRAP, int WA, int WB, int WS, int MINB>
__global__ void __launch_bounds__(32, MINB) fused_csc_scale2_mma_kernel(const __grid_constant__ Fused3Params P) {
    __shared__ __align__(16) uint32_t ring[GMATB_F4I_RING];
    fused5m_band<DST, WRAP, WA, WB, WS>(P, ring);
}

}  // namespace gmatb
