// scale_fused3.cuh -- v3 of the headline kernel (scale_fused.cuh): YUV 4:2:0 / packed RGB ->
// packed RGB colour conversion FUSED with an exact 2:1 four-tap resample.  Identical
// arithmetic (every lane performs the reference's roundings in the reference's order:
// csc_core.cuh, resample_core.cuh), leaner instruction stream.
//
// Cost model measured on B200 (tools/probe2.cu -> profiles/r1c_probe2_operand_forms.txt):
// scalar FP32 ops issue at 1/clk/SMSP; packed FFMA2/FADD2/FMUL2, IMAD, IDP4A, PRMT, SEL, MOV,
// IADD3, LOP3 and I2IP at 1 per 2 clk; SHFL at 1 per 4 clk; both fused kernels run at
// ~0.82 x (2 x two-cycle instructions + scalar-FP32 instructions) cycles.  A packed op
// therefore buys no pipe time over two scalar ones, and the ~40 % of v1's stream that was not
// arithmetic (halo-column loads + conversion executed by every warp for lanes 0/31, SEL/MOV
// traffic around the shuffles, per-iteration 64-bit address arithmetic with row clamps,
// rematerialised parameters under a 96-register cap) cost as much as the arithmetic.  v3:
//   * strips overlap by one lane on each side: lanes 0 and 31 of a warp convert a strip only
//     to hand its edge column to lanes 1 / 30 (30 owning lanes = 240 source columns per warp,
//     4K = 16 warps per row).  Every lane's halo is simply the shuffle result: no halo loads,
//     no selects.  At the frame's left/right edge the provider lane is the out-of-frame
//     strip: it loads the edge strip and byte-replicates the edge column in the raw words;
//   * one per-lane pointer per plane + warp-uniform 32-bit row offsets that advance by
//     constant steps (clamped rows exist only at the first/last row pair of the frame);
//   * row-pair registers ping-pong between two explicitly unrolled loop bodies (no copies);
//   * constants arrive as 8-byte aligned pairs that FFMA2/FMUL2 take straight from uniform
//     registers.
// TAPS2 (weights exactly {0, .5, .5, 0} on both axes: default bicubic at 2:1): no halo, 32 owning
// lanes, and each pass is one rounding of a plain sum (see step()).
#pragma once
#include "scale_fused.cuh"

namespace gmatb {

struct alignas(8) Fused3Params {
    Img src, dst;
    float cm45[2];     // (m4, m5)   chroma -> G terms
    float cm72[2];     // (m7, m2)   U -> B term, V -> R term
    float m0, m1, m3, m6;   // luma gains of the three rows; m1 == 0 (run-time zero addend, csc_core.cuh)
    float wx[4], wy[4];
    NormK nk;
    float factor;      // 255 or 65535
    float factor_q;    // factor / 4 (exact), TAPS2
    float factor_s, factor_qs;   // the same times 2^norm_shift: yuv sources run the chains on scaled samples (quant_norm2d)
    int band;
    int dstW, dstH;
};

template <int L, int SBITS> struct Raw3;
template <int L> struct Raw3<L, 8>  { uint2 yt, yb, c0; };
template <int L> struct Raw3<L, 16> { uint4 yt, yb, c0; };
template <> struct Raw3<L_RGB3, 8>  { uint2 t[3], b[3]; };

// out-of-frame provider strips: make column 7 (left provider) / column 0 (right provider) a copy of
// the frame's edge column, in the raw words
template <int L, int SBITS>
__device__ __forceinline__ void edge_replicate(Raw3<L, SBITS> &R, bool lrep, bool rrep) {
    if constexpr (L == L_RGB3) {
        if (lrep) {   // pixel 7 (bytes 21..23) := pixel 0 (bytes 0..2)
            R.t[2].y = prmt(R.t[2].y, R.t[0].x, 0x6540u); R.b[2].y = prmt(R.b[2].y, R.b[0].x, 0x6540u);
        }
        if (rrep) {   // pixel 0 := pixel 7
            R.t[0].x = prmt(R.t[0].x, R.t[2].y, 0x3765u); R.b[0].x = prmt(R.b[0].x, R.b[2].y, 0x3765u);
        }
    } else if constexpr (SBITS == 8) {
        if (lrep) {
            R.yt.y = prmt(R.yt.y, R.yt.x, 0x4210u); R.yb.y = prmt(R.yb.y, R.yb.x, 0x4210u);
            if (L == L_NV12) R.c0.y = prmt(R.c0.y, R.c0.x, 0x5410u);
            else { R.c0.x = prmt(R.c0.x, R.c0.x, 0x0210u); R.c0.y = prmt(R.c0.y, R.c0.y, 0x0210u); }
        }
        if (rrep) {
            R.yt.x = prmt(R.yt.x, R.yt.y, 0x3217u); R.yb.x = prmt(R.yb.x, R.yb.y, 0x3217u);
            if (L == L_NV12) R.c0.x = prmt(R.c0.x, R.c0.y, 0x3276u);
            else { R.c0.x = prmt(R.c0.x, R.c0.x, 0x3213u); R.c0.y = prmt(R.c0.y, R.c0.y, 0x3213u); }
        }
    } else {
        if (lrep) {
            R.yt.w = prmt(R.yt.w, R.yt.x, 0x5410u); R.yb.w = prmt(R.yb.w, R.yb.x, 0x5410u);
            if (L == L_NV12) R.c0.w = R.c0.x;
            else { R.c0.y = prmt(R.c0.y, R.c0.x, 0x5410u); R.c0.w = prmt(R.c0.w, R.c0.z, 0x5410u); }
        }
        if (rrep) {
            R.yt.x = prmt(R.yt.x, R.yt.w, 0x3276u); R.yb.x = prmt(R.yb.x, R.yb.w, 0x3276u);
            if (L == L_NV12) R.c0.x = R.c0.w;
            else { R.c0.x = prmt(R.c0.x, R.c0.y, 0x3276u); R.c0.z = prmt(R.c0.z, R.c0.w, 0x3276u); }
        }
    }
}

// the 8 (top,bottom) column pairs of one row pair as normalised samples
template <int L, int SBITS>
__device__ __forceinline__ void produce3(const Raw3<L, SBITS> &R, const Fused3Params &P, f2 (&C)[8][3]) {
    if constexpr (L == L_RGB3) {
        RawRowRGB rr;
#pragma unroll
        for (int i = 0; i < 3; i++) { rr.t[i] = R.t[i]; rr.b[i] = R.b[i]; }
        rgb_column<0>(rr, P.nk, C[0]); rgb_column<1>(rr, P.nk, C[1]); rgb_column<2>(rr, P.nk, C[2]); rgb_column<3>(rr, P.nk, C[3]);
        rgb_column<4>(rr, P.nk, C[4]); rgb_column<5>(rr, P.nk, C[5]); rgb_column<6>(rr, P.nk, C[6]); rgb_column<7>(rr, P.nk, C[7]);
    } else {
        constexpr bool FMAFORM = SBITS == 16;
        constexpr float CB = -(GMATB_MAGIC + (SBITS == 8 ? 128.f : 32768.f));
        constexpr float YB = -(GMATB_MAGIC + (SBITS == 8 ? 16.f : 4096.f));
        RawRow<SBITS> rr; rr.yt = R.yt; rr.yb = R.yb; rr.c0 = R.c0;
        float yt[8], yb[8], um[4], vm[4];
        fused_unpack<L>(rr, yt, yb, um, vm);
        const f2 k45 = *reinterpret_cast<const f2 *>(P.cm45), k72 = *reinterpret_cast<const f2 *>(P.cm72);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const f2 uv = add2(pk(um[j], vm[j]), bc(CB));
            float fu, fv, t1g, t2g, t1b, t2r;
            upk(uv, fu, fv);
            if (FMAFORM) {      // r = FFMA(fv, mC, FFMA(fy, mA, FMUL(fu, mB)))  (P010/P016 kernels of the reference)
                t1g = __fmul_rn(fu, P.cm45[0]); t1b = __fmul_rn(fu, P.cm72[0]); t2g = t2r = 0.f;
            } else {            // r = FADD(FFMA(fy, mA, FMUL(fu, mB)), FMUL(fv, mC))
                upk(mul2(uv, k45), t1g, t2g);
                upk(mul2(uv, k72), t1b, t2r);
            }
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int col = 2 * j + h;
                const f2 fy2 = add2(pk(yt[col], yb[col]), bc(YB));
                f2 xr = fma2(fy2, bc(P.m0), bc(P.m1));      // m1 is a run-time 0.0f: RN(fy*m0), as FFMA(fy, m0, +-0)
                f2 xg = fma2(fy2, bc(P.m3), bc(t1g));
                f2 xb = fma2(fy2, bc(P.m6), bc(t1b));
                if (FMAFORM) { xr = fma2(bc(fv), bc(P.cm72[1]), xr); xg = fma2(bc(fv), bc(P.cm45[1]), xg); }
                else { xr = add2(xr, bc(t2r)); xg = add2(xg, bc(t2g)); }
                C[col][0] = quant_norm2d(xr, P.nk); C[col][1] = quant_norm2d(xg, P.nk); C[col][2] = quant_norm2d(xb, P.nk);
            }
        }
    }
}

template <int L, int SBITS, int DST, bool TAPS2, bool WRAP>
struct Fused3 {
    typedef Raw3<L, SBITS> Row;
    static constexpr int SB = SBITS / 8;

    // one row pair k: convert `cur`, finish output row k-1, start output row k.  `refill(cur)` issues the
    // loads of the row pair two steps ahead into the same buffer as soon as its raw words are consumed
    // (measured DRAM latency under load is ~2 steps of a warp's instruction stream at 4 warps/SMSP).
    template <typename Refill>
    static __device__ __forceinline__ void step(const Fused3Params &P, Row &cur, f2 (&acc)[2][3], f2 (&hb_prev)[2][3],
                                                bool store, uint8_t *pd, int alpha_i, Refill refill) {
        f2 C[8][3];
        if (SBITS == 8) {
            const Row now = cur;
            refill(cur);
            produce3<L, SBITS>(now, P, C);
        } else {             // 16-bit rows are 12 registers each: refill once the raw words are consumed (no third copy)
            produce3<L, SBITS>(cur, P, C);
            refill(cur);
        }
        f2 PL[3] = {0ull, 0ull, 0ull}, PR[3] = {0ull, 0ull, 0ull};
        if (!TAPS2) {
#pragma unroll
            for (int c = 0; c < 3; c++) { PL[c] = shfl_up2(C[7][c]); PR[c] = shfl_dn2(C[0][c]); }
        }
        // horizontal results of the row pair, re-paired for the vertical pass: HT / HB = the top / bottom row of two
        // neighbouring output columns (two register moves per pair: the vertical chain, the factor multiply and the
        // truncation then run packed like everything before them -- a homogeneous FFMA2 stream, see quant_norm2d)
        f2 HT[2][3], HB[2][3];
        float s2[4][3];                       // TAPS2: top + bottom of each output column (no re-pairing: measured 1.7 % faster scalar)
#pragma unroll
        for (int xp = 0; xp < 2; xp++)
#pragma unroll
            for (int c = 0; c < 3; c++) {
                float t0, b0, t1, b1;
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    const int xo = 2 * xp + q;
                    const f2 p0 = xo == 0 ? PL[c] : C[2 * xo - 1][c];
                    const f2 p3 = xo == 3 ? PR[c] : C[2 * xo + 2][c];
                    // TAPS2: weights are exactly {0, .5, .5, 0}: FFMA(.5, p2, FMUL(.5, p1)) == RN(p1 + p2) / 2 (scaling by a
                    // power of two commutes with rounding), so the halvings are deferred to the final factor
                    const f2 h = TAPS2 ? add2(C[2 * xo][c], C[2 * xo + 1][c]) : hpass<false>(P.wx, p0, C[2 * xo][c], C[2 * xo + 1][c], p3);
                    if (q == 0) upk(h, t0, b0); else upk(h, t1, b1);
                }
                if (TAPS2) { s2[2 * xp][c] = __fadd_rn(b0, t0); s2[2 * xp + 1][c] = __fadd_rn(b1, t1); }
                else { HT[xp][c] = pk(t0, t1); HB[xp][c] = pk(b0, b1); }
            }
        if (store) {
            int o[4][3];
            const float fac = L == L_RGB3 ? (TAPS2 ? P.factor_q : P.factor) : (TAPS2 ? P.factor_qs : P.factor_s);
#pragma unroll
            for (int xp = 0; xp < 2; xp++)
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    // TAPS2: the output row taps only this pair: RN(2h_top + 2h_bottom) = 4 x the reference's vertical
                    // result, exactly, written in the same step (pd addresses row k)
                    if (TAPS2) {
                        o[2 * xp][c] = trunc_i(__fmul_rn(s2[2 * xp][c], fac)); o[2 * xp + 1][c] = trunc_i(__fmul_rn(s2[2 * xp + 1][c], fac));
                    } else {
                        const f2 v = fma2(bc(P.wy[3]), HT[xp][c], acc[xp][c]);
                        upki(mul2_rz(mul2(v, bc(fac)), bc(GMATB_TWO_M149)), o[2 * xp][c], o[2 * xp + 1][c]);
                    }
                    if (WRAP) {
                        o[2 * xp][c] = max(o[2 * xp][c], 0) & (SBITS == 8 ? 0xFF : 0xFFFF);
                        o[2 * xp + 1][c] = max(o[2 * xp + 1][c], 0) & (SBITS == 8 ? 0xFF : 0xFFFF);
                    }
                }
            constexpr bool SW = dst_swap(DST);
#define CH(i, c) o[i][SW ? 2 - (c) : (c)]
            if (DST == D_RGB24 || DST == D_BGR24) {
                stg32(pd,     pack4_u8(CH(0, 0), CH(0, 1), CH(0, 2), CH(1, 0)));
                stg32(pd + 4, pack4_u8(CH(1, 1), CH(1, 2), CH(2, 0), CH(2, 1)));
                stg32(pd + 8, pack4_u8(CH(2, 2), CH(3, 0), CH(3, 1), CH(3, 2)));
            } else if (DST == D_RGBA || DST == D_BGRA) {
                stg128(pd, make_uint4(pack4_u8(CH(0, 0), CH(0, 1), CH(0, 2), alpha_i), pack4_u8(CH(1, 0), CH(1, 1), CH(1, 2), alpha_i),
                                      pack4_u8(CH(2, 0), CH(2, 1), CH(2, 2), alpha_i), pack4_u8(CH(3, 0), CH(3, 1), CH(3, 2), alpha_i)));
            } else if (DST == D_RGB48 || DST == D_BGR48) {
                stg64(pd,      make_uint2(pack2_u16(CH(0, 0), CH(0, 1)), pack2_u16(CH(0, 2), CH(1, 0))));
                stg64(pd + 8,  make_uint2(pack2_u16(CH(1, 1), CH(1, 2)), pack2_u16(CH(2, 0), CH(2, 1))));
                stg64(pd + 16, make_uint2(pack2_u16(CH(2, 2), CH(3, 0)), pack2_u16(CH(3, 1), CH(3, 2))));
            } else {
                stg128(pd, make_uint4(pack2_u16(CH(0, 0), CH(0, 1)), pack2_u16(CH(0, 2), alpha_i),
                                      pack2_u16(CH(1, 0), CH(1, 1)), pack2_u16(CH(1, 2), alpha_i)));
                stg128(pd + 16, make_uint4(pack2_u16(CH(2, 0), CH(2, 1)), pack2_u16(CH(2, 2), alpha_i),
                                           pack2_u16(CH(3, 0), CH(3, 1)), pack2_u16(CH(3, 2), alpha_i)));
            }
#undef CH
        }
#pragma unroll
        for (int xp = 0; xp < 2; xp++)
#pragma unroll
            for (int c = 0; c < 3; c++) {
                if (TAPS2) continue;
                f2 t = mul2(bc(P.wy[1]), HT[xp][c]);
                t = fma2(bc(P.wy[0]), hb_prev[xp][c], t);
                t = fma2(bc(P.wy[2]), HB[xp][c], t);
                acc[xp][c] = t;
                hb_prev[xp][c] = HB[xp][c];
            }
    }
};

// The whole band loop; EDGE = this warp holds an out-of-frame provider strip (first / last warp of a row).
// (bx, fz) = strip block and frame of this warp, [yo_begin, yo_end) = its output rows (the integer kernel of
// scale_fused4i.cuh hands the rest of a band over to this loop with its own yo_begin).
template <int L, int SBITS, int DST, bool TAPS2, bool WRAP, bool EDGE>
__device__ __forceinline__ void fused3_band(const Fused3Params &P, const int bx, const long long fz, const int yo_begin, const int yo_end) {
    typedef Fused3<L, SBITS, DST, TAPS2, WRAP> F;
    typedef typename F::Row Row;
    constexpr int OWN = TAPS2 ? 32 : 30;
    constexpr int SB = SBITS / 8;
    constexpr int SPP = L == L_RGB3 ? 3 : SB;          // source bytes per pixel in plane 0
    const int lane = threadIdx.x;
    const int nstrips = P.src.w >> 3;
    const int strip = bx * OWN + lane - (TAPS2 ? 0 : 1);
    const bool owner = (TAPS2 || (lane >= 1 && lane <= 30)) && strip < nstrips;
    const int sl = min(max(strip, 0), nstrips - 1);
    const int H = P.src.h, HC = H >> 1;

    const uint8_t *py = P.src.pl[0].p + fz * P.src.pl[0].bstride + (size_t)sl * (8 * SPP);
    const uint8_t *pu = py, *pv = py;
    if (L == L_NV12) pu = P.src.pl[1].p + fz * P.src.pl[1].bstride + (size_t)sl * (8 * SB);
    if (L == L_I420) {
        pu = P.src.pl[1].p + fz * P.src.pl[1].bstride + (size_t)sl * (4 * SB);
        pv = P.src.pl[2].p + fz * P.src.pl[2].bstride + (size_t)sl * (4 * SB);
    }
    const unsigned pitch_y = P.src.pl[0].pitch, pitch_c = P.src.pl[1].pitch, pitch_c2 = P.src.pl[2].pitch;
    // 4-tap: pair k finishes output row k-1 and starts row k: pairs yo_begin-1 .. yo_end are consumed, the first two only
    // prime the accumulators, pd addresses row k-1 while pair k is processed (never dereferenced before row yo_begin).
    // TAPS2: output row k taps pair k only: pairs yo_begin .. yo_end-1, pd addresses row k.
    const int kfirst = TAPS2 ? yo_begin : yo_begin - 1, klast = TAPS2 ? yo_end - 1 : yo_end;
    const int kstore = TAPS2 ? kfirst : kfirst + 2;          // first pair whose step writes a row
    uint8_t *pd = P.dst.pl[0].p + fz * P.dst.pl[0].bstride + ((long long)kfirst - (TAPS2 ? 0 : 1)) * (long long)P.dst.pl[0].pitch
                + (long long)(owner ? strip : 0) * (4 * dst_bpp(DST));
    const unsigned pitch_d = P.dst.pl[0].pitch;
    const bool lrep = EDGE && strip < 0, rrep = EDGE && strip >= nstrips;

    auto load_at = [&](Row &R, unsigned ot, unsigned ob, unsigned oc, unsigned oc2) {
        if constexpr (L == L_RGB3) {
#pragma unroll
            for (int i = 0; i < 3; i++) { R.t[i] = ldg64(py + ot + 8 * i); R.b[i] = ldg64(py + ob + 8 * i); }
        } else if constexpr (SBITS == 8) {
            R.yt = ldg64(py + ot); R.yb = ldg64(py + ob);
            if (L == L_NV12) R.c0 = ldg64(pu + oc);
            else { R.c0.x = ldg32(pu + oc); R.c0.y = ldg32(pv + oc2); }
        } else {
            R.yt = ldg128(py + ot); R.yb = ldg128(py + ob);
            if (L == L_NV12) R.c0 = ldg128(pu + oc);
            else { const uint2 u = ldg64(pu + oc), v = ldg64(pv + oc2); R.c0 = make_uint4(u.x, u.y, v.x, v.y); }
        }
        if (EDGE) edge_replicate<L, SBITS>(R, lrep, rrep);
    };
    // any row pair, rows clamped to the frame (pairs -1 and HC replicate the first / last row)
    auto load_clamped = [&](int k, Row &R) {
        const unsigned rt = (unsigned)min(max(2 * k, 0), H - 1), rb = (unsigned)min(max(2 * k + 1, 0), H - 1);
        const unsigned rc = (unsigned)min(max(k, 0), HC - 1);
        load_at(R, rt * pitch_y, rb * pitch_y, rc * pitch_c, rc * pitch_c2);
    };

    f2 hb_prev[2][3], acc[2][3];          // vertical state of the 4 output columns, two columns per register pair
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int c = 0; c < 3; c++) { hb_prev[i][c] = 0ull; acc[i][c] = 0ull; }

    // constant alpha of 4-channel outputs: the chain applied to a constant 1.0 image
    int alpha_i = 0;
    if (dst_alpha(DST)) {
        // the reference's intermediate alpha is 255 in either depth (yuv2rgb_cuda.cu:89)
        const float one = SBITS == 8 ? 1.0f : 255.0f / 65535.0f;
        float ah = __fmul_rn(P.wx[1], one);
        ah = __fmaf_rn(P.wx[0], one, ah); ah = __fmaf_rn(P.wx[2], one, ah); ah = __fmaf_rn(P.wx[3], one, ah);
        float av = __fmul_rn(P.wy[1], ah);
        av = __fmaf_rn(P.wy[0], ah, av); av = __fmaf_rn(P.wy[2], ah, av); av = __fmaf_rn(P.wy[3], ah, av);
        alpha_i = trunc_i(__fmul_rn(av, P.factor));
    }

    // A holds pair k, B pair k+1; each step refills its own buffer with the pair two steps ahead.
    Row A, B;
    int k = kfirst;
    load_clamped(k, A);
    load_clamped(k + 1, B);
    // generic trip (rolled; the first pair and the last few of the band): clamped rows, buffers swapped by copy
    auto slow_trip = [&]() {
        F::step(P, A, acc, hb_prev, owner && k >= kstore, pd, alpha_i,
                [&](Row &R) { if (k + 2 <= klast) load_clamped(k + 2, R); });
        const Row t = A; A = B; B = t;
        k++; pd += pitch_d;
    };
    // warp-uniform offsets of the pair the steady-state loop loads next (k+2); no clamps there
    unsigned ot = 0, ob = 0, oc = 0, oc2 = 0;
    const unsigned sy = 2 * pitch_y;
    auto refill = [&](Row &R) { load_at(R, ot, ob, oc, oc2); ot += sy; ob += sy; oc += pitch_c; oc2 += pitch_c2; };
#pragma unroll 1
    while (k <= klast) {
        if (k >= yo_begin && k + 3 <= min(klast, HC - 1)) {
            // steady state: pairs k+2 and k+3 are interior, A and B ping-pong
            ot = (unsigned)(2 * k + 4) * pitch_y; ob = ot + pitch_y; oc = (unsigned)(k + 2) * pitch_c; oc2 = (unsigned)(k + 2) * pitch_c2;
#pragma unroll 1
            do {
                F::step(P, A, acc, hb_prev, owner && k >= kstore, pd, alpha_i, refill);
                pd += pitch_d;
                F::step(P, B, acc, hb_prev, owner, pd, alpha_i, refill);
                pd += pitch_d;
                k += 2;
            } while (k + 3 <= min(klast, HC - 1));
        } else {
            slow_trip();         // the first pair and the last three or four of the band (one copy of the code)
        }
    }
}

template <int L, int SBITS, int DST, bool TAPS2, bool WRAP, int MINB>
__global__ void __launch_bounds__(32, MINB) fused_csc_scale2_v3_kernel(const Fused3Params P) {
    constexpr int OWN = TAPS2 ? 32 : 30;
    const bool edge_warp = !TAPS2 && (blockIdx.x == 0 || (int)(blockIdx.x + 1) * OWN >= (P.src.w >> 3));
    const int yo_begin = blockIdx.y * P.band, yo_end = min(yo_begin + P.band, P.dstH);
    if (edge_warp) fused3_band<L, SBITS, DST, TAPS2, WRAP, true>(P, blockIdx.x, blockIdx.z, yo_begin, yo_end);
    else           fused3_band<L, SBITS, DST, TAPS2, WRAP, false>(P, blockIdx.x, blockIdx.z, yo_begin, yo_end);
}

}  // namespace gmatb
