// scale_fused4i.cuh -- the headline kernel, exact-integer form: 8-bit YUV 4:2:0 -> packed RGB colour
// conversion FUSED with an exact 2:1 four-tap resample whose weights are small dyadic rationals,
// (WA, WB, WB, WA) / 2^WS on both axes.  That is R-B bicubic at exactly 2:1 (fx = fy = 0.5) with
// param0 = 0.75 (A = -0.75: (-3, 19, 19, -3)/32, the BASELINE C2 headline), 0.5 ((-1, 9, 9, -1)/16) and
// 1.0 ((-1, 5, 5, -1)/8).  Same results as scale_fused3.cuh, bit for bit, at ~2/3 of its instructions.
//
// Why integers are exact here.  The reference's resize stage (vf_scale_cuda.cu:1040-1074, restated in
// resample_core.cuh) reads the quantised intermediate j in 0..255 as p = RN(j/255), runs two 4-tap float
// chains and stores trunc(255 v).  With dyadic weights the exact value of 255 v is N / 2^(2 WS),
//     N = sum_y sum_x W[y] W[x] j[y][x]        (an integer, |N| < 2^19),
// and the float chain carries at most 11.4 half-ulp roundings of values < 2: |255 v_float - N/2^(2WS)| <
// 1.9e-4 (measured maximum over 2e8 neighbourhoods: 6.1e-5; tests/test_oracle.py).  So whenever N is NOT a
// multiple of 2^(2WS) the exact value is at least 2^-(2WS) >= 9.8e-4 away from an integer and
//     trunc(255 v_float) = N >> 2 WS.
// When N IS a positive multiple of 2^(2WS) the float chain lands on either side of the integer and only
// the chain itself can tell: those outputs (1/1024 of them on noise, all of them on flat areas) are
//   * recomputed on the spot with the float chain (fused4i_fix) from the quantised bytes of the last four
//     steps, which every lane keeps in a shared-memory ring, when a warp step has few of them;
//   * handed to the float kernel's band loop (fused3_band) for the rest of the band when a step has many
//     (flat or synthetic content: the integer form has no advantage there).
// N = 0 needs no care (trunc gives 0 on both sides), nor does N > 255 * 2^(2WS) when the store saturates.
//
// Work decomposition: as v3 (one warp per CTA, a lane owns 8 source columns and walks down the frame one
// row pair per step; strips overlap by one lane per side).  Per step and lane:
//   CSC of 8x2 pixels on FFMA2 (the reference's chain, csc_core.cuh), truncation by RZ multiply, saturating
//   I2IP packs into planar bytes (4 columns per register: the quantised intermediate image itself);
//   halo bytes from the neighbouring lanes (one SHFL per register, no conversion);
//   horizontal pass = one IDP4A per (row, output, channel) on a PRMT window of 4 bytes; the bottom row's
//   result seeds the top row's accumulator (Q = ht + hb), the vertical pass is one IMAD and one more IDP4A
//   with the weights pre-multiplied by WA:  N[k-1] = P[k-1] + WA ht[k],  P[k] = WB Q[k] + WA hb[k-1].
#pragma once
#include "scale_fused3.cuh"

namespace gmatb {

#define GMATB_F4I_DENSE 8    /* more ambiguous outputs than this in one warp step: the band continues in float */
#define GMATB_F4I_RING  (4 * 32 * 12)   /* words: the packed rows of the last 4 steps of every lane */

// d = c + sum_i a.u8[i] * b.s8[i]
__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c) {
    int d; asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d;
}
__host__ __device__ constexpr uint32_t s8x4(int b0, int b1, int b2, int b3) {
    return (uint32_t)(b0 & 0xFF) | ((uint32_t)(b1 & 0xFF) << 8) | ((uint32_t)(b2 & 0xFF) << 16) | ((uint32_t)(b3 & 0xFF) << 24);
}

// One ambiguous output -- channel c of output column xo (0..3) of lane `src_lane`, output row k-1, found while
// pair k is processed -- recomputed with the float chain (the operations of scale_fused3.cuh /
// resample_core.cuh) from the quantised bytes the last steps left in the shared-memory ring.  The whole warp
// takes part: lane r (0..3) does the horizontal chain of source row 2(k-1)-1+r (`row` = that row's words in the
// ring: b[k-2], t[k-1], b[k-1], t[k]), then everybody the vertical one.  Returns the output value.
template <bool WRAP>
__device__ __forceinline__ int fused4i_fix(const Fused3Params &P, const uint32_t *row, int src_lane, int xo, int c) {
    // the window of 4 bytes starts at byte 3 + 2 xo of (left lane's columns 4..7 | own 0..3 | own 4..7 | right lane's 0..3)
    const int a = src_lane * 12 + 2 * c;
    const int olo = xo == 0 ? a - 11 : xo == 3 ? a + 1 : a;
    const int ohi = xo == 0 ? a : xo == 3 ? a + 12 : a + 1;
    const uint32_t w = __funnelshift_r(row[olo], row[ohi], (xo & 1) ? 8 : 24);
    const float p0 = norm_inrange(byte_magic<0>(w), P.nk), p1 = norm_inrange(byte_magic<1>(w), P.nk);
    const float p2 = norm_inrange(byte_magic<2>(w), P.nk), p3 = norm_inrange(byte_magic<3>(w), P.nk);
    float h = __fmul_rn(P.wx[1], p1);
    h = __fmaf_rn(P.wx[0], p0, h); h = __fmaf_rn(P.wx[2], p2, h); h = __fmaf_rn(P.wx[3], p3, h);
    const float h0 = __shfl_sync(0xffffffffu, h, 0), h1 = __shfl_sync(0xffffffffu, h, 1);
    const float h2 = __shfl_sync(0xffffffffu, h, 2), h3 = __shfl_sync(0xffffffffu, h, 3);
    float v = __fmul_rn(P.wy[1], h1);
    v = __fmaf_rn(P.wy[0], h0, v); v = __fmaf_rn(P.wy[2], h2, v); v = __fmaf_rn(P.wy[3], h3, v);
    const int o = trunc_i(__fmul_rn(v, P.factor));
    return WRAP ? (max(o, 0) & 0xFF) : min(max(o, 0), 255);
}

// 8x2 pixels -> the quantised intermediate image as planar bytes: R?[c][0] = columns 0..3, R?[c][1] = 4..7
template <int L>
__device__ __forceinline__ void produce4i(const Raw3<L, 8> &R, const Fused3Params &P, uint32_t (&Rt)[3][2], uint32_t (&Rb)[3][2]) {
    constexpr float CB = -(GMATB_MAGIC + 128.f), YB = -(GMATB_MAGIC + 16.f);
    RawRow<8> rr; rr.yt = R.yt; rr.yb = R.yb; rr.c0 = R.c0;
    float yt[8], yb[8], um[4], vm[4];
    fused_unpack<L>(rr, yt, yb, um, vm);
    const f2 k45 = *reinterpret_cast<const f2 *>(P.cm45), k72 = *reinterpret_cast<const f2 *>(P.cm72);
    const f2 z = bc(GMATB_TWO_M149);
    int it[8][3], ib[8][3];
#pragma unroll
    for (int j = 0; j < 4; j++) {
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
        Rt[c][0] = pack4_u8(it[0][c], it[1][c], it[2][c], it[3][c]); Rt[c][1] = pack4_u8(it[4][c], it[5][c], it[6][c], it[7][c]);
        Rb[c][0] = pack4_u8(ib[0][c], ib[1][c], ib[2][c], ib[3][c]); Rb[c][1] = pack4_u8(ib[4][c], ib[5][c], ib[6][c], ib[7][c]);
    }
}

template <int L, int DST, bool WRAP, int WA, int WB, int WS>
struct Fused4i {
    typedef Raw3<L, 8> Row;
    static constexpr uint32_t W4 = s8x4(WA, WB, WB, WA);
    static constexpr uint32_t WA4 = s8x4(WA * WA, WA * WB, WA * WB, WA * WA);
    static constexpr int SH = 2 * WS;
    static_assert(WA * WB >= -128 && WA * WB <= 127 && WB <= 127 && WA >= -128, "weights must fit s8");

    // One row pair k: finishes output row k-1 (N), starts row k (Pacc, Ta); `slot` = this lane's 12 words of ring
    // slot k & 3.  Returns the smallest (N << (32 - SH)) + index of the lane: < 12 <=> some output of this lane
    // has N = 0 (mod 2^SH).
    template <typename Refill>
    static __device__ __forceinline__ uint32_t step(const Fused3Params &P, Row &cur, int (&Pacc)[4][3], int (&Ta)[4][3],
                                                    int (&N)[4][3], bool store, uint8_t *pd, int alpha_i, uint32_t *slot, Refill refill) {
        const Row now = cur;
        refill(cur);
        uint32_t Rt[3][2], Rb[3][2];
        produce4i<L>(now, P, Rt, Rb);
        uint4 *s4 = reinterpret_cast<uint4 *>(slot);
        s4[0] = make_uint4(Rt[0][0], Rt[0][1], Rt[1][0], Rt[1][1]);
        s4[1] = make_uint4(Rt[2][0], Rt[2][1], Rb[0][0], Rb[0][1]);
        s4[2] = make_uint4(Rb[1][0], Rb[1][1], Rb[2][0], Rb[2][1]);
        uint32_t umin = 0xFFFFFFFFu;
        int o[4][3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const uint32_t hlt = __shfl_up_sync(0xffffffffu, Rt[c][1], 1), hrt = __shfl_down_sync(0xffffffffu, Rt[c][0], 1);
            const uint32_t hlb = __shfl_up_sync(0xffffffffu, Rb[c][1], 1), hrb = __shfl_down_sync(0xffffffffu, Rb[c][0], 1);
            uint32_t wt[4], wb[4];
            wt[0] = prmt(hlt, Rt[c][0], 0x6543u); wt[1] = prmt(Rt[c][0], Rt[c][1], 0x4321u);
            wt[2] = prmt(Rt[c][0], Rt[c][1], 0x6543u); wt[3] = prmt(Rt[c][1], hrt, 0x4321u);
            wb[0] = prmt(hlb, Rb[c][0], 0x6543u); wb[1] = prmt(Rb[c][0], Rb[c][1], 0x4321u);
            wb[2] = prmt(Rb[c][0], Rb[c][1], 0x6543u); wb[3] = prmt(Rb[c][1], hrb, 0x4321u);
#pragma unroll
            for (int xo = 0; xo < 4; xo++) {
                const int hb = dp4a_us(wb[xo], W4, 0);
                const int q = dp4a_us(wt[xo], W4, hb);
                const int n = dp4a_us(wt[xo], WA4, Pacc[xo][c]);
                Pacc[xo][c] = q * WB + Ta[xo][c];
                Ta[xo][c] = hb * WA;
                N[xo][c] = n;
                umin = min(umin, ((uint32_t)n << (32 - SH)) + (uint32_t)(xo * 3 + c));
                o[xo][c] = n >> SH;
                if (WRAP) o[xo][c] = max(o[xo][c], 0) & 0xFF;
            }
        }
        if (store) {
            constexpr bool SW = dst_swap(DST);
#define CH(i, c) o[i][SW ? 2 - (c) : (c)]
            if (DST == D_RGB24 || DST == D_BGR24) {
                stg32(pd,     pack4_u8(CH(0, 0), CH(0, 1), CH(0, 2), CH(1, 0)));
                stg32(pd + 4, pack4_u8(CH(1, 1), CH(1, 2), CH(2, 0), CH(2, 1)));
                stg32(pd + 8, pack4_u8(CH(2, 2), CH(3, 0), CH(3, 1), CH(3, 2)));
            } else {
                stg128(pd, make_uint4(pack4_u8(CH(0, 0), CH(0, 1), CH(0, 2), alpha_i), pack4_u8(CH(1, 0), CH(1, 1), CH(1, 2), alpha_i),
                                      pack4_u8(CH(2, 0), CH(2, 1), CH(2, 2), alpha_i), pack4_u8(CH(3, 0), CH(3, 1), CH(3, 2), alpha_i)));
            }
#undef CH
        }
        return umin;
    }
};

// ONE copy of the loop for every warp (edge strips and clamped rows are handled with warp-uniform branches and
// uniform-register address arithmetic): the first version had separate edge / interior and steady / boundary
// loops like v3 and stalled on instruction fetch (70 KB of hot code against a 32 KB L1.5 instruction cache:
// smsp__average_warps_issue_stalled_no_instruction 3.7 per issue).
template <int L, int DST, bool WRAP, int WA, int WB, int WS>
__device__ __forceinline__ void fused4i_band(const Fused3Params &P, uint32_t *ring) {
    typedef Fused4i<L, DST, WRAP, WA, WB, WS> F;
    typedef typename F::Row Row;
    constexpr int OWN = 30, SH = F::SH;
    const int lane = threadIdx.x;
    const int nstrips = P.src.w >> 3;
    const int strip = blockIdx.x * OWN + lane - 1;
    const bool owner = lane >= 1 && lane <= 30 && strip < nstrips;
    const int sl = min(max(strip, 0), nstrips - 1);
    const long long fz = blockIdx.z;
    const int yo_begin = blockIdx.y * P.band;
    const int yo_end = min(yo_begin + P.band, P.dstH);
    const int H = P.src.h, HC = H >> 1;
    const bool edge = blockIdx.x == 0 || (int)(blockIdx.x + 1) * OWN >= nstrips;     // warp-uniform

    const uint8_t *py = P.src.pl[0].p + fz * P.src.pl[0].bstride + (size_t)sl * 8;
    const uint8_t *pu = py, *pv = py;
    if (L == L_NV12) pu = P.src.pl[1].p + fz * P.src.pl[1].bstride + (size_t)sl * 8;
    if (L == L_I420) {
        pu = P.src.pl[1].p + fz * P.src.pl[1].bstride + (size_t)sl * 4;
        pv = P.src.pl[2].p + fz * P.src.pl[2].bstride + (size_t)sl * 4;
    }
    const unsigned pitch_y = P.src.pl[0].pitch, pitch_c = P.src.pl[1].pitch, pitch_c2 = P.src.pl[2].pitch;
    // pair k finishes output row k-1 and starts row k: pairs yo_begin-1 .. yo_end are consumed, the first two only
    // prime the accumulators; pd addresses row k-1 while pair k is processed
    const int kfirst = yo_begin - 1, klast = yo_end, kstore = kfirst + 2;
    uint8_t *pd = P.dst.pl[0].p + fz * P.dst.pl[0].bstride + ((long long)kfirst - 1) * (long long)P.dst.pl[0].pitch
                + (long long)(owner ? strip : 0) * (4 * dst_bpp(DST));
    const unsigned pitch_d = P.dst.pl[0].pitch;
    const bool lrep = strip < 0, rrep = strip >= nstrips;

    // Warp-uniform byte offsets of the rows of the pair that is loaded next.  Rows are clamped to the frame (pairs
    // -1 and HC replicate the first / last row): inside the loop the offsets advance by whole pairs, except into
    // pair HC.
    unsigned ot, ob, oc, oc2;
    auto seek = [&](int kk) {
        ot = (unsigned)min(max(2 * kk, 0), H - 1) * pitch_y; ob = (unsigned)min(max(2 * kk + 1, 0), H - 1) * pitch_y;
        const unsigned rc = (unsigned)min(max(kk, 0), HC - 1);
        oc = rc * pitch_c; oc2 = rc * pitch_c2;
    };
    auto load_here = [&](Row &R) {
        R.yt = ldg64(py + ot); R.yb = ldg64(py + ob);
        if (L == L_NV12) R.c0 = ldg64(pu + oc);
        else { R.c0.x = ldg32(pu + oc); R.c0.y = ldg32(pv + oc2); }
        if (edge) edge_replicate<L, 8>(R, lrep, rrep);
    };
    const unsigned sy = 2 * pitch_y;
    auto load_next = [&](Row &R, int kk) {      // loads pair kk >= 1 (the offsets are on it), then moves them to pair kk+1
        load_here(R);
        if (kk + 1 < HC) { ot += sy; ob += sy; oc += pitch_c; oc2 += pitch_c2; }
        else ot = ob;                           // pair HC: rows H-1, H-1, chroma row HC-1
    };

    int Pacc[4][3], Ta[4][3], N[4][3];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int c = 0; c < 3; c++) { Pacc[i][c] = 0; Ta[i][c] = 0; }

    int alpha_i = 0;
    if (dst_alpha(DST)) {   // the chain applied to the reference's constant intermediate alpha (scale_fused3.cuh)
        float ah = __fmul_rn(P.wx[1], 1.0f);
        ah = __fmaf_rn(P.wx[0], 1.0f, ah); ah = __fmaf_rn(P.wx[2], 1.0f, ah); ah = __fmaf_rn(P.wx[3], 1.0f, ah);
        float av = __fmul_rn(P.wy[1], ah);
        av = __fmaf_rn(P.wy[0], ah, av); av = __fmaf_rn(P.wy[2], ah, av); av = __fmaf_rn(P.wy[3], ah, av);
        alpha_i = trunc_i(__fmul_rn(av, P.factor));
    }

    // The outputs of row k-1 with N = 0 (mod 2^SH), N != 0: recompute each with the float chain (few), or report
    // a dense step (many).  Entered by the whole warp after the stores of the step.
    auto ambiguous = [&](bool stored, int k) -> bool {
        uint32_t mask = 0;
#pragma unroll
        for (int xo = 0; xo < 4; xo++)
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const uint32_t n = (uint32_t)N[xo][c];
                const bool f = stored && (n & (0u - n)) >= (1u << SH);           // lowest set bit >= 2^SH; false for n == 0
                mask |= (uint32_t)f << (xo * 3 + c);
            }
        const int total = __reduce_add_sync(0xffffffffu, __popc(mask));
        if (total == 0) return false;
        if (total > GMATB_F4I_DENSE) return true;
        __syncwarp();                       // this step's ring stores and output stores are visible to the whole warp
        uint32_t act = __ballot_sync(0xffffffffu, mask != 0u);
        uint8_t *orow = P.dst.pl[0].p + fz * P.dst.pl[0].bstride + (long long)(k - 1) * (long long)pitch_d
                      + ((long long)blockIdx.x * OWN - 1) * (4 * dst_bpp(DST));
        const int r = lane & 3;
        const uint32_t *rrow = ring + ((k - 2 + ((r + 1) >> 1)) & 3) * (32 * 12) + ((r & 1) ? 0 : 6);
        while (act) {
            const int l = __ffs(act) - 1;
            act &= act - 1;
            uint32_t mk = __shfl_sync(0xffffffffu, mask, l);
            while (mk) {
                const int i = __ffs(mk) - 1;
                mk &= mk - 1;
                const int xo = i / 3, c = i - 3 * xo;
                const int o = fused4i_fix<WRAP>(P, rrow, l, xo, c);
                if (lane == 0) orow[(l * 4 + xo) * dst_bpp(DST) + (dst_swap(DST) ? 2 - c : c)] = (uint8_t)o;
            }
        }
        return false;
    };

    Row A, B;
    int k = kfirst;
    seek(k); load_here(A);
    seek(k + 1); load_here(B);
    seek(k + 2);
    bool dense = false;
    uint32_t *lane_ring = ring + lane * 12;
#pragma unroll 1
    for (;;) {
        {
            const bool st = owner && k >= kstore;
            const uint32_t um = F::step(P, A, Pacc, Ta, N, st, pd, alpha_i, lane_ring + (k & 3) * (32 * 12),
                                        [&](Row &R) { if (k + 2 <= klast) load_next(R, k + 2); });
            if (__any_sync(0xffffffffu, st && um < 12u) && ambiguous(st, k)) { dense = true; break; }
            pd += pitch_d;
            if (++k > klast) break;
        }
        {
            const bool st = owner && k >= kstore;
            const uint32_t um = F::step(P, B, Pacc, Ta, N, st, pd, alpha_i, lane_ring + (k & 3) * (32 * 12),
                                        [&](Row &R) { if (k + 2 <= klast) load_next(R, k + 2); });
            if (__any_sync(0xffffffffu, st && um < 12u) && ambiguous(st, k)) { dense = true; break; }
            pd += pitch_d;
            if (++k > klast) break;
        }
    }
    // a dense step at pair k: output rows k-1 .. yo_end-1 are (re)done by the float loop (its edge-strip form is
    // correct for every warp)
    if (dense) fused3_band<L, 8, DST, false, WRAP, true>(P, blockIdx.x, fz, k - 1, yo_end);
}

template <int L, int DST, bool WRAP, int WA, int WB, int WS, int MINB>
__global__ void __launch_bounds__(32, MINB) fused_csc_scale2_int_kernel(const __grid_constant__ Fused3Params P) {
    __shared__ __align__(16) uint32_t ring[GMATB_F4I_RING];
    fused4i_band<L, DST, WRAP, WA, WB, WS>(P, ring);
}

}  // namespace gmatb
