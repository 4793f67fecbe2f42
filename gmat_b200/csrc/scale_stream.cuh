// scale_stream.cuh -- ANY-RATIO fused colour conversion + 4-tap resample for 8-bit yuv 4:2:0 -> 8-bit packed rgb
// (R-B arithmetic: bicubic / Lanczos; every ratio that is not the exact 2:1 of scale_fused3.cuh: 1080p -> 720p,
// 4K -> 720p, 1080p -> 4K ...).  Same operations, in the same order, as the generic tile kernel
// (scale_generic.cuh) and the reference's two-kernel pipeline (csc_core.cuh, resample_core.cuh): bit-identical.
//
// The generic kernel stages a destination tile's whole source window in shared memory as floats (three
// stages, two CTA barriers, ~165 B of shared-memory traffic per destination pixel at 1.5:1) and ran at 6-11 % of
// the HBM roofline.  Here the VERTICAL direction streams through registers like the 2:1 kernel, and shared memory
// holds ONE row pair of converted samples, only to give the horizontal taps their run-time offsets:
//   * one warp per CTA owns a strip of up to 256 source columns and walks down a band of the frame one row PAIR
//     (one chroma row) per step; lane l loads and converts columns 8l .. 8l+7 of both rows exactly as
//     scale_fused3.cuh does (64-bit loads two steps ahead, packed (top, bottom) CSC, quantise + normalise);
//   * the 8 x 3 packed samples go to a 264-entry row buffer per channel (STS.128; entry = column - X0 + 2, two
//     replicated entries on either side of the frame stand for the clamped taps), __syncwarp;
//   * horizontal pass: the warp's outputs are dealt to the lanes round robin (output lane + 32 i, i < NOUT); each
//     takes its 4 consecutive taps per channel from the row buffer at the offset of its column (one LDS.64 per
//     tap: top and bottom row together) and runs the packed chain with its own 4 weights (registers);
//   * vertical pass: the horizontal results of the 3 previous source rows stay in registers; every output row whose
//     window ends at one of the two rows just produced is finished on the spot: 4-tap chain, scale, truncate, pack;
//   * stores: rgba one word per lane; rgb24: the four pixels of a lane quad are re-dealt into three words with one
//     SHFL + one PRMT, 96 contiguous bytes per warp store.
// Which source strip and which outputs a warp owns is planned on the host (run_stream in scale.cu) from the
// same position table the kernel's taps come from: first output a multiple of 4, strip start a multiple of 8.
#pragma once
#include <type_traits>
#include "scale_fused3.cuh"

namespace gmatb {

#define GMATB_STREAM_BUF 264        /* entries per channel of the row buffer */

struct StreamParams {
    Fused3Params F;                 // images, matrix, normalisation constants (wx / wy / band unused)
    const float4 *cx, *cy;          // per output column / row: the 4 weights
    const int *px, *py;             // per output column / row: position of the first tap
    const int4 *plan;               // per warp: x = first source column X0, y = first output column, z = outputs, w = converting lanes
    int band;                       // output rows per CTA
    int wrap;
};

static __device__ __noinline__ void stream_store_px3(uint8_t *pp, uint32_t pw) {
    pp[0] = (uint8_t)pw; pp[1] = (uint8_t)(pw >> 8); pp[2] = (uint8_t)(pw >> 16);
}

template <int L, int DST, int NOUT, int DEAL, int RA, int MINB>
__global__ void __launch_bounds__(32, MINB) fused_csc_scale_stream_kernel(const __grid_constant__ StreamParams P) {
    typedef Raw3<L, 8> Row;
    constexpr int BPP = dst_bpp(DST);
    constexpr bool SW = dst_swap(DST);
    __shared__ __align__(16) f2 buf[3][GMATB_STREAM_BUF];
    const int lane = threadIdx.x;
    const long long fz = blockIdx.z;
    const Fused3Params &F = P.F;
    const int W = F.src.w, H = F.src.h, HC = (H + 1) >> 1;
    const int4 pl = P.plan[blockIdx.x];
    const int X0 = pl.x, xoA = pl.y, nout = pl.z, nconv = pl.w;

    // ---- this lane's outputs ---------------------------------------------------------------------------------
    // DEAL 0: output lane + 32 i (round robin).  DEAL 1 (NOUT = 5): outputs 2 lane, 2 lane + 1 (i = 0, 1), 64 + the same
    // (i = 2, 3) and 128 + lane (i = 4): at ratios near 1.5 neighbouring lanes' taps are then 3 row-buffer entries apart
    // -- an odd stride: the 16 lanes of an LDS.64 phase hit 16 different bank pairs -- instead of 1.5 (2-way conflicts,
    // 36 M of 89 M shared-memory wavefronts under ncu).
    auto out_index = [&](int i) { return DEAL == 0 ? lane + 32 * i : i < 4 ? 64 * (i >> 1) + 2 * lane + (i & 1) : 128 + lane; };
    static_assert(DEAL == 0 || NOUT == 5, "the paired deal is laid out for 5 outputs per lane");
    float4 wx[NOUT];
    int off[NOUT];
#pragma unroll
    for (int i = 0; i < NOUT; i++) {
        const int xo = xoA + min(out_index(i), nout - 1);
        wx[i] = __ldg(P.cx + xo);
        off[i] = __ldg(P.px + xo) - X0 + 2;
    }
    const int yo_begin = blockIdx.y * P.band, yo_end = min(yo_begin + P.band, F.dstH);
    const int v_first = __ldg(P.py + yo_begin), v_last = __ldg(P.py + yo_end - 1) + 3;
    const int kp_start = v_first >> 1, kp_last = v_last >> 1;             // row pairs kp_start .. kp_last

    // ---- source rows: this lane's 8 columns -------------------------------------------------------------------
    const int cs = min(X0 + 8 * lane, (W - 1) & ~7);                     // lanes past the row end re-read its last chunk (never tapped)
    const uint8_t *py_ = F.src.pl[0].p + fz * F.src.pl[0].bstride + cs;
    const uint8_t *pu = F.src.pl[1].p + fz * F.src.pl[1].bstride + (L == L_NV12 ? cs : cs >> 1);
    const uint8_t *pv = L == L_I420 ? F.src.pl[2].p + fz * F.src.pl[2].bstride + (cs >> 1) : pu;
    const unsigned pitch_y = F.src.pl[0].pitch, pitch_c = F.src.pl[1].pitch, pitch_c2 = F.src.pl[2].pitch;
    auto load_pair = [&](int kp, Row &R) {       // rows 2kp, 2kp+1 and chroma row kp, each clamped to the frame
        const unsigned rt = (unsigned)min(max(2 * kp, 0), H - 1), rb = (unsigned)min(max(2 * kp + 1, 0), H - 1);
        const unsigned rc = (unsigned)min(max(kp, 0), HC - 1);
        R.yt = ldg64(py_ + rt * pitch_y); R.yb = ldg64(py_ + rb * pitch_y);
        if (L == L_NV12) R.c0 = ldg64(pu + rc * pitch_c);
        else { R.c0.x = ldg32(pu + rc * pitch_c); R.c0.y = ldg32(pv + rc * pitch_c2); }
    };

    // ---- destination ------------------------------------------------------------------------------------------
    const unsigned pitch_d = F.dst.pl[0].pitch;
    uint8_t *pd0 = F.dst.pl[0].p + fz * F.dst.pl[0].bstride + (size_t)xoA * BPP;
    // rgb24: lane q of a quad stores word q of the quad's 12 bytes (q < 3): bytes from its own pixel and the next lane's
    const int q = lane & 3;
    const uint32_t selq = q == 0 ? 0x4210u : q == 1 ? 0x5421u : 0x6542u;
    // alpha of 4-channel outputs: the chain over the constant 255 the reference's CSC writes, p = 1.0 (scale_generic.cuh)
    float ah[NOUT];
#pragma unroll
    for (int i = 0; i < NOUT; i++) {
        const float one = RA ? 255.0f : 1.0f;
        ah[i] = gen_chain(wx[i].x, wx[i].y, wx[i].z, wx[i].w, one, one, one, one);
    }

    // horizontal results of the three source rows before the current pair, oldest first.  (They move down by two rows
    // per step with register copies: with compile-time ring slots instead the step exists in four variants and the
    // loop, 46 KB of code, thrashed the 32 KB instruction cache: 1.4 no-instruction stalls per issue under ncu.)
    float hist[3][NOUT][3];
#pragma unroll
    for (int s = 0; s < 3; s++)
#pragma unroll
        for (int i = 0; i < NOUT; i++) { hist[s][i][0] = 0.f; hist[s][i][1] = 0.f; hist[s][i][2] = 0.f; }

    // the next two output rows' window ends and weights (warp-uniform loads, issued well ahead of their use)
    int yo = yo_begin;
    int pend0 = v_first + 3, pend1 = yo + 1 < yo_end ? __ldg(P.py + yo + 1) + 3 : 0x7fffffff;
    float4 wy0 = __ldg(P.cy + yo), wy1 = __ldg(P.cy + min(yo + 1, yo_end - 1));

    // finish output row yo: its window is rows r0 (oldest) .. r3 (newest)
    auto emit = [&](const float (&r0)[NOUT][3], const float (&r1)[NOUT][3], const float (&r2)[NOUT][3], const float (&r3)[NOUT][3]) {
        const float4 w = wy0;
        uint8_t *prow = pd0 + (size_t)yo * pitch_d;
        uint32_t pw[NOUT];
        int o[NOUT][4];
#pragma unroll
        for (int i = 0; i < NOUT; i++) {
#pragma unroll
            for (int c = 0; c < 3; c++) {
                // R-A weights are (0, 1-f, f, 0): FFMA(0, p, t) = t exactly, so the outer taps are skipped, not changed
                const float t = RA ? __fmaf_rn(w.z, r2[i][c], __fmul_rn(w.y, r1[i][c]))
                                   : gen_chain(w.x, w.y, w.z, w.w, r0[i][c], r1[i][c], r2[i][c], r3[i][c]);
                // R-A: rint + saturate: 1.5 * 2^23 + t rounds to nearest even, its low bits are the integer; the pack saturates
                if (RA) o[i][c] = __float_as_int(__fadd_rn(t, 12582912.0f)) - 0x4B400000;
                // fmaxf(NaN, -1) = -1: a NaN (0/0 Lanczos coefficients, scale_generic.cuh) stores 0 like cvt.rzi.u32.f32;
                // negative values are sign-magnitude integers < 0: the saturating pack clamps both ends
                else o[i][c] = trunc_i(fmaxf(__fmul_rn(t, F.factor), -1.0f));
            }
            o[i][3] = 255;
            if (BPP == 4) {
                const float av = gen_chain(w.x, w.y, w.z, w.w, ah[i], ah[i], ah[i], ah[i]);
                if (RA) o[i][3] = __float_as_int(__fadd_rn(av, 12582912.0f)) - 0x4B400000;
                else o[i][3] = trunc_i(fmaxf(__fmul_rn(av, F.factor), -1.0f));
            }
        }
        if (!RA && P.wrap) {                          // GMATB_SWS_PARITY_WRAP (tests): values >= 256 wrap instead of saturating
#pragma unroll
            for (int i = 0; i < NOUT; i++)
#pragma unroll
                for (int c = 0; c < 4; c++) o[i][c] = max(o[i][c], 0) & 0xFF;
        }
#pragma unroll
        for (int i = 0; i < NOUT; i++) pw[i] = pack4_u8(SW ? o[i][2] : o[i][0], o[i][1], SW ? o[i][0] : o[i][2], o[i][3]);      // saturating
#pragma unroll
        for (int i = 0; i < NOUT; i++) {
            const int oi = out_index(i);                               // index of the pixel among the warp's outputs
            if (BPP == 4) {
                if (oi < nout) stg32(prow + (size_t)oi * 4, pw[i]);
            } else if (DEAL == 1 && i < 4) {
                // lanes 2m, 2m+1 hold pixels 4m .. 4m+3 of the group: the even lane stores words 0 and 1, the odd lane word 2
                if (i & 1) continue;                                   // handled with i - 1
                const uint32_t nx = __shfl_down_sync(0xffffffffu, pw[i], 1);
                const int gb = 64 * (i >> 1) + 2 * (lane & ~1);        // first pixel of the group of four
                if (gb + 3 < nout) {
                    uint8_t *pg = prow + (size_t)gb * 3;
                    stg32(pg + ((lane & 1) ? 8 : 0), (lane & 1) ? prmt(pw[i], pw[i + 1], 0x6542u) : prmt(pw[i], pw[i + 1], 0x4210u));
                    if (!(lane & 1)) stg32(pg + 4, prmt(pw[i + 1], nx, 0x5421u));
                } else {                                               // ragged last group of the frame's last warp
                    if (oi < nout) stream_store_px3(prow + (size_t)oi * 3, pw[i]);
                    if (oi + 1 < nout) stream_store_px3(prow + (size_t)(oi + 1) * 3, pw[i + 1]);
                }
            } else {
                const uint32_t nx = __shfl_down_sync(0xffffffffu, pw[i], 1);
                const int qb = oi & ~3;                                // first pixel of the quad
                if (qb + 3 < nout) {
                    if (q < 3) stg32(prow + (size_t)qb * 3 + 4 * q, prmt(pw[i], nx, selq));
                } else if (oi < nout) stream_store_px3(prow + (size_t)oi * 3, pw[i]);      // ragged last quad of the frame's last warp
            }
        }
        ++yo;
        pend0 = pend1; wy0 = wy1;
        if (yo + 1 < yo_end) { pend1 = __ldg(P.py + yo + 1) + 3; wy1 = __ldg(P.cy + yo + 1); }
        else pend1 = 0x7fffffff;
    };

    // Bank conflicts: a lane's 8 entries are 64 B, so the STS.128 of column pair j would hit 2 bank groups from the 8 lanes
    // of a quarter warp (4-way conflict).  Lane l therefore works through its column pairs in the order j + rot (mod 4),
    // rot = (l >> 1) & 3 -- its raw words are rotated by 2 rot bytes once per step, nothing else changes -- which puts
    // the 8 stores of a quarter warp into 8 different bank groups.
    const int rot = (lane >> 1) & 3;
    f2 *wpos[4];
#pragma unroll
    for (int j = 0; j < 4; j++) wpos[j] = &buf[0][8 * lane + 2 + 2 * ((j + rot) & 3)];
    const bool lpad = X0 == 0, rpad = X0 + 8 * nconv >= W;             // warp-uniform: the strip touches the frame's left / right edge
    const f2 k45 = *reinterpret_cast<const f2 *>(F.cm45), k72 = *reinterpret_cast<const f2 *>(F.cm72);
    constexpr float CB = -(GMATB_MAGIC + 128.f), YB = -(GMATB_MAGIC + 16.f);

    // one row pair kp: convert, publish, horizontal pass, the output rows that end in it
    auto step = [&](const Row &now, int kp) {
        RawRow<8> rr;
        {
            auto rot64 = [&](uint2 v, int bits) {           // rotate the 8 bytes right by `bits` (0, 16, 32, 48)
                const uint32_t lo = (bits & 32) ? v.y : v.x, hi = (bits & 32) ? v.x : v.y;
                return make_uint2(__funnelshift_r(lo, hi, bits & 31), __funnelshift_r(hi, lo, bits & 31));
            };
            rr.yt = rot64(now.yt, 16 * rot); rr.yb = rot64(now.yb, 16 * rot);
            if (L == L_NV12) rr.c0 = rot64(now.c0, 16 * rot);
            else rr.c0 = make_uint2(__funnelshift_r(now.c0.x, now.c0.x, 8 * rot), __funnelshift_r(now.c0.y, now.c0.y, 8 * rot));
        }
        float yt[8], yb[8], um[4], vm[4];
        fused_unpack<L>(rr, yt, yb, um, vm);
        __syncwarp();                                  // the previous step's taps have been read
#pragma unroll
        for (int j = 0; j < 4; j++) {                  // two columns (one chroma sample) at a time, published at once
            const f2 uv = add2(pk(um[j], vm[j]), bc(CB));
            float t1g, t2g, t1b, t2r;
            upk(mul2(uv, k45), t1g, t2g);
            upk(mul2(uv, k72), t1b, t2r);
            f2 S[2][3];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int col = 2 * j + h;
                const f2 fy2 = add2(pk(yt[col], yb[col]), bc(YB));
                f2 xr = fma2(fy2, bc(F.m0), bc(F.m1));      // m1 is a run-time 0.0f: RN(fy*m0), as FFMA(fy, m0, +-0)
                f2 xg = fma2(fy2, bc(F.m3), bc(t1g));
                const f2 xb = fma2(fy2, bc(F.m6), bc(t1b));
                xr = add2(xr, bc(t2r)); xg = add2(xg, bc(t2g));
                if (RA) {            // integer-valued samples, clamped (gen_sample<1> of scale_generic.cuh)
                    auto qa = [](f2 x) {
                        float a, b;
                        upk(add2(add2_rz(x, bc(GMATB_MAGIC)), bc(-GMATB_MAGIC)), a, b);
                        return pk(fminf(fmaxf(a, 0.f), 255.f), fminf(fmaxf(b, 0.f), 255.f));
                    };
                    S[h][0] = qa(xr); S[h][1] = qa(xg); S[h][2] = qa(xb);
                } else {
                    S[h][0] = quant_norm2(xr, F.nk); S[h][1] = quant_norm2(xg, F.nk); S[h][2] = quant_norm2(xb, F.nk);
                }
            }
            if (lane < nconv) {
#pragma unroll
                for (int c = 0; c < 3; c++)
                    *reinterpret_cast<ulonglong2 *>(wpos[j] + c * GMATB_STREAM_BUF) = make_ulonglong2(S[0][c], S[1][c]);
                if (lpad && lane == 0 && j == 0) {           // lane 0 has rot = 0: this is column 0
#pragma unroll
                    for (int c = 0; c < 3; c++) *reinterpret_cast<ulonglong2 *>(&buf[c][0]) = make_ulonglong2(S[0][c], S[0][c]);
                }
            }
        }
        __syncwarp();
        if (rpad) {                                    // columns W, W+1 = column W-1
            if (lane < 3) { const f2 e = buf[lane][W - 1 - X0 + 2]; buf[lane][W - X0 + 2] = e; buf[lane][W - X0 + 3] = e; }
            __syncwarp();
        }
        // ---- horizontal pass -------------------------------------------------------------------------------------
        float htop[NOUT][3], hbot[NOUT][3];
#pragma unroll
        for (int i = 0; i < NOUT; i++)
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const f2 *t = &buf[c][off[i]];
                const f2 hh = RA ? fma2(bc(wx[i].z), t[2], mul2(bc(wx[i].y), t[1]))
                                 : gen_chain2(wx[i].x, wx[i].y, wx[i].z, wx[i].w, t[0], t[1], t[2], t[3]);
                upk(hh, htop[i][c], hbot[i][c]);
            }
        // ---- vertical pass: output rows whose window ends at row 2kp, then at row 2kp+1 ----------------------------
        while (pend0 == 2 * kp) emit(hist[0], hist[1], hist[2], htop);
        while (pend0 == 2 * kp + 1) emit(hist[1], hist[2], htop, hbot);
#pragma unroll
        for (int i = 0; i < NOUT; i++)
#pragma unroll
            for (int c = 0; c < 3; c++) { hist[0][i][c] = hist[2][i][c]; hist[1][i][c] = htop[i][c]; hist[2][i][c] = hbot[i][c]; }
    };

    // rows are loaded one pair ahead: the copy at the end of a step is the first use of the loaded words
    Row cur, nxt;
    load_pair(kp_start, cur);
#pragma unroll 1
    for (int kp = kp_start; kp <= kp_last; kp++) {
        if (kp < kp_last) load_pair(kp + 1, nxt);
        step(cur, kp);
        cur = nxt;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// The same machine for ONE PLANE of 8- or 16-bit samples with CH = 1 or 2 interleaved components, or 4 of 8 bits (the Y and
// UV planes of nv12 / yuv420p / p010 / p016 -> same format scaling, and rgb0 / bgr0 / rgba / bgra -> same format: what the scale_cuda filter and the yuv -> yuv branch of sws_scale
// do, swscale_cuda.c:372-476; any ratio, 2:1 included).  No colour conversion: a sample is RN(j / max) (R-B) or j (R-A).
// A lane owns 8 pixels of both rows of a pair; outputs leave through a small shared-memory row so that the warp
// stores whole words.
struct PlaneStreamParams {
    Plane src, dst;
    int W, H, dstW, dstH;
    NormK nk;
    const float4 *cx, *cy;
    const int *px, *py;
    const int4 *plan;
    int band, wrap;
};

template <int CH, int SBITS, int NOUT, int DEAL, int RA, int MINB>
__global__ void __launch_bounds__(32, MINB) plane_scale_stream_kernel(const __grid_constant__ PlaneStreamParams P) {
    constexpr int BP = CH * SBITS / 8;                // bytes per pixel: 1, 2, 2, 4
    constexpr int NW = 2 * BP;                        // words of a lane's 8 pixels
    constexpr int SMAX = SBITS == 8 ? 255 : 65535;
    __shared__ __align__(16) f2 buf[CH][GMATB_STREAM_BUF];
    __shared__ __align__(16) uint8_t orow[32 * NOUT * BP];
    const int lane = threadIdx.x;
    const long long fz = blockIdx.z;
    const int W = P.W, H = P.H;
    const int4 pl = P.plan[blockIdx.x];
    const int X0 = pl.x, xoA = pl.y, nout = pl.z, nconv = pl.w;
    // DEAL 1 (NOUT = 5, ratios near 1.5): outputs 2 lane, 2 lane + 1, 64 + the same, 128 + lane -- an odd lane stride in the
    // row buffer, no LDS bank conflicts (see the yuv kernel)
    auto out_index = [&](int i) { return DEAL == 0 ? lane + 32 * i : i < 4 ? 64 * (i >> 1) + 2 * lane + (i & 1) : 128 + lane; };
    static_assert(DEAL == 0 || NOUT == 5, "the paired deal is laid out for 5 outputs per lane");
    float4 wx[NOUT];
    int off[NOUT];
#pragma unroll
    for (int i = 0; i < NOUT; i++) {
        const int xo = xoA + min(out_index(i), nout - 1);
        wx[i] = __ldg(P.cx + xo);
        off[i] = __ldg(P.px + xo) - X0 + 2;
    }
    const int yo_begin = blockIdx.y * P.band, yo_end = min(yo_begin + P.band, P.dstH);
    const int v_first = __ldg(P.py + yo_begin), v_last = __ldg(P.py + yo_end - 1) + 3;
    const int kp_start = v_first >> 1, kp_last = v_last >> 1;

    const int cs = min(X0 + 8 * lane, (W - 1) & ~7);
    const uint8_t *ps = P.src.p + fz * P.src.bstride + (size_t)cs * BP;
    const unsigned pitch_s = P.src.pitch, pitch_d = P.dst.pitch;
    struct Rows { uint32_t t[NW], b[NW]; };
    auto load_row = [&](const uint8_t *q, uint32_t (&w)[NW]) {
        if (NW == 2) { const uint2 a = ldg64(q); w[0] = a.x; w[1] = a.y; }
        else if (NW == 6) {                            // 3 components of 8 bits: 24 bytes, 8-byte aligned
#pragma unroll
            for (int k = 0; k < 3; k++) { const uint2 a = ldg64(q + 8 * k); w[(2 * k) % NW] = a.x; w[(2 * k + 1) % NW] = a.y; }
        } else {
#pragma unroll
            for (int k = 0; k < NW / 4; k++) {
                const uint4 a = ldg128(q + 16 * k);
                w[4 * k] = a.x; w[4 * k + 1] = a.y; w[4 * k + 2] = a.z; w[4 * k + 3] = a.w;
            }
        }
    };
    auto load_pair = [&](int kp, Rows &R) {
        const unsigned rt = (unsigned)min(max(2 * kp, 0), H - 1), rb = (unsigned)min(max(2 * kp + 1, 0), H - 1);
        load_row(ps + rt * pitch_s, R.t); load_row(ps + rb * pitch_s, R.b);
    };
    uint8_t *pd0 = P.dst.p + fz * P.dst.bstride + (size_t)xoA * BP;
    const float factor = (float)SMAX;

    float hist[3][NOUT][CH];
#pragma unroll
    for (int s = 0; s < 3; s++)
#pragma unroll
        for (int i = 0; i < NOUT; i++)
#pragma unroll
            for (int c = 0; c < CH; c++) hist[s][i][c] = 0.f;
    int yo = yo_begin;
    int pend0 = v_first + 3, pend1 = yo + 1 < yo_end ? __ldg(P.py + yo + 1) + 3 : 0x7fffffff;
    float4 wy0 = __ldg(P.cy + yo), wy1 = __ldg(P.cy + min(yo + 1, yo_end - 1));

    auto emit = [&](const float (&r0)[NOUT][CH], const float (&r1)[NOUT][CH], const float (&r2)[NOUT][CH], const float (&r3)[NOUT][CH]) {
        const float4 w = wy0;
        __syncwarp();                                  // the previous row's copy-out has read the staging row
#pragma unroll
        for (int i = 0; i < NOUT; i++) {
            int o[CH];
#pragma unroll
            for (int c = 0; c < CH; c++) {
                const float t = RA ? __fmaf_rn(w.z, r2[i][c], __fmul_rn(w.y, r1[i][c]))
                                   : gen_chain(w.x, w.y, w.z, w.w, r0[i][c], r1[i][c], r2[i][c], r3[i][c]);
                if (RA) o[c] = __float_as_int(__fadd_rn(t, 12582912.0f)) - 0x4B400000;
                else o[c] = trunc_i(fmaxf(__fmul_rn(t, factor), -1.0f));
                if (!RA && P.wrap) o[c] = max(o[c], 0) & SMAX;      // GMATB_SWS_PARITY_WRAP (tests)
            }
            const int oi = out_index(i);
            if (SBITS == 8 && CH == 3) {
                const uint32_t pw = pack4_u8(o[0], o[1 % CH], o[2 % CH], 0);      // saturating
                orow[3 * oi] = (uint8_t)pw; orow[3 * oi + 1] = (uint8_t)(pw >> 8); orow[3 * oi + 2] = (uint8_t)(pw >> 16);
            } else if (SBITS == 8 && CH == 4) {
                reinterpret_cast<uint32_t *>(orow)[oi] = pack4_u8(o[0], o[1 % CH], o[2 % CH], o[3 % CH]);      // saturating
            } else if (SBITS == 8) {
                const uint32_t pw = pack4_u8(o[0], o[CH - 1], 0, 0);      // saturating
                if (CH == 1) orow[oi] = (uint8_t)pw;
                else reinterpret_cast<unsigned short *>(orow)[oi] = (unsigned short)pw;
            } else {
                const uint32_t pw = pack2_u16(o[0], o[CH - 1]);
                if (CH == 1) reinterpret_cast<unsigned short *>(orow)[oi] = (unsigned short)pw;
                else reinterpret_cast<uint32_t *>(orow)[oi] = pw;
            }
        }
        __syncwarp();
        uint8_t *prow = pd0 + (size_t)yo * pitch_d;
        const int nbytes = nout * BP;
#pragma unroll
        for (int t = 0; t < (NOUT * BP + 3) / 4; t++) {
            const int wd = lane + 32 * t;
            if (4 * wd + 3 < nbytes) stg32(prow + 4 * wd, reinterpret_cast<const uint32_t *>(orow)[wd]);
            else {
#pragma unroll 1
                for (int b = 4 * wd; b < nbytes; b++) prow[b] = orow[b];
            }
        }
        ++yo;
        pend0 = pend1; wy0 = wy1;
        if (yo + 1 < yo_end) { pend1 = __ldg(P.py + yo + 1) + 3; wy1 = __ldg(P.cy + yo + 1); }
        else pend1 = 0x7fffffff;
    };

    // bank-conflict-free publication: as in the yuv kernel, lane l walks its 4 pixel pairs in the order j + rot
    const int rot = (lane >> 1) & 3;
    f2 *wpos[4];
#pragma unroll
    for (int j = 0; j < 4; j++) wpos[j] = &buf[0][8 * lane + 2 + 2 * ((j + rot) & 3)];
    const bool lpad = X0 == 0, rpad = X0 + 8 * nconv >= W;

    auto sample2 = [&](float mt, float mb) -> f2 {     // (top, bottom) magic floats -> samples
        if (RA) return add2(pk(mt, mb), bc(-GMATB_MAGIC));
        return norm2_inrange(mt, mb, P.nk);
    };
    // rotate a lane's 8 pixels left by 2 rot pixels, in its raw words
    auto rotate = [&](uint32_t (&w)[NW]) {
        if (NW == 2) {                                 // 8-bit, 1 component: a pair is 16 bits
            const int bits = 16 * rot;
            const uint32_t lo = (bits & 32) ? w[1] : w[0], hi = (bits & 32) ? w[0] : w[1];
            w[0] = __funnelshift_r(lo, hi, bits & 31); w[1] = __funnelshift_r(hi, lo, bits & 31);
        } else if (NW == 6) {                          // a pair is 6 bytes: by 3 words (two pairs), then by 1 word + 16 bits (one pair)
            uint32_t a[NW];
#pragma unroll
            for (int k = 0; k < NW; k++) a[k] = (rot & 2) ? w[(k + 3) % NW] : w[k];
#pragma unroll
            for (int k = 0; k < NW; k++) w[k] = (rot & 1) ? __funnelshift_r(a[(k + 1) % NW], a[(k + 2) % NW], 16) : a[k];
        } else {                                       // a pair is NW / 4 whole words: two conditional stages
            uint32_t a[NW];
#pragma unroll
            for (int k = 0; k < NW; k++) a[k] = (rot & 2) ? w[(k + NW / 2) % NW] : w[k];
#pragma unroll
            for (int k = 0; k < NW; k++) w[k] = (rot & 1) ? a[(k + NW / 4) % NW] : a[k];
        }
    };
    auto step = [&](const Rows &now, int kp) {
        uint32_t wt[NW], wb[NW];
#pragma unroll
        for (int k = 0; k < NW; k++) { wt[k] = now.t[k]; wb[k] = now.b[k]; }
        rotate(wt); rotate(wb);
        __syncwarp();                                  // the previous step's taps have been read
#pragma unroll
        for (int j = 0; j < 4; j++) {                  // pixels 2j, 2j+1 (after rotation)
            f2 S[2][CH];
            if (SBITS == 8 && CH == 1) {
                const uint32_t a = wt[j >> 1], b = wb[j >> 1];
                if (j & 1) { S[0][0] = sample2(byte_magic<2>(a), byte_magic<2>(b)); S[1][0] = sample2(byte_magic<3>(a), byte_magic<3>(b)); }
                else       { S[0][0] = sample2(byte_magic<0>(a), byte_magic<0>(b)); S[1][0] = sample2(byte_magic<1>(a), byte_magic<1>(b)); }
            } else if (SBITS == 8 && CH == 3) {        // bytes 6j .. 6j+5: component c of pixel 2j+h is byte 6j + 3h + c
#pragma unroll
                for (int h = 0; h < 2; h++)
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        const int B = 6 * j + 3 * h + c;
                        const uint32_t a = wt[(B >> 2) % NW], b = wb[(B >> 2) % NW];
                        S[h][c % CH] = (B & 3) == 0 ? sample2(byte_magic<0>(a), byte_magic<0>(b)) : (B & 3) == 1 ? sample2(byte_magic<1>(a), byte_magic<1>(b))
                                     : (B & 3) == 2 ? sample2(byte_magic<2>(a), byte_magic<2>(b)) : sample2(byte_magic<3>(a), byte_magic<3>(b));
                    }
            } else if (SBITS == 8 && CH == 4) {        // two words: the 4 components of pixel 2j, of pixel 2j+1
                const uint32_t a0 = wt[(2 * j) % NW], b0 = wb[(2 * j) % NW], a1 = wt[(2 * j + 1) % NW], b1 = wb[(2 * j + 1) % NW];
                S[0][0] = sample2(byte_magic<0>(a0), byte_magic<0>(b0)); S[0][1 % CH] = sample2(byte_magic<1>(a0), byte_magic<1>(b0));
                S[0][2 % CH] = sample2(byte_magic<2>(a0), byte_magic<2>(b0)); S[0][3 % CH] = sample2(byte_magic<3>(a0), byte_magic<3>(b0));
                S[1][0] = sample2(byte_magic<0>(a1), byte_magic<0>(b1)); S[1][1 % CH] = sample2(byte_magic<1>(a1), byte_magic<1>(b1));
                S[1][2 % CH] = sample2(byte_magic<2>(a1), byte_magic<2>(b1)); S[1][3 % CH] = sample2(byte_magic<3>(a1), byte_magic<3>(b1));
            } else if (SBITS == 8) {                   // one word: c0 c1 of pixel 2j, c0 c1 of pixel 2j+1
                const uint32_t a = wt[j], b = wb[j];
                S[0][0] = sample2(byte_magic<0>(a), byte_magic<0>(b)); S[0][CH - 1] = sample2(byte_magic<1>(a), byte_magic<1>(b));
                S[1][0] = sample2(byte_magic<2>(a), byte_magic<2>(b)); S[1][CH - 1] = sample2(byte_magic<3>(a), byte_magic<3>(b));
            } else if (CH == 1) {                      // one word: pixels 2j, 2j+1 as halves
                const uint32_t a = wt[j], b = wb[j];
                S[0][0] = sample2(half_magic<0>(a), half_magic<0>(b)); S[1][0] = sample2(half_magic<1>(a), half_magic<1>(b));
            } else {                                   // two words: (c0, c1) of pixel 2j, (c0, c1) of pixel 2j+1
                const uint32_t a0 = wt[(2 * j) % NW], b0 = wb[(2 * j) % NW], a1 = wt[(2 * j + 1) % NW], b1 = wb[(2 * j + 1) % NW];
                S[0][0] = sample2(half_magic<0>(a0), half_magic<0>(b0)); S[0][CH - 1] = sample2(half_magic<1>(a0), half_magic<1>(b0));
                S[1][0] = sample2(half_magic<0>(a1), half_magic<0>(b1)); S[1][CH - 1] = sample2(half_magic<1>(a1), half_magic<1>(b1));
            }
            if (lane < nconv) {
#pragma unroll
                for (int c = 0; c < CH; c++)
                    *reinterpret_cast<ulonglong2 *>(wpos[j] + c * GMATB_STREAM_BUF) = make_ulonglong2(S[0][c], S[1][c]);
                if (lpad && lane == 0 && j == 0) {
#pragma unroll
                    for (int c = 0; c < CH; c++) *reinterpret_cast<ulonglong2 *>(&buf[c][0]) = make_ulonglong2(S[0][c], S[0][c]);
                }
            }
        }
        __syncwarp();
        if (rpad) {
            if (lane < CH) { const f2 e = buf[lane][W - 1 - X0 + 2]; buf[lane][W - X0 + 2] = e; buf[lane][W - X0 + 3] = e; }
            __syncwarp();
        }
        float htop[NOUT][CH], hbot[NOUT][CH];
#pragma unroll
        for (int i = 0; i < NOUT; i++)
#pragma unroll
            for (int c = 0; c < CH; c++) {
                const f2 *t = &buf[c][off[i]];
                const f2 hh = RA ? fma2(bc(wx[i].z), t[2], mul2(bc(wx[i].y), t[1]))
                                 : gen_chain2(wx[i].x, wx[i].y, wx[i].z, wx[i].w, t[0], t[1], t[2], t[3]);
                upk(hh, htop[i][c], hbot[i][c]);
            }
        while (pend0 == 2 * kp) emit(hist[0], hist[1], hist[2], htop);
        while (pend0 == 2 * kp + 1) emit(hist[1], hist[2], htop, hbot);
#pragma unroll
        for (int i = 0; i < NOUT; i++)
#pragma unroll
            for (int c = 0; c < CH; c++) { hist[0][i][c] = hist[2][i][c]; hist[1][i][c] = htop[i][c]; hist[2][i][c] = hbot[i][c]; }
    };

    Rows cur, nxt;
    load_pair(kp_start, cur);
#pragma unroll 1
    for (int kp = kp_start; kp <= kp_last; kp++) {
        if (kp < kp_last) load_pair(kp + 1, nxt);
        step(cur, kp);
        cur = nxt;
    }
}

}  // namespace gmatb
