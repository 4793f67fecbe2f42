// scale_fused_lut.cuh -- second generation of the fused CSC + 2:1 resample kernel for
// 8-bit sources.  Same arithmetic, same results as scale_fused.cuh, different machine
// mapping of its most expensive step.
//
// Profiling the first version (profiles/) showed the FP32 pipe saturated, and 47 % of its
// work was one step: turning each quantised colour sample j = trunc(r) into the float the
// reference's texture fetch returns, p = RN(j/255), clamped -- 4 FMA-pipe operations per
// sample, two of which (.sat) have no packed form.  Here that step leaves the FP32 pipe:
//
//   m    = FADD2.RZ(4r, magic)                 bits(m) = const + floor(r)
//   addr = SHF(bits, lane)                      one ALU op
//   p    = LDS [addr]                           one shared-memory load per sample
//
// The table holds RN(clamp(j,0,255)/255) for every value floor(r) can take (negative results
// read 0.0, results above 255 read 1.0), so the clamp is free.  To make the lookup
// conflict-free for ANY data it is replicated once per lane: word address = index*32 + lane,
// i.e. each lane only ever touches its own shared-memory bank -- exactly one wavefront per
// LDS whatever the pixel values are.  1024 entries x 32 lanes x 4 B = 128 KB of the SM's
// 227 KB, which is why this kernel is PERSISTENT: one 512-thread CTA per SM builds the table
// once and its 16 warps then pull (frame, band, strip) work items from a grid-stride queue
// until the batch is done.
#pragma once
#include "scale_fused.cuh"

namespace gmatb {

#define LUT_FIRST 384                  // table entry 0  <=>  floor(r) = -384  (B of BT.2020 reaches -299)
#define LUT_N     1024                 // entries: floor(r) in [-384, 640)
#define LUT_BYTES (LUT_N * 32 * 4 + 128)

struct FusedLutParams {
    Fused2Params f;                // f.M holds 4 x the CSC matrix (see below)
    int warps_x, nbands, batch;    // work-item grid
};

// p(top), p(bottom) of one colour component of one column, from r4 = 4*r.
//
// The colour-conversion chain is run with the matrix scaled by 4 (a power of two: every
// intermediate is exactly 4x the reference's, roundings included), so that the truncating
// add can use the magic constant 2^25, whose bit pattern 0x4C000000 has NO bits inside the
// 25 positions that survive a left shift by 7:
//   m    = RZ(4r + 2^25 + 4*(384 + base/128))       ulp(2^25) = 4  ->  m = 2^25 + 4*(floor(r) + 384 + base/128)
//   bits = 0x4C000000 + floor(r) + 384 + base/128
//   addr = funnel-shift-left((bits : lane*4 << 25), 7) = (floor(r)+384)*128 + base + lane*4
// i.e. the absolute shared-memory address of this lane's copy of the entry, in ONE ALU
// instruction (SHF) -- no integer multiply-add on the FP32 pipe, no clamp: negative results
// read the 0.0 entries, results above 255 the 1.0 entries.
__device__ __forceinline__ f2 lut_norm2(f2 r4, f2 magic2, unsigned lane_hi) {
    const f2 m = add2_rz(r4, magic2);
    int b0, b1;
    upki(m, b0, b1);
    unsigned a0, a1;
    asm("shf.l.wrap.b32 %0, %1, %2, 7;" : "=r"(a0) : "r"(lane_hi), "r"(b0));
    asm("shf.l.wrap.b32 %0, %1, %2, 7;" : "=r"(a1) : "r"(lane_hi), "r"(b1));
    float p0, p1;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(p0) : "r"(a0));
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(p1) : "r"(a1));
    return pk(p0, p1);
}

template <bool SPARSE>
__device__ __forceinline__ void lut_column(float ytm, float ybm, const ChromaTerms &t, const Fused2Params &P,
                                           f2 magic2, unsigned lane_hi, f2 (&out)[3]) {
    constexpr float YB = -(GMATB_MAGIC + 16.f);
    f2 r, g, b;
    csc_pair_f<SPARSE>(add2(pk(ytm, ybm), bc(YB)), t, P.M, r, g, b);      // P.M = 4*M  ->  4r
    out[0] = lut_norm2(r, magic2, lane_hi); out[1] = lut_norm2(g, magic2, lane_hi); out[2] = lut_norm2(b, magic2, lane_hi);
}

template <int L, int DST, bool TAPS2>
__global__ void __launch_bounds__(512, 1) fused_csc_scale2_lut_kernel(const FusedLutParams Q) {
    extern __shared__ float lut_raw[];
    const Fused2Params &P = Q.f;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // 128-byte aligned table base (so that base/128 can ride in the magic constant)
    const unsigned raw = (unsigned)__cvta_generic_to_shared(lut_raw);
    const unsigned base = (raw + 127u) & ~127u;
    float *lut = lut_raw + ((base - raw) >> 2);         // [LUT_N][32]
    // ---- build the table (once per CTA): entry e <-> floor(r) = e - LUT_FIRST ---------------------
    for (int i = threadIdx.x; i < LUT_N * 32; i += blockDim.x) {
        const int j = min(max((i >> 5) - LUT_FIRST, 0), 255);
        lut[i] = __fdiv_rn((float)j, 255.0f);          // what the texture unit returns for texel j
    }
    __syncthreads();
    const f2 magic2 = bc(33554432.0f + 4.0f * (float)(LUT_FIRST + (base >> 7)));   // 2^25 + 4*(LUT_FIRST + base/128), exact
    const unsigned lane_hi = (unsigned)(lane * 4) << 25;

    const int W = P.src.w;
    constexpr float CB = -(GMATB_MAGIC + 128.f);
    const long long total = (long long)Q.warps_x * Q.nbands * Q.batch;
    for (long long item = (long long)blockIdx.x * (blockDim.x >> 5) + warp; item < total;
         item += (long long)gridDim.x * (blockDim.x >> 5)) {
        const int wx = (int)(item % Q.warps_x);
        const int band = (int)((item / Q.warps_x) % Q.nbands);
        const long long fz = item / ((long long)Q.warps_x * Q.nbands);
        const int x0 = (wx * 32 + lane) * 8;
        const bool active = x0 < W;
        const int xs = active ? x0 : W - 8;
        const int yo_begin = band * P.band;
        const int yo_end = min(yo_begin + P.band, P.dstH);
        const bool ledge = xs == 0, redge = xs + 8 == W;
        const bool need_extra = (lane == 0 && !ledge) || (lane == 31 && !redge);
        const int xe = lane == 0 ? max(xs - 1, 0) : min(xs + 8, W - 1);

        float hb_prev[4][3], acc[4][3];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int c = 0; c < 3; c++) { hb_prev[i][c] = 0.f; acc[i][c] = 0.f; }
        int alpha_i = 0;
        if (dst_alpha(DST)) {
            float ah = __fmul_rn(P.wx[1], 1.0f);
            ah = __fmaf_rn(P.wx[0], 1.0f, ah); ah = __fmaf_rn(P.wx[2], 1.0f, ah); ah = __fmaf_rn(P.wx[3], 1.0f, ah);
            float av = __fmul_rn(P.wy[1], ah);
            av = __fmaf_rn(P.wy[0], ah, av); av = __fmaf_rn(P.wy[2], ah, av); av = __fmaf_rn(P.wy[3], ah, av);
            alpha_i = trunc_i(__fmul_rn(av, P.factor));
        }
        RawRow<8> cur, nxt;
        fused_load<L>(P, fz, xs, yo_begin - 1, cur);
        for (int k = yo_begin - 1; k <= yo_end; k++) {
            if (k < yo_end) fused_load<L>(P, fz, xs, k + 1, nxt);
            f2 E[3] = {0ull, 0ull, 0ull};
            if (!TAPS2 && need_extra) {
                const int H = P.src.h;
                const int rt = min(max(2 * k, 0), H - 1), rb = min(max(2 * k + 1, 0), H - 1);
                const int rc = min(max(k, 0), (H >> 1) - 1);
                const uint8_t *py = P.src.pl[0].p + fz * P.src.pl[0].bstride;
                unsigned a = py[(size_t)rt * P.src.pl[0].pitch + xe], b = py[(size_t)rb * P.src.pl[0].pitch + xe], u, v;
                if (L == L_NV12) {
                    const uint8_t *qc = P.src.pl[1].p + fz * P.src.pl[1].bstride + (size_t)rc * P.src.pl[1].pitch + (xe >> 1) * 2;
                    u = qc[0]; v = qc[1];
                } else {
                    u = P.src.pl[1].p[fz * P.src.pl[1].bstride + (size_t)rc * P.src.pl[1].pitch + (xe >> 1)];
                    v = P.src.pl[2].p[fz * P.src.pl[2].bstride + (size_t)rc * P.src.pl[2].pitch + (xe >> 1)];
                }
                float fu, fv;
                upk(add2(pk(__uint_as_float(0x4B000000u | u), __uint_as_float(0x4B000000u | v)), bc(CB)), fu, fv);
                ChromaTerms t = chroma_terms<true>(fu, fv, P.M);
                lut_column<true>(__uint_as_float(0x4B000000u | a), __uint_as_float(0x4B000000u | b), t, P, magic2, lane_hi, E);
            }
            float yt[8], yb[8], um[4], vm[4];
            fused_unpack<L>(cur, yt, yb, um, vm);
            f2 C[8][3];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                float fu, fv;
                upk(add2(pk(um[j], vm[j]), bc(CB)), fu, fv);
                ChromaTerms t = chroma_terms<true>(fu, fv, P.M);
                lut_column<true>(yt[2 * j], yb[2 * j], t, P, magic2, lane_hi, C[2 * j]);
                lut_column<true>(yt[2 * j + 1], yb[2 * j + 1], t, P, magic2, lane_hi, C[2 * j + 1]);
            }
            f2 PL[3] = {0ull, 0ull, 0ull}, PR[3] = {0ull, 0ull, 0ull};
            if (!TAPS2) {
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    f2 up = shfl_up2(C[7][c]), dn = shfl_dn2(C[0][c]);
                    PL[c] = ledge ? C[0][c] : (lane == 0 ? E[c] : up);
                    PR[c] = redge ? C[7][c] : (lane == 31 ? E[c] : dn);
                }
            }
            float ht[4][3], hbm[4][3];
#pragma unroll
            for (int xo = 0; xo < 4; xo++)
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    f2 p0 = xo == 0 ? PL[c] : C[2 * xo - 1][c];
                    f2 p3 = xo == 3 ? PR[c] : C[2 * xo + 2][c];
                    f2 h = hpass<TAPS2>(P.wx, p0, C[2 * xo][c], C[2 * xo + 1][c], p3);
                    upk(h, ht[xo][c], hbm[xo][c]);
                }
            const int yo = k - 1;
            if (yo >= yo_begin && active) {
                int o[4][3];
#pragma unroll
                for (int xo = 0; xo < 4; xo++)
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        float v = TAPS2 ? acc[xo][c] : __fmaf_rn(P.wy[3], ht[xo][c], acc[xo][c]);
                        o[xo][c] = trunc_i(__fmul_rn(v, P.factor));
                        if (P.wrap) o[xo][c] = max(o[xo][c], 0) & 0xFF;
                    }
                constexpr bool SW = dst_swap(DST);
                uint8_t *pd = P.dst.pl[0].p + fz * P.dst.pl[0].bstride + (size_t)yo * P.dst.pl[0].pitch
                            + (size_t)(x0 >> 1) * dst_bpp(DST);
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
#pragma unroll
            for (int xo = 0; xo < 4; xo++)
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    float t = __fmul_rn(P.wy[1], ht[xo][c]);
                    if (!TAPS2) t = __fmaf_rn(P.wy[0], hb_prev[xo][c], t);
                    t = __fmaf_rn(P.wy[2], hbm[xo][c], t);
                    acc[xo][c] = t;
                    hb_prev[xo][c] = hbm[xo][c];
                }
            cur = nxt;
        }
    }
}

}  // namespace gmatb
