// experiments/scale_fused_v1.cuh -- the first form of the fused 2:1 kernel (767 Gpx/s on C2, A = -0.75):
// a warp's outer lanes load + convert one extra halo column themselves and every lane selects between
// the shuffle result, the halo column and its own edge column.  Superseded by scale_fused3.cuh
// (overlapped strips, 975-994 Gpx/s); kept for the record, not compiled into the library.
// Needs the helpers of gmat_b200/csrc/scale_fused.cuh.
#pragma once
#include "../gmat_b200/csrc/scale_fused.cuh"

namespace gmatb {

// ---- v1-only helpers (moved here from scale_fused.cuh together with the kernel) ----
struct Fused2Params {
    Img src, dst;
    Mat9 M;
    float wx[4], wy[4];
    NormK nk;
    float factor;      // 255 or 65535
    int band;          // output rows per band
    int wrap;          // GMATB_SWS_PARITY_WRAP
    int dstW, dstH;
};


template <int L>
__device__ __forceinline__ void fused_load(const Fused2Params &P, long long fz, int xs, int k, RawRow<8> &R) {
    const int H = P.src.h;
    const int rt = min(max(2 * k, 0), H - 1), rb = min(max(2 * k + 1, 0), H - 1);
    const int rc = min(max(k, 0), (H >> 1) - 1);
    const uint8_t *py = P.src.pl[0].p + fz * P.src.pl[0].bstride;
    R.yt = ldg64(py + (size_t)rt * P.src.pl[0].pitch + xs);
    R.yb = ldg64(py + (size_t)rb * P.src.pl[0].pitch + xs);
    if (L == L_NV12) {
        R.c0 = ldg64(P.src.pl[1].p + fz * P.src.pl[1].bstride + (size_t)rc * P.src.pl[1].pitch + xs);
    } else {
        R.c0.x = ldg32(P.src.pl[1].p + fz * P.src.pl[1].bstride + (size_t)rc * P.src.pl[1].pitch + (xs >> 1));
        R.c0.y = ldg32(P.src.pl[2].p + fz * P.src.pl[2].bstride + (size_t)rc * P.src.pl[2].pitch + (xs >> 1));
    }
}
template <int L>
__device__ __forceinline__ void fused_load(const Fused2Params &P, long long fz, int xs, int k, RawRow<16> &R) {
    const int H = P.src.h;
    const int rt = min(max(2 * k, 0), H - 1), rb = min(max(2 * k + 1, 0), H - 1);
    const int rc = min(max(k, 0), (H >> 1) - 1);
    const uint8_t *py = P.src.pl[0].p + fz * P.src.pl[0].bstride;
    R.yt = ldg128(py + (size_t)rt * P.src.pl[0].pitch + xs * 2);
    R.yb = ldg128(py + (size_t)rb * P.src.pl[0].pitch + xs * 2);
    if (L == L_NV12) {
        R.c0 = ldg128(P.src.pl[1].p + fz * P.src.pl[1].bstride + (size_t)rc * P.src.pl[1].pitch + xs * 2);
    } else {
        uint2 u = ldg64(P.src.pl[1].p + fz * P.src.pl[1].bstride + (size_t)rc * P.src.pl[1].pitch + xs);
        uint2 v = ldg64(P.src.pl[2].p + fz * P.src.pl[2].bstride + (size_t)rc * P.src.pl[2].pitch + xs);
        R.c0 = make_uint4(u.x, u.y, v.x, v.y);
    }
}


__device__ __forceinline__ void fused_load_rgb(const Fused2Params &P, long long fz, int xs, int k, RawRowRGB &R) {
    const int H = P.src.h;
    const int rt = min(max(2 * k, 0), H - 1), rb = min(max(2 * k + 1, 0), H - 1);
    const uint8_t *py = P.src.pl[0].p + fz * P.src.pl[0].bstride + (size_t)xs * 3;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        R.t[i] = ldg64(py + (size_t)rt * P.src.pl[0].pitch + 8 * i);
        R.b[i] = ldg64(py + (size_t)rb * P.src.pl[0].pitch + 8 * i);
    }
}
template <int L, int SBITS> struct RowSel { typedef RawRow<SBITS> type; };
template <> struct RowSel<L_RGB3, 8> { typedef RawRowRGB type; };
template <int L> __device__ __forceinline__ void fused_load(const Fused2Params &P, long long fz, int xs, int k, RawRowRGB &R) {
    fused_load_rgb(P, fz, xs, k, R);
}

// one source column (top,bottom) -> normalised (r,g,b) pairs
template <int SBITS, bool SPARSE>
__device__ __forceinline__ void fused_column(float ytm, float ybm, const ChromaTerms &t, const Fused2Params &P, f2 (&out)[3]) {
    constexpr float YB = -(GMATB_MAGIC + (SBITS == 8 ? 16.f : 4096.f));
    f2 r, g, b;
    csc_pair_f<SPARSE, SBITS == 16>(add2(pk(ytm, ybm), bc(YB)), t, P.M, r, g, b);
    out[0] = quant_norm2(r, P.nk); out[1] = quant_norm2(g, P.nk); out[2] = quant_norm2(b, P.nk);
}


// the 8 (top,bottom) column pairs of one iteration as normalised samples
template <int L, int SBITS, bool SPARSE, typename Row>
__device__ __forceinline__ void fused_produce(const Row &cur, const Fused2Params &P, f2 (&C)[8][3]) {
    constexpr float CB = -(GMATB_MAGIC + (SBITS == 8 ? 128.f : 32768.f));
    float yt[8], yb[8], um[4], vm[4];
    fused_unpack<L>(cur, yt, yb, um, vm);
#pragma unroll
    for (int j = 0; j < 4; j++) {
        float fu, fv;
        upk(add2(pk(um[j], vm[j]), bc(CB)), fu, fv);
        ChromaTerms t = chroma_terms<SPARSE, SBITS == 16>(fu, fv, P.M);
        fused_column<SBITS, SPARSE>(yt[2 * j], yb[2 * j], t, P, C[2 * j]);
        fused_column<SBITS, SPARSE>(yt[2 * j + 1], yb[2 * j + 1], t, P, C[2 * j + 1]);
    }
}
template <int L, int SBITS, bool SPARSE>
__device__ __forceinline__ void fused_produce(const RawRowRGB &cur, const Fused2Params &P, f2 (&C)[8][3]) {
    rgb_column<0>(cur, P.nk, C[0]); rgb_column<1>(cur, P.nk, C[1]); rgb_column<2>(cur, P.nk, C[2]); rgb_column<3>(cur, P.nk, C[3]);
    rgb_column<4>(cur, P.nk, C[4]); rgb_column<5>(cur, P.nk, C[5]); rgb_column<6>(cur, P.nk, C[6]); rgb_column<7>(cur, P.nk, C[7]);
}


// raw samples of the one halo column a warp's outer lanes fetch themselves (prefetched one
// iteration ahead like the strip itself: the consumer must not wait on an L2 round trip).
//   8-bit yuv : w0 = top luma | bottom luma << 8 | U << 16 | V << 24
//   16-bit yuv: w0 = top luma | bottom luma << 16, w1 = U | V << 16
//   rgb       : w0 / w1 = the 3 bytes of the top / bottom pixel
struct ExtraRaw { uint32_t w0, w1; };
template <int L, int SBITS>
__device__ __forceinline__ void extra_load(const Fused2Params &P, long long fz, int xe, int k, ExtraRaw &X) {
    constexpr int SB = SBITS / 8;
    const int H = P.src.h;
    const int rt = min(max(2 * k, 0), H - 1), rb = min(max(2 * k + 1, 0), H - 1);
    const uint8_t *py = P.src.pl[0].p + fz * P.src.pl[0].bstride;
    if (L == L_RGB3) {
        const uint8_t *qa = py + (size_t)rt * P.src.pl[0].pitch + (size_t)xe * 3;
        const uint8_t *qb = py + (size_t)rb * P.src.pl[0].pitch + (size_t)xe * 3;
        X.w0 = qa[0] | (qa[1] << 8) | (qa[2] << 16);
        X.w1 = qb[0] | (qb[1] << 8) | (qb[2] << 16);
        return;
    }
    const int rc = min(max(k, 0), (H >> 1) - 1);
    const uint8_t *qa = py + (size_t)rt * P.src.pl[0].pitch + xe * SB;
    const uint8_t *qb = py + (size_t)rb * P.src.pl[0].pitch + xe * SB;
    uint32_t a, b, u, v;
    if (SBITS == 8) { a = *qa; b = *qb; }
    else { a = *reinterpret_cast<const uint16_t *>(qa); b = *reinterpret_cast<const uint16_t *>(qb); }
    if (L == L_NV12) {
        const uint8_t *qc = P.src.pl[1].p + fz * P.src.pl[1].bstride + (size_t)rc * P.src.pl[1].pitch + (xe >> 1) * 2 * SB;
        if (SBITS == 8) { X.w0 = a | (b << 8) | ((uint32_t)*reinterpret_cast<const uint16_t *>(qc) << 16); X.w1 = 0; }
        else { X.w0 = a | (b << 16); X.w1 = *reinterpret_cast<const uint32_t *>(qc); }
    } else {
        const uint8_t *qu = P.src.pl[1].p + fz * P.src.pl[1].bstride + (size_t)rc * P.src.pl[1].pitch + (xe >> 1) * SB;
        const uint8_t *qv = P.src.pl[2].p + fz * P.src.pl[2].bstride + (size_t)rc * P.src.pl[2].pitch + (xe >> 1) * SB;
        if (SBITS == 8) { u = *qu; v = *qv; X.w0 = a | (b << 8) | (u << 16) | (v << 24); X.w1 = 0; }
        else { u = *reinterpret_cast<const uint16_t *>(qu); v = *reinterpret_cast<const uint16_t *>(qv); X.w0 = a | (b << 16); X.w1 = u | (v << 16); }
    }
}

// DST: D_* code from csc.cu (packed rgb).  TAPS2: the outer weights of both axes are
// exactly zero (default bicubic, A = 0, at 2:1), FFMA(0, p, t) == t is skipped.
template <int L, int SBITS, int DST, bool TAPS2, bool WRAP, int MINB>
__global__ void __launch_bounds__(32, MINB) fused_csc_scale2_kernel(const Fused2Params P) {
    constexpr bool SPARSE = true;     // the host only selects this kernel for matrices with m[1] == m[8] == 0
    const int lane = threadIdx.x;
    const int W = P.src.w;
    const int x0 = (blockIdx.x * 32 + lane) * 8;
    const bool active = x0 < W;
    const int xs = active ? x0 : W - 8;
    const long long fz = blockIdx.z;
    const int yo_begin = blockIdx.y * P.band;
    const int yo_end = min(yo_begin + P.band, P.dstH);
    const bool ledge = xs == 0, redge = xs + 8 == W;
    const bool need_extra = (lane == 0 && !ledge) || (lane == 31 && !redge);
    const int xe = lane == 0 ? max(xs - 1, 0) : min(xs + 8, W - 1);
    constexpr int SB = SBITS / 8;
    constexpr float CB = -(GMATB_MAGIC + (SBITS == 8 ? 128.f : 32768.f));

    float hb_prev[4][3];   // horizontal results of the previous pair's bottom row
    float acc[4][3];       // partial vertical sums of the output row in flight
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int c = 0; c < 3; c++) { hb_prev[i][c] = 0.f; acc[i][c] = 0.f; }

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

    typename RowSel<L, SBITS>::type cur, nxt;
    fused_load<L>(P, fz, xs, yo_begin - 1, cur);
    ExtraRaw ecur = {0, 0}, enxt = {0, 0};
    if (!TAPS2 && need_extra) extra_load<L, SBITS>(P, fz, xe, yo_begin - 1, ecur);

#pragma unroll 2
    for (int k = yo_begin - 1; k <= yo_end; k++) {
        if (k < yo_end) {
            fused_load<L>(P, fz, xs, k + 1, nxt);
            if (!TAPS2 && need_extra) extra_load<L, SBITS>(P, fz, xe, k + 1, enxt);
        }
        // ---- extra (halo) column for the warp's outer lanes ------------------------
        f2 E[3] = {0ull, 0ull, 0ull};
        if (!TAPS2 && need_extra && L == L_RGB3) {
            E[0] = norm2_inrange(byte_magic<0>(ecur.w0), byte_magic<0>(ecur.w1), P.nk);
            E[1] = norm2_inrange(byte_magic<1>(ecur.w0), byte_magic<1>(ecur.w1), P.nk);
            E[2] = norm2_inrange(byte_magic<2>(ecur.w0), byte_magic<2>(ecur.w1), P.nk);
        } else if (!TAPS2 && need_extra) {
            float fu, fv, ya, yb;
            if (SBITS == 8) {
                ya = byte_magic<0>(ecur.w0); yb = byte_magic<1>(ecur.w0);
                upk(add2(pk(byte_magic<2>(ecur.w0), byte_magic<3>(ecur.w0)), bc(CB)), fu, fv);
            } else {
                ya = half_magic<0>(ecur.w0); yb = half_magic<1>(ecur.w0);
                upk(add2(pk(half_magic<0>(ecur.w1), half_magic<1>(ecur.w1)), bc(CB)), fu, fv);
            }
            ChromaTerms t = chroma_terms<SPARSE, SBITS == 16>(fu, fv, P.M);
            fused_column<SBITS, SPARSE>(ya, yb, t, P, E);
        }
        // ---- colour conversion of the 8x2 block ------------------------------------
        f2 C[8][3];
        fused_produce<L, SBITS, SPARSE>(cur, P, C);
        // ---- halo columns ------------------------------------------------------------
        f2 PL[3] = {0ull, 0ull, 0ull}, PR[3] = {0ull, 0ull, 0ull};
        if (!TAPS2) {
#pragma unroll
            for (int c = 0; c < 3; c++) {
                f2 up = shfl_up2(C[7][c]), dn = shfl_dn2(C[0][c]);
                PL[c] = ledge ? C[0][c] : (lane == 0 ? E[c] : up);
                PR[c] = redge ? C[7][c] : (lane == 31 ? E[c] : dn);
            }
        }
        // ---- horizontal pass (packed over the row pair) ----------------------------
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
        // ---- vertical pass: finish output row k-1, start output row k ---------------
        const int yo = k - 1;
        if (yo >= yo_begin && active) {
            int o[4][3];
#pragma unroll
            for (int xo = 0; xo < 4; xo++)
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    float v = TAPS2 ? acc[xo][c] : __fmaf_rn(P.wy[3], ht[xo][c], acc[xo][c]);
                    o[xo][c] = trunc_i(__fmul_rn(v, P.factor));
                    if (WRAP) o[xo][c] = max(o[xo][c], 0) & (SBITS == 8 ? 0xFF : 0xFFFF);
                }
            constexpr bool SW = dst_swap(DST);
            uint8_t *pd = P.dst.pl[0].p + fz * P.dst.pl[0].bstride + (size_t)yo * P.dst.pl[0].pitch
                        + (size_t)(x0 >> 1) * dst_bpp(DST);
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
        ecur = enxt;
    }
}


}  // namespace gmatb
