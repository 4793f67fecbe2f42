// scale_generic.cuh -- any-ratio, any-size scaling kernel (everything that is not the fused
// exact-2:1 case: other ratios, rgb->rgb, rgb->yuv prescale and per-plane yuv scaling).
//
// A CTA produces a tile of the destination in three stages over shared memory:
//   A  the source window the tile taps (coordinates clamped to the image like the reference's
//      texture fetches) is converted ONCE per source pixel to the sample the reference's resize
//      stage would read: a normalised packed component, or -- for yuv 4:2:0 sources -- the
//      quantised, normalised result of the colour conversion (P-ORDER, swscale_cuda.c:352-370).
//      yuv sources are converted a 2x2 block at a time (one chroma sample, two packed
//      (left,right) pixel pairs on FFMA2, 16/32-bit loads); window cells whose block touches the
//      frame edge take the per-pixel path with clamped coordinates.
//   B  the horizontal 4-tap chain, once per (window row, output column);
//   C  the vertical chain, the store conversion and the store.
// Samples sit in shared memory as 1, 2 or 4 floats per pixel (3-component pixels are padded to
// 4: one LDS.128 per tap, the chain runs on two f32x2 pairs).  Both chains are in the reference's
// operand order (resample_core.cuh), so the result is bit-identical to the fused kernel's.
#pragma once
#include "csc_core.cuh"
#include "resample_core.cuh"

namespace gmatb {

enum { GS_NV12 = 0, GS_I420 = 1, GS_PACKED = 2 };

struct GenParams {
    Img src, dst;
    Mat9 M;
    NormK nk;
    float factor, vmax;
    int wrap;
    const float4 *cx, *cy;
    const int *px, *py;
    int dstW, dstH;
    int tile_w, tile_h;      // destination tile (tile_w = 1 << tile_shift)
    int tile_shift;
    int win_w, win_h;        // allocated window (upper bound, even)
    int src_kind;            // GS_*
    int ch;                  // components resampled (3 for yuv sources)
    int dst_code;            // D_* for yuv sources; ignored for packed (same layout as src)
    int aligned;             // yuv sources: planes allow the 2x2 block loads
    int dst_vec;             // 4-component destination pixels may be stored as one word
};

template <int ST> struct SmemPix;
template <> struct SmemPix<1> { typedef float T; };
template <> struct SmemPix<2> { typedef float2 T; };
template <> struct SmemPix<4> { typedef float4 T; };

__device__ __forceinline__ float gen_chain(float w0, float w1, float w2, float w3, float p0, float p1, float p2, float p3) {
    float t = __fmul_rn(w1, p1);
    t = __fmaf_rn(w0, p0, t); t = __fmaf_rn(w2, p2, t); t = __fmaf_rn(w3, p3, t);
    return t;
}
__device__ __forceinline__ f2 gen_chain2(float w0, float w1, float w2, float w3, f2 p0, f2 p1, f2 p2, f2 p3) {
    f2 t = mul2(bc(w1), p1);
    t = fma2(bc(w0), p0, t); t = fma2(bc(w2), p2, t); t = fma2(bc(w3), p3, t);
    return t;
}
template <typename T> struct GenTap;
template <> struct GenTap<float> {
    static __device__ __forceinline__ float run(const float4 &w, float a, float b, float c, float d) { return gen_chain(w.x, w.y, w.z, w.w, a, b, c, d); }
};
template <> struct GenTap<float2> {
    static __device__ __forceinline__ float2 run(const float4 &w, float2 a, float2 b, float2 c, float2 d) {
        float2 r; upk(gen_chain2(w.x, w.y, w.z, w.w, pk(a.x, a.y), pk(b.x, b.y), pk(c.x, c.y), pk(d.x, d.y)), r.x, r.y); return r;
    }
};
template <> struct GenTap<float4> {
    static __device__ __forceinline__ float4 run(const float4 &w, float4 a, float4 b, float4 c, float4 d) {
        float4 r;
        upk(gen_chain2(w.x, w.y, w.z, w.w, pk(a.x, a.y), pk(b.x, b.y), pk(c.x, c.y), pk(d.x, d.y)), r.x, r.y);
        upk(gen_chain2(w.x, w.y, w.z, w.w, pk(a.z, a.w), pk(b.z, b.w), pk(c.z, c.w), pk(d.z, d.w)), r.z, r.w);
        return r;
    }
};

// quantised CSC result -> the sample the resize stage reads: R-B normalised, R-A integer-valued and clamped
template <int RA>
__device__ __forceinline__ float gen_sample(float r, const GenParams &P) {
    const float j = __fadd_rn(__fadd_rz(r, GMATB_MAGIC), -GMATB_MAGIC);   // trunc for r >= 0
    if (RA) return fminf(fmaxf(j, 0.f), P.vmax);
    return fma_sat(j, P.nk.khi, __fmul_rn(j, P.nk.klo));
}

// SRC: GS_*; SBITS: sample width of the source; CH: components per pixel (yuv sources: 3);
// RA: 0 = R-B arithmetic (normalised samples, truncating store), 1 = R-A / nearest arithmetic
// (integer-valued samples, rint + saturate).
template <int SRC, int SBITS, int CH, int RA>
__global__ void __launch_bounds__(256) generic_scale_kernel(const GenParams P) {
    constexpr int ST = CH == 3 ? 4 : CH;
    typedef typename SmemPix<ST>::T Pix;
    extern __shared__ float4 smem_raw[];
    const int tid = threadIdx.x;
    const int tx = tid & 31, tyy = tid >> 5;                 // 32 x 8 threads: no integer division in the loops
    const long long fz = blockIdx.z;
    const int xo0 = blockIdx.x * P.tile_w, yo0 = blockIdx.y * P.tile_h;
    const int tw = min(P.tile_w, P.dstW - xo0), th = min(P.tile_h, P.dstH - yo0);
    // Every table read this thread will need is issued here, ahead of stage A: most of the L1 is carved out
    // as shared memory, so these come from L2, and a load issued after a barrier stalls the whole CTA.
    // tile_w is a power of two <= 32 and divides the CTA size: a thread keeps its output column.
    const int tws = P.tile_shift;
    const int xo = tid & (P.tile_w - 1), ty0 = tid >> tws;
    const bool col_ok = xo < tw;
    const int pxa = __ldg(P.px + xo0), pxb = __ldg(P.px + xo0 + tw - 1), pya = __ldg(P.py + yo0), pyb = __ldg(P.py + yo0 + th - 1);
    const float4 wxv = __ldg(P.cx + xo0 + (col_ok ? xo : 0));
    const int pxo = __ldg(P.px + xo0 + (col_ok ? xo : 0));
    const float4 wyv0 = __ldg(P.cy + yo0 + min(ty0, th - 1));
    const int pyo0 = __ldg(P.py + yo0 + min(ty0, th - 1));
    // window origin rounded down (x to 4, y to 2) so that yuv sources convert in aligned 4x2 groups
    const int wx0 = pxa & ~3, wy0 = pya & ~1;
    const int ww = pxb + 4 - wx0, wh = pyb + 4 - wy0;
    const int wws = P.win_w;
    // window: 1- and 2-component pixels interleaved [win_h][win_w]; 3- and 4-component pixels PLANAR
    // [CH][win_h][win_w] (a lane's 4 consecutive columns are one conflict-free STS.128 per component, and a
    // tap run is 4 consecutive floats); horizontal results always as Pix [win_h][tile_w]
    Pix *Pw = reinterpret_cast<Pix *>(smem_raw);
    float *Pf = reinterpret_cast<float *>(smem_raw);
    const size_t plane = (size_t)P.win_w * P.win_h;
    Pix *Hs = ST == 4 ? reinterpret_cast<Pix *>(Pf + (((size_t)CH * plane + 3) & ~(size_t)3)) : Pw + plane;
    constexpr int SB = SBITS / 8;
    const int W = P.src.w, H = P.src.h;

    // ---- stage A ------------------------------------------------------------------
    if (SRC == GS_PACKED) {
        const uint8_t *base = P.src.pl[0].p + fz * P.src.pl[0].bstride;
        for (int wy = tyy; wy < wh; wy += 8) {
            const int sy = min(max(wy0 + wy, 0), H - 1);
            const uint8_t *row = base + (size_t)sy * P.src.pl[0].pitch;
            for (int wx = tx; wx < ww; wx += 32) {
                const int sx = min(max(wx0 + wx, 0), W - 1);
                const uint8_t *q = row + (size_t)sx * CH * SB;
                float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int c = 0; c < CH; c++) {
                    // int -> float through the 2^23 magic number (exact; the I2F unit is 8x slower than the FP32 pipe)
                    const unsigned jv = SBITS == 8 ? (unsigned)q[c] : (unsigned)reinterpret_cast<const uint16_t *>(q)[c];
                    const float j = __uint_as_float(0x4B000000u | jv) - GMATB_MAGIC;
                    o[c] = RA ? j : fma_sat(j, P.nk.khi, __fmul_rn(j, P.nk.klo));
                }
                if (ST == 4) {
#pragma unroll
                    for (int c = 0; c < CH; c++) Pf[c * plane + (size_t)wy * wws + wx] = o[c];
                } else {
                    Pix v;
                    if (ST == 1) *reinterpret_cast<float *>(&v) = o[0];
                    else *reinterpret_cast<float2 *>(&v) = make_float2(o[0], o[1]);
                    Pw[(size_t)wy * wws + wx] = v;
                }
            }
        }
    } else {
        constexpr bool FMAF = SBITS == 16;                      // the reference's P016 template, see csc_core.cuh
        const float low = SBITS == 8 ? 16.f : 4096.f, mid = SBITS == 8 ? 128.f : 32768.f;
        const uint8_t *by = P.src.pl[0].p + fz * P.src.pl[0].bstride;
        const uint8_t *bu = P.src.pl[1].p + fz * P.src.pl[1].bstride;
        const uint8_t *bv = SRC == GS_I420 ? P.src.pl[2].p + fz * P.src.pl[2].bstride : nullptr;
        const int ngx = (ww + 3) >> 2, nby = (wh + 1) >> 1;
        constexpr float YB = -(GMATB_MAGIC + (SBITS == 8 ? 16.f : 4096.f)), CB = -(GMATB_MAGIC + (SBITS == 8 ? 128.f : 32768.f));
        // a warp walks row pairs, a lane a group of 4 columns x 2 rows (two chroma samples).  Two groups
        // per trip: both groups' loads are issued before either is converted (one DRAM round trip, not two)
        struct Raw { uint32_t y0[SB], y1[SB], c[SB], c2[SB]; };
        auto inside = [&](int X, int Y) { return P.aligned && X >= 0 && X + 3 < W && Y >= 0 && Y + 1 < H; };
        auto fetch = [&](int X, int Y, Raw &R) {
            const uint8_t *q0 = by + (size_t)Y * P.src.pl[0].pitch + (size_t)X * SB;
            const uint8_t *q1 = q0 + P.src.pl[0].pitch;
#pragma unroll
            for (int k = 0; k < SB; k++) { R.y0[k] = __ldg(reinterpret_cast<const uint32_t *>(q0) + k); R.y1[k] = __ldg(reinterpret_cast<const uint32_t *>(q1) + k); }
            if (SRC == GS_NV12) {
                const uint8_t *qc = bu + (size_t)(Y >> 1) * P.src.pl[1].pitch + (size_t)X * SB;
#pragma unroll
                for (int k = 0; k < SB; k++) { R.c[k] = __ldg(reinterpret_cast<const uint32_t *>(qc) + k); R.c2[k] = 0; }
            } else {
                // the U and V planes have their own strides (linesize[1] / linesize[2])
                const size_t co = (size_t)(Y >> 1) * P.src.pl[1].pitch + (size_t)(X >> 1) * SB;
                const size_t cv = (size_t)(Y >> 1) * P.src.pl[2].pitch + (size_t)(X >> 1) * SB;
                if (SBITS == 8) { R.c[0] = __ldg(reinterpret_cast<const uint16_t *>(bu + co)); R.c2[0] = __ldg(reinterpret_cast<const uint16_t *>(bv + cv)); }
                else { R.c[0] = __ldg(reinterpret_cast<const uint32_t *>(bu + co)); R.c2[0] = __ldg(reinterpret_cast<const uint32_t *>(bv + cv)); R.c[SB - 1] = R.c[0]; R.c2[SB - 1] = R.c2[0]; }
            }
        };
        auto convert = [&](const Raw &R, int lx, int ly) {
            float yt[4], yb[4], um[2], vm[2];
            if (SBITS == 8) {
                yt[0] = byte_magic<0>(R.y0[0]); yt[1] = byte_magic<1>(R.y0[0]); yt[2] = byte_magic<2>(R.y0[0]); yt[3] = byte_magic<3>(R.y0[0]);
                yb[0] = byte_magic<0>(R.y1[0]); yb[1] = byte_magic<1>(R.y1[0]); yb[2] = byte_magic<2>(R.y1[0]); yb[3] = byte_magic<3>(R.y1[0]);
                if (SRC == GS_NV12) { um[0] = byte_magic<0>(R.c[0]); vm[0] = byte_magic<1>(R.c[0]); um[1] = byte_magic<2>(R.c[0]); vm[1] = byte_magic<3>(R.c[0]); }
                else { um[0] = byte_magic<0>(R.c[0]); um[1] = byte_magic<1>(R.c[0]); vm[0] = byte_magic<0>(R.c2[0]); vm[1] = byte_magic<1>(R.c2[0]); }
            } else {
                yt[0] = half_magic<0>(R.y0[0]); yt[1] = half_magic<1>(R.y0[0]); yt[2] = half_magic<0>(R.y0[SB - 1]); yt[3] = half_magic<1>(R.y0[SB - 1]);
                yb[0] = half_magic<0>(R.y1[0]); yb[1] = half_magic<1>(R.y1[0]); yb[2] = half_magic<0>(R.y1[SB - 1]); yb[3] = half_magic<1>(R.y1[SB - 1]);
                if (SRC == GS_NV12) { um[0] = half_magic<0>(R.c[0]); vm[0] = half_magic<1>(R.c[0]); um[1] = half_magic<0>(R.c[SB - 1]); vm[1] = half_magic<1>(R.c[SB - 1]); }
                else { um[0] = half_magic<0>(R.c[0]); um[1] = half_magic<1>(R.c[0]); vm[0] = half_magic<0>(R.c2[0]); vm[1] = half_magic<1>(R.c2[0]); }
            }
            float o[2][3][4];                              // [row][component][column]
#pragma unroll
            for (int h = 0; h < 2; h++) {                  // the two chroma samples of the group
                float fu, fv;
                upk(add2(pk(um[h], vm[h]), bc(CB)), fu, fv);
                const ChromaTerms t = chroma_terms<false, FMAF>(fu, fv, P.M);
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    const f2 fy2 = add2(pk(r ? yb[2 * h] : yt[2 * h], r ? yb[2 * h + 1] : yt[2 * h + 1]), bc(YB));
                    f2 cc[3];
                    csc_pair_f<false, FMAF>(fy2, t, P.M, cc[0], cc[1], cc[2]);
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        if (RA) {
                            float a, b; upk(cc[c], a, b);
                            o[r][c][2 * h] = gen_sample<RA>(a, P); o[r][c][2 * h + 1] = gen_sample<RA>(b, P);
                        } else {
                            // packed quantise + normalise (resample_core.cuh): the same value as gen_sample, RN(j/max)
                            upk(quant_norm2(cc[c], P.nk), o[r][c][2 * h], o[r][c][2 * h + 1]);
                        }
                    }
                }
            }
            // (the allocated window is a multiple of 4 wide and of 2 high: no bounds checks needed)
#pragma unroll
            for (int r = 0; r < 2; r++)
#pragma unroll
                for (int c = 0; c < 3; c++)
                    *reinterpret_cast<float4 *>(Pf + c * plane + (size_t)(ly + r) * wws + lx) = make_float4(o[r][c][0], o[r][c][1], o[r][c][2], o[r][c][3]);
        };
        for (int byi = tyy; byi < nby; byi += 16)
        for (int gxi = tx; gxi < ngx; gxi += 32) {
            const int X = wx0 + 4 * gxi, Ya = wy0 + 2 * byi, Yb = Ya + 16;
            const bool fa = inside(X, Ya), fb = byi + 8 < nby && inside(X, Yb);
            Raw Ra, Rb;
            if (fa) fetch(X, Ya, Ra);
            if (fb) fetch(X, Yb, Rb);
            if (fa) convert(Ra, 4 * gxi, 2 * byi);
            if (fb) convert(Rb, 4 * gxi, 2 * byi + 16);
#pragma unroll 1
            for (int half = 0; half < 2; half++) {
            if (half ? (fb || byi + 8 >= nby) : fa) continue;
            const int lx = 4 * gxi, ly = 2 * byi + 16 * half;
            {
                // group on (or over) the frame edge: each window cell on its own with clamped coordinates
#pragma unroll 1
                for (int k = 0; k < 8; k++) {
                    const int cx_ = lx + (k & 3), cy_ = ly + (k >> 2);
                    if (cx_ >= ww || cy_ >= wh) continue;
                    const int sx = min(max(wx0 + cx_, 0), W - 1), sy = min(max(wy0 + cy_, 0), H - 1);
                    unsigned y, u, v;
                    const uint8_t *qy = by + (size_t)sy * P.src.pl[0].pitch + (size_t)sx * SB;
                    y = SBITS == 8 ? *qy : *reinterpret_cast<const uint16_t *>(qy);
                    const int ccx = sx >> 1, ccy = sy >> 1;
                    if (SRC == GS_NV12) {
                        const uint8_t *q = bu + (size_t)ccy * P.src.pl[1].pitch + (size_t)ccx * 2 * SB;
                        if (SBITS == 8) { u = q[0]; v = q[1]; }
                        else { u = reinterpret_cast<const uint16_t *>(q)[0]; v = reinterpret_cast<const uint16_t *>(q)[1]; }
                    } else {
                        const size_t co = (size_t)ccy * P.src.pl[1].pitch + (size_t)ccx * SB;
                        const size_t cv = (size_t)ccy * P.src.pl[2].pitch + (size_t)ccx * SB;
                        if (SBITS == 8) { u = bu[co]; v = bv[cv]; }
                        else { u = *reinterpret_cast<const uint16_t *>(bu + co); v = *reinterpret_cast<const uint16_t *>(bv + cv); }
                    }
                    const float fy = __uint_as_float(0x4B000000u | y) - (GMATB_MAGIC + low);
                    const float fu = __uint_as_float(0x4B000000u | u) - (GMATB_MAGIC + mid);
                    const float fv = __uint_as_float(0x4B000000u | v) - (GMATB_MAGIC + mid);
                    // scalar form of csc_pair_f (same IEEE operations)
                    const float *m = P.M.m;
                    float r, g, b;
                    if (FMAF) {
                        r = __fmaf_rn(fv, m[2], __fmaf_rn(fy, m[0], __fmul_rn(fu, m[1])));
                        g = __fmaf_rn(fv, m[5], __fmaf_rn(fy, m[3], __fmul_rn(fu, m[4])));
                        b = __fmaf_rn(fv, m[8], __fmaf_rn(fy, m[6], __fmul_rn(fu, m[7])));
                    } else {
                        r = __fadd_rn(__fmaf_rn(fy, m[0], __fmul_rn(fu, m[1])), __fmul_rn(fv, m[2]));
                        g = __fadd_rn(__fmaf_rn(fy, m[3], __fmul_rn(fu, m[4])), __fmul_rn(fv, m[5]));
                        b = __fadd_rn(__fmaf_rn(fy, m[6], __fmul_rn(fu, m[7])), __fmul_rn(fv, m[8]));
                    }
                    Pf[(size_t)cy_ * wws + cx_] = gen_sample<RA>(r, P);
                    Pf[plane + (size_t)cy_ * wws + cx_] = gen_sample<RA>(g, P);
                    Pf[2 * plane + (size_t)cy_ * wws + cx_] = gen_sample<RA>(b, P);
                }
            }
            }
        }
    }
    __syncthreads();
    // ---- stage B: horizontal ------------------------------------------------------
    if (col_ok) {
        if (ST == 4) {
            const float *q = Pf + (pxo - wx0);
            for (int wy = ty0; wy < wh; wy += 256 >> tws) {
                const float *r = q + (size_t)wy * wws;
                float4 hv;
                // components 0 and 1 ride one f32x2 chain, the rest scalar chains
                upk(gen_chain2(wxv.x, wxv.y, wxv.z, wxv.w, pk(r[0], r[plane]), pk(r[1], r[plane + 1]), pk(r[2], r[plane + 2]), pk(r[3], r[plane + 3])), hv.x, hv.y);
                hv.z = gen_chain(wxv.x, wxv.y, wxv.z, wxv.w, r[2 * plane], r[2 * plane + 1], r[2 * plane + 2], r[2 * plane + 3]);
                hv.w = CH == 4 ? gen_chain(wxv.x, wxv.y, wxv.z, wxv.w, r[3 * plane], r[3 * plane + 1], r[3 * plane + 2], r[3 * plane + 3]) : 0.f;
                *reinterpret_cast<float4 *>(&Hs[(wy << tws) + xo]) = hv;
            }
        } else {
            const Pix *q = Pw + (pxo - wx0);
            for (int wy = ty0; wy < wh; wy += 256 >> tws) {
                const Pix *r = q + (size_t)wy * wws;
                Hs[(wy << tws) + xo] = GenTap<Pix>::run(wxv, r[0], r[1], r[2], r[3]);
            }
        }
    }
    __syncthreads();
    // ---- stage C: vertical + store ------------------------------------------------
    const int out_ch = SRC == GS_PACKED ? CH : dst_bpp(P.dst_code) / (dst_is16(P.dst_code) ? 2 : 1);
    const int smax = SBITS == 8 ? 255 : 65535;
    for (int ty = ty0; ty < th; ty += 256 >> tws) {
        if (col_ok) {
            const float4 w = ty == ty0 ? wyv0 : __ldg(P.cy + yo0 + ty);
            const Pix *h = Hs + (size_t)((ty == ty0 ? pyo0 : __ldg(P.py + yo0 + ty)) - wy0) * P.tile_w + xo;
            const Pix tv = GenTap<Pix>::run(w, h[0], h[P.tile_w], h[2 * P.tile_w], h[3 * P.tile_w]);
            const float *tf = reinterpret_cast<const float *>(&tv);
            int o[CH];
#pragma unroll
            for (int c = 0; c < CH; c++) {
                const float t = tf[c];
                int v;
                if (RA) v = (int)fminf(fmaxf(rintf(t), 0.f), P.vmax);
                else {
                    // fmaxf(NaN, -1) = -1: a NaN (the reference's Lanczos coefficients are 0/0 when the phase
                    // is ~1e-8 but not 0: all four __sinf taps flush to zero) stores 0, like cvt.rzi.u32.f32(NaN)
                    v = trunc_i(fmaxf(__fmul_rn(t, P.factor), -1.0f));
                    v = P.wrap ? (max(v, 0) & smax) : clamp_i(v, smax);
                }
                o[c] = v;
            }
            uint8_t *pd = P.dst.pl[0].p + fz * P.dst.pl[0].bstride + (size_t)(yo0 + ty) * P.dst.pl[0].pitch
                        + (size_t)(xo0 + xo) * out_ch * SB;
            if (SRC == GS_PACKED) {
#pragma unroll
                for (int c = 0; c < CH; c++) {
                    if (SBITS == 8) pd[c] = o[c]; else reinterpret_cast<uint16_t *>(pd)[c] = o[c];
                }
            } else {
                // alpha of 4-channel outputs: the same chain over the constant 255 the reference's CSC
                // writes (yuv2rgb_cuda.cu:89) -- i.e. p = 1.0 in R-B, 255 in R-A
                int a = 255;
                if (out_ch == 4) {
                    const float one = RA ? 255.f : (SBITS == 8 ? 1.0f : 255.0f / 65535.0f);
                    const float ah = gen_chain(wxv.x, wxv.y, wxv.z, wxv.w, one, one, one, one);
                    const float av = gen_chain(w.x, w.y, w.z, w.w, ah, ah, ah, ah);
                    if (RA) a = (int)fminf(fmaxf(rintf(av), 0.f), P.vmax);
                    else { a = trunc_i(fmaxf(__fmul_rn(av, P.factor), -1.0f)); a = P.wrap ? (max(a, 0) & smax) : clamp_i(a, smax); }
                }
                const bool sw = dst_swap(P.dst_code);
                const int c0 = sw ? o[CH - 1] : o[0], c2 = sw ? o[0] : o[CH - 1];
                if (SBITS == 8) {
                    if (out_ch == 4 && P.dst_vec) *reinterpret_cast<uint32_t *>(pd) = (uint32_t)c0 | ((uint32_t)o[CH > 1 ? 1 : 0] << 8) | ((uint32_t)c2 << 16) | ((uint32_t)a << 24);
                    else { pd[0] = c0; pd[1] = o[CH > 1 ? 1 : 0]; pd[2] = c2; if (out_ch == 4) pd[3] = a; }
                } else {
                    uint16_t *q = reinterpret_cast<uint16_t *>(pd);
                    if (out_ch == 4 && P.dst_vec) *reinterpret_cast<uint2 *>(pd) = make_uint2((uint32_t)c0 | ((uint32_t)o[CH > 1 ? 1 : 0] << 16), (uint32_t)c2 | ((uint32_t)a << 16));
                    else { q[0] = c0; q[1] = o[CH > 1 ? 1 : 0]; q[2] = c2; if (out_ch == 4) q[3] = a; }
                }
            }
        }
    }
}

}  // namespace gmatb
