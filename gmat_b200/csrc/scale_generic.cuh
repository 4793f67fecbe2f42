// scale_generic.cuh -- any-ratio, any-size scaling kernel (the safety net behind the
// fused 2:1 kernel, and the path for rgb->rgb, rgb->yuv and per-plane yuv scaling).
//
// A CTA produces a tile of the destination.  Stage A fills shared memory with the
// source window the tile taps (coordinates clamped to the image like the reference's
// texture fetches), already converted to the sample the resize stage of the reference
// would read: either a packed source component, or -- for yuv 4:2:0 sources -- the
// quantised result of the colour conversion (P-ORDER, swscale_cuda.c:352-370).
// Stage B applies the horizontal 4-tap chain once per (window row, output column),
// stage C the vertical chain; both in the reference's operand order (resample_core.cuh).
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
    int tile_w, tile_h;      // destination tile
    int win_w, win_h;        // allocated window (upper bound)
    int src_kind;            // GS_*
    int ch;                  // components resampled (3 for yuv sources)
    int dst_code;            // D_* for yuv sources; ignored for packed (same layout as src)
    int sparse;
};

// SBITS: sample width of the source.  RA: 0 = R-B arithmetic (normalised samples,
// truncating store), 1 = R-A / nearest arithmetic (integer-valued samples, rint + saturate).
template <int SBITS, int RA>
__global__ void __launch_bounds__(256) generic_scale_kernel(const GenParams P) {
    extern __shared__ float smem[];
    const int tid = threadIdx.x;
    const long long fz = blockIdx.z;
    const int xo0 = blockIdx.x * P.tile_w, yo0 = blockIdx.y * P.tile_h;
    const int tw = min(P.tile_w, P.dstW - xo0), th = min(P.tile_h, P.dstH - yo0);
    const int wx0 = P.px[xo0], wy0 = P.py[yo0];
    const int ww = P.px[xo0 + tw - 1] + 4 - wx0, wh = P.py[yo0 + th - 1] + 4 - wy0;
    const int CH = P.ch;
    float *Pw = smem;                                   // [wh][ww][CH]
    float *Hs = smem + (size_t)P.win_w * P.win_h * CH;  // [wh][tile_w][CH]
    constexpr int SB = SBITS / 8;
    const int W = P.src.w, H = P.src.h;

    // ---- stage A ------------------------------------------------------------------
    const int tx = tid & 31, tyy = tid >> 5;                 // 32 x 8 threads: no integer division in the loops
    for (int wy = tyy; wy < wh; wy += 8)
    for (int wx = tx; wx < ww; wx += 32) {
        const int i = wy * ww + wx;
        const int sx = min(max(wx0 + wx, 0), W - 1), sy = min(max(wy0 + wy, 0), H - 1);
        float *o = Pw + (size_t)i * CH;
        if (P.src_kind == GS_PACKED) {
            const uint8_t *q = P.src.pl[0].p + fz * P.src.pl[0].bstride + (size_t)sy * P.src.pl[0].pitch + (size_t)sx * CH * SB;
            for (int c = 0; c < CH; c++) {
                // int -> float through the 2^23 magic number (exact; the I2F unit is 8x slower than the FP32 pipe)
                const unsigned jv = SBITS == 8 ? (unsigned)q[c] : (unsigned)reinterpret_cast<const uint16_t *>(q)[c];
                const float j = __uint_as_float(0x4B000000u | jv) - GMATB_MAGIC;
                o[c] = RA ? j : fma_sat(j, P.nk.khi, __fmul_rn(j, P.nk.klo));
            }
        } else {
            unsigned y, u, v;
            const uint8_t *qy = P.src.pl[0].p + fz * P.src.pl[0].bstride + (size_t)sy * P.src.pl[0].pitch + (size_t)sx * SB;
            y = SBITS == 8 ? *qy : *reinterpret_cast<const uint16_t *>(qy);
            const int cx = sx >> 1, cy = sy >> 1;
            if (P.src_kind == GS_NV12) {
                const uint8_t *q = P.src.pl[1].p + fz * P.src.pl[1].bstride + (size_t)cy * P.src.pl[1].pitch + (size_t)cx * 2 * SB;
                if (SBITS == 8) { u = q[0]; v = q[1]; }
                else { u = reinterpret_cast<const uint16_t *>(q)[0]; v = reinterpret_cast<const uint16_t *>(q)[1]; }
            } else {
                const uint8_t *qu = P.src.pl[1].p + fz * P.src.pl[1].bstride + (size_t)cy * P.src.pl[1].pitch + (size_t)cx * SB;
                const uint8_t *qv = P.src.pl[2].p + fz * P.src.pl[2].bstride + (size_t)cy * P.src.pl[2].pitch + (size_t)cx * SB;
                if (SBITS == 8) { u = *qu; v = *qv; }
                else { u = *reinterpret_cast<const uint16_t *>(qu); v = *reinterpret_cast<const uint16_t *>(qv); }
            }
            const float low = SBITS == 8 ? 16.f : 4096.f, mid = SBITS == 8 ? 128.f : 32768.f;
            const float fy = __uint_as_float(0x4B000000u | y) - (GMATB_MAGIC + low);
            const float fu = __uint_as_float(0x4B000000u | u) - (GMATB_MAGIC + mid);
            const float fv = __uint_as_float(0x4B000000u | v) - (GMATB_MAGIC + mid);
            // scalar form of csc_pair_f (same IEEE operations)
            const float *m = P.M.m;
            float r, g, b;
            if (SBITS == 16) {   // FMA form (the reference's P016 template), see csc_core.cuh
                r = __fmaf_rn(fv, m[2], __fmaf_rn(fy, m[0], __fmul_rn(fu, m[1])));
                g = __fmaf_rn(fv, m[5], __fmaf_rn(fy, m[3], __fmul_rn(fu, m[4])));
                b = __fmaf_rn(fv, m[8], __fmaf_rn(fy, m[6], __fmul_rn(fu, m[7])));
            } else {
                r = __fadd_rn(__fmaf_rn(fy, m[0], __fmul_rn(fu, m[1])), __fmul_rn(fv, m[2]));
                g = __fadd_rn(__fmaf_rn(fy, m[3], __fmul_rn(fu, m[4])), __fmul_rn(fv, m[5]));
                b = __fadd_rn(__fmaf_rn(fy, m[6], __fmul_rn(fu, m[7])), __fmul_rn(fv, m[8]));
            }
            float rgb[3] = {r, g, b};
            for (int c = 0; c < 3; c++) {
                float j = __fadd_rn(__fadd_rz(rgb[c], GMATB_MAGIC), -GMATB_MAGIC);   // trunc for r >= 0
                if (RA) o[c] = fminf(fmaxf(j, 0.f), P.vmax);
                else o[c] = fma_sat(j, P.nk.khi, __fmul_rn(j, P.nk.klo));
            }
        }
    }
    __syncthreads();
    // ---- stage B: horizontal ------------------------------------------------------
    for (int wy = tyy; wy < wh; wy += 8)
    for (int xo = tx; xo < tw; xo += 32) {
        const float4 w = P.cx[xo0 + xo];
        const float *p = Pw + ((size_t)wy * ww + (P.px[xo0 + xo] - wx0)) * CH;
        float *h = Hs + ((size_t)wy * P.tile_w + xo) * CH;
        for (int c = 0; c < CH; c++) {
            float t = __fmul_rn(w.y, p[CH + c]);
            t = __fmaf_rn(w.x, p[c], t);
            t = __fmaf_rn(w.z, p[2 * CH + c], t);
            t = __fmaf_rn(w.w, p[3 * CH + c], t);
            h[c] = t;
        }
    }
    __syncthreads();
    // ---- stage C: vertical + store ------------------------------------------------
    const int out_ch = P.src_kind == GS_PACKED ? CH : dst_bpp(P.dst_code) / (dst_is16(P.dst_code) ? 2 : 1);
    const int smax = SBITS == 8 ? 255 : 65535;
    for (int ty = tyy; ty < th; ty += 8)
    for (int xo = tx; xo < tw; xo += 32) {
        const float4 w = P.cy[yo0 + ty];
        const float *h = Hs + ((size_t)(P.py[yo0 + ty] - wy0) * P.tile_w + xo) * CH;
        const size_t rs = (size_t)P.tile_w * CH;
        int o[4];
        for (int c = 0; c < CH; c++) {
            float t = __fmul_rn(w.y, h[rs + c]);
            t = __fmaf_rn(w.x, h[c], t);
            t = __fmaf_rn(w.z, h[2 * rs + c], t);
            t = __fmaf_rn(w.w, h[3 * rs + c], t);
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
        if (P.src_kind == GS_PACKED) {
            for (int c = 0; c < CH; c++) {
                if (SBITS == 8) pd[c] = o[c]; else reinterpret_cast<uint16_t *>(pd)[c] = o[c];
            }
        } else {
            // alpha of 4-channel outputs: the same chain over the constant 255 the reference's CSC
            // writes (yuv2rgb_cuda.cu:89) -- i.e. p = 1.0 in R-B, 255 in R-A
            int a = 255;
            if (out_ch == 4) {
                const float4 wxv = P.cx[xo0 + xo];
                const float one = RA ? 255.f : (SBITS == 8 ? 1.0f : 255.0f / 65535.0f);
                float ah = __fmul_rn(wxv.y, one);
                ah = __fmaf_rn(wxv.x, one, ah); ah = __fmaf_rn(wxv.z, one, ah); ah = __fmaf_rn(wxv.w, one, ah);
                float av = __fmul_rn(w.y, ah);
                av = __fmaf_rn(w.x, ah, av); av = __fmaf_rn(w.z, ah, av); av = __fmaf_rn(w.w, ah, av);
                if (RA) a = (int)fminf(fmaxf(rintf(av), 0.f), P.vmax);
                else { a = trunc_i(fmaxf(__fmul_rn(av, P.factor), -1.0f)); a = P.wrap ? (max(a, 0) & smax) : clamp_i(a, smax); }
            }
            const bool sw = dst_swap(P.dst_code);
            const int c0 = sw ? o[2] : o[0], c2 = sw ? o[0] : o[2];
            if (SBITS == 8) { pd[0] = c0; pd[1] = o[1]; pd[2] = c2; if (out_ch == 4) pd[3] = a; }
            else { uint16_t *q = reinterpret_cast<uint16_t *>(pd); q[0] = c0; q[1] = o[1]; q[2] = c2; if (out_ch == 4) q[3] = a; }
        }
    }
}

}  // namespace gmatb
