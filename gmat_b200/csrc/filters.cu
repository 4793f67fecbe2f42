// filters.cu -- crop / flip / rotate / gaussian / median on packed 8-bit images.
//
// These replace the CV-CUDA operators the reference's libavfilter wrappers submit
// (vf_crop_nvcv.c:277 cvcudaCustomCropSubmit, vf_flip_nvcv.c:251 cvcudaFlipSubmit,
// vf_rotate_nvcv.c:275 cvcudaRotateSubmit, vf_smooth_nvcv.c:290/294
// cvcudaGaussianSubmit / cvcudaMedianBlurSubmit).  CV-CUDA is a closed third-party
// dependency of the reference (not in /root/reference), so the arithmetic of rotate /
// gaussian / median is pinned by OUR restatement (oracle/gmat_oracle.c, SURVEY 8c
// P-FILTERS); crop and flip are exact copies.  Every float operation below is written
// with an explicit-rounding intrinsic so that the oracle's C code (compiled with
// -ffp-contract=off) performs the same IEEE operations in the same order.
#include <cmath>
#include <cstring>
#include <algorithm>
#include <vector>
#include "common.cuh"

namespace gmatb {

struct PImg {          // packed image, bpp bytes per pixel
    uint8_t *p; int pitch; long long bstride; int w, h, bpp;
};

static int packed_bpp(int fmt) {
    switch (fmt) {
    case GMATB_FMT_RGB24: case GMATB_FMT_BGR24: return 3;
    case GMATB_FMT_RGBA: case GMATB_FMT_BGRA: case GMATB_FMT_RGB0: case GMATB_FMT_BGR0:
    case GMATB_FMT_0RGB: case GMATB_FMT_0BGR: return 4;
    default: return 0;
    }
}
static bool to_pimg(const GmatbImage *g, PImg *o) {
    if (!g || !g->data[0] || g->width <= 0 || g->height <= 0) return false;
    o->bpp = packed_bpp(g->format);
    if (!o->bpp) return false;
    o->p = (uint8_t *)g->data[0]; o->pitch = g->linesize[0]; o->w = g->width; o->h = g->height;
    o->bstride = g->batch > 1 ? g->batch_stride[0] : 0;
    return true;
}
static inline int nbatch(const GmatbImage *g) { return g->batch > 1 ? g->batch : 1; }

// ---------------------------------------------------------------------------
// crop: dst(x,y) = src(x+cx, y+cy).  16 destination bytes per thread; the source
// run starts at an arbitrary byte, so it is fetched as aligned words and realigned
// with a funnel shift (5 x LDG.32 + 1 x STG.128 per 16 bytes).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) crop_kernel(PImg s, PImg d, int cx, int cy, int dst_vec) {
    const int y = blockIdx.y * 8 + threadIdx.y;
    const int b0 = (blockIdx.x * 32 + threadIdx.x) * 16;      // first destination byte of this thread
    const int rowb = d.w * d.bpp;
    if (y >= d.h || b0 >= rowb) return;
    const long long fz = blockIdx.z;
    const uint8_t *ps = s.p + fz * s.bstride + (size_t)(y + cy) * s.pitch + (size_t)cx * s.bpp + b0;
    uint8_t *pd = d.p + fz * d.bstride + (size_t)y * d.pitch + b0;
    if (dst_vec && b0 + 16 <= rowb) {
        const uintptr_t a = (uintptr_t)ps;
        // a 16-byte aligned window origin (e.g. the centred 1080p window of a 4K rgb24 frame: 2880 bytes into the row) is a plain
        // 128-bit copy; the branch is uniform over the launch
        if ((a & 15) == 0) { stg128(pd, ldg128(ps)); return; }
        const uint32_t *q = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
        const unsigned sh = (unsigned)(a & 3) * 8;
        uint32_t w0 = __ldcs(q), w1 = __ldcs(q + 1), w2 = __ldcs(q + 2), w3 = __ldcs(q + 3);
        uint32_t w4 = sh ? __ldcs(q + 4) : 0u;
        uint4 o;
        o.x = __funnelshift_r(w0, w1, sh); o.y = __funnelshift_r(w1, w2, sh);
        o.z = __funnelshift_r(w2, w3, sh); o.w = __funnelshift_r(w3, w4, sh);
        stg128(pd, o);
    } else {
        const int n = min(16, rowb - b0);
        for (int i = 0; i < n; i++) pd[i] = ps[i];
    }
}

// ---------------------------------------------------------------------------
// flip: code 0 = vertical (rows), > 0 = horizontal (columns), < 0 = both
// (cvcudaFlip / cv::flip convention used by vf_flip_nvcv.c:77-80)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) flip_rows_kernel(PImg s, PImg d, int vec) {
    const int y = blockIdx.y * 8 + threadIdx.y;
    const int b0 = (blockIdx.x * 32 + threadIdx.x) * 16;
    const int rowb = d.w * d.bpp;
    if (y >= d.h || b0 >= rowb) return;
    const long long fz = blockIdx.z;
    const uint8_t *ps = s.p + fz * s.bstride + (size_t)(s.h - 1 - y) * s.pitch + b0;
    uint8_t *pd = d.p + fz * d.bstride + (size_t)y * d.pitch + b0;
    if (vec && b0 + 16 <= rowb) stg128(pd, ldg128(ps));
    else { const int n = min(16, rowb - b0); for (int i = 0; i < n; i++) pd[i] = ps[i]; }
}

// mirror of 4 packed 3-byte pixels (12 bytes, words a0 a1 a2): p3 p2 p1 p0, 4 PRMTs
__device__ __forceinline__ void mirror4_rgb(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t &o0, uint32_t &o1, uint32_t &o2) {
    o0 = __byte_perm(a2, a1, 0x6321);                           // a9 a10 a11 a6
    o1 = __byte_perm(__byte_perm(a1, a2, 0x0043), a0, 0x3710);  // a7 a8 a3 a4
    o2 = __byte_perm(a1, a0, 0x6541);                           // a5 a0 a1 a2
}

// vec: 2 = 16 pixels per thread through 128-bit loads/stores (3-byte pixels: width % 16 == 0),
//      1 = 4 pixels per thread through 32-bit words (3-byte pixels, width % 4 == 0) / one 128-bit word (4-byte pixels),
//      0 = bytes.  PX = pixels per thread.
template <int BPP, int PX>
__global__ void __launch_bounds__(256) flip_cols_kernel(PImg s, PImg d, int also_rows, int vec) {
    const int y = blockIdx.y * 8 + threadIdx.y;
    const int x0 = (blockIdx.x * 32 + threadIdx.x) * PX;
    if (y >= d.h || x0 >= d.w) return;
    const long long fz = blockIdx.z;
    const int sy = also_rows ? s.h - 1 - y : y;
    const uint8_t *srow = s.p + fz * s.bstride + (size_t)sy * s.pitch;
    uint8_t *pd = d.p + fz * d.bstride + (size_t)y * d.pitch + (size_t)x0 * BPP;
    if (BPP == 3 && PX == 16) {          // host guarantees width % 16 == 0 and 16-byte aligned rows
        const uint8_t *q = srow + (size_t)(s.w - 16 - x0) * 3;
        const uint4 A = ldg128(q), B = ldg128(q + 16), Cw = ldg128(q + 32);
        uint4 o0, o1, o2;
        mirror4_rgb(Cw.y, Cw.z, Cw.w, o0.x, o0.y, o0.z);
        mirror4_rgb(B.z, B.w, Cw.x, o0.w, o1.x, o1.y);
        mirror4_rgb(A.w, B.x, B.y, o1.z, o1.w, o2.x);
        mirror4_rgb(A.x, A.y, A.z, o2.y, o2.z, o2.w);
        stg128(pd, o0); stg128(pd + 16, o1); stg128(pd + 32, o2);
        return;
    }
    if (BPP == 3 && vec && x0 + 4 <= d.w) {
        const uint8_t *q = srow + (size_t)(s.w - 4 - x0) * 3;
        uint32_t o0, o1, o2;
        mirror4_rgb(ldg32(q), ldg32(q + 4), ldg32(q + 8), o0, o1, o2);
        stg32(pd, o0); stg32(pd + 4, o1); stg32(pd + 8, o2);
        return;
    }
    if (BPP == 4 && vec && x0 + 4 <= d.w) {
        uint4 v = ldg128(srow + (size_t)(s.w - 4 - x0) * 4);
        stg128(pd, make_uint4(v.w, v.z, v.y, v.x));
        return;
    }
    for (int i = 0; i < PX && x0 + i < d.w; i++) {
        const uint8_t *q = srow + (size_t)(s.w - 1 - x0 - i) * BPP;
#pragma unroll
        for (int c = 0; c < BPP; c++) pd[i * BPP + c] = q[c];
    }
}

// ---------------------------------------------------------------------------
// rotate (P-FILTERS): for each destination pixel
//   dx = x - shift_x, dy = y - shift_y                         (double)
//   sx = (float)(dx*c - dy*s),  sy = (float)(dx*s + dy*c)      (double products, one subtraction/addition)
//   written only if -0.5 < sx < W and -0.5 < sy < H, else 0
// ---------------------------------------------------------------------------
struct RotParams { double c, s, shx, shy; int interp; };

__device__ __forceinline__ float cubic_w(float d) {   // Keys kernel, A = -0.75, d >= 0
    const float A = -0.75f;
    if (d <= 1.0f) return __fmaf_rn(__fmul_rn(__fmaf_rn(A + 2.0f, d, -(A + 3.0f)), d), d, 1.0f);
    if (d < 2.0f)  return __fmaf_rn(__fmaf_rn(__fmaf_rn(A, d, -5.0f * A), d, 8.0f * A), d, -4.0f * A);
    return 0.0f;
}

template <int BPP>
__global__ void __launch_bounds__(256) rotate_kernel(PImg s, PImg d, RotParams R) {
    // a warp covers a 16x2 block of the destination (a compact source footprint: fewer cache lines per
    // gather than a 32x1 row), the CTA 32x8
    const int lane = threadIdx.x, warp = threadIdx.y;
    const int x = blockIdx.x * 32 + (warp & 1) * 16 + (lane & 15);
    const int y = blockIdx.y * 8 + (warp >> 1) * 2 + (lane >> 4);
    if (x >= d.w || y >= d.h) return;
    const long long fz = blockIdx.z;
    const double dx = __dsub_rn((double)x, R.shx), dy = __dsub_rn((double)y, R.shy);
    const float sx = (float)__dsub_rn(__dmul_rn(dx, R.c), __dmul_rn(dy, R.s));
    const float sy = (float)__dadd_rn(__dmul_rn(dx, R.s), __dmul_rn(dy, R.c));
    const uint8_t *ps = s.p + fz * s.bstride;
    uint8_t *pd = d.p + fz * d.bstride + (size_t)y * d.pitch + (size_t)x * BPP;
    int out[BPP];
#pragma unroll
    for (int c = 0; c < BPP; c++) out[c] = 0;
    const int W = s.w, H = s.h;
    if (sx > -0.5f && sx < (float)W && sy > -0.5f && sy < (float)H) {
        if (R.interp == GMATB_INTERP_NEAREST) {
            const int x1 = min(__float2int_rz(__fadd_rn(sx, 0.5f)), W - 1);
            const int y1 = min(__float2int_rz(__fadd_rn(sy, 0.5f)), H - 1);
            const uint8_t *q = ps + (size_t)y1 * s.pitch + (size_t)x1 * BPP;
#pragma unroll
            for (int c = 0; c < BPP; c++) out[c] = q[c];
        } else if (R.interp == GMATB_INTERP_CUBIC) {
            const float fxf = floorf(sx), fyf = floorf(sy);
            const int ix = (int)fxf, iy = (int)fyf;
            float wx[4], wy[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                wx[i] = cubic_w(fabsf(__fsub_rn(sx, __fadd_rn(fxf, (float)(i - 1)))));
                wy[i] = cubic_w(fabsf(__fsub_rn(sy, __fadd_rn(fyf, (float)(i - 1)))));
            }
            float acc[BPP];
#pragma unroll
            for (int c = 0; c < BPP; c++) acc[c] = 0.0f;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int yy = min(max(iy - 1 + j, 0), H - 1);
                float rowacc[BPP];
#pragma unroll
                for (int c = 0; c < BPP; c++) rowacc[c] = 0.0f;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int xx = min(max(ix - 1 + i, 0), W - 1);
                    const uint8_t *q = ps + (size_t)yy * s.pitch + (size_t)xx * BPP;
#pragma unroll
                    for (int c = 0; c < BPP; c++) rowacc[c] = __fmaf_rn(wx[i], (float)q[c], rowacc[c]);
                }
#pragma unroll
                for (int c = 0; c < BPP; c++) acc[c] = __fmaf_rn(wy[j], rowacc[c], acc[c]);
            }
#pragma unroll
            for (int c = 0; c < BPP; c++) out[c] = min(max(__float2int_rn(acc[c]), 0), 255);
        } else {   // linear (and area)
            const int x1 = __float2int_rz(sx), y1 = __float2int_rz(sy);
            const int x2 = x1 + 1, y2 = y1 + 1;
            const int x2r = min(x2, W - 1), y2r = min(y2, H - 1);
            const float ax = __fsub_rn((float)x2, sx), bx = __fsub_rn(sx, (float)x1);
            const float ay = __fsub_rn((float)y2, sy), by = __fsub_rn(sy, (float)y1);
            const float w00 = __fmul_rn(ax, ay), w01 = __fmul_rn(bx, ay), w10 = __fmul_rn(ax, by), w11 = __fmul_rn(bx, by);
            // the two taps of a row are 2*BPP contiguous bytes: fetch them as aligned words + funnel shift
            // instead of 2*BPP byte loads (the gather is LSU-bound)
            float p00[BPP], p01[BPP], p10[BPP], p11[BPP];
            const bool wordable = (((uintptr_t)ps | (uintptr_t)s.pitch) & 3) == 0;
#pragma unroll
            for (int rr = 0; rr < 2; rr++) {
                const uint8_t *row = ps + (size_t)(rr ? y2r : y1) * s.pitch;
                float (&pa)[BPP] = rr ? p10 : p00;
                float (&pb)[BPP] = rr ? p11 : p01;
                if (wordable && x2r == x2 && x1 >= 0) {
                    const uintptr_t a = (uintptr_t)(row + (size_t)x1 * BPP);
                    const uint32_t *q = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
                    const unsigned sh = (unsigned)(a & 3) * 8;
                    const uint32_t w0 = __ldg(q), w1 = __ldg(q + 1);
                    const uint32_t w2 = (BPP == 4 || sh == 24) ? __ldg(q + 2) : 0u;       // bytes 6..7 of a 3-byte pair only when misaligned by 3
                    const uint32_t v0 = __funnelshift_r(w0, w1, sh), v1 = __funnelshift_r(w1, w2, sh);
                    // byte -> float through PRMT magic numbers (the I2F unit is 8x slower than the FP32 pipe)
                    if (BPP == 3) {
                        pa[0] = byte_magic<0>(v0) - GMATB_MAGIC; pa[1] = byte_magic<1>(v0) - GMATB_MAGIC; pa[2] = byte_magic<2>(v0) - GMATB_MAGIC;
                        pb[0] = byte_magic<3>(v0) - GMATB_MAGIC; pb[1] = byte_magic<0>(v1) - GMATB_MAGIC; pb[2] = byte_magic<1>(v1) - GMATB_MAGIC;
                    } else {
                        pa[0] = byte_magic<0>(v0) - GMATB_MAGIC; pa[1] = byte_magic<1>(v0) - GMATB_MAGIC;
                        pa[2] = byte_magic<2>(v0) - GMATB_MAGIC; pa[BPP - 1] = byte_magic<3>(v0) - GMATB_MAGIC;
                        pb[0] = byte_magic<0>(v1) - GMATB_MAGIC; pb[1] = byte_magic<1>(v1) - GMATB_MAGIC;
                        pb[2] = byte_magic<2>(v1) - GMATB_MAGIC; pb[BPP - 1] = byte_magic<3>(v1) - GMATB_MAGIC;
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < BPP; c++) { pa[c] = (float)row[(size_t)x1 * BPP + c]; pb[c] = (float)row[(size_t)x2r * BPP + c]; }
                }
            }
#pragma unroll
            for (int c = 0; c < BPP; c++) {
                float a = __fmul_rn(p00[c], w00);
                a = __fmaf_rn(p01[c], w01, a);
                a = __fmaf_rn(p10[c], w10, a);
                a = __fmaf_rn(p11[c], w11, a);
                out[c] = min(max(__float2int_rn(a), 0), 255);
            }
        }
    }
#pragma unroll
    for (int c = 0; c < BPP; c++) pd[c] = out[c];
}

// ---------------------------------------------------------------------------
// border index mapping (NVCV_BORDER_* semantics = OpenCV's)
// returns -1 for BORDER_CONSTANT out-of-range (value 0)
// ---------------------------------------------------------------------------
__host__ __device__ inline int border_idx(int i, int n, int mode) {
    if ((unsigned)i < (unsigned)n) return i;
    switch (mode) {
    case GMATB_BORDER_REPLICATE: return i < 0 ? 0 : n - 1;
    case GMATB_BORDER_REFLECT: {       // fedcba|abcdefgh|hgfedcb
        if (n == 1) return 0;
        int p = 2 * n; i %= p; if (i < 0) i += p;
        return i < n ? i : p - 1 - i;
    }
    case GMATB_BORDER_REFLECT101: {    // gfedcb|abcdefgh|gfedcba
        if (n == 1) return 0;
        int p = 2 * n - 2; i %= p; if (i < 0) i += p;
        return i < n ? i : p - i;
    }
    case GMATB_BORDER_WRAP: { i %= n; if (i < 0) i += n; return i; }
    default: return -1;
    }
}

}  // namespace gmatb
#include "gauss_stream.cuh"
#include "median3_stream.cuh"
#include "median5_stream.cuh"
#include "rotate_linear.cuh"
namespace gmatb {

// ---------------------------------------------------------------------------
// gaussian: separable, fp32.  t(y,x) = sum_i kx[i]*src(y, x+i-rx) accumulated left to
// right with FFMA starting from 0; out = rint(sum_j ky[j]*t(y+j-ry, x)), saturated.
// CTA tile 32x16 destination pixels; source tile (+halo) staged in shared memory as
// bytes, horizontal results as floats.
// ---------------------------------------------------------------------------
#define GAUSS_MAXK 31
struct GaussParams { float kx[GAUSS_MAXK], ky[GAUSS_MAXK]; int kw, kh, border; };

// Produces the destination columns [xa, xb) only (the whole row, or a frame-edge strip beside the
// streaming kernel's interior).
template <int BPP>
__global__ void __launch_bounds__(256) gaussian_kernel(PImg s, PImg d, GaussParams G, int xa, int xb) {
    extern __shared__ unsigned char gsm[];
    const int TW = 32, TH = 16;
    const int rx = G.kw / 2, ry = G.kh / 2;
    const int sw = TW + G.kw - 1, shh = TH + G.kh - 1;
    uint8_t *tile = gsm;                                                  // [shh][sw][BPP]
    float *hbuf = reinterpret_cast<float *>(gsm + (((size_t)shh * sw * BPP + 15) & ~(size_t)15));   // [shh][TW][BPP]
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int x0 = xa + blockIdx.x * TW, y0 = blockIdx.y * TH;
    const long long fz = blockIdx.z;
    const uint8_t *ps = s.p + fz * s.bstride;
    for (int i = tid; i < sw * shh; i += 256) {
        const int ty = i / sw, tx = i - ty * sw;
        const int sx = border_idx(x0 + tx - rx, s.w, G.border), sy = border_idx(y0 + ty - ry, s.h, G.border);
        uint8_t *o = tile + (size_t)i * BPP;
        if (sx < 0 || sy < 0) {
#pragma unroll
            for (int c = 0; c < BPP; c++) o[c] = 0;
        } else {
            const uint8_t *q = ps + (size_t)sy * s.pitch + (size_t)sx * BPP;
#pragma unroll
            for (int c = 0; c < BPP; c++) o[c] = q[c];
        }
    }
    __syncthreads();
    for (int i = tid; i < shh * TW; i += 256) {
        const int ty = i / TW, tx = i - ty * TW;
        float acc[BPP];
#pragma unroll
        for (int c = 0; c < BPP; c++) acc[c] = 0.0f;
        const uint8_t *q = tile + ((size_t)ty * sw + tx) * BPP;
        for (int k = 0; k < G.kw; k++) {
#pragma unroll
            for (int c = 0; c < BPP; c++) acc[c] = __fmaf_rn(G.kx[k], (float)q[k * BPP + c], acc[c]);
        }
#pragma unroll
        for (int c = 0; c < BPP; c++) hbuf[(size_t)i * BPP + c] = acc[c];
    }
    __syncthreads();
    for (int i = tid; i < TH * TW; i += 256) {
        const int ty = i / TW, tx = i - ty * TW;
        if (x0 + tx >= xb || y0 + ty >= d.h) continue;
        float acc[BPP];
#pragma unroll
        for (int c = 0; c < BPP; c++) acc[c] = 0.0f;
        for (int k = 0; k < G.kh; k++) {
            const float *h = hbuf + ((size_t)(ty + k) * TW + tx) * BPP;
#pragma unroll
            for (int c = 0; c < BPP; c++) acc[c] = __fmaf_rn(G.ky[k], h[c], acc[c]);
        }
        uint8_t *pd = d.p + fz * d.bstride + (size_t)(y0 + ty) * d.pitch + (size_t)(x0 + tx) * BPP;
#pragma unroll
        for (int c = 0; c < BPP; c++) pd[c] = min(max(__float2int_rn(acc[c]), 0), 255);
    }
}

// ---------------------------------------------------------------------------
// median: exact kw x kh median per channel, replicate border.  The window is
// staged in shared memory; each thread ranks the centre candidates by counting.
// For N = kw*kh values the median is the element of rank N/2 (0-based, ascending).
// ---------------------------------------------------------------------------
template <int BPP>
__global__ void __launch_bounds__(256) median_kernel(PImg s, PImg d, int kw, int kh) {
    extern __shared__ unsigned char msm[];
    const int TW = 32, TH = 8;
    const int rx = kw / 2, ry = kh / 2;
    const int sw = TW + kw - 1, shh = TH + kh - 1;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const long long fz = blockIdx.z;
    const uint8_t *ps = s.p + fz * s.bstride;
    for (int i = tid; i < sw * shh; i += 256) {
        const int ty = i / sw, tx = i - ty * sw;
        const int sx = min(max(x0 + tx - rx, 0), s.w - 1), sy = min(max(y0 + ty - ry, 0), s.h - 1);
        const uint8_t *q = ps + (size_t)sy * s.pitch + (size_t)sx * BPP;
#pragma unroll
        for (int c = 0; c < BPP; c++) msm[(size_t)i * BPP + c] = q[c];
    }
    __syncthreads();
    const int tx = threadIdx.x, ty = threadIdx.y;
    if (x0 + tx >= d.w || y0 + ty >= d.h) return;
    const int N = kw * kh, target = N / 2;
    uint8_t *pd = d.p + fz * d.bstride + (size_t)(y0 + ty) * d.pitch + (size_t)(x0 + tx) * BPP;
#pragma unroll
    for (int c = 0; c < BPP; c++) {
        // 256-bin search by bisection on the value: count(v <= m) > target  <=>  median <= m
        int lo = 0, hi = 255;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            int cnt = 0;
            for (int j = 0; j < kh; j++) {
                const uint8_t *row = msm + ((size_t)(ty + j) * sw + tx) * BPP + c;
                for (int i = 0; i < kw; i++) cnt += (row[i * BPP] <= mid);
            }
            if (cnt > target) hi = mid; else lo = mid + 1;
        }
        pd[c] = lo;
    }
}


// ---------------------------------------------------------------------------
// median, small windows (3x3, 5x5): pruned Batcher odd-even merge-sort SELECTION network on
// u16x2 lanes (VIMNMX.U16x2 is a native Blackwell instruction): each thread ranks the windows
// of TWO horizontally adjacent pixels at once, one network per channel, entirely in registers.
// The comparator lists are generated by tools/gen_median_net.py (comparators that cannot reach
// the median output are dropped: 40 min/max ops for 3x3, 202 for 5x5).
// ---------------------------------------------------------------------------
#define CS2(x, y)  do { const unsigned lo_ = __vminu2(x, y), hi_ = __vmaxu2(x, y); x = lo_; y = hi_; } while (0)
#define CMIN(x, y) x = __vminu2(x, y)
#define CMAX(x, y) x = __vmaxu2(x, y)
#include "median_nets.inc"

template <int BPP, int KW, int KH>
__global__ void __launch_bounds__(256) median_net_kernel(PImg s, PImg d) {
    extern __shared__ unsigned char msm[];
    constexpr int TW = 64, TH = 8, N = KW * KH;
    constexpr int rx = KW / 2, ry = KH / 2, sw = TW + KW - 1, shh = TH + KH - 1;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const long long fz = blockIdx.z;
    const uint8_t *ps = s.p + fz * s.bstride;
    for (int i = tid; i < sw * shh; i += 256) {
        const int ty = i / sw, tx = i - ty * sw;
        const int sx = min(max(x0 + tx - rx, 0), s.w - 1), sy = min(max(y0 + ty - ry, 0), s.h - 1);
        const uint8_t *q = ps + (size_t)sy * s.pitch + (size_t)sx * BPP;
#pragma unroll
        for (int c = 0; c < BPP; c++) msm[(size_t)i * BPP + c] = q[c];
    }
    __syncthreads();
    const int tx = threadIdx.x * 2, ty = threadIdx.y;
    if (x0 + tx >= d.w || y0 + ty >= d.h) return;
    uint8_t *pd = d.p + fz * d.bstride + (size_t)(y0 + ty) * d.pitch + (size_t)(x0 + tx) * BPP;
    const bool second = x0 + tx + 1 < d.w;
#pragma unroll
    for (int c = 0; c < BPP; c++) {
        unsigned a[N];
#pragma unroll
        for (int j = 0; j < KH; j++) {
            const uint8_t *row = msm + ((size_t)(ty + j) * sw + tx) * BPP + c;
#pragma unroll
            for (int i = 0; i < KW; i++) a[j * KW + i] = (unsigned)row[i * BPP] | ((unsigned)row[(i + 1) * BPP] << 16);
        }
        unsigned m;
        if (N == 9) { MEDIAN_NET_9(a); m = a[MEDIAN_OUT_9]; }
        else        { MEDIAN_NET_25(a); m = a[MEDIAN_OUT_25]; }
        pd[c] = m & 0xFFu;
        if (second) pd[BPP + c] = (m >> 16) & 0xFFu;
    }
}

}  // namespace gmatb

using namespace gmatb;

static bool same_geom(const PImg &a, const PImg &b) { return a.w == b.w && a.h == b.h && a.bpp == b.bpp; }
static bool al16(const PImg &a) { return !(((uintptr_t)a.p | (uintptr_t)a.pitch | (uintptr_t)a.bstride) & 15); }

extern "C" int gmatb_crop(const GmatbImage *src, const GmatbImage *dst, int x, int y, void *stream) {
    PImg s, d;
    if (!to_pimg(src, &s) || !to_pimg(dst, &d) || s.bpp != d.bpp || nbatch(src) != nbatch(dst)) return GMATB_ERR_INVAL;
    // same check as vf_crop_nvcv.c:151-154, with a proper negative error code
    if (x < 0 || y < 0 || x + d.w > s.w || y + d.h > s.h) return GMATB_ERR_INVAL;
    dim3 b(32, 8), g((d.w * d.bpp + 511) / 512, (d.h + 7) / 8, nbatch(src));
    crop_kernel<<<g, b, 0, (cudaStream_t)stream>>>(s, d, x, y, al16(d) ? 1 : 0);
    count_launch();
    return set_cuda_error(cudaGetLastError());
}

extern "C" int gmatb_flip(const GmatbImage *src, const GmatbImage *dst, int code, void *stream) {
    PImg s, d;
    if (!to_pimg(src, &s) || !to_pimg(dst, &d) || !same_geom(s, d) || nbatch(src) != nbatch(dst)) return GMATB_ERR_INVAL;
    const int vec = al16(s) && al16(d);
    dim3 b(32, 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (code == 0) {
        dim3 g((d.w * d.bpp + 511) / 512, (d.h + 7) / 8, nbatch(src));
        flip_rows_kernel<<<g, b, 0, st>>>(s, d, vec);
    } else {
        const int v4 = d.w % 4 == 0 && (d.bpp == 4 ? vec : (s.pitch % 4 == 0 && d.pitch % 4 == 0 && ((uintptr_t)s.p | (uintptr_t)d.p | (uintptr_t)s.bstride | (uintptr_t)d.bstride) % 4 == 0));
        if (d.bpp == 3 && vec && d.w % 16 == 0) {
            dim3 g((d.w + 511) / 512, (d.h + 7) / 8, nbatch(src));
            flip_cols_kernel<3, 16><<<g, b, 0, st>>>(s, d, code < 0, 2);
        } else {
            dim3 g((d.w + 127) / 128, (d.h + 7) / 8, nbatch(src));
            if (d.bpp == 3) flip_cols_kernel<3, 4><<<g, b, 0, st>>>(s, d, code < 0, v4);
            else            flip_cols_kernel<4, 4><<<g, b, 0, st>>>(s, d, code < 0, v4);
        }
    }
    count_launch();
    return set_cuda_error(cudaGetLastError());
}

extern "C" int gmatb_rotate(const GmatbImage *src, const GmatbImage *dst, double angle_deg, double shift_x, double shift_y,
                            int interp, void *stream) {
    PImg s, d;
    if (!to_pimg(src, &s) || !to_pimg(dst, &d) || !same_geom(s, d) || nbatch(src) != nbatch(dst)) return GMATB_ERR_INVAL;
    if (interp < GMATB_INTERP_NEAREST || interp > GMATB_INTERP_AREA) return GMATB_ERR_INVAL;
    if (interp == GMATB_INTERP_AREA) {       // NVCV_INTERP_AREA has no meaning for a same-size rotation; CV-CUDA's choice is not pinned
        static bool said = false;
        if (!said) { said = true; gmatb_log("gmat_b200: rotate interp=area runs as linear"); }
    }
    RotParams R;
    const double rad = angle_deg * 3.14159265358979323846 / 180.0;
    R.c = cos(rad); R.s = sin(rad); R.shx = shift_x; R.shy = shift_y; R.interp = interp;
    const uintptr_t sal = (uintptr_t)s.p | (uintptr_t)s.pitch | (uintptr_t)s.bstride;
    const uintptr_t dal = (uintptr_t)d.p | (uintptr_t)d.pitch | (uintptr_t)d.bstride;
    if ((interp == GMATB_INTERP_LINEAR || interp == GMATB_INTERP_AREA) && d.w % 4 == 0 && s.w >= 2 &&
        (sal & 3) == 0 && (dal & (d.bpp == 3 ? 3 : 15)) == 0) {
        dim3 b4(8, 32), g4((d.w / 4 + 7) / 8, (d.h + 31) / 32, nbatch(src));
        // TMA-staged gather: the bounding box of a 32 x 32 tile's footprint, + the tap, + slack for the word reads
        {
            const double ac = fabs(R.c), as = fabs(R.s);
            const int ext = (int)ceil(31.0 * (ac + as)) + 4;
            const unsigned box_x = (unsigned)((ext * d.bpp + 8 + 15 + 15) / 16 * 16), box_y = (unsigned)ext;   // + the 16-byte alignment of the origin
            CUtensorMap tm;
            if (s.w >= 64 && s.h >= 64 && box_x <= 256 && box_y <= 256 &&
                make_tensor_map_3d(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, s.p, (unsigned long long)s.w * s.bpp, s.h, nbatch(src), s.pitch,
                                   nbatch(src) > 1 ? s.bstride : 0, box_x, box_y)) {
                const size_t smem = (size_t)box_x * box_y;
                if (d.bpp == 3) rotate_linear_tma_kernel<3><<<g4, b4, smem, (cudaStream_t)stream>>>(tm, s, d, R, (int)box_x, (int)box_y);
                else            rotate_linear_tma_kernel<4><<<g4, b4, smem, (cudaStream_t)stream>>>(tm, s, d, R, (int)box_x, (int)box_y);
                count_launch();
                return set_cuda_error(cudaGetLastError());
            }
        }
        if (d.bpp == 3) rotate_linear4_kernel<3><<<g4, b4, 0, (cudaStream_t)stream>>>(s, d, R);
        else            rotate_linear4_kernel<4><<<g4, b4, 0, (cudaStream_t)stream>>>(s, d, R);
        count_launch();
        return set_cuda_error(cudaGetLastError());
    }
    dim3 b(32, 8), g((d.w + 31) / 32, (d.h + 7) / 8, nbatch(src));
    if (d.bpp == 3) rotate_kernel<3><<<g, b, 0, (cudaStream_t)stream>>>(s, d, R);
    else            rotate_kernel<4><<<g, b, 0, (cudaStream_t)stream>>>(s, d, R);
    count_launch();
    return set_cuda_error(cudaGetLastError());
}

// weights exactly as the oracle computes them: double exp, normalised in double, cast to float
static void gauss_weights(int k, double sigma, float *out) {
    if (sigma <= 0.0) sigma = 0.3 * ((k - 1) * 0.5 - 1.0) + 0.8;
    double w[GAUSS_MAXK], sum = 0.0;
    const int r = k / 2;
    for (int i = 0; i < k; i++) { const double x = (double)(i - r); w[i] = exp(-(x * x) / (2.0 * sigma * sigma)); sum += w[i]; }
    for (int i = 0; i < k; i++) out[i] = (float)(w[i] / sum);
}

extern "C" int gmatb_gaussian(const GmatbImage *src, const GmatbImage *dst, int kw, int kh, double sigma_x, double sigma_y,
                              int border, void *stream) {
    PImg s, d;
    if (!to_pimg(src, &s) || !to_pimg(dst, &d) || !same_geom(s, d) || nbatch(src) != nbatch(dst)) return GMATB_ERR_INVAL;
    if (kw < 1 || kh < 1 || kw > GAUSS_MAXK || kh > GAUSS_MAXK || !(kw & 1) || !(kh & 1)) return GMATB_ERR_INVAL;
    if (kw > s.w || kh > s.h) return GMATB_ERR_INVAL;                 // vf_smooth_nvcv.c:172-175
    if (border < GMATB_BORDER_CONSTANT || border > GMATB_BORDER_REFLECT101) return GMATB_ERR_INVAL;
    GaussParams G;
    memset(&G, 0, sizeof(G));
    G.kw = kw; G.kh = kh; G.border = border;
    // OpenCV / CV-CUDA rule: sigmaY <= 0 takes sigmaX; a sigma <= 0 is derived from its kernel size
    gauss_weights(kw, sigma_x, G.kx);
    gauss_weights(kh, sigma_y > 0.0 ? sigma_y : sigma_x, G.ky);
    const int sw = 32 + kw - 1, sh = 16 + kh - 1;
    const size_t smem = (((size_t)sh * sw * d.bpp + 15) & ~(size_t)15) + (size_t)sh * 32 * d.bpp * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = nbatch(src);
    // generic tile kernel on destination columns [xa, xb)
    auto tiles = [&](int xa, int xb) -> cudaError_t {
        if (xb <= xa) return cudaSuccess;
        dim3 b(32, 8), g((xb - xa + 31) / 32, (d.h + 15) / 16, nb);
        cudaError_t e = cudaSuccess;
        if (d.bpp == 3) {
            if (smem > 48 * 1024) e = cudaFuncSetAttribute(gaussian_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) gaussian_kernel<3><<<g, b, smem, st>>>(s, d, G, xa, xb);
        } else {
            if (smem > 48 * 1024) e = cudaFuncSetAttribute(gaussian_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) gaussian_kernel<4><<<g, b, smem, st>>>(s, d, G, xa, xb);
        }
        if (e == cudaSuccess) { count_launch(); e = cudaGetLastError(); }
        return e;
    };
    const bool words = ((((uintptr_t)s.p | (uintptr_t)s.pitch | (uintptr_t)s.bstride | (uintptr_t)d.p | (uintptr_t)d.pitch | (uintptr_t)d.bstride) & 3) == 0);
    const int rx = kw / 2;
    // streaming kernel: 4-pixel strips whose window [4t - rx, 4t + 4 + rx) lies inside the row
    const int t0 = (rx + 3) / 4, t1 = (d.w - rx - 4) / 4 + 1;
    if ((kw == 3 || kw == 5 || kw == 7) && (kh == 3 || kh == 5 || kh == 7) && words && t1 > t0) {
        GaussS S;
        memset(&S, 0, sizeof(S));
        for (int i = 0; i < kw; i++) S.kx[i] = G.kx[i];
        for (int i = 0; i < kh; i++) S.ky[i] = G.ky[i];
        S.border = border;
        const int wx = t1 - t0;                       // threads per row
        long long want = 148LL * 16 * 32 * 4;         // threads
        int bands = (int)((want + (long long)wx * nb - 1) / ((long long)wx * nb));
        bands = std::max(1, std::min(bands, (d.h + 31) / 32));
        S.band = (d.h + bands - 1) / bands;
        bands = (d.h + S.band - 1) / S.band;
        dim3 g2((wx + 127) / 128, bands, nb);
        const int minb = d.bpp == 4 ? 4 : 5;        // 4: 128 registers, loads two rows ahead (gauss_stream.cuh)
#define GS(B, KW_, KH_) do { if (minb >= 5) gauss_stream_kernel<B, KW_, KH_, 5><<<g2, 128, 0, st>>>(s.p, s.pitch, s.bstride, d.p, d.pitch, d.bstride, d.h, t0, t1, S); \
        else gauss_stream_kernel<B, KW_, KH_, 4><<<g2, 128, 0, st>>>(s.p, s.pitch, s.bstride, d.p, d.pitch, d.bstride, d.h, t0, t1, S); } while (0)
#define GK(B, KW_) do { if (kh == 3) GS(B, KW_, 3); else if (kh == 5) GS(B, KW_, 5); else GS(B, KW_, 7); } while (0)
#define GB(B) do { if (kw == 3) GK(B, 3); else if (kw == 5) GK(B, 5); else GK(B, 7); } while (0)
        if (kw == 5 && kh == 5 && d.bpp == 4) {
            // 4-byte pixels, the smooth filter's common window: rows arrive through a cp.async ring in shared memory
            // (gauss_stream.cuh; measured 341 -> 363 Gpx/s at 4K; 3-byte pixels measured no gain and keep the register form)
            constexpr int RING = 8;
            const size_t rsm = (size_t)RING * gauss_ring_words(4, 5) * 128 * 4;
            gauss_ring_kernel<4, 5, 5, RING><<<g2, 128, rsm, st>>>(s.p, s.pitch, s.bstride, d.p, d.pitch, d.bstride, d.h, t0, t1, S);
        } else if (d.bpp == 3) GB(3); else GB(4);
#undef GS
#undef GK
#undef GB
        count_launch();
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = tiles(0, 4 * t0);            // left frame-edge strip
        if (e == cudaSuccess) e = tiles(4 * t1, d.w);          // right frame-edge strip
        return set_cuda_error(e);
    }
    return set_cuda_error(tiles(0, d.w));
}

extern "C" int gmatb_median(const GmatbImage *src, const GmatbImage *dst, int kw, int kh, void *stream) {
    PImg s, d;
    if (!to_pimg(src, &s) || !to_pimg(dst, &d) || !same_geom(s, d) || nbatch(src) != nbatch(dst)) return GMATB_ERR_INVAL;
    // odd windows up to GMATB_MEDIAN_MAXK (the AVFilter option table advertises the same range)
    if (kw < 1 || kh < 1 || kw > GMATB_MEDIAN_MAXK || kh > GMATB_MEDIAN_MAXK || !(kw & 1) || !(kh & 1) || kw > s.w || kh > s.h) return GMATB_ERR_INVAL;
    const int wbytes = d.w * d.bpp;
    if (kw == 3 && kh == 3 && (wbytes % 16) == 0 && al16(s) && al16(d)) {
        // streaming u16x2 kernel (median3_stream.cuh): a thread owns 16 byte columns of two row bands
        Med3Params P;
        P.sp = s.p; P.dp = d.p; P.spitch = s.pitch; P.dpitch = d.pitch; P.sbs = s.bstride; P.dbs = d.bstride;
        P.wb = wbytes; P.H = d.h;
        const int nbt = nbatch(src), gx = (wbytes / 16 + 127) / 128;
        // rows per band: long bands amortise the two priming rows, but keep ~4 waves of 148 SMs x 4 CTAs
        int rows = 32;
        while (rows > 4 && (long long)gx * ((d.h + 2 * rows - 1) / (2 * rows)) * nbt < 148LL * 4 * 4) rows >>= 1;
        P.rows = rows;
        dim3 g3(gx, (d.h + 2 * rows - 1) / (2 * rows), nbt);
        if (d.bpp == 3) median3_stream_kernel<3><<<g3, 128, 0, (cudaStream_t)stream>>>(P);
        else            median3_stream_kernel<4><<<g3, 128, 0, (cudaStream_t)stream>>>(P);
        count_launch();
        return set_cuda_error(cudaGetLastError());
    }
    if (kw == 5 && kh == 5 && al16(s) && al16(d) && d.h >= 1) {
        // streaming u16x2 kernel (median5_stream.cuh): a thread owns NOUT pixels of two row bands
        // 8 pixels per thread when the row allows (halo columns 12 / 8 instead of 8 / 4), else 4
        const int nout = wbytes % (8 * d.bpp) == 0 ? 8 : wbytes % (4 * d.bpp) == 0 ? 4 : 0;
        if (nout) {
            Med3Params P;
            P.sp = s.p; P.dp = d.p; P.spitch = s.pitch; P.dpitch = d.pitch; P.sbs = s.bstride; P.dbs = d.bstride;
            P.wb = wbytes; P.H = d.h;
            const int S = nout * d.bpp, nbt = nbatch(src), gx = (wbytes / S + 127) / 128;
            int rows = 32;
            while (rows > 4 && (long long)gx * ((d.h + 2 * rows - 1) / (2 * rows)) * nbt < 148LL * 2 * 4) rows >>= 1;
            P.rows = rows;
            dim3 g5(gx, (d.h + 2 * rows - 1) / (2 * rows), nbt);
            cudaStream_t st = (cudaStream_t)stream;
            if (nout == 8) { if (d.bpp == 3) median5_stream_kernel<3, 8><<<g5, 128, 0, st>>>(P); else median5_stream_kernel<4, 8><<<g5, 128, 0, st>>>(P); }
            else           { if (d.bpp == 3) median5_stream_kernel<3, 4><<<g5, 128, 0, st>>>(P); else median5_stream_kernel<4, 4><<<g5, 128, 0, st>>>(P); }
            count_launch();
            return set_cuda_error(cudaGetLastError());
        }
    }
    if ((kw == 3 && kh == 3) || (kw == 5 && kh == 5)) {
        const size_t sm2 = (size_t)(64 + kw - 1) * (8 + kh - 1) * d.bpp;
        dim3 b2(32, 8), g2((d.w + 63) / 64, (d.h + 7) / 8, nbatch(src));
        cudaStream_t st = (cudaStream_t)stream;
        if (kw == 3) { if (d.bpp == 3) median_net_kernel<3, 3, 3><<<g2, b2, sm2, st>>>(s, d); else median_net_kernel<4, 3, 3><<<g2, b2, sm2, st>>>(s, d); }
        else         { if (d.bpp == 3) median_net_kernel<3, 5, 5><<<g2, b2, sm2, st>>>(s, d); else median_net_kernel<4, 5, 5><<<g2, b2, sm2, st>>>(s, d); }
        count_launch();
        return set_cuda_error(cudaGetLastError());
    }
    const size_t smem = (size_t)(32 + kw - 1) * (8 + kh - 1) * d.bpp;
    dim3 b(32, 8), g((d.w + 31) / 32, (d.h + 7) / 8, nbatch(src));
    if (d.bpp == 3) median_kernel<3><<<g, b, smem, (cudaStream_t)stream>>>(s, d, kw, kh);
    else            median_kernel<4><<<g, b, smem, (cudaStream_t)stream>>>(s, d, kw, kh);
    count_launch();
    return set_cuda_error(cudaGetLastError());
}
