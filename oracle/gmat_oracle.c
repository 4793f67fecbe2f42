/*
 * oracle/gmat_oracle.c -- CPU restatement of the pixel-transform hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under gmat_b200/ links, imports or calls this
 * file; it is used by tests/, by __graft_entry__.smoke() and by bench.py's
 * cpu_baseline leg as the checker.  Plain scalar C, one pixel at a time, compiled
 * with -ffp-contract=off so that every float operation below is exactly one IEEE
 * binary32 operation, in the order the reference's CUDA kernels perform them.
 *
 * What it restates, and how it is pinned:
 *   colour conversion  -- NVIDIA/GMAT ffmpeg-gpu/libswscale/cuda/yuv2rgb_cuda.cu:72-106
 *        (yuv2rgb_for_pixel), :653-702 (RgbToY/U/V, RgbToYuvKernel), :782-848 (matrices),
 *        yuv2yuv_cuda.cu:56-63 (bit-depth repack), rgb2rgb_cuda_kernel.cu:6-22, in the
 *        operation order of the SASS nvcc 12.9 emits for sm_100a.  PINNED: golden vectors
 *        under tests/golden/ were produced by the reference's own kernels (oracle O1 =
 *        oracle/_ref/libref_gpuscale.so) on a B200, tests/golden/make_golden.py.
 *   resample "R-B"     -- ffmpeg-gpu/libavfilter/vf_scale_cuda.cu:948-1074 (the only in-tree
 *        CUDA bicubic/Lanczos; CV-CUDA, which the reference actually calls, is a closed
 *        third-party binary -- nvcv 0.3.1-beta -- absent from /root/reference).  PINNED:
 *        golden vectors produced by those kernels (oracle O2) on a B200.
 *   resample "R-A" (bilinear / nearest), rotate, gaussian, median
 *        -- CV-CUDA operators called at swscale_cuda.c:326,369, vf_rotate_nvcv.c:275,
 *        vf_smooth_nvcv.c:290,294.  PARITY UNPINNED: no source, no binary, no reference
 *        test pins them; this file is the specification (SURVEY.md 8c P-FILTERS).
 *   crop / flip        -- exact copies (vf_crop_nvcv.c:277, vf_flip_nvcv.c:251).
 *   format_cuda        -- ffmpeg-gpu/libavfilter/format_cuda_kernel.cu:32-96 (matrix selection),
 *        :257-297 (NV12 -> planar float), :516-570 (planar float -> NV12), SASS order.  PINNED:
 *        golden vectors produced by that file compiled unmodified (oracle O3 =
 *        oracle/_ref/libref_format_cuda.so) on a B200, tests/golden/make_golden_format.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../include/gmat_b200.h"

/* ------------------------------------------------------------------ helpers */
static int nplanes(int fmt)
{
    switch (fmt) {
    case GMATB_FMT_NV12: case GMATB_FMT_P010LE: case GMATB_FMT_P016LE: return 2;
    case GMATB_FMT_YUV420P: case GMATB_FMT_YUV420P10LE: case GMATB_FMT_YUV420P16LE: return 3;
    default: return 1;
    }
}
static int is16(int fmt)
{
    switch (fmt) {
    case GMATB_FMT_P010LE: case GMATB_FMT_P016LE: case GMATB_FMT_YUV420P10LE: case GMATB_FMT_YUV420P16LE:
    case GMATB_FMT_RGB48LE: case GMATB_FMT_BGR48LE: case GMATB_FMT_RGBA64LE: case GMATB_FMT_BGRA64LE: return 1;
    default: return 0;
    }
}
static int rgb_ch(int fmt)
{
    switch (fmt) {
    case GMATB_FMT_RGB24: case GMATB_FMT_BGR24: case GMATB_FMT_RGB48LE: case GMATB_FMT_BGR48LE: return 3;
    default: return 4;
    }
}
static int rgb_swapped(int fmt)
{
    return fmt == GMATB_FMT_BGR24 || fmt == GMATB_FMT_BGRA || fmt == GMATB_FMT_BGR0 ||
           fmt == GMATB_FMT_BGR48LE || fmt == GMATB_FMT_BGRA64LE;
}
static const uint8_t *plane(const GmatbImage *g, int p, int frame)
{
    return (const uint8_t *)g->data[p] + (g->batch > 1 ? (size_t)frame * g->batch_stride[p] : 0);
}
static int iclamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

static void get_yuv(const GmatbImage *s, int frame, int x, int y, unsigned *Y, unsigned *U, unsigned *V)
{
    const int b16 = is16(s->format), bs = b16 ? 2 : 1;
    const uint8_t *py = plane(s, 0, frame) + (size_t)y * s->linesize[0] + (size_t)x * bs;
    *Y = b16 ? *(const uint16_t *)py : *py;
    const int cx = x >> 1, cy = y >> 1;
    if (nplanes(s->format) == 2) {
        const uint8_t *q = plane(s, 1, frame) + (size_t)cy * s->linesize[1] + (size_t)cx * 2 * bs;
        if (b16) { *U = ((const uint16_t *)q)[0]; *V = ((const uint16_t *)q)[1]; } else { *U = q[0]; *V = q[1]; }
    } else {
        const uint8_t *qu = plane(s, 1, frame) + (size_t)cy * s->linesize[1] + (size_t)cx * bs;
        const uint8_t *qv = plane(s, 2, frame) + (size_t)cy * s->linesize[2] + (size_t)cx * bs;
        if (b16) { *U = *(const uint16_t *)qu; *V = *(const uint16_t *)qv; } else { *U = *qu; *V = *qv; }
    }
}

/* ------------------------------------------------------------------ matrices
 * yuv2rgb_cuda.cu:782-848, same expressions / types / order (floats for the entries,
 * double for the range scale, cast to float). */
static void constants(int cspace, float *wr, float *wb, int *black, int *white, int *max)
{
    *black = 16; *white = 235; *max = 255;
    switch (cspace) {
    case GMATB_SPC_BT709:     *wr = 0.2126f; *wb = 0.0722f; break;
    case GMATB_SPC_FCC:       *wr = 0.30f;   *wb = 0.11f;   break;
    case GMATB_SPC_SMPTE240M: *wr = 0.212f;  *wb = 0.087f;  break;
    case GMATB_SPC_BT2020_NCL: case GMATB_SPC_BT2020_CL:
        *wr = 0.2627f; *wb = 0.0593f; *black = 64 << 6; *white = 940 << 6; *max = (1 << 16) - 1; break;
    default:                  *wr = 0.2990f; *wb = 0.1140f; break;
    }
}
void orc_matrix_yuv2rgb(int cspace, float m[9])
{
    float wr, wb; int black, white, max;
    constants(cspace, &wr, &wb, &black, &white, &max);
    float mat[3][3] = {
        {1.0f, 0.0f, (1.0f - wr) / 0.5f},
        {1.0f, -wb * (1.0f - wb) / 0.5f / (1 - wb - wr), -wr * (1 - wr) / 0.5f / (1 - wb - wr)},
        {1.0f, (1.0f - wb) / 0.5f, 0.0f},
    };
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            m[i * 3 + j] = (float)(1.0 * max / (white - black) * mat[i][j]);
}
void orc_matrix_rgb2yuv(int cspace, float m[9])
{
    float wr, wb; int black, white, max;
    constants(cspace, &wr, &wb, &black, &white, &max);
    float mat[3][3] = {
        {wr, 1.0f - wb - wr, wb},
        {-0.5f * wr / (1.0f - wb), -0.5f * (1 - wb - wr) / (1.0f - wb), 0.5f},
        {0.5f, -0.5f * (1.0f - wb - wr) / (1.0f - wr), -0.5f * wb / (1.0f - wr)},
    };
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            m[i * 3 + j] = (float)(1.0 * (white - black) / max * mat[i][j]);
}

/* ------------------------------------------------------------------ colour conversion
 * yuv2rgb_for_pixel (yuv2rgb_cuda.cu:72-106) as compiled: t1 = FMUL(fu,mB); t2 = FMUL(fv,mC);
 * x = FFMA(fy,mA,t1); r = FADD(x,t2); r<0 -> 0; min(r,max); F2I.U32.TRUNC. */
/* Two roundings exist in the reference, because nvcc contracts the same source expression
 * differently per kernel (SASS of oracle/_ref/libref_gpuscale.so):
 *   fma_form 0  yuv2rgb_odd_kernel / yuv02rgb_odd_kernel (NV12, I420 -> packed rgb):  FADD(FFMA(fy,mA,t1), FMUL(fv,mC))
 *   fma_form 1  yuv2rgb_kernel (P010/P016) and yuv2rgb_planar_kernel (planar float):  FFMA(fv,mC, FFMA(fy,mA,t1)) */
static float csc_chain(float fy, float fu, float fv, const float *row, int fma_form)
{
    const float t1 = fu * row[1];
    const float x = fmaf(fy, row[0], t1);
    if (fma_form) return fmaf(fv, row[2], x);
    const float t2 = fv * row[2];
    return x + t2;
}
static unsigned quantise(float r, float maxf)
{
    if (r < 0.0f) r = 0.0f;         /* FSETP.GEU + FSEL */
    if (r > maxf) r = maxf;         /* FMNMX */
    return (unsigned)r;             /* truncation */
}
static void yuv_to_rgb_q2(unsigned Y, unsigned U, unsigned V, int b16, const float m[9], unsigned rgb[3], int fma_form)
{
    const int low = b16 ? 4096 : 16, mid = b16 ? 32768 : 128;
    const float maxf = b16 ? 65535.0f : 255.0f;
    const float fy = (float)((int)Y - low), fu = (float)((int)U - mid), fv = (float)((int)V - mid);
    for (int c = 0; c < 3; c++) rgb[c] = quantise(csc_chain(fy, fu, fv, m + 3 * c, fma_form), maxf);
}
static void yuv_to_rgb_q(unsigned Y, unsigned U, unsigned V, int b16, const float m[9], unsigned rgb[3])
{
    yuv_to_rgb_q2(Y, U, V, b16, m, rgb, b16);      /* 16-bit sources: the P016 template's form */
}
static void put_rgb(uint8_t *p, int fmt, int src16, const unsigned rgb[3])
{
    unsigned c[3] = {rgb[0], rgb[1], rgb[2]};
    const int d16 = is16(fmt), ch = rgb_ch(fmt), sw = rgb_swapped(fmt);
    for (int i = 0; i < 3; i++) {
        if (src16 && !d16) c[i] >>= 8;            /* yuv2rgb_cuda.cu:91-94 */
        if (!src16 && d16) c[i] <<= 8;            /* :95-99 */
    }
    const unsigned o0 = sw ? c[2] : c[0], o2 = sw ? c[0] : c[2];
    if (d16) {
        uint16_t *q = (uint16_t *)p;
        q[0] = o0; q[1] = c[1]; q[2] = o2;
        if (ch == 4) q[3] = 255;                   /* DEFAULT_ALPHA added to a 16-bit lane, :89 */
    } else {
        p[0] = o0; p[1] = c[1]; p[2] = o2;
        if (ch == 4) p[3] = 255;
    }
}

int orc_yuv2rgb(const GmatbImage *s, const GmatbImage *d, const float m[9])
{
    const int nb = s->batch > 1 ? s->batch : 1;
    const int bpp = rgb_ch(d->format) * (is16(d->format) ? 2 : 1);
    for (int f = 0; f < nb; f++)
        for (int y = 0; y < s->height; y++)
            for (int x = 0; x < s->width; x++) {
                unsigned Y, U, V, rgb[3];
                get_yuv(s, f, x, y, &Y, &U, &V);
                yuv_to_rgb_q(Y, U, V, is16(s->format), m, rgb);
                put_rgb((uint8_t *)plane(d, 0, f) + (size_t)y * d->linesize[0] + (size_t)x * bpp, d->format, is16(s->format), rgb);
            }
    return 0;
}

/* normalize_pixel (yuv2rgb_cuda.cu:381-389): (c - shift) / norm, IEEE division */
int orc_yuv2rgb_planar_f32(const GmatbImage *s, const GmatbImage *d, const float m[9], float norm, const float shift[3])
{
    const int nb = s->batch > 1 ? s->batch : 1;
    for (int f = 0; f < nb; f++)
        for (int y = 0; y < s->height; y++)
            for (int x = 0; x < s->width; x++) {
                unsigned Y, U, V, rgb[3];
                get_yuv(s, f, x, y, &Y, &U, &V);
                yuv_to_rgb_q2(Y, U, V, 0, m, rgb, 1);        /* yuv2rgb_planar_kernel: FMA form */
                for (int c = 0; c < 3; c++) {
                    float *q = (float *)((uint8_t *)plane(d, c, f) + (size_t)y * d->linesize[c]) + x;
                    *q = ((float)rgb[c] - (shift ? shift[c] : 0.0f)) / norm;
                }
            }
    return 0;
}

/* RgbToY/U/V (yuv2rgb_cuda.cu:653-669) as compiled:
 * t = FMUL(g,m1); t = FFMA(r,m0,t); t = FFMA(b,m2,t); t = FADD(t, low|mid); F2I.U32.TRUNC */
static unsigned rgb_to_comp(float r, float g, float b, const float *row, float off, int maxv)
{
    float t = g * row[1];
    t = fmaf(r, row[0], t);
    t = fmaf(b, row[2], t);
    t = t + off;
    if (t < 0.0f) return 0;
    unsigned v = (unsigned)t;
    return v > (unsigned)maxv ? (unsigned)maxv : v;
}
static void get_rgb(const GmatbImage *s, int frame, int x, int y, int want16, int rgb[3])
{
    x = iclamp(x, 0, s->width - 1); y = iclamp(y, 0, s->height - 1);
    const int b16 = is16(s->format), ch = rgb_ch(s->format), sw = rgb_swapped(s->format);
    const uint8_t *p = plane(s, 0, frame) + (size_t)y * s->linesize[0] + (size_t)x * ch * (b16 ? 2 : 1);
    int c[3];
    for (int i = 0; i < 3; i++) {
        c[i] = b16 ? ((const uint16_t *)p)[i] : p[i];
        if (b16 && !want16) c[i] >>= 8;
    }
    rgb[0] = sw ? c[2] : c[0]; rgb[1] = c[1]; rgb[2] = sw ? c[0] : c[2];
}
int orc_rgb2yuv(const GmatbImage *s, const GmatbImage *d, const float m[9])
{
    const int nb = s->batch > 1 ? s->batch : 1;
    const int d16 = is16(d->format), bs = d16 ? 2 : 1;
    const float low = d16 ? 4096.f : 16.f, mid = d16 ? 32768.f : 128.f;
    const int maxv = d16 ? 65535 : 255;
    for (int f = 0; f < nb; f++) {
        for (int y = 0; y < s->height; y++)
            for (int x = 0; x < s->width; x++) {
                int c[3];
                get_rgb(s, f, x, y, d16, c);
                unsigned v = rgb_to_comp((float)c[0], (float)c[1], (float)c[2], m, low, maxv);
                uint8_t *q = (uint8_t *)plane(d, 0, f) + (size_t)y * d->linesize[0] + (size_t)x * bs;
                if (d16) *(uint16_t *)q = v; else *q = v;
            }
        for (int cy = 0; cy < (s->height + 1) / 2; cy++)
            for (int cx = 0; cx < (s->width + 1) / 2; cx++) {
                int sum[3] = {0, 0, 0};
                for (int j = 0; j < 2; j++)
                    for (int i = 0; i < 2; i++) {
                        int c[3];
                        get_rgb(s, f, 2 * cx + i, 2 * cy + j, d16, c);   /* clamped at odd edges */
                        sum[0] += c[0]; sum[1] += c[1]; sum[2] += c[2];
                    }
                const float r = (float)(sum[0] / 4), g = (float)(sum[1] / 4), b = (float)(sum[2] / 4);   /* :685-687 */
                unsigned u = rgb_to_comp(r, g, b, m + 3, mid, maxv), v = rgb_to_comp(r, g, b, m + 6, mid, maxv);
                if (nplanes(d->format) == 2) {
                    uint8_t *q = (uint8_t *)plane(d, 1, f) + (size_t)cy * d->linesize[1] + (size_t)cx * 2 * bs;
                    if (d16) { ((uint16_t *)q)[0] = u; ((uint16_t *)q)[1] = v; } else { q[0] = u; q[1] = v; }
                } else {
                    uint8_t *qu = (uint8_t *)plane(d, 1, f) + (size_t)cy * d->linesize[1] + (size_t)cx * bs;
                    uint8_t *qv = (uint8_t *)plane(d, 2, f) + (size_t)cy * d->linesize[2] + (size_t)cx * bs;
                    if (d16) { *(uint16_t *)qu = u; *(uint16_t *)qv = v; } else { *qu = u; *qv = v; }
                }
            }
    }
    return 0;
}

/* yuv2yuv_cuda.cu:56-63 bit-depth rules; depth 10 = MSB-aligned in 16 bits */
static int yuv_depth(int fmt)
{
    switch (fmt) {
    case GMATB_FMT_P010LE: case GMATB_FMT_YUV420P10LE: return 10;
    case GMATB_FMT_P016LE: case GMATB_FMT_YUV420P16LE: return 16;
    default: return 8;
    }
}
static unsigned conv_depth(unsigned x, int sd, int dd)
{
    if (sd == dd) return x;
    if (sd == 8 && dd == 10) return (x | (x << 8)) & 0xFFC0u;
    if (sd == 8 && dd == 16) return (x | (x << 8)) & 0xFFFFu;
    if (dd == 8) return x >> 8;
    if (sd == 10 && dd == 16) return x | (x >> 10);
    return x & 0xFFC0u;
}
int orc_yuv2yuv(const GmatbImage *s, const GmatbImage *d)
{
    const int nb = s->batch > 1 ? s->batch : 1;
    const int sd = yuv_depth(s->format), dd = yuv_depth(d->format), dbs = dd == 8 ? 1 : 2;
    for (int f = 0; f < nb; f++) {
        for (int y = 0; y < s->height; y++)
            for (int x = 0; x < s->width; x++) {
                unsigned Y, U, V;
                get_yuv(s, f, x, y, &Y, &U, &V);
                Y = conv_depth(Y, sd, dd);
                uint8_t *q = (uint8_t *)plane(d, 0, f) + (size_t)y * d->linesize[0] + (size_t)x * dbs;
                if (dbs == 2) *(uint16_t *)q = Y; else *q = Y;
            }
        for (int cy = 0; cy < (s->height + 1) / 2; cy++)
            for (int cx = 0; cx < (s->width + 1) / 2; cx++) {
                unsigned Y, U, V;
                get_yuv(s, f, iclamp(2 * cx, 0, s->width - 1), iclamp(2 * cy, 0, s->height - 1), &Y, &U, &V);
                U = conv_depth(U, sd, dd); V = conv_depth(V, sd, dd);
                if (nplanes(d->format) == 2) {
                    uint8_t *q = (uint8_t *)plane(d, 1, f) + (size_t)cy * d->linesize[1] + (size_t)cx * 2 * dbs;
                    if (dbs == 2) { ((uint16_t *)q)[0] = U; ((uint16_t *)q)[1] = V; } else { q[0] = U; q[1] = V; }
                } else {
                    uint8_t *qu = (uint8_t *)plane(d, 1, f) + (size_t)cy * d->linesize[1] + (size_t)cx * dbs;
                    uint8_t *qv = (uint8_t *)plane(d, 2, f) + (size_t)cy * d->linesize[2] + (size_t)cx * dbs;
                    if (dbs == 2) { *(uint16_t *)qu = U; *(uint16_t *)qv = V; } else { *qu = U; *qv = V; }
                }
            }
    }
    return 0;
}

int orc_rgb24tobgr24(const GmatbImage *s, const GmatbImage *d)
{
    const int nb = s->batch > 1 ? s->batch : 1;
    for (int f = 0; f < nb; f++)
        for (int y = 0; y < s->height; y++) {
            const uint8_t *p = plane(s, 0, f) + (size_t)y * s->linesize[0];
            uint8_t *q = (uint8_t *)plane(d, 0, f) + (size_t)y * d->linesize[0];
            for (int x = 0; x < s->width; x++) { q[3 * x] = p[3 * x + 2]; q[3 * x + 1] = p[3 * x + 1]; q[3 * x + 2] = p[3 * x]; }
        }
    return 0;
}

/* ------------------------------------------------------------------ resample
 * vf_scale_cuda.cu:970-981 bicubic_coeffs in SASS order (see gmat_b200/csrc/resample_core.cuh) */
static void bicubic_coeffs(float x, float A, float w[4])
{
    const float x1 = x + 1.0f, omx = -x + 1.0f;
    const float Ap2 = A + 2.0f, Ap3 = A + 3.0f, A5 = A * 5.0f, A4 = A * 4.0f;
    float s = fmaf(x1, A, -A5);
    s = x1 * s;
    s = fmaf(A, 8.0f, s);
    w[0] = fmaf(x1, s, -A4);
    float t = fmaf(x, Ap2, -Ap3);
    t = x * t;
    w[1] = fmaf(x, t, 1.0f);
    float u = fmaf(Ap2, omx, -Ap3);
    u = omx * u;
    w[2] = fmaf(omx, u, 1.0f);
    w[3] = ((-w[0] + 1.0f) + -w[1]) + -w[2];
}
/* lanczos_coeffs :948-968.  The reference uses the GPU's __sinf (MUFU.SIN), which has no
 * bit-exact CPU equivalent: this version uses sinf() and is a TOLERANCE reference only
 * (<= 2e-6 per coefficient); bit-exact tests take the table from the device instead
 * (gmatb_sws_get_filter) and feed it to orc_resample / orc_sws_scale. */
static float lanczos_tap(float t)
{
    if (t == 0.0f) return 1.0f;
    return (sinf(t) * sinf(t * 0.5f)) / ((t * t) * 0.5f);
}
static void lanczos_coeffs(float x, float w[4])
{
    const float pi = 3.141592654f;
    w[0] = lanczos_tap(pi * (x + 1.0f)); w[1] = lanczos_tap(pi * x);
    w[2] = lanczos_tap(pi * (x + -1.0f)); w[3] = lanczos_tap(pi * (x + -2.0f));
    const float sum = ((w[0] + w[1]) + w[2]) + w[3];
    for (int i = 0; i < 4; i++) w[i] = w[i] / sum;
}
/* algo: 0 bicubic, 1 lanczos, 2 bilinear (R-A), 3 nearest; coeffs: 4 floats per entry */
void orc_filter_table(int algo, int src_n, int dst_n, float A, float *coeffs, int *pos)
{
    const float scale = (float)src_n / (float)dst_n;
    for (int o = 0; o < dst_n; o++) {
        float *w = coeffs + 4 * o;
        if (algo == 3) {
            const float xi = ((float)o + 0.5f) * scale;
            w[0] = 0; w[1] = 1; w[2] = 0; w[3] = 0;
            pos[o] = (int)floorf(xi) - 1;
            continue;
        }
        const float xi = fmaf((float)o + 0.5f, scale, -0.5f);
        const float pf = floorf(xi);
        const float f = xi + -pf;
        if (algo == 0) bicubic_coeffs(f, A, w);
        else if (algo == 1) lanczos_coeffs(f, w);
        else { w[0] = 0; w[1] = 1.0f + -f; w[2] = f; w[3] = 0; }
        pos[o] = (int)pf - 1;
    }
}

static float chain4(const float w[4], float p0, float p1, float p2, float p3)
{
    float t = w[1] * p1;           /* apply_coeffs :983-992 as contracted: the 2nd product is the plain FMUL */
    t = fmaf(w[0], p0, t);
    t = fmaf(w[2], p2, t);
    t = fmaf(w[3], p3, t);
    return t;
}
/* texture normalised read: texel / 255 (or 65535), correctly rounded */
static float tex_norm(unsigned j, int b16) { return (float)j / (b16 ? 65535.0f : 255.0f); }

static unsigned finish(float v, int b16, int ra, int wrap)
{
    const int maxv = b16 ? 65535 : 255;
    if (ra) { float r = rintf(v); if (r < 0) r = 0; if (r > (float)maxv) r = (float)maxv; return (unsigned)r; }
    float o = v * (b16 ? 65535.0f : 255.0f);
    if (!(o >= 0.0f)) return 0;      /* cvt.rzi.u32.f32: negatives and NaN -> 0 */
    unsigned u = (unsigned)o;
    if (wrap) return u & (unsigned)maxv;   /* the reference stores the low bits */
    return u > (unsigned)maxv ? (unsigned)maxv : u;
}

/* Generic separable resample of a buffer of samples `smp` (sw x sh x ch floats:
 * normalised for R-B, integer-valued for R-A) into integers. */
static void resample_samples(const float *smp, int sw, int sh, int ch, int dw, int dh,
                             const float *cx, const int *px, const float *cy, const int *py,
                             int b16, int ra, int wrap, unsigned *out /* dw*dh*ch */)
{
    float *h = (float *)malloc(sizeof(float) * (size_t)sh * dw * ch);
    for (int y = 0; y < sh; y++)
        for (int xo = 0; xo < dw; xo++)
            for (int c = 0; c < ch; c++) {
                float p[4];
                for (int k = 0; k < 4; k++) p[k] = smp[((size_t)y * sw + iclamp(px[xo] + k, 0, sw - 1)) * ch + c];
                h[((size_t)y * dw + xo) * ch + c] = chain4(cx + 4 * xo, p[0], p[1], p[2], p[3]);
            }
    for (int yo = 0; yo < dh; yo++)
        for (int xo = 0; xo < dw; xo++)
            for (int c = 0; c < ch; c++) {
                float p[4];
                for (int k = 0; k < 4; k++) p[k] = h[((size_t)iclamp(py[yo] + k, 0, sh - 1) * dw + xo) * ch + c];
                out[((size_t)yo * dw + xo) * ch + c] = finish(chain4(cy + 4 * yo, p[0], p[1], p[2], p[3]), b16, ra, wrap);
            }
    free(h);
}

/* packed plane -> packed plane, same component count / depth (rgb->rgb, yuv planes) */
int orc_resample_packed(const uint8_t *src, int spitch, int sw, int sh, uint8_t *dst, int dpitch, int dw, int dh,
                        int ch, int b16, const float *cx, const int *px, const float *cy, const int *py, int ra, int wrap)
{
    float *smp = (float *)malloc(sizeof(float) * (size_t)sw * sh * ch);
    unsigned *out = (unsigned *)malloc(sizeof(unsigned) * (size_t)dw * dh * ch);
    for (int y = 0; y < sh; y++)
        for (int x = 0; x < sw; x++)
            for (int c = 0; c < ch; c++) {
                unsigned j = b16 ? ((const uint16_t *)(src + (size_t)y * spitch))[x * ch + c] : src[(size_t)y * spitch + x * ch + c];
                smp[((size_t)y * sw + x) * ch + c] = ra ? (float)j : tex_norm(j, b16);
            }
    resample_samples(smp, sw, sh, ch, dw, dh, cx, px, cy, py, b16, ra, wrap, out);
    for (int y = 0; y < dh; y++)
        for (int x = 0; x < dw * ch; x++) {
            if (b16) ((uint16_t *)(dst + (size_t)y * dpitch))[x] = out[(size_t)y * dw * ch + x];
            else dst[(size_t)y * dpitch + x] = out[(size_t)y * dw * ch + x];
        }
    free(smp); free(out);
    return 0;
}

/* yuv 4:2:0 -> packed rgb with scaling: CSC at source resolution, quantise, then resample
 * (P-ORDER, swscale_cuda.c:352-370).  One frame (frame index f of a batch). */
int orc_yuv2rgb_scale(const GmatbImage *s, const GmatbImage *d, const float m[9],
                      const float *cx, const int *px, const float *cy, const int *py, int ra, int wrap)
{
    const int nb = s->batch > 1 ? s->batch : 1;
    const int b16 = is16(s->format), sw = s->width, sh = s->height, dw = d->width, dh = d->height;
    const int ch = rgb_ch(d->format), sw_ = rgb_swapped(d->format);
    float *smp = (float *)malloc(sizeof(float) * (size_t)sw * sh * 4);
    unsigned *out = (unsigned *)malloc(sizeof(unsigned) * (size_t)dw * dh * 4);
    for (int f = 0; f < nb; f++) {
        for (int y = 0; y < sh; y++)
            for (int x = 0; x < sw; x++) {
                unsigned Y, U, V, rgb[3];
                get_yuv(s, f, x, y, &Y, &U, &V);
                yuv_to_rgb_q(Y, U, V, b16, m, rgb);
                float *o = smp + ((size_t)y * sw + x) * 4;
                for (int c = 0; c < 3; c++) o[c] = ra ? (float)rgb[c] : tex_norm(rgb[c], b16);
                o[3] = ra ? 255.0f : tex_norm(255, 0);      /* the intermediate's alpha is 255 (8-bit: 1.0; :89) */
                if (!ra && b16) o[3] = tex_norm(255, 1);
            }
        resample_samples(smp, sw, sh, 4, dw, dh, cx, px, cy, py, b16, ra, wrap, out);
        for (int y = 0; y < dh; y++)
            for (int x = 0; x < dw; x++) {
                const unsigned *o = out + ((size_t)y * dw + x) * 4;
                uint8_t *q = (uint8_t *)plane(d, 0, f) + (size_t)y * d->linesize[0] + (size_t)x * ch * (b16 ? 2 : 1);
                const unsigned c0 = sw_ ? o[2] : o[0], c2 = sw_ ? o[0] : o[2];
                if (b16) { uint16_t *w = (uint16_t *)q; w[0] = c0; w[1] = o[1]; w[2] = c2; if (ch == 4) w[3] = o[3]; }
                else { q[0] = c0; q[1] = o[1]; q[2] = c2; if (ch == 4) q[3] = o[3]; }
            }
    }
    free(smp); free(out);
    return 0;
}

/* two-term quotient used by the CUDA kernels vs the correctly rounded division:
 * returns the number of samples j in [0, max] where they differ (must be 0) */
int orc_check_norm(int b16)
{
    const double inv = 1.0 / (b16 ? 65535.0 : 255.0);
    const float khi = (float)inv, klo = (float)(inv - (double)khi);
    int bad = 0;
    for (unsigned j = 0; j <= (b16 ? 65535u : 255u); j++) {
        float t = (float)j * klo;
        float p = fmaf((float)j, khi, t);
        if (p > 1.0f) p = 1.0f;
        if (p != tex_norm(j, b16)) bad++;
    }
    return bad;
}

/* ------------------------------------------------------------------ filters */
static int packed_bpp(int fmt) { return (fmt == GMATB_FMT_RGB24 || fmt == GMATB_FMT_BGR24) ? 3 : 4; }

int orc_crop(const GmatbImage *s, const GmatbImage *d, int cx, int cy)
{
    const int nb = s->batch > 1 ? s->batch : 1, bpp = packed_bpp(s->format);
    for (int f = 0; f < nb; f++)
        for (int y = 0; y < d->height; y++)
            memcpy((uint8_t *)plane(d, 0, f) + (size_t)y * d->linesize[0],
                   plane(s, 0, f) + (size_t)(y + cy) * s->linesize[0] + (size_t)cx * bpp, (size_t)d->width * bpp);
    return 0;
}
int orc_flip(const GmatbImage *s, const GmatbImage *d, int code)
{
    const int nb = s->batch > 1 ? s->batch : 1, bpp = packed_bpp(s->format);
    for (int f = 0; f < nb; f++)
        for (int y = 0; y < d->height; y++)
            for (int x = 0; x < d->width; x++) {
                const int sy = (code <= 0) ? s->height - 1 - y : y;
                const int sx = (code != 0) ? s->width - 1 - x : x;
                memcpy((uint8_t *)plane(d, 0, f) + (size_t)y * d->linesize[0] + (size_t)x * bpp,
                       plane(s, 0, f) + (size_t)sy * s->linesize[0] + (size_t)sx * bpp, bpp);
            }
    return 0;
}

static float cubic_w(float d)
{
    const float A = -0.75f;
    if (d <= 1.0f) return fmaf(fmaf(A + 2.0f, d, -(A + 3.0f)) * d, d, 1.0f);
    if (d < 2.0f)  return fmaf(fmaf(fmaf(A, d, -5.0f * A), d, 8.0f * A), d, -4.0f * A);
    return 0.0f;
}
static int sat_rn(float v) { float r = rintf(v); return r < 0 ? 0 : (r > 255 ? 255 : (int)r); }

int orc_rotate(const GmatbImage *s, const GmatbImage *d, double angle_deg, double shx, double shy, int interp)
{
    const int nb = s->batch > 1 ? s->batch : 1, bpp = packed_bpp(s->format);
    const double rad = angle_deg * 3.14159265358979323846 / 180.0;
    const double c = cos(rad), sn = sin(rad);
    const int W = s->width, H = s->height;
    for (int f = 0; f < nb; f++)
        for (int y = 0; y < d->height; y++)
            for (int x = 0; x < d->width; x++) {
                const double dx = (double)x - shx, dy = (double)y - shy;
                const float sx = (float)(dx * c - dy * sn), sy = (float)(dx * sn + dy * c);
                uint8_t *q = (uint8_t *)plane(d, 0, f) + (size_t)y * d->linesize[0] + (size_t)x * bpp;
                const uint8_t *ps = plane(s, 0, f);
                int out[4] = {0, 0, 0, 0};
                if (sx > -0.5f && sx < (float)W && sy > -0.5f && sy < (float)H) {
                    if (interp == GMATB_INTERP_NEAREST) {
                        int x1 = (int)(sx + 0.5f), y1 = (int)(sy + 0.5f);
                        if (x1 > W - 1) x1 = W - 1;
                        if (y1 > H - 1) y1 = H - 1;
                        for (int k = 0; k < bpp; k++) out[k] = ps[(size_t)y1 * s->linesize[0] + (size_t)x1 * bpp + k];
                    } else if (interp == GMATB_INTERP_CUBIC) {
                        const float fxf = floorf(sx), fyf = floorf(sy);
                        const int ix = (int)fxf, iy = (int)fyf;
                        float wx[4], wy[4], acc[4] = {0, 0, 0, 0};
                        for (int i = 0; i < 4; i++) {
                            wx[i] = cubic_w(fabsf(sx - (fxf + (float)(i - 1))));
                            wy[i] = cubic_w(fabsf(sy - (fyf + (float)(i - 1))));
                        }
                        for (int j = 0; j < 4; j++) {
                            const int yy = iclamp(iy - 1 + j, 0, H - 1);
                            float rowacc[4] = {0, 0, 0, 0};
                            for (int i = 0; i < 4; i++) {
                                const int xx = iclamp(ix - 1 + i, 0, W - 1);
                                for (int k = 0; k < bpp; k++)
                                    rowacc[k] = fmaf(wx[i], (float)ps[(size_t)yy * s->linesize[0] + (size_t)xx * bpp + k], rowacc[k]);
                            }
                            for (int k = 0; k < bpp; k++) acc[k] = fmaf(wy[j], rowacc[k], acc[k]);
                        }
                        for (int k = 0; k < bpp; k++) out[k] = sat_rn(acc[k]);
                    } else {
                        const int x1 = (int)sx, y1 = (int)sy, x2 = x1 + 1, y2 = y1 + 1;
                        const int x2r = x2 < W - 1 ? x2 : W - 1, y2r = y2 < H - 1 ? y2 : H - 1;
                        const float ax = (float)x2 - sx, bx = sx - (float)x1, ay = (float)y2 - sy, by = sy - (float)y1;
                        const float w00 = ax * ay, w01 = bx * ay, w10 = ax * by, w11 = bx * by;
                        for (int k = 0; k < bpp; k++) {
                            float a = (float)ps[(size_t)y1 * s->linesize[0] + (size_t)x1 * bpp + k] * w00;
                            a = fmaf((float)ps[(size_t)y1 * s->linesize[0] + (size_t)x2r * bpp + k], w01, a);
                            a = fmaf((float)ps[(size_t)y2r * s->linesize[0] + (size_t)x1 * bpp + k], w10, a);
                            a = fmaf((float)ps[(size_t)y2r * s->linesize[0] + (size_t)x2r * bpp + k], w11, a);
                            out[k] = sat_rn(a);
                        }
                    }
                }
                for (int k = 0; k < bpp; k++) q[k] = out[k];
            }
    return 0;
}

static int border_idx(int i, int n, int mode)
{
    if (i >= 0 && i < n) return i;
    switch (mode) {
    case GMATB_BORDER_REPLICATE: return i < 0 ? 0 : n - 1;
    case GMATB_BORDER_REFLECT: { if (n == 1) return 0; int p = 2 * n; i %= p; if (i < 0) i += p; return i < n ? i : p - 1 - i; }
    case GMATB_BORDER_REFLECT101: { if (n == 1) return 0; int p = 2 * n - 2; i %= p; if (i < 0) i += p; return i < n ? i : p - i; }
    case GMATB_BORDER_WRAP: { i %= n; if (i < 0) i += n; return i; }
    default: return -1;
    }
}
static void gauss_weights(int k, double sigma, float *out)
{
    if (sigma <= 0.0) sigma = 0.3 * ((k - 1) * 0.5 - 1.0) + 0.8;
    double w[64], sum = 0.0;
    const int r = k / 2;
    for (int i = 0; i < k; i++) { const double x = (double)(i - r); w[i] = exp(-(x * x) / (2.0 * sigma * sigma)); sum += w[i]; }
    for (int i = 0; i < k; i++) out[i] = (float)(w[i] / sum);
}
int orc_gaussian(const GmatbImage *s, const GmatbImage *d, int kw, int kh, double sx_, double sy_, int border)
{
    const int nb = s->batch > 1 ? s->batch : 1, bpp = packed_bpp(s->format);
    float kx[64], ky[64];
    gauss_weights(kw, sx_, kx);
    gauss_weights(kh, sy_ > 0.0 ? sy_ : sx_, ky);
    const int W = s->width, H = s->height, rx = kw / 2, ry = kh / 2;
    float *t = (float *)malloc(sizeof(float) * (size_t)(H + kh) * W * bpp);
    for (int f = 0; f < nb; f++) {
        const uint8_t *ps = plane(s, 0, f);
        /* horizontal pass for every (possibly border-mapped) row y in [-ry, H+ry) */
        for (int yy = -ry; yy < H + ry; yy++) {
            const int sy = border_idx(yy, H, border);
            for (int x = 0; x < W; x++)
                for (int c = 0; c < bpp; c++) {
                    float acc = 0.0f;
                    for (int k = 0; k < kw; k++) {
                        const int sx = border_idx(x + k - rx, W, border);
                        const float v = (sx < 0 || sy < 0) ? 0.0f : (float)ps[(size_t)sy * s->linesize[0] + (size_t)sx * bpp + c];
                        acc = fmaf(kx[k], v, acc);
                    }
                    t[((size_t)(yy + ry) * W + x) * bpp + c] = acc;
                }
        }
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++)
                for (int c = 0; c < bpp; c++) {
                    float acc = 0.0f;
                    for (int k = 0; k < kh; k++) acc = fmaf(ky[k], t[((size_t)(y + k) * W + x) * bpp + c], acc);
                    ((uint8_t *)plane(d, 0, f))[(size_t)y * d->linesize[0] + (size_t)x * bpp + c] = sat_rn(acc);
                }
    }
    free(t);
    return 0;
}
static int cmp_u8(const void *a, const void *b) { return (int)*(const uint8_t *)a - (int)*(const uint8_t *)b; }
int orc_median(const GmatbImage *s, const GmatbImage *d, int kw, int kh)
{
    const int nb = s->batch > 1 ? s->batch : 1, bpp = packed_bpp(s->format);
    const int W = s->width, H = s->height, rx = kw / 2, ry = kh / 2, N = kw * kh;
    uint8_t win[256];
    for (int f = 0; f < nb; f++) {
        const uint8_t *ps = plane(s, 0, f);
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++)
                for (int c = 0; c < bpp; c++) {
                    int n = 0;
                    for (int j = 0; j < kh; j++)
                        for (int i = 0; i < kw; i++)
                            win[n++] = ps[(size_t)iclamp(y + j - ry, 0, H - 1) * s->linesize[0] + (size_t)iclamp(x + i - rx, 0, W - 1) * bpp + c];
                    qsort(win, N, 1, cmp_u8);
                    ((uint8_t *)plane(d, 0, f))[(size_t)y * d->linesize[0] + (size_t)x * bpp + c] = win[N / 2];
                }
    }
    return 0;
}

/* ---------------------------------------------------------------- format_cuda */
/* GetConstants (format_cuda_kernel.cu:32-63): the colourspace whose matrix the filter uploads */
int orc_format_colorspace(int av_colorspace)
{
    switch (av_colorspace) {
    case GMATB_SPC_FCC: case GMATB_SPC_BT470BG: case GMATB_SPC_SMPTE240M: return av_colorspace;
    case GMATB_SPC_BT2020_NCL: case GMATB_SPC_BT2020_CL: return GMATB_SPC_BT2020_NCL;
    default: return GMATB_SPC_BT709;      /* BT709, SMPTE170M, unspecified, ... */
    }
}

/* float -> uint8_t as nvcc compiles it there: F2I.U32.TRUNC (negative -> 0) then the low byte */
static unsigned trunc_u8_wrap(float v)
{
    if (!(v > 0.0f)) return 0;
    if (v >= 4294967296.0f) return 255;     /* F2I.U32 saturates at 0xFFFFFFFF */
    return (unsigned)v & 0xFFu;
}

/* RgbpToYuvKernel<uchar2, RGBAF32, float2> (format_cuda_kernel.cu:516-570).  m = matRgb2Yuv.
 * Even width and height (the reference returns early for the last odd column/row). */
int orc_format_rgbpf32_to_nv12(const GmatbImage *s, const GmatbImage *d, const float m[9])
{
    const int nb = s->batch > 1 ? s->batch : 1;
    if ((s->width | s->height) & 1) return -1;
    for (int f = 0; f < nb; f++)
        for (int y = 0; y < s->height; y += 2)
            for (int x = 0; x < s->width; x += 2) {
                float raw[3][2][2], v[3][2][2];
                for (int c = 0; c < 3; c++)
                    for (int r = 0; r < 2; r++)
                        for (int h = 0; h < 2; h++) {
                            raw[c][r][h] = ((const float *)(plane(s, c, f) + (size_t)(y + r) * s->linesize[c]))[x + h];
                            v[c][r][h] = raw[c][r][h] * 255.0f;
                        }
                uint8_t *py = (uint8_t *)plane(d, 0, f) + (size_t)y * d->linesize[0] + x;
                for (int r = 0; r < 2; r++)
                    for (int h = 0; h < 2; h++) {
                        const float b = (r == 1 && h == 1) ? v[1][1][1] : v[2][r][h];      /* :560 passes g for b */
                        float t = v[1][r][h] * m[1];
                        t = fmaf(v[0][r][h], m[0], t);
                        t = fmaf(b, m[2], t);
                        py[(size_t)r * d->linesize[0] + h] = (uint8_t)trunc_u8_wrap(t + 16.0f);
                    }
                const float rm = (((v[0][0][0] + v[0][0][1]) + v[0][1][0]) + v[0][1][1]) * 0.25f;
                const float gm = (((v[1][0][0] + v[1][0][1]) + v[1][1][0]) + v[1][1][1]) * 0.25f;
                const float bm = fmaf(raw[2][1][1], 255.0f, (v[2][0][0] + v[2][0][1]) + v[2][1][0]) * 0.25f;   /* contracted by nvcc */
                float u = gm * m[4]; u = fmaf(rm, m[3], u); u = fmaf(bm, m[5], u);
                float w = gm * m[7]; w = fmaf(rm, m[6], w); w = fmaf(bm, m[8], w);
                uint8_t *pc = (uint8_t *)plane(d, 1, f) + (size_t)(y >> 1) * d->linesize[1] + x;
                pc[0] = (uint8_t)trunc_u8_wrap(u + 128.0f);
                pc[1] = (uint8_t)trunc_u8_wrap(w + 128.0f);
            }
    return 0;
}
