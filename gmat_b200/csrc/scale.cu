// scale.cu -- the scaling context behind SwsContext's CUDA path: what
// libswscale/cuda/swscale_cuda.c does with CV-CUDA (init :112-271, per-frame :273-479,
// free :86-110), re-done with our own kernels and no intermediate HBM round trip on the
// yuv->rgb path.  See include/gmat_b200.h for the C ABI.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "scale_fused3.cuh"
#include "scale_bilinear2.cuh"
#include "scale_generic.cuh"
#include "scale_stream.cuh"
#include "scale_plane2.cuh"

namespace gmatb {
int yuv2rgb_launch(const GmatbImage *, const GmatbImage *, const Mat9 &, cudaStream_t);
int yuv2rgb_planar_launch(const GmatbImage *, const GmatbImage *, const Mat9 &, float, const float *, cudaStream_t);
int rgb2yuv_launch(const GmatbImage *, const GmatbImage *, const Mat9 &, cudaStream_t);
int yuv2yuv_launch(const GmatbImage *, const GmatbImage *, cudaStream_t);
int rgb24swap_launch(const GmatbImage *, const GmatbImage *, cudaStream_t);
bool csc_pair_supported(int kind, int srcFormat, int dstFormat);

enum { K_YUV2RGB = 0, K_RGB2YUV = 1, K_YUV2YUV = 2, K_RGB2RGB = 3 };
enum { PATH_UNSCALED = 0, PATH_FUSED2 = 1, PATH_GENERIC = 2 };

static bool is_yuv(int f) {
    return f == GMATB_FMT_NV12 || f == GMATB_FMT_YUV420P || f == GMATB_FMT_P010LE || f == GMATB_FMT_P016LE ||
           f == GMATB_FMT_YUV420P10LE || f == GMATB_FMT_YUV420P16LE;
}
static bool is_rgb(int f) {
    switch (f) {
    case GMATB_FMT_RGB24: case GMATB_FMT_BGR24: case GMATB_FMT_RGBA: case GMATB_FMT_BGRA:
    case GMATB_FMT_RGB0: case GMATB_FMT_BGR0: case GMATB_FMT_0RGB: case GMATB_FMT_0BGR:
    case GMATB_FMT_RGB48LE: case GMATB_FMT_BGR48LE: case GMATB_FMT_RGBA64LE: case GMATB_FMT_BGRA64LE:
    case GMATB_FMT_RGBPF32LE: case GMATB_FMT_RGBAPF32LE: return true;
    default: return false;
    }
}
static int fmt_bits(int f) {
    switch (f) {
    case GMATB_FMT_P010LE: case GMATB_FMT_P016LE: case GMATB_FMT_YUV420P10LE: case GMATB_FMT_YUV420P16LE:
    case GMATB_FMT_RGB48LE: case GMATB_FMT_BGR48LE: case GMATB_FMT_RGBA64LE: case GMATB_FMT_BGRA64LE: return 16;
    case GMATB_FMT_RGBPF32LE: case GMATB_FMT_RGBAPF32LE: return 32;
    default: return 8;
    }
}
static int rgb_channels(int f) {
    switch (f) {
    case GMATB_FMT_RGB24: case GMATB_FMT_BGR24: case GMATB_FMT_RGB48LE: case GMATB_FMT_BGR48LE: return 3;
    default: return 4;
    }
}
static int rgb_dst_code(int fmt) {
    switch (fmt) {
    case GMATB_FMT_RGB24: return D_RGB24;   case GMATB_FMT_BGR24: return D_BGR24;
    case GMATB_FMT_RGBA: case GMATB_FMT_RGB0: return D_RGBA;
    case GMATB_FMT_BGRA: case GMATB_FMT_BGR0: return D_BGRA;
    case GMATB_FMT_RGB48LE: return D_RGB48; case GMATB_FMT_BGR48LE: return D_BGR48;
    case GMATB_FMT_RGBA64LE: return D_RGBA64; case GMATB_FMT_BGRA64LE: return D_BGRA64;
    default: return -1;
    }
}
}  // namespace gmatb

using namespace gmatb;

struct GmatbSws {
    int srcW, srcH, srcFmt, dstW, dstH, dstFmt, flags, cspace;
    double param[2];
    cudaStream_t stream;
    int kind, path, algo, ra;
    float A;
    Mat9 M;
    // filter banks: [0] luma/packed axes, [1] chroma axes (yuv->yuv only)
    float4 *cx[2], *cy[2];
    int *px[2], *py[2];
    std::vector<int> hpx[2], hpy[2];
    float wx[4], wy[4];
    bool taps2;
    int iw;   // 0, or 1..3: both axes have the dyadic weights of the exact-integer kernel (scale_fused4i.cuh)
    // any-ratio streaming kernel (scale_stream.cuh): which strip / outputs each warp owns; built on first use
    float pw2[2][8]; int pw2_state[2];      // exact-2:1 plane kernel (scale_plane2.cuh): the 4 + 4 weights per filter bank; 0 unknown, 1 ready
    int4 *splan[2]; int splan_n[2], splan_nout[2], splan_state[2];   // per filter bank; state: 0 not built, 1 ready, -1 does not apply
    // scratch
    void *tmp; size_t tmp_size;
    void *stage_src, *stage_dst; size_t stage_src_size, stage_dst_size;
    // gmatb_sws_scale_host: copy-in / copy-out streams and the events that link them to `stream`; owned by the
    // context (one context is driven by one thread at a time, like a CPU SwsContext), created on first use on the
    // device that is current then
    struct HostPipe *pipe;
};

// (a, b, b, a) / 2^s as exact floats on both axes: the integer kernel's instantiations
static int dyadic_set(const float *wx, const float *wy) {
    static const struct { int a, b, s; } sets[3] = {{-3, 19, 5}, {-1, 9, 4}, {-1, 5, 3}};
    for (int i = 0; i < 3; i++) {
        const float fa = (float)sets[i].a / (float)(1 << sets[i].s), fb = (float)sets[i].b / (float)(1 << sets[i].s);
        if (wx[0] == fa && wx[3] == fa && wx[1] == fb && wx[2] == fb && wy[0] == fa && wy[3] == fa && wy[1] == fb && wy[2] == fb) return i + 1;
    }
    return 0;
}

#define GMATB_PIPE_EVENTS 16
struct HostPipe;
static void host_pipe_free(HostPipe *hp);

static int build_axis(int algo, int srcN, int dstN, float A, float4 **dc, int **dp, std::vector<int> *hp, cudaStream_t st) {
    if (cudaMalloc(dc, sizeof(float4) * dstN) != cudaSuccess || cudaMalloc(dp, sizeof(int) * dstN) != cudaSuccess)
        return GMATB_ERR_NOMEM;
    filter_table_kernel<<<(dstN + 127) / 128, 128, 0, st>>>(algo, srcN, dstN, A, *dc, *dp);
    count_launch();
    hp->resize(dstN);
    cudaError_t e = cudaMemcpyAsync(hp->data(), *dp, sizeof(int) * dstN, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    return set_cuda_error(e);
}

extern "C" GmatbSws *gmatb_sws_create(int srcW, int srcH, int srcFormat, int dstW, int dstH, int dstFormat,
                                      int flags, const double *param, int colorspace) {
    if (srcW <= 0 || srcH <= 0 || dstW <= 0 || dstH <= 0) return nullptr;
    const bool sy = is_yuv(srcFormat), sr = is_rgb(srcFormat), dy = is_yuv(dstFormat), dr = is_rgb(dstFormat);
    if (!(sy || sr) || !(dy || dr)) return nullptr;
    GmatbSws *c = new GmatbSws();
    memset(c, 0, offsetof(GmatbSws, hpx));
    c->tmp = c->stage_src = c->stage_dst = nullptr;
    c->tmp_size = c->stage_src_size = c->stage_dst_size = 0;
    c->pipe = nullptr;
    c->pw2_state[0] = c->pw2_state[1] = 0;
    for (int i = 0; i < 2; i++) { c->splan[i] = nullptr; c->splan_n[i] = c->splan_nout[i] = c->splan_state[i] = 0; }
    c->srcW = srcW; c->srcH = srcH; c->srcFmt = srcFormat; c->dstW = dstW; c->dstH = dstH; c->dstFmt = dstFormat;
    c->flags = flags; c->cspace = colorspace; c->stream = 0;
    c->param[0] = param ? param[0] : GMATB_SWS_PARAM_DEFAULT;
    c->param[1] = param ? param[1] : GMATB_SWS_PARAM_DEFAULT;
    c->kind = sy ? (dr ? K_YUV2RGB : K_YUV2YUV) : (dy ? K_RGB2YUV : K_RGB2RGB);
    if (c->kind == K_RGB2YUV) gmatb_csc_matrix_rgb2yuv(colorspace, c->M.m);
    else gmatb_csc_matrix_yuv2rgb(colorspace, c->M.m);

    if (srcW == dstW && srcH == dstH) {   // sws_init_context_cuda: unscaled (utils.c:2048-2055)
        c->path = PATH_UNSCALED;
        if (c->kind != K_RGB2RGB && !csc_pair_supported(c->kind, srcFormat, dstFormat)) { delete c; return nullptr; }
        if (c->kind == K_RGB2RGB && srcFormat != dstFormat &&
            !((srcFormat == GMATB_FMT_RGB24 && dstFormat == GMATB_FMT_BGR24) ||
              (srcFormat == GMATB_FMT_BGR24 && dstFormat == GMATB_FMT_RGB24))) { delete c; return nullptr; }
        return c;
    }
    // ---- scaled --------------------------------------------------------------------
    if (fmt_bits(srcFormat) == 32 || fmt_bits(dstFormat) == 32) { delete c; return nullptr; }
    if (c->kind == K_YUV2RGB && (rgb_dst_code(dstFormat) < 0 || fmt_bits(srcFormat) != fmt_bits(dstFormat))) { delete c; return nullptr; }
    if (c->kind == K_RGB2RGB && srcFormat != dstFormat) { delete c; return nullptr; }
    if (c->kind == K_RGB2YUV && (fmt_bits(srcFormat) != 8 || fmt_bits(dstFormat) != 8 || !csc_pair_supported(K_RGB2YUV, srcFormat, dstFormat))) { delete c; return nullptr; }
    if (c->kind == K_YUV2YUV && ((srcW | srcH | dstW | dstH) & 1)) { delete c; return nullptr; }

    // algorithm from the SWS_* bit (the reference intends this mapping, swscale_cuda.c:69-74;
    // its own call site passes the wrong field, :305, and always gets LINEAR)
    // map_resize_algo tests BILINEAR, then BICUBIC, then AREA (swscale_cuda.c:68-73); LANCZOS / POINT are ours.
    // SWS_AREA: the reference maps it to NVCV_INTERP_AREA (CV-CUDA, closed); we have no pinned definition of it and
    // run bilinear -- said once per context on stderr, never silently.
    if (flags & (GMATB_SWS_BILINEAR | GMATB_SWS_FAST_BILINEAR)) c->algo = RS_BILINEAR;
    else if (flags & GMATB_SWS_BICUBIC) c->algo = RS_BICUBIC;
    else if (flags & GMATB_SWS_AREA) { c->algo = RS_BILINEAR; gmatb_log("gmat_b200: SWS_AREA is not implemented, using SWS_BILINEAR"); }
    else if (flags & GMATB_SWS_LANCZOS) c->algo = RS_LANCZOS;
    else if (flags & GMATB_SWS_POINT) c->algo = RS_NEAREST;
    else c->algo = RS_BILINEAR;   // no scaler bit: map_resize_algo's fallthrough (NVCV_INTERP_LINEAR)
    c->ra = (c->algo == RS_BILINEAR || c->algo == RS_NEAREST);
    const bool pdef = (c->param[0] == GMATB_SWS_PARAM_DEFAULT);
    c->A = pdef ? 0.0f : -(float)c->param[0];   // vf_scale_cuda.cu:972

    int rc = build_axis(c->algo, srcW, dstW, c->A, &c->cx[0], &c->px[0], &c->hpx[0], 0);
    if (!rc) rc = build_axis(c->algo, srcH, dstH, c->A, &c->cy[0], &c->py[0], &c->hpy[0], 0);
    if (!rc && c->kind == K_YUV2YUV) {
        rc = build_axis(c->algo, srcW / 2, dstW / 2, c->A, &c->cx[1], &c->px[1], &c->hpx[1], 0);
        if (!rc) rc = build_axis(c->algo, srcH / 2, dstH / 2, c->A, &c->cy[1], &c->py[1], &c->hpy[1], 0);
    }
    if (rc) { gmatb_sws_free(c); return nullptr; }

    c->path = PATH_GENERIC;
    const bool sparse = c->M.m[1] == 0.f && c->M.m[8] == 0.f;
    if (c->kind == K_YUV2RGB && !c->ra && sparse && srcW == 2 * dstW && srcH == 2 * dstH && (srcW % 8) == 0) {
        float4 hx, hy;
        cudaMemcpy(&hx, c->cx[0], sizeof(hx), cudaMemcpyDeviceToHost);
        cudaMemcpy(&hy, c->cy[0], sizeof(hy), cudaMemcpyDeviceToHost);
        c->wx[0] = hx.x; c->wx[1] = hx.y; c->wx[2] = hx.z; c->wx[3] = hx.w;
        c->wy[0] = hy.x; c->wy[1] = hy.y; c->wy[2] = hy.z; c->wy[3] = hy.w;
        c->taps2 = hx.x == 0.f && hx.w == 0.f && hy.x == 0.f && hy.w == 0.f && hx.y == .5f && hx.z == .5f && hy.y == .5f && hy.z == .5f;
        c->iw = (fmt_bits(srcFormat) == 8 && dstW <= 8191 && dstH <= 131071) ? dyadic_set(c->wx, c->wy) : 0;
        c->path = PATH_FUSED2;
    }
    // packed 3-byte rgb -> same format at exactly 2:1: the fused kernel without its colour conversion
    if (c->kind == K_RGB2RGB && !c->ra && (srcFormat == GMATB_FMT_RGB24 || srcFormat == GMATB_FMT_BGR24) &&
        srcW == 2 * dstW && srcH == 2 * dstH && (srcW % 8) == 0) {
        float4 hx, hy;
        cudaMemcpy(&hx, c->cx[0], sizeof(hx), cudaMemcpyDeviceToHost);
        cudaMemcpy(&hy, c->cy[0], sizeof(hy), cudaMemcpyDeviceToHost);
        c->wx[0] = hx.x; c->wx[1] = hx.y; c->wx[2] = hx.z; c->wx[3] = hx.w;
        c->wy[0] = hy.x; c->wy[1] = hy.y; c->wy[2] = hy.z; c->wy[3] = hy.w;
        c->taps2 = hx.x == 0.f && hx.w == 0.f && hy.x == 0.f && hy.w == 0.f && hx.y == .5f && hx.z == .5f && hy.y == .5f && hy.z == .5f;
        c->path = PATH_FUSED2;
    }
    // bilinear at exactly 2:1 from 8-bit yuv: the integer fast path (scale_bilinear2.cuh)
    if (c->kind == K_YUV2RGB && c->algo == RS_BILINEAR && sparse && srcW == 2 * dstW && srcH == 2 * dstH && (srcW % 8) == 0 &&
        fmt_bits(srcFormat) == 8 && rgb_dst_code(dstFormat) <= D_BGRA)
        c->path = PATH_FUSED2;
    return c;
}

extern "C" void gmatb_sws_free(GmatbSws *c) {
    if (!c) return;
    for (int i = 0; i < 2; i++) {
        cudaFree(c->cx[i]); cudaFree(c->cy[i]); cudaFree(c->px[i]); cudaFree(c->py[i]);
    }
    cudaFree(c->tmp); cudaFree(c->stage_src); cudaFree(c->stage_dst); cudaFree(c->splan[0]); cudaFree(c->splan[1]);
    host_pipe_free(c->pipe);
    delete c;
}
extern "C" void gmatb_sws_set_stream(GmatbSws *c, void *stream) { if (c) c->stream = (cudaStream_t)stream; }
extern "C" int gmatb_sws_path(const GmatbSws *c) { return c ? c->path : GMATB_ERR_INVAL; }

extern "C" int gmatb_sws_get_filter(GmatbSws *c, int axis, float *coeffs, int *pos) {
    if (!c || c->path == PATH_UNSCALED || axis < 0 || axis > 1) return GMATB_ERR_INVAL;
    const int n = axis == 0 ? c->dstW : c->dstH;
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess && coeffs) e = cudaMemcpy(coeffs, axis == 0 ? c->cx[0] : c->cy[0], sizeof(float4) * n, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && pos) e = cudaMemcpy(pos, axis == 0 ? c->px[0] : c->py[0], sizeof(int) * n, cudaMemcpyDeviceToHost);
    return set_cuda_error(e);
}

// ------------------------------------------------------------------------------------
namespace gmatb {

static NormK norm_k(int bits) {
    const double inv = 1.0 / (bits == 8 ? 255.0 : 65535.0);
    NormK k;
    k.khi = (float)inv;
    k.klo = (float)(inv - (double)k.khi);
    // geometric-series form (resample_core.cuh): 1/255 = 0x010101 * 2^-24 * (1 + 2^-24 + ...)
    k.c1 = bits == 8 ? 65793.0f / 16777216.0f : 1.0f / 65536.0f;
    k.c0 = -8388608.0f * k.c1;
    k.c2 = bits == 8 ? 1.0f / 16777216.0f : 1.0f / 65536.0f + 1.0f / 4294967296.0f;
    // denormal-domain form (quant_norm2d): c1 * 2^(149 - shift) = 0x010101 * 2^111 (8 bit), 2^127 (16 bit)
    k.kd = ldexpf(bits == 8 ? 65793.0f : 1.0f, bits == 8 ? 111 : 127);
    k.jmax = bits == 8 ? 255 : 65535;
    return k;
}

// MINB = resident warps per SM the register allocation must allow (__launch_bounds__(32, MINB)):
// 16 -> 128 registers, the v3 kernel needs ~120 to keep every constant and both row-pair buffers
// resident (at 96 it spills and rematerialises parameters inside the loop).  The 2-tap instantiation fits 80
// registers, but 16 / 20 / 24 resident warps all measure 1.50-1.54 Tpx/s: it is fma-pipe bound, not latency bound.
// Re-measured on the all-packed kernel (round 2, 64 4K frames): MINB 12 (168 registers) 1010, 16 (128) 1023, 20 (96, spills) 986 Gpx/s.
#ifndef GMATB_FUSED3_MINB
#define GMATB_FUSED3_MINB 16
#endif
template <int L, int SBITS, int DST>
static void launch_fused_t(bool taps2, bool wrap, dim3 g, cudaStream_t st, const Fused3Params &P) {
#define K(T, W) fused_csc_scale2_v3_kernel<L, SBITS, DST, T, W, GMATB_FUSED3_MINB><<<g, 32, 0, st>>>(P)
    if (wrap) { if (taps2) K(true, true); else K(false, true); }
    else      { if (taps2) K(true, false); else K(false, false); }
#undef K
}
template <int L, int SBITS>
static int launch_fused_d(int dc, bool taps2, bool wrap, dim3 g, cudaStream_t st, const Fused3Params &P) {
    if (SBITS == 8) {
        switch (dc) {
        case D_RGB24: launch_fused_t<L, SBITS, D_RGB24>(taps2, wrap, g, st, P); break;
        case D_BGR24: launch_fused_t<L, SBITS, D_BGR24>(taps2, wrap, g, st, P); break;
        case D_RGBA:  launch_fused_t<L, SBITS, D_RGBA>(taps2, wrap, g, st, P); break;
        case D_BGRA:  launch_fused_t<L, SBITS, D_BGRA>(taps2, wrap, g, st, P); break;
        default: return GMATB_ERR_UNSUPPORTED;
        }
    } else {
        switch (dc) {
        case D_RGB48:  launch_fused_t<L, SBITS, D_RGB48>(taps2, wrap, g, st, P); break;
        case D_BGR48:  launch_fused_t<L, SBITS, D_BGR48>(taps2, wrap, g, st, P); break;
        case D_RGBA64: launch_fused_t<L, SBITS, D_RGBA64>(taps2, wrap, g, st, P); break;
        case D_BGRA64: launch_fused_t<L, SBITS, D_BGRA64>(taps2, wrap, g, st, P); break;
        default: return GMATB_ERR_UNSUPPORTED;
        }
    }
    count_launch();
    return set_cuda_error(cudaGetLastError());
}

// exact-integer instantiations (8-bit yuv sources, dyadic weights): scale_int.cu
int fused_int_launch(bool semi, int dc, int iw, bool wrap, dim3 g, cudaStream_t st, const Fused3Params &P);
int fused_mma_launch(int dc, int iw, bool wrap, dim3 g, cudaStream_t st, const Fused3Params &P);
// any-ratio streaming kernel (8-bit yuv -> 8-bit packed rgb): scale_stream.cu
int stream_launch(bool semi, int dc, int nout, int deal, int ra, dim3 g, cudaStream_t st, const StreamParams &P);
int plane_stream_launch(int ch, int bits, int nout, int deal, int ra, dim3 g, cudaStream_t st, const PlaneStreamParams &P);

static bool planes_aligned(const Img &a, int np, int al) {
    for (int i = 0; i < np; i++)
        if (((uintptr_t)a.pl[i].p | (uintptr_t)a.pl[i].pitch | (uintptr_t)a.pl[i].bstride) & (al - 1)) return false;
    return true;
}

static int run_fused(GmatbSws *c, const GmatbImage *src, const GmatbImage *dst, bool *done) {
    *done = false;
    Fused3Params P;
    const int np = is_yuv(src->format) ? fmt_planes(src->format) : 1;
    if (!to_img(src, &P.src, np) || !to_img(dst, &P.dst, 1)) return GMATB_ERR_INVAL;
    const int bits = fmt_bits(src->format);
    const int dc = rgb_dst_code(dst->format);
    const bool semi = np == 2;
    const bool rgbsrc = c->kind == K_RGB2RGB;
    if (c->ra) {      // bilinear 2:1 integer kernel
        if (!planes_aligned(P.src, np, 16) || !planes_aligned(P.dst, 1, (dc == D_RGB24 || dc == D_BGR24) ? 4 : 16)) return 0;
        Bl2Params Q;
        Q.src = P.src; Q.dst = P.dst;
        Q.cm45[0] = c->M.m[4]; Q.cm45[1] = c->M.m[5]; Q.cm72[0] = c->M.m[7]; Q.cm72[1] = c->M.m[2];
        Q.m0 = c->M.m[0]; Q.m1 = c->M.m[1]; Q.m3 = c->M.m[3]; Q.m6 = c->M.m[6];
        const int nbatch = src->batch > 1 ? src->batch : 1;
        const int wx = (c->srcW / 8 + 31) / 32;
        // bands: ~12 waves of 148 SMs x 24 warps (measured on B200, C2 x 64 frames: 3 waves 1952, 6 -> 2016,
        // 12 -> 2092, 24 -> 2078 Gpx/s: short bands even out the tail), but no shorter than 8 row pairs
        const int blw = 12;
        int nbl = (int)((148LL * 24 * blw + (long long)wx * nbatch - 1) / ((long long)wx * nbatch));
        nbl = std::max(1, std::min(nbl, (c->dstH + 7) / 8));
        Q.band = (c->dstH + nbl - 1) / nbl;
        nbl = (c->dstH + Q.band - 1) / Q.band;
        dim3 g(wx, nbl, nbatch);
#define BL(Lx) do { switch (dc) { \
            case D_RGB24: fused_csc_bilinear2_stream_kernel<Lx, D_RGB24, 32><<<g, 32, 0, c->stream>>>(Q); break; \
            case D_BGR24: fused_csc_bilinear2_stream_kernel<Lx, D_BGR24, 32><<<g, 32, 0, c->stream>>>(Q); break; \
            case D_RGBA:  fused_csc_bilinear2_stream_kernel<Lx, D_RGBA, 32><<<g, 32, 0, c->stream>>>(Q); break; \
            default:      fused_csc_bilinear2_stream_kernel<Lx, D_BGRA, 32><<<g, 32, 0, c->stream>>>(Q); break; } } while (0)
        if (semi) BL(L_NV12); else BL(L_I420);
#undef BL
        count_launch();
        int rc0 = set_cuda_error(cudaGetLastError());
        *done = (rc0 == 0);
        return rc0;
    }
    // vector-access preconditions; otherwise the generic kernel takes the frame
    if (!planes_aligned(P.src, 1, bits == 8 ? 8 : 16)) return 0;
    if (!rgbsrc && (semi ? !planes_aligned(P.src, 2, bits == 8 ? 8 : 16) : !(planes_aligned(P.src, 3, bits == 8 ? 4 : 8)))) return 0;
    const int dal = (dc == D_RGB24 || dc == D_BGR24) ? 4 : (dc == D_RGB48 || dc == D_BGR48) ? 8 : 16;
    if (!planes_aligned(P.dst, 1, dal)) return 0;
    P.cm45[0] = c->M.m[4]; P.cm45[1] = c->M.m[5]; P.cm72[0] = c->M.m[7]; P.cm72[1] = c->M.m[2];
    P.m0 = c->M.m[0]; P.m1 = c->M.m[1]; P.m3 = c->M.m[3]; P.m6 = c->M.m[6];
    for (int i = 0; i < 4; i++) { P.wx[i] = c->wx[i]; P.wy[i] = c->wy[i]; }
    P.nk = norm_k(bits);
    P.factor = bits == 8 ? 255.f : 65535.f;
    P.factor_q = P.factor * 0.25f;
    // yuv sources carry their samples scaled by 2^-norm_shift through the chains (quant_norm2d); weights so small that a
    // scaled product could leave the normal range do not occur in any filter table, but are refused all the same
    P.factor_s = ldexpf(P.factor, norm_shift(bits)); P.factor_qs = ldexpf(P.factor_q, norm_shift(bits));
    for (int i = 0; i < 4; i++) {
        if ((c->wx[i] != 0.f && fabsf(c->wx[i]) < 9.094947e-13f) || (c->wy[i] != 0.f && fabsf(c->wy[i]) < 9.094947e-13f)) return 0;
    }
    P.dstW = c->dstW; P.dstH = c->dstH;
    const int batch = src->batch > 1 ? src->batch : 1;
    // 4-tap: strips overlap by one lane per side (30 owning lanes per warp); 2-tap: 32
    const int own = c->taps2 ? 32 : 30;
    const int warps_x = (c->srcW / 8 + own - 1) / own;
    // enough warps to fill 148 SMs x 16 resident warps a few times over, but bands no
    // shorter than 32 output rows (each band converts 2 extra row pairs to prime its accumulators; measured:
    // 8 -> 32 rows is +7 % on 4-frame launches, neutral on 64-frame ones)
    const int fwaves = 8;   // measured: 2 -> 952, 4 -> 975, 8 -> 994, 16 -> 968 Gpx/s
    long long want = 148LL * 16 * fwaves;
    int nb = (int)((want + (long long)warps_x * batch - 1) / ((long long)warps_x * batch));
    const int min_band = 32;
    nb = std::max(1, std::min(nb, (c->dstH + min_band - 1) / min_band));
    P.band = (c->dstH + nb - 1) / nb;
    nb = (c->dstH + P.band - 1) / P.band;
    dim3 g(warps_x, nb, batch);
    int rc;
    const bool wrap = (c->flags & GMATB_SWS_PARITY_WRAP) != 0;
    if (rgbsrc) {
        // same component order in and out: run the RGB24 instantiation for both rgb24 and bgr24
        launch_fused_t<L_RGB3, 8, D_RGB24>(c->taps2, wrap, g, c->stream, P);
        count_launch();
        rc = set_cuda_error(cudaGetLastError());
    } else if (c->iw && semi && (c->flags & GMATB_SWS_MMA_CHAIN) && dc <= D_BGRA) {
        rc = fused_mma_launch(dc, c->iw, wrap, g, c->stream, P);
    } else if (c->iw && (c->flags & GMATB_SWS_INT_CHAIN)) {
        rc = fused_int_launch(semi, dc, c->iw, wrap, g, c->stream, P);
    } else if (semi) rc = bits == 8 ? launch_fused_d<L_NV12, 8>(dc, c->taps2, wrap, g, c->stream, P) : launch_fused_d<L_NV12, 16>(dc, c->taps2, wrap, g, c->stream, P);
    else      rc = bits == 8 ? launch_fused_d<L_I420, 8>(dc, c->taps2, wrap, g, c->stream, P) : launch_fused_d<L_I420, 16>(dc, c->taps2, wrap, g, c->stream, P);
    *done = (rc == 0);
    return rc;
}

// The warps of the streaming kernel: greedy cut of the output columns.  A warp converts source columns [X0, X0 + 8 lanes)
// (X0 a multiple of 8, at most 256 columns) and owns outputs [xo, xo + n): every tap of every owned output, clamped to the
// frame, must lie in its strip (taps left of column 0 / right of column W-1 are the replicated pad entries).
static bool build_stream_plan(GmatbSws *c, int bank, int W, int srcH, int dW) {
    const std::vector<int> &px = c->hpx[bank];
    if (W < 16 || srcH < 2 || dW < 1 || (int)px.size() != dW) return false;
    const double r = (double)W / dW;
    if (r > 48.0) return false;
    const int nout = r >= 2.5 ? 3 : 5;
    std::vector<int4> plan;
    int xo = 0;
    while (xo < dW) {
        const int X0 = std::max(px[xo], 0) & ~7;
        int n = 0, last = X0;
        const int cap = std::min(32 * nout, dW - xo);
        while (n < cap) {
            const int right = std::min(px[xo + n] + 3, W - 1);
            if (right > X0 + 255) break;
            last = std::max(last, right);
            n++;
        }
        if (xo + n < dW) n &= ~3;                 // every warp but the last starts the next one on a multiple of 4
        if (n <= 0) return false;
        last = X0;
        for (int i = 0; i < n; i++) last = std::max(last, std::min(px[xo + i] + 3, W - 1));
        plan.push_back(make_int4(X0, xo, n, (last - X0) / 8 + 1));
        xo += n;
    }
    if (cudaMalloc(&c->splan[bank], plan.size() * sizeof(int4)) != cudaSuccess) return false;
    if (cudaMemcpy(c->splan[bank], plan.data(), plan.size() * sizeof(int4), cudaMemcpyHostToDevice) != cudaSuccess) return false;
    c->splan_n[bank] = (int)plan.size(); c->splan_nout[bank] = nout;
    return true;
}

// 8-bit yuv 4:2:0 -> 8-bit packed rgb at any ratio, R-B or R-A arithmetic: the streaming kernel
static int run_stream(GmatbSws *c, const GmatbImage *src, const GmatbImage *dst, bool *done) {
    *done = false;
    if (c->splan_state[0] < 0) return 0;
    const int dc = rgb_dst_code(dst->format);
    if (fmt_bits(src->format) != 8 || dc < 0 || dc > D_BGRA || !(c->M.m[1] == 0.f && c->M.m[8] == 0.f)) { c->splan_state[0] = -1; return 0; }
    if (c->splan_state[0] == 0) {
        if (!build_stream_plan(c, 0, c->srcW, c->srcH, c->dstW)) { c->splan_state[0] = -1; return 0; }
        c->splan_state[0] = 1;
    }
    StreamParams P;
    memset(&P, 0, sizeof(P));
    const int np = fmt_planes(src->format);
    if (!to_img(src, &P.F.src, np) || !to_img(dst, &P.F.dst, 1)) return GMATB_ERR_INVAL;
    const bool semi = np == 2;
    if (!planes_aligned(P.F.src, 1, 8) || (semi ? !planes_aligned(P.F.src, 2, 8) : !planes_aligned(P.F.src, 3, 4))) return 0;
    if (!planes_aligned(P.F.dst, 1, 4)) return 0;
    // the last chunk of a row is read whole: the pitch must cover it
    if (P.F.src.pl[0].pitch < ((c->srcW + 7) & ~7)) return 0;
    if (semi ? P.F.src.pl[1].pitch < ((c->srcW + 7) & ~7) : (P.F.src.pl[1].pitch < ((c->srcW + 7) & ~7) / 2 || P.F.src.pl[2].pitch < ((c->srcW + 7) & ~7) / 2)) return 0;
    P.F.cm45[0] = c->M.m[4]; P.F.cm45[1] = c->M.m[5]; P.F.cm72[0] = c->M.m[7]; P.F.cm72[1] = c->M.m[2];
    P.F.m0 = c->M.m[0]; P.F.m1 = c->M.m[1]; P.F.m3 = c->M.m[3]; P.F.m6 = c->M.m[6];
    P.F.nk = norm_k(8);
    P.F.factor = 255.f;
    P.F.dstW = c->dstW; P.F.dstH = c->dstH;
    P.cx = c->cx[0]; P.cy = c->cy[0]; P.px = c->px[0]; P.py = c->py[0];
    P.plan = c->splan[0];
    P.wrap = (c->flags & GMATB_SWS_PARITY_WRAP) ? 1 : 0;
    const int batch = src->batch > 1 ? src->batch : 1;
    // bands: enough warps for a few waves of 148 SMs x 12-16 resident warps, no shorter than 16 output rows
    long long want = 148LL * 16 * 6;
    int nb = (int)((want + (long long)c->splan_n[0] * batch - 1) / ((long long)c->splan_n[0] * batch));
    nb = std::max(1, std::min(nb, (c->dstH + 15) / 16));
    P.band = (c->dstH + nb - 1) / nb;
    nb = (c->dstH + P.band - 1) / P.band;
    dim3 g(c->splan_n[0], nb, batch);
    // paired deal of the outputs when two outputs advance ~3 source columns (ratios 1.25 .. 1.75): see the kernel
    const double r = (double)c->srcW / c->dstW;
    const int deal = (c->splan_nout[0] == 5 && r >= 1.25 && r <= 1.75) ? 1 : 0;
    int rc = stream_launch(semi, dc, c->splan_nout[0], deal, c->ra ? 1 : 0, g, c->stream, P);
    *done = (rc == 0);
    return rc;
}

// generic kernel on one "image" (a yuv frame -> rgb, or one packed plane -> same layout)
static int run_generic(GmatbSws *c, int bank, const Img &s, const Img &d, int dW, int dH, int src_kind, int ch,
                       int bits, int dst_code, int batch) {
    GenParams P;
    P.src = s; P.dst = d; P.M = c->M; P.nk = norm_k(bits);
    P.factor = bits == 8 ? 255.f : 65535.f; P.vmax = P.factor;
    P.wrap = (c->flags & GMATB_SWS_PARITY_WRAP) ? 1 : 0;
    P.cx = c->cx[bank]; P.cy = c->cy[bank]; P.px = c->px[bank]; P.py = c->py[bank];
    P.dstW = dW; P.dstH = dH; P.src_kind = src_kind; P.ch = ch; P.dst_code = dst_code;
    const std::vector<int> &hx = c->hpx[bank], &hy = c->hpy[bank];
    const int st = ch == 3 ? 4 : ch;              // floats per pixel in shared memory
    int tw = 32, th = 64;
    size_t smem = 0;
    for (;;) {   // shrink the tile until the window fits in shared memory; several CTAs per SM wanted
        int mw = 0, mh = 0;
        // the kernel rounds the window origin down (x to 4, y to 2: chroma grid, word loads): measure from there
        for (int x = 0; x < dW; x += tw) mw = std::max(mw, hx[std::min(x + tw, dW) - 1] + 4 - (hx[x] & ~3));
        for (int y = 0; y < dH; y += th) mh = std::max(mh, hy[std::min(y + th, dH) - 1] + 4 - (hy[y] & ~1));
        mw = (mw + 3) & ~3; mh = (mh + 1) & ~1;
        // window: planar ch floats per pixel for 3/4 components (rounded to 16 bytes), else st; horizontal results: st
        smem = (st == 4 ? (((size_t)mw * mh * ch + 3) & ~(size_t)3) : (size_t)mw * mh * st) * sizeof(float) + (size_t)mh * tw * st * sizeof(float);
        P.win_w = mw; P.win_h = mh;
        const size_t budget = 64 * 1024;
        if (smem <= budget || (tw == 1 && th == 1)) break;
        if (th > 8) th /= 2; else if (tw >= 2 * th && tw > 8) tw /= 2; else if (th > 1) th /= 2; else tw /= 2;
    }
    if (smem > 200 * 1024) return GMATB_ERR_UNSUPPORTED;
    P.tile_w = tw; P.tile_h = th;
    P.tile_shift = 0;
    while ((1 << P.tile_shift) < tw) P.tile_shift++;
    const int sb = bits / 8;
    auto al = [](const Plane &pl, int a) { return ((((uintptr_t)pl.p) | (uintptr_t)pl.pitch | (uintptr_t)pl.bstride) & (uintptr_t)(a - 1)) == 0; };
    P.aligned = src_kind != GS_PACKED && al(s.pl[0], 4 * sb) && al(s.pl[1], src_kind == GS_NV12 ? 4 * sb : 2 * sb) &&
                (src_kind == GS_NV12 || al(s.pl[2], 2 * sb));
    P.dst_vec = al(d.pl[0], 4 * sb);
    dim3 g((dW + tw - 1) / tw, (dH + th - 1) / th, batch > 1 ? batch : 1);
    cudaError_t e = cudaSuccess;
#define GO(S, B, C, R) do { \
        if (smem > 48 * 1024) e = cudaFuncSetAttribute(generic_scale_kernel<S, B, C, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e == cudaSuccess) generic_scale_kernel<S, B, C, R><<<g, 256, smem, c->stream>>>(P); } while (0)
#define GO_R(S, B, C) do { if (c->ra) GO(S, B, C, 1); else GO(S, B, C, 0); } while (0)
#define GO_B(S, C) do { if (bits == 8) GO_R(S, 8, C); else GO_R(S, 16, C); } while (0)
    if (src_kind == GS_NV12) GO_B(GS_NV12, 3);
    else if (src_kind == GS_I420) GO_B(GS_I420, 3);
    else if (ch == 1) GO_B(GS_PACKED, 1);
    else if (ch == 2) GO_B(GS_PACKED, 2);
    else if (ch == 3) GO_B(GS_PACKED, 3);
    else if (ch == 4) GO_B(GS_PACKED, 4);
    else return GMATB_ERR_UNSUPPORTED;
#undef GO_B
#undef GO_R
#undef GO
    if (e != cudaSuccess) return set_cuda_error(e);
    count_launch();
    return set_cuda_error(cudaGetLastError());
}

static int ensure(void **p, size_t *have, size_t need) {
    if (*have >= need) return 0;
    cudaFree(*p); *p = nullptr; *have = 0;
    if (cudaMalloc(p, need) != cudaSuccess) return GMATB_ERR_NOMEM;
    *have = need;
    return 0;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// describe a tightly laid out (pitch aligned to 256) image of `fmt` in a scratch buffer
static size_t layout_image(GmatbImage *g, int fmt, int w, int h, int batch, uint8_t *base) {
    memset(g, 0, sizeof(*g));
    g->format = fmt; g->width = w; g->height = h; g->batch = batch;
    size_t off = 0, frame = 0;
    const int bytes = fmt_bits(fmt) / 8;
    int np = fmt_planes(fmt);
    size_t psize[4] = {0, 0, 0, 0};
    if (is_yuv(fmt)) {
        g->linesize[0] = (int)align_up((size_t)w * bytes, 256);
        psize[0] = (size_t)g->linesize[0] * h;
        const int cw = (w + 1) / 2, chh = (h + 1) / 2;
        if (np == 2) { g->linesize[1] = (int)align_up((size_t)cw * 2 * bytes, 256); psize[1] = (size_t)g->linesize[1] * chh; }
        else { g->linesize[1] = g->linesize[2] = (int)align_up((size_t)cw * bytes, 256); psize[1] = psize[2] = (size_t)g->linesize[1] * chh; }
    } else {
        np = 1;
        g->linesize[0] = (int)align_up((size_t)w * rgb_channels(fmt) * bytes, 256);
        psize[0] = (size_t)g->linesize[0] * h;
    }
    for (int i = 0; i < np; i++) frame += psize[i];
    for (int i = 0; i < np; i++) { g->data[i] = base ? base + off : nullptr; g->batch_stride[i] = (long long)frame; off += psize[i]; }
    return frame * (batch > 1 ? batch : 1);
}

static void fix_nv12_uv(GmatbImage *g) {
    // the reference's NV12/P016 kernels take one pointer and find UV at src + H*pitch
    // (yuv2rgb_cuda.cu:226); accept callers that pass only data[0]
    if ((g->format == GMATB_FMT_NV12 || g->format == GMATB_FMT_P010LE || g->format == GMATB_FMT_P016LE) && !g->data[1] && g->data[0]) {
        g->data[1] = (uint8_t *)g->data[0] + (size_t)g->height * g->linesize[0];
        g->linesize[1] = g->linesize[0];
        g->batch_stride[1] = g->batch_stride[0];
    }
}

static int plane_resample(GmatbSws *c, int bank, const GmatbImage *s, const GmatbImage *d, int plane, int pw, int ph,
                          int dw, int dh, int ch, int bits) {
    Img si, di;
    memset(&si, 0, sizeof(si)); memset(&di, 0, sizeof(di));
    si.w = pw; si.h = ph; di.w = dw; di.h = dh;
    si.pl[0].p = (uint8_t *)s->data[plane]; si.pl[0].pitch = s->linesize[plane]; si.pl[0].bstride = s->batch > 1 ? s->batch_stride[plane] : 0;
    di.pl[0].p = (uint8_t *)d->data[plane]; di.pl[0].pitch = d->linesize[plane]; di.pl[0].bstride = d->batch > 1 ? d->batch_stride[plane] : 0;
    if (!si.pl[0].p || !di.pl[0].p) return GMATB_ERR_INVAL;
    // exactly 2:1, R-B: the register-streaming plane kernel (scale_plane2.cuh)
    // (its taps are at 2 o - 1 .. 2 o + 2 on both axes: true of bicubic / Lanczos / bilinear at 2:1, not of nearest)
    const bool plane_ok = ch == 1 || ch == 2 || ((ch == 3 || ch == 4) && bits == 8);
    if (plane_ok && ch != 3 && !(c->flags & GMATB_SWS_TILE_KERNEL) && c->algo != RS_NEAREST && c->hpx[bank][0] == -1 && c->hpy[bank][0] == -1 && pw == 2 * dw && ph == 2 * dh && (pw % 8) == 0 &&
        planes_aligned(si, 1, ch * bits == 8 ? 8 : 16) && planes_aligned(di, 1, std::min(16, 4 * ch * bits / 8))) {
        if (!c->pw2_state[bank]) {
            float4 hx, hy;
            if (cudaMemcpy(&hx, c->cx[bank], sizeof(hx), cudaMemcpyDeviceToHost) != cudaSuccess ||
                cudaMemcpy(&hy, c->cy[bank], sizeof(hy), cudaMemcpyDeviceToHost) != cudaSuccess) return set_cuda_error(cudaGetLastError());
            const float w[8] = {hx.x, hx.y, hx.z, hx.w, hy.x, hy.y, hy.z, hy.w};
            memcpy(c->pw2[bank], w, sizeof(w));
            c->pw2_state[bank] = 1;
        }
        Plane2Params Q;
        memset(&Q, 0, sizeof(Q));
        Q.src = si.pl[0]; Q.dst = di.pl[0];
        Q.W = pw; Q.H = ph; Q.dstW = dw; Q.dstH = dh;
        for (int i = 0; i < 4; i++) { Q.wx[i] = c->pw2[bank][i]; Q.wy[i] = c->pw2[bank][4 + i]; }
        Q.nk = norm_k(bits);
        Q.wrap = (c->flags & GMATB_SWS_PARITY_WRAP) ? 1 : 0;
        const int batch = s->batch > 1 ? s->batch : 1;
        const int warps_x = (pw / 8 + 29) / 30;
        long long want = 148LL * 16 * 8;
        int nb = (int)((want + (long long)warps_x * batch - 1) / ((long long)warps_x * batch));
        nb = std::max(1, std::min(nb, (dh + 31) / 32));
        Q.band = (dh + nb - 1) / nb;
        nb = (dh + Q.band - 1) / Q.band;
        dim3 g(warps_x, nb, batch);
#define P2(C_, B_) do { if (c->ra) plane_scale2_kernel<C_, B_, 1><<<g, 32, 0, c->stream>>>(Q); else plane_scale2_kernel<C_, B_, 0><<<g, 32, 0, c->stream>>>(Q); } while (0)
        if (bits == 8) { if (ch == 1) P2(1, 8); else if (ch == 2) P2(2, 8); else P2(4, 8); }
        else           { if (ch == 1) P2(1, 16); else P2(2, 16); }
#undef P2
        count_launch();
        return set_cuda_error(cudaGetLastError());
    }
    // 8- and 16-bit planes of 1 or 2 components: the streaming kernel (scale_stream.cuh)
    const int bp = ch * bits / 8;
    if (plane_ok && !(c->flags & GMATB_SWS_TILE_KERNEL) && c->splan_state[bank] >= 0 &&
        planes_aligned(si, 1, (bp == 1 || bp == 3) ? 8 : 16) && planes_aligned(di, 1, 4) && si.pl[0].pitch >= ((pw + 7) & ~7) * bp) {
        if (c->splan_state[bank] == 0) c->splan_state[bank] = build_stream_plan(c, bank, pw, ph, dw) ? 1 : -1;
        if (c->splan_state[bank] == 1) {
            PlaneStreamParams P;
            memset(&P, 0, sizeof(P));
            P.src = si.pl[0]; P.dst = di.pl[0];
            P.W = pw; P.H = ph; P.dstW = dw; P.dstH = dh;
            P.nk = norm_k(bits);
            P.cx = c->cx[bank]; P.cy = c->cy[bank]; P.px = c->px[bank]; P.py = c->py[bank];
            P.plan = c->splan[bank];
            P.wrap = (c->flags & GMATB_SWS_PARITY_WRAP) ? 1 : 0;
            const int batch = s->batch > 1 ? s->batch : 1;
            long long want = 148LL * 16 * 6;
            int nb = (int)((want + (long long)c->splan_n[bank] * batch - 1) / ((long long)c->splan_n[bank] * batch));
            nb = std::max(1, std::min(nb, (dh + 15) / 16));
            P.band = (dh + nb - 1) / nb;
            nb = (dh + P.band - 1) / P.band;
            const double r = (double)pw / dw;
            const int deal = (c->splan_nout[bank] == 5 && r >= 1.25 && r <= 1.75) ? 1 : 0;
            return plane_stream_launch(ch, bits, c->splan_nout[bank], deal, c->ra ? 1 : 0, dim3(c->splan_n[bank], nb, batch), c->stream, P);
        }
    }
    return run_generic(c, bank, si, di, dw, dh, GS_PACKED, ch, bits, 0, s->batch);
}

static int scale_batch(GmatbSws *c, const GmatbImage *src_in, const GmatbImage *dst_in) {
    if (!c || !src_in || !dst_in) return GMATB_ERR_INVAL;
    GmatbImage src = *src_in, dst = *dst_in;
    if (src.width != c->srcW || src.height != c->srcH || dst.width != c->dstW || dst.height != c->dstH ||
        src.format != c->srcFmt || dst.format != c->dstFmt) return GMATB_ERR_INVAL;
    if ((src.batch > 1 ? src.batch : 1) != (dst.batch > 1 ? dst.batch : 1)) return GMATB_ERR_INVAL;
    fix_nv12_uv(&src); fix_nv12_uv(&dst);
    const int batch = src.batch > 1 ? src.batch : 1;

    if (c->path == PATH_UNSCALED) {
        switch (c->kind) {
        case K_YUV2RGB:
            if (fmt_bits(dst.format) == 32) {
                if (dst.format == GMATB_FMT_RGBPF32LE && !dst.data[1]) {   // single pointer, planes stacked (yuv2rgb_cuda.cu:420-422)
                    for (int p = 1; p < 3; p++) {
                        dst.data[p] = (uint8_t *)dst.data[0] + (size_t)p * dst.height * dst.linesize[0];
                        dst.linesize[p] = dst.linesize[0]; dst.batch_stride[p] = dst.batch_stride[0];
                    }
                }
                return yuv2rgb_planar_launch(&src, &dst, c->M, 255.0f, nullptr, c->stream);
            }
            return yuv2rgb_launch(&src, &dst, c->M, c->stream);
        case K_RGB2YUV: return rgb2yuv_launch(&src, &dst, c->M, c->stream);
        case K_YUV2YUV: return yuv2yuv_launch(&src, &dst, c->stream);
        default:
            if (src.format != dst.format) return rgb24swap_launch(&src, &dst, c->stream);
            {
                const size_t row = (size_t)src.width * rgb_channels(src.format) * (fmt_bits(src.format) / 8);
                for (int i = 0; i < batch; i++) {
                    cudaError_t e = cudaMemcpy2DAsync((uint8_t *)dst.data[0] + i * dst.batch_stride[0], dst.linesize[0],
                                                      (const uint8_t *)src.data[0] + i * src.batch_stride[0], src.linesize[0],
                                                      row, src.height, cudaMemcpyDeviceToDevice, c->stream);
                    if (e != cudaSuccess) return set_cuda_error(e);
                }
                return 0;
            }
        }
    }

    const int bits = fmt_bits(src.format);
    if (c->kind == K_YUV2RGB) {
        if (c->path == PATH_FUSED2) {
            bool done = false;
            int rc = run_fused(c, &src, &dst, &done);
            if (rc || done) return rc;
        }
        if (!(c->flags & GMATB_SWS_TILE_KERNEL)) {
            bool done = false;
            int rc = run_stream(c, &src, &dst, &done);
            if (rc || done) return rc;
        }
        Img s, d;
        const int np = fmt_planes(src.format);
        if (!to_img(&src, &s, np) || !to_img(&dst, &d, 1)) return GMATB_ERR_INVAL;
        return run_generic(c, 0, s, d, c->dstW, c->dstH, np == 2 ? GS_NV12 : GS_I420, 3, bits, rgb_dst_code(dst.format), batch);
    }
    if (c->kind == K_RGB2RGB) {
        if (c->path == PATH_FUSED2) {
            bool done = false;
            int rc = run_fused(c, &src, &dst, &done);
            if (rc || done) return rc;
        }
        return plane_resample(c, 0, &src, &dst, 0, c->srcW, c->srcH, c->dstW, c->dstH, rgb_channels(src.format), bits);
    }
    if (c->kind == K_RGB2YUV) {   // resize first, convert at destination size (swscale_cuda.c:312-341)
        GmatbImage tmp;
        size_t need = layout_image(&tmp, src.format, c->dstW, c->dstH, batch, nullptr);
        if (ensure(&c->tmp, &c->tmp_size, need)) return GMATB_ERR_NOMEM;
        layout_image(&tmp, src.format, c->dstW, c->dstH, batch, (uint8_t *)c->tmp);
        int rc = plane_resample(c, 0, &src, &tmp, 0, c->srcW, c->srcH, c->dstW, c->dstH, rgb_channels(src.format), bits);
        if (rc) return rc;
        return rgb2yuv_launch(&tmp, &dst, c->M, c->stream);
    }
    // K_YUV2YUV: repack to the destination format at source size, then scale each plane
    // (swscale_cuda.c:372-476)
    GmatbImage cur = src;
    if (src.format != dst.format) {
        GmatbImage tmp;
        size_t need = layout_image(&tmp, dst.format, c->srcW, c->srcH, batch, nullptr);
        if (ensure(&c->tmp, &c->tmp_size, need)) return GMATB_ERR_NOMEM;
        layout_image(&tmp, dst.format, c->srcW, c->srcH, batch, (uint8_t *)c->tmp);
        int rc = yuv2yuv_launch(&src, &tmp, c->stream);
        if (rc) return rc;
        cur = tmp;
    }
    const int dbits = fmt_bits(dst.format);
    int rc = plane_resample(c, 0, &cur, &dst, 0, c->srcW, c->srcH, c->dstW, c->dstH, 1, dbits);
    if (rc) return rc;
    if (fmt_planes(dst.format) == 2)
        return plane_resample(c, 1, &cur, &dst, 1, c->srcW / 2, c->srcH / 2, c->dstW / 2, c->dstH / 2, 2, dbits);
    rc = plane_resample(c, 1, &cur, &dst, 1, c->srcW / 2, c->srcH / 2, c->dstW / 2, c->dstH / 2, 1, dbits);
    if (rc) return rc;
    return plane_resample(c, 1, &cur, &dst, 2, c->srcW / 2, c->srcH / 2, c->dstW / 2, c->dstH / 2, 1, dbits);
}

// byte span [lo, hi) covered by an image's planes (all frames)
static void image_span(const GmatbImage *g, uintptr_t *lo, uintptr_t *hi) {
    *lo = ~(uintptr_t)0; *hi = 0;
    const int batch = g->batch > 1 ? g->batch : 1;
    const int np = is_yuv(g->format) ? fmt_planes(g->format) : (fmt_bits(g->format) == 32 ? fmt_planes(g->format) : 1);
    for (int p = 0; p < np; p++) {
        if (!g->data[p]) continue;
        int rows = g->height;
        if (is_yuv(g->format) && p > 0) rows = (g->height + 1) / 2;
        uintptr_t a = (uintptr_t)g->data[p];
        uintptr_t b = a + (size_t)(batch - 1) * g->batch_stride[p] + (size_t)rows * g->linesize[p];
        *lo = std::min(*lo, a); *hi = std::max(*hi, b);
    }
}

}  // namespace gmatb

extern "C" int gmatb_sws_scale_batch(GmatbSws *c, const GmatbImage *src, const GmatbImage *dst) {
    return scale_batch(c, src, dst);
}

extern "C" int gmatb_sws_scale(GmatbSws *c, const uint8_t *const src[4], const int srcStride[4],
                               uint8_t *const dst[4], const int dstStride[4]) {
    if (!c || !src || !dst || !srcStride || !dstStride) return GMATB_ERR_INVAL;
    GmatbImage s, d;
    memset(&s, 0, sizeof(s)); memset(&d, 0, sizeof(d));
    for (int i = 0; i < 4; i++) { s.data[i] = (void *)src[i]; s.linesize[i] = srcStride[i]; d.data[i] = dst[i]; d.linesize[i] = dstStride[i]; }
    s.width = c->srcW; s.height = c->srcH; s.format = c->srcFmt; s.batch = 1;
    d.width = c->dstW; d.height = c->dstH; d.format = c->dstFmt; d.batch = 1;
    return scale_batch(c, &s, &d);
}

// HOST frames in, HOST frames out.  The device staging buffers mirror the host layout
// byte for byte (same strides).  The batch is cut into chunks of frames that flow through
// three streams -- copy-in, convert, copy-out -- linked by events, so that the PCIe uplink,
// the kernels and the PCIe downlink overlap (the link is full duplex; the reference API has
// no host-buffer call at all: its callers cudaMemcpy around sws_scale themselves).
struct HostPipe {
    cudaStream_t in, out;
    cudaEvent_t ev_in[GMATB_PIPE_EVENTS], ev_k[GMATB_PIPE_EVENTS];
    int nev;
    bool ok;
};
static HostPipe *host_pipe(GmatbSws *c) {
    if (c->pipe) return c->pipe;
    HostPipe *hp = new HostPipe();
    hp->in = hp->out = nullptr; hp->nev = 0;
    hp->ok = cudaStreamCreateWithFlags(&hp->in, cudaStreamNonBlocking) == cudaSuccess &&
             cudaStreamCreateWithFlags(&hp->out, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < GMATB_PIPE_EVENTS && hp->ok; i++) {
        hp->ok = cudaEventCreateWithFlags(&hp->ev_in[i], cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&hp->ev_k[i], cudaEventDisableTiming) == cudaSuccess;
        if (hp->ok) hp->nev = i + 1;
    }
    c->pipe = hp;
    return hp;
}
static void host_pipe_free(HostPipe *hp) {
    if (!hp) return;
    for (int i = 0; i < hp->nev; i++) { cudaEventDestroy(hp->ev_in[i]); cudaEventDestroy(hp->ev_k[i]); }
    if (hp->in) cudaStreamDestroy(hp->in);
    if (hp->out) cudaStreamDestroy(hp->out);
    delete hp;
}

extern "C" int gmatb_sws_scale_host(GmatbSws *c, const GmatbImage *src_host, const GmatbImage *dst_host) {
    if (!c || !src_host || !dst_host) return GMATB_ERR_INVAL;
    GmatbImage s = *src_host, d = *dst_host;
    fix_nv12_uv(&s); fix_nv12_uv(&d);
    const int n = s.batch > 1 ? s.batch : 1;
    if ((d.batch > 1 ? d.batch : 1) != n) return GMATB_ERR_INVAL;
    uintptr_t slo, shi, dlo, dhi;
    image_span(&s, &slo, &shi); image_span(&d, &dlo, &dhi);
    if (shi <= slo || dhi <= dlo) return GMATB_ERR_INVAL;
    // keep the low address bits so that alignment-dependent fast paths still apply
    const size_t spad = slo & 255, dpad = dlo & 255;
    if (ensure(&c->stage_src, &c->stage_src_size, (shi - slo) + 256) || ensure(&c->stage_dst, &c->stage_dst_size, (dhi - dlo) + 256))
        return GMATB_ERR_NOMEM;
    uint8_t *ds = (uint8_t *)c->stage_src + spad, *dd = (uint8_t *)c->stage_dst + dpad;
    GmatbImage sdev = s, ddev = d;
    for (int p = 0; p < 4; p++) {
        if (s.data[p]) sdev.data[p] = ds + ((uintptr_t)s.data[p] - slo);
        if (d.data[p]) ddev.data[p] = dd + ((uintptr_t)d.data[p] - dlo);
    }
    HostPipe *hp = host_pipe(c);
    // chunking needs frames that are whole, equally spaced byte ranges (the FrameBatch / frame-pool layout)
    bool chunkable = hp->ok && n > 1;
    const long long sfs = s.batch_stride[0], dfs = d.batch_stride[0];
    for (int p = 0; p < 4 && chunkable; p++) {
        if (s.data[p] && s.batch_stride[p] != sfs) chunkable = false;
        if (d.data[p] && d.batch_stride[p] != dfs) chunkable = false;
    }
    if (chunkable && ((long long)(shi - slo) > sfs * n || (long long)(dhi - dlo) > dfs * n)) chunkable = false;
    if (!chunkable) {
        cudaError_t e = cudaMemcpyAsync(ds, (const void *)slo, shi - slo, cudaMemcpyHostToDevice, c->stream);
        if (e != cudaSuccess) return set_cuda_error(e);
        int rc = scale_batch(c, &sdev, &ddev);
        if (rc) return rc;
        e = cudaMemcpyAsync((void *)dlo, dd, dhi - dlo, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        return set_cuda_error(e);
    }
    int per = (n + 7) / 8;                       // ~8 chunks in flight (<= GMATB_PIPE_EVENTS: every chunk has its own events)
    if (per < 1) per = 1;
    const int nchunks = (n + per - 1) / per;
    cudaError_t e = cudaSuccess;
    int rc = 0;
    for (int k = 0; k < nchunks && e == cudaSuccess && !rc; k++) {
        const int f0 = k * per, fn = std::min(per, n - f0);
        const size_t so = (size_t)f0 * sfs, sb = (f0 + fn == n) ? (shi - slo) - so : (size_t)fn * sfs;
        const size_t dofs = (size_t)f0 * dfs, db = (f0 + fn == n) ? (dhi - dlo) - dofs : (size_t)fn * dfs;
        e = cudaMemcpyAsync(ds + so, (const uint8_t *)slo + so, sb, cudaMemcpyHostToDevice, hp->in);
        if (e == cudaSuccess) e = cudaEventRecord(hp->ev_in[k], hp->in);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(c->stream, hp->ev_in[k], 0);
        if (e != cudaSuccess) break;
        GmatbImage sc = sdev, dc = ddev;
        sc.batch = fn; dc.batch = fn;
        for (int p = 0; p < 4; p++) {
            if (sc.data[p]) sc.data[p] = (uint8_t *)sc.data[p] + (size_t)f0 * s.batch_stride[p];
            if (dc.data[p]) dc.data[p] = (uint8_t *)dc.data[p] + (size_t)f0 * d.batch_stride[p];
        }
        rc = scale_batch(c, &sc, &dc);
        if (rc) break;
        e = cudaEventRecord(hp->ev_k[k], c->stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(hp->out, hp->ev_k[k], 0);
        if (e == cudaSuccess) e = cudaMemcpyAsync((uint8_t *)dlo + dofs, dd + dofs, db, cudaMemcpyDeviceToHost, hp->out);
    }
    cudaError_t e2 = cudaStreamSynchronize(hp->in);
    cudaError_t e3 = cudaStreamSynchronize(c->stream);
    cudaError_t e4 = cudaStreamSynchronize(hp->out);
    if (rc) return rc;
    if (e == cudaSuccess) e = e2;
    if (e == cudaSuccess) e = e3;
    if (e == cudaSuccess) e = e4;
    return set_cuda_error(e);
}
