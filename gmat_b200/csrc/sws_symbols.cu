// sws_symbols.cu -- five of the nine libswscale-internal symbols (include/gmat_b200_sws.h):
// the unscaled converters libswscale's convert_unscaled wrappers call directly.
#include <atomic>
#include <cstring>
#include "common.cuh"
#include "../../include/gmat_b200_sws.h"

static std::atomic<int> g_cspace{0};

static void fill(GmatbImage *g, const uint8_t *const p[], const int stride[], int w, int h, int fmt) {
    memset(g, 0, sizeof(*g));
    const int np = gmatb::fmt_planes(fmt);
    for (int i = 0; i < np && i < 4; i++) { g->data[i] = (void *)p[i]; g->linesize[i] = stride[i]; }
    g->width = w; g->height = h; g->format = fmt; g->batch = 1;
    // the reference's NV12 kernels only ever look at src[0]: UV at src[0] + H*pitch, pitch rounded
    // up to even (yuv2rgb_cuda.cu:213,226).  Honour a caller that passes no second plane.
    if (np == 2 && !g->data[1] && g->data[0]) {
        g->data[1] = (uint8_t *)g->data[0] + (size_t)h * stride[0];
        g->linesize[1] = (stride[0] + 1) / 2 * 2;
    }
    if (fmt == GMATB_FMT_RGBPF32LE && !g->data[1] && g->data[0]) {   // planes stacked (yuv2rgb_cuda.cu:420-422)
        for (int i = 1; i < 3; i++) { g->data[i] = (uint8_t *)g->data[0] + (size_t)i * h * stride[0]; g->linesize[i] = stride[0]; }
    }
}

extern "C" {

void gmatb_set_process_colorspace(int cs) { g_cspace.store(cs); }
int  gmatb_get_process_colorspace(void) { return g_cspace.load(); }

int yuv2rgb_cuda(const uint8_t *src[], int srcStride[], uint8_t *dst[], int dstStride[],
                 int w, int h, int srcFormat, int dstFormat, void *stream) {
    GmatbImage s, d;
    fill(&s, src, srcStride, w, h, srcFormat);
    fill(&d, dst, dstStride, w, h, dstFormat);
    return gmatb_yuv2rgb(&s, &d, g_cspace.load(), stream) == 0 ? 0 : -1;
}
int rgb2yuv_cuda(const uint8_t *src[], int srcStride[], uint8_t *dst[], int dstStride[],
                 int w, int h, int srcFormat, int dstFormat, void *stream) {
    GmatbImage s, d;
    fill(&s, src, srcStride, w, h, srcFormat);
    fill(&d, dst, dstStride, w, h, dstFormat);
    return gmatb_rgb2yuv(&s, &d, g_cspace.load(), stream) == 0 ? 0 : -1;
}
int yuv2yuv_cuda(const uint8_t *src[], int srcStride[], uint8_t *dst[], int dstStride[],
                 int w, int h, int srcFormat, int dstFormat, void *stream) {
    GmatbImage s, d;
    fill(&s, src, srcStride, w, h, srcFormat);
    fill(&d, dst, dstStride, w, h, dstFormat);
    return gmatb_yuv2yuv(&s, &d, stream) == 0 ? 0 : -1;
}
void rgb24tobgr24_cuda(const uint8_t *src[], uint8_t *dst[], int srcStride[], int dstStride[], int w, int h, void *stream) {
    GmatbImage s, d;
    fill(&s, src, srcStride, w, h, GMATB_FMT_RGB24);
    fill(&d, dst, dstStride, w, h, GMATB_FMT_BGR24);
    gmatb_rgb24tobgr24(&s, &d, stream);
}
void rgb2rgb_init_cuda(void) {}

}  // extern "C"
