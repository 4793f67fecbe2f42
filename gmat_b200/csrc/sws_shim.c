/*
 * sws_shim.c -- the four SwsContext-taking symbols of the libswscale boundary
 * (include/gmat_b200_sws.h).  Plain C, compiled against the reference's own
 * libswscale/swscale_internal.h (SwsContext is an internal struct), calling the
 * kernel layer only through the C ABI of include/gmat_b200.h.
 *
 * State lives in the SwsContext fields GMAT added for CV-CUDA
 * (swscale_internal.h:687-691): cv_resize_handle holds our GmatbSws*.
 */
#include "libswscale/swscale.h"
#include "libswscale/swscale_internal.h"
#include "libavutil/error.h"
#include "libavutil/log.h"

#include "../../include/gmat_b200.h"
#include "../../include/gmat_b200_sws.h"

/* av_log lives in libavutil, which is always present when this object is linked into
 * ffmpeg-gpu; weak so that the library can also be loaded on its own by the ABI tests. */
#pragma weak av_log
#define LOG(...) do { if (av_log) av_log(__VA_ARGS__); } while (0)

void ff_yuv2rgb_init_tables_cuda(SwsContext *c)
{
    /* reference: set_mat_yuv2rgb_cuda / set_mat_rgb2yuv_cuda(c->cspace) into process-global
     * constant memory (swscale_cuda.c:76-84).  c->cspace is never assigned by the reference
     * (always 0 -> BT.601 limited); whatever it holds is honoured. */
    gmatb_set_process_colorspace((int)c->cspace);
}

int ff_sws_init_swscale_cuda(SwsContext *c)
{
    GmatbSws *g = gmatb_sws_create(c->srcW, c->srcH, (int)c->srcFormat, c->dstW, c->dstH, (int)c->dstFormat,
                                   c->flags, c->param, (int)c->cspace);
    if (!g) {
        LOG(c, AV_LOG_ERROR, "gmat_b200: unsupported conversion %dx%d fmt %d -> %dx%d fmt %d\n",
               c->srcW, c->srcH, (int)c->srcFormat, c->dstW, c->dstH, (int)c->dstFormat);
        return AVERROR(EINVAL);
    }
    c->cv_resize_handle = g;
    ff_yuv2rgb_init_tables_cuda(c);
    return 0;
}

int ff_swscale_cuda(SwsContext *c, const uint8_t *src[], int srcStride[], int srcSliceY, int srcSliceH,
                    uint8_t *dst[], int dstStride[], int dstSliceY, int dstSliceH)
{
    GmatbSws *g = (GmatbSws *)c->cv_resize_handle;
    int ret;
    if (!g)
        return AVERROR(EINVAL);
    /* like the reference only whole frames are converted */
    gmatb_sws_set_stream(g, (void *)c->cuda_stream);
    ret = gmatb_sws_scale(g, (const uint8_t *const *)src, srcStride, (uint8_t *const *)dst, dstStride);
    if (ret < 0) {
        LOG(c, AV_LOG_ERROR, "gmat_b200: scale failed (%d, cuda %d: %s)\n", ret,
               gmatb_last_cuda_error(), gmatb_last_cuda_error_string());
        return AVERROR_EXTERNAL;
    }
    return 0;
}

int ff_sws_free_swscale_cuda(SwsContext *c)
{
    if (c->convert_unscaled)
        return 0;
    gmatb_sws_free((GmatbSws *)c->cv_resize_handle);
    c->cv_resize_handle = NULL;
    return 0;
}
