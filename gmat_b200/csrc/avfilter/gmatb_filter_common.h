/*
 * gmatb_filter_common.h -- glue shared by vf_{crop,rotate,flip,smooth}_cuda.c.
 *
 * Same shape as the reference's CV-CUDA filters (libavfilter/vf_rotate_nvcv.c:103-113
 * query_formats, :151-204 config_props, :206-291 filter_frame; doc/
 * FFmpeg_GPU_Filter_Implementation.md): CUDA frames only, a fresh AVHWFramesContext for
 * the output pool, cuCtxPushCurrent / Pop around the launch, work enqueued on the device
 * context's stream and never synchronised.  Differences, all deliberate:
 *   - no CV-CUDA: the launch goes to the C ABI of include/gmat_b200.h;
 *   - only the public libavutil/hwcontext_cuda.h is needed (the reference includes
 *     hwcontext_cuda_internal.h and with it the un-vendored nv-codec-headers);
 *   - errors are negative AVERROR codes (the reference returns positive EINVAL in places,
 *     vf_crop_nvcv.c:115,153) and `in` is freed on every path.
 */
#ifndef GMATB_FILTER_COMMON_H
#define GMATB_FILTER_COMMON_H

#include <cuda.h>
#include <string.h>

#include "libavfilter/avfilter.h"
#include "libavfilter/formats.h"
#include "libavfilter/internal.h"
#include "libavutil/buffer.h"
#include "libavutil/error.h"
#include "libavutil/frame.h"
#include "libavutil/hwcontext.h"
#include "libavutil/hwcontext_cuda.h"
#include "libavutil/log.h"
#include "libavutil/opt.h"
#include "libavutil/pixdesc.h"
#include "libavutil/pixfmt.h"

#include "gmat_b200.h"

typedef struct GmatbFilterBase {
    const AVClass *class;
    AVBufferRef *hw_frames_ctx;      /* output pool */
    enum AVPixelFormat sw_fmt;
} GmatbFilterBase;

/* formats the kernel layer takes: packed 8-bit rgb, 3 or 4 bytes per pixel
 * (same list as vf_rotate_nvcv.c:92-101) */
static int gmatb_sw_format_ok(enum AVPixelFormat f)
{
    return f == AV_PIX_FMT_RGB24 || f == AV_PIX_FMT_BGR24 || f == AV_PIX_FMT_0RGB32 || f == AV_PIX_FMT_0BGR32 ||
           f == AV_PIX_FMT_RGBA || f == AV_PIX_FMT_BGRA || f == AV_PIX_FMT_RGB0 || f == AV_PIX_FMT_BGR0;
}

static int gmatb_query_formats(AVFilterContext *ctx)
{
    static const enum AVPixelFormat pix_fmts[] = { AV_PIX_FMT_CUDA, AV_PIX_FMT_NONE };
    AVFilterFormats *l = ff_make_format_list((const int *)pix_fmts);
    if (!l)
        return AVERROR(ENOMEM);
    return ff_set_common_formats(ctx, l);
}

/* clone the input pool's geometry (or w x h when given) into a new output pool */
static int gmatb_config_output(AVFilterLink *outlink, GmatbFilterBase *s, int out_w, int out_h)
{
    AVFilterContext *ctx = outlink->src;
    AVFilterLink *inlink = ctx->inputs[0];
    AVHWFramesContext *in_frames, *out_frames;
    AVBufferRef *out_ref;
    int ret;

    if (!inlink->hw_frames_ctx) {
        av_log(ctx, AV_LOG_ERROR, "a CUDA hardware frames context is required on the input\n");
        return AVERROR(EINVAL);
    }
    in_frames = (AVHWFramesContext *)inlink->hw_frames_ctx->data;
    if (!gmatb_sw_format_ok(in_frames->sw_format)) {
        av_log(ctx, AV_LOG_ERROR, "unsupported sw_format %s (packed 8-bit rgb only)\n", av_get_pix_fmt_name(in_frames->sw_format));
        return AVERROR(ENOSYS);
    }
    out_ref = av_hwframe_ctx_alloc(in_frames->device_ref);
    if (!out_ref)
        return AVERROR(ENOMEM);
    out_frames = (AVHWFramesContext *)out_ref->data;
    out_frames->format = AV_PIX_FMT_CUDA;
    out_frames->sw_format = s->sw_fmt = in_frames->sw_format;
    out_frames->width = out_w > 0 ? out_w : in_frames->width;
    out_frames->height = out_h > 0 ? out_h : in_frames->height;
    ret = av_hwframe_ctx_init(out_ref);
    if (ret < 0) {
        av_buffer_unref(&out_ref);
        return ret;
    }
    av_buffer_unref(&s->hw_frames_ctx);
    s->hw_frames_ctx = out_ref;
    outlink->hw_frames_ctx = av_buffer_ref(s->hw_frames_ctx);
    if (!outlink->hw_frames_ctx)
        return AVERROR(ENOMEM);
    if (out_w > 0) outlink->w = out_w;
    if (out_h > 0) outlink->h = out_h;
    return 0;
}

static void gmatb_describe(GmatbImage *g, const AVFrame *f, enum AVPixelFormat sw_fmt)
{
    memset(g, 0, sizeof(*g));
    g->data[0] = f->data[0];
    g->linesize[0] = f->linesize[0];
    g->width = f->width;
    g->height = f->height;
    g->format = (int)sw_fmt;           /* GMATB_FMT_* are AVPixelFormat values */
    g->batch = 1;
}

typedef int (*gmatb_launch_fn)(AVFilterContext *ctx, const GmatbImage *src, const GmatbImage *dst, void *stream);

static int gmatb_filter_frame(AVFilterLink *inlink, AVFrame *in, gmatb_launch_fn launch)
{
    AVFilterContext *ctx = inlink->dst;
    GmatbFilterBase *s = ctx->priv;
    AVFilterLink *outlink = ctx->outputs[0];
    AVHWFramesContext *frames = (AVHWFramesContext *)inlink->hw_frames_ctx->data;
    AVCUDADeviceContext *hw = frames->device_ctx->hwctx;
    AVFrame *out = av_frame_alloc();
    GmatbImage gi, go;
    CUcontext dummy;
    int ret, pushed = 0;

    if (!out) {
        ret = AVERROR(ENOMEM);
        goto fail;
    }
    if (cuCtxPushCurrent(hw->cuda_ctx) != CUDA_SUCCESS) {
        ret = AVERROR_EXTERNAL;
        goto fail;
    }
    pushed = 1;
    ret = av_hwframe_get_buffer(s->hw_frames_ctx, out, 0);
    if (ret < 0)
        goto fail;
    gmatb_describe(&gi, in, s->sw_fmt);
    gmatb_describe(&go, out, s->sw_fmt);
    ret = launch(ctx, &gi, &go, (void *)hw->stream);
    if (ret < 0) {
        av_log(ctx, AV_LOG_ERROR, "gmat_b200 kernel launch failed: %d (cuda %d: %s)\n", ret,
               gmatb_last_cuda_error(), gmatb_last_cuda_error_string());
        ret = ret == GMATB_ERR_INVAL ? AVERROR(EINVAL) : AVERROR_EXTERNAL;
        goto fail;
    }
    cuCtxPopCurrent(&dummy);
    pushed = 0;
    ret = av_frame_copy_props(out, in);
    if (ret < 0)
        goto fail;
    av_frame_free(&in);
    return ff_filter_frame(outlink, out);
fail:
    if (pushed)
        cuCtxPopCurrent(&dummy);
    av_frame_free(&in);
    av_frame_free(&out);
    return ret;
}

static av_cold void gmatb_uninit(AVFilterContext *ctx)
{
    GmatbFilterBase *s = ctx->priv;
    av_buffer_unref(&s->hw_frames_ctx);
}

#define GMATB_FLAGS (AV_OPT_FLAG_VIDEO_PARAM | AV_OPT_FLAG_FILTERING_PARAM)

#endif /* GMATB_FILTER_COMMON_H */
