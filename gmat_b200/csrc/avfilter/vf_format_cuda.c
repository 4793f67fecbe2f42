/*
 * format_cuda -- replaces the reference's format_cuda filter (libavfilter/vf_format_cuda.c + its
 * format_cuda_kernel.cu) on the gmat_b200 kernel layer.
 *
 * Same surface: option `pix_fmt` (vf_format_cuda.c:69-72, required, :87-91), supported targets nv12 and
 * rgbpf32le (:76-79); input sw_format nv12 -> rgbpf32le output via nv12_to_rgbpf32 (:185-195), rgbpf32le
 * input -> nv12 via rgbpf32_to_nv12 (:197-200), colourspace taken from the frame (in->colorspace);
 * the output pool is sized to the link for rgbpf32le ("tighten", :100-101,160-161) and to the input pool
 * otherwise.  Differences, all deliberate: an unsupported pix_fmt fails in init (the reference computes
 * `ret` at :103-110 and then returns 0), `in` is freed and the CUDA context popped on every path, only the
 * public libavutil/hwcontext_cuda.h is needed, and every plane pointer / pitch of both frames is honoured
 * (the reference's rgbpf32_to_nv12 assumes contiguous planes at height*pitch).
 */
#include "gmatb_filter_common.h"

typedef struct FormatCudaContext {
    GmatbFilterBase base;
    char *pix_fmt;
    enum AVPixelFormat in_fmt, out_fmt;
} FormatCudaContext;

#define OFFSET(x) offsetof(FormatCudaContext, x)
static const AVOption format_cuda_options[] = {
    { "pix_fmt", "Target pixel format (nv12 or rgbpf32le)", OFFSET(pix_fmt), AV_OPT_TYPE_STRING, .flags = GMATB_FLAGS },
    { NULL }
};
AVFILTER_DEFINE_CLASS(format_cuda);

static av_cold int format_init(AVFilterContext *ctx)
{
    FormatCudaContext *s = ctx->priv;
    if (!s->pix_fmt) {
        av_log(ctx, AV_LOG_ERROR, "No output pixel format specified.\n");
        return AVERROR(EINVAL);
    }
    s->out_fmt = av_get_pix_fmt(s->pix_fmt);
    if (s->out_fmt != AV_PIX_FMT_NV12 && s->out_fmt != AV_PIX_FMT_RGBPF32LE) {
        av_log(ctx, AV_LOG_ERROR, "Unsupported target pixel format '%s' (nv12, rgbpf32le).\n", s->pix_fmt);
        return AVERROR(EINVAL);
    }
    return 0;
}

static int format_config_props(AVFilterLink *outlink)
{
    AVFilterContext *ctx = outlink->src;
    AVFilterLink *inlink = ctx->inputs[0];
    FormatCudaContext *s = ctx->priv;
    AVHWFramesContext *in_frames, *out_frames;
    AVBufferRef *out_ref;
    const int tighten = s->out_fmt == AV_PIX_FMT_RGBPF32LE;
    int ret;

    if (!inlink->hw_frames_ctx) {
        av_log(ctx, AV_LOG_ERROR, "a CUDA hardware frames context is required on the input\n");
        return AVERROR(EINVAL);
    }
    in_frames = (AVHWFramesContext *)inlink->hw_frames_ctx->data;
    s->in_fmt = in_frames->sw_format;
    if (!((s->in_fmt == AV_PIX_FMT_NV12 && s->out_fmt == AV_PIX_FMT_RGBPF32LE) ||
          (s->in_fmt == AV_PIX_FMT_RGBPF32LE && s->out_fmt == AV_PIX_FMT_NV12))) {
        av_log(ctx, AV_LOG_ERROR, "Unsupported input/output pixel format combination.\n");
        return AVERROR(EINVAL);
    }
    out_ref = av_hwframe_ctx_alloc(in_frames->device_ref);
    if (!out_ref)
        return AVERROR(ENOMEM);
    out_frames = (AVHWFramesContext *)out_ref->data;
    out_frames->format = AV_PIX_FMT_CUDA;
    out_frames->sw_format = s->base.sw_fmt = s->out_fmt;
    out_frames->width = tighten ? inlink->w : in_frames->width;
    out_frames->height = tighten ? inlink->h : in_frames->height;
    ret = av_hwframe_ctx_init(out_ref);
    if (ret < 0) {
        av_buffer_unref(&out_ref);
        return ret;
    }
    av_buffer_unref(&s->base.hw_frames_ctx);
    s->base.hw_frames_ctx = out_ref;
    outlink->hw_frames_ctx = av_buffer_ref(s->base.hw_frames_ctx);
    if (!outlink->hw_frames_ctx)
        return AVERROR(ENOMEM);
    return 0;
}

static void format_describe(GmatbImage *g, const AVFrame *f, enum AVPixelFormat sw_fmt, int w, int h)
{
    const int np = sw_fmt == AV_PIX_FMT_NV12 ? 2 : 3;
    memset(g, 0, sizeof(*g));
    for (int i = 0; i < np; i++) {
        g->data[i] = f->data[i];
        g->linesize[i] = f->linesize[i];
    }
    g->width = w;
    g->height = h;
    g->format = (int)sw_fmt;
    g->batch = 1;
}

static int format_filter_frame(AVFilterLink *inlink, AVFrame *in)
{
    AVFilterContext *ctx = inlink->dst;
    FormatCudaContext *s = ctx->priv;
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
    ret = av_hwframe_get_buffer(s->base.hw_frames_ctx, out, 0);
    if (ret < 0)
        goto fail;
    format_describe(&gi, in, s->in_fmt, in->width, in->height);
    format_describe(&go, out, s->out_fmt, in->width, in->height);
    if (s->out_fmt == AV_PIX_FMT_RGBPF32LE)     /* vf_format_cuda.c:193: norm 255, no shift, R,G,B plane order */
        ret = gmatb_format_nv12_to_rgbpf32(&gi, &go, (int)in->colorspace, 255.0f, NULL, 0, (void *)hw->stream);
    else                                        /* vf_format_cuda.c:198 */
        ret = gmatb_format_rgbpf32_to_nv12(&gi, &go, (int)in->colorspace, (void *)hw->stream);
    if (ret < 0) {
        av_log(ctx, AV_LOG_ERROR, "gmat_b200 format kernel failed: %d (cuda %d: %s)\n", ret,
               gmatb_last_cuda_error(), gmatb_last_cuda_error_string());
        ret = ret == GMATB_ERR_INVAL ? AVERROR(EINVAL) : ret == GMATB_ERR_UNSUPPORTED ? AVERROR(ENOSYS) : AVERROR_EXTERNAL;
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

static const AVFilterPad format_cuda_inputs[] = {
    { .name = "default", .type = AVMEDIA_TYPE_VIDEO, .filter_frame = format_filter_frame },
};
static const AVFilterPad format_cuda_outputs[] = {
    { .name = "default", .type = AVMEDIA_TYPE_VIDEO, .config_props = format_config_props },
};

const AVFilter ff_vf_format_cuda = {
    .name           = "format_cuda",
    .description    = NULL_IF_CONFIG_SMALL("Convert CUDA frames between NV12 and planar float RGB (gmat_b200 kernels)"),
    FILTER_INPUTS(format_cuda_inputs),
    FILTER_OUTPUTS(format_cuda_outputs),
    .priv_class     = &format_cuda_class,
    .priv_size      = sizeof(FormatCudaContext),
    .init           = format_init,
    .uninit         = gmatb_uninit,
    FILTER_QUERY_FUNC(gmatb_query_formats),
    .flags_internal = FF_FILTER_FLAG_HWFRAME_AWARE,
};
