/*
 * scale_cuda -- the reference's YUV-domain / rgb0 CUDA resizer (libavfilter/vf_scale_cuda.c + vf_scale_cuda.cu)
 * on the gmat_b200 kernel layer.
 *
 * Same surface: options w, h (expressions), interp_algo nearest|bilinear|bicubic|lanczos (default = bicubic,
 * :300-318), format, passthrough (default 1), param (default SCALE_CUDA_PARAM_DEFAULT, bicubic A = -param,
 * vf_scale_cuda.cu:972), force_original_aspect_ratio, force_divisible_by (:598-616); dimensions through
 * ff_scale_eval_dimensions / ff_scale_adjust_dimensions (:374-381); passthrough when geometry and format match
 * (:255-258); a same-size same-format conversion with the default algorithm becomes nearest (:266-268).
 * The work is one gmatb_sws context per link: its bicubic / Lanczos / nearest arithmetic IS vf_scale_cuda.cu's
 * (resample definition R-B: bit-exact against the reference's Subsample_* kernels, tests/test_gpu_scale.py);
 * bilinear is half-pixel-centre fp32 bilinear, not the 9-bit-weight hardware texture filter of the reference.
 * Formats: the 4:2:0 formats and the 4-byte rgb formats of the reference's list (:45-54); yuv444p / yuv444p16
 * are refused with ENOSYS.  Only public libav* headers + <cuda.h> are needed.
 */
#include <float.h>

#include "gmatb_filter_common.h"
#include "libavfilter/scale_eval.h"
#include "libavfilter/video.h"

#define SCALE_CUDA_PARAM_DEFAULT 999999.0f      /* libavfilter/vf_scale_cuda.h:26 */

enum { INTERP_ALGO_DEFAULT, INTERP_ALGO_NEAREST, INTERP_ALGO_BILINEAR, INTERP_ALGO_BICUBIC, INTERP_ALGO_LANCZOS, INTERP_ALGO_COUNT };

typedef struct ScaleCudaContext {
    GmatbFilterBase base;
    enum AVPixelFormat in_fmt, out_fmt, format;
    char *w_expr, *h_expr;
    int passthrough, interp_algo, force_original_aspect_ratio, force_divisible_by;
    float param;
    GmatbSws *sws;
    int in_w, in_h;
} ScaleCudaContext;

static int scale_format_supported(enum AVPixelFormat f)
{
    return f == AV_PIX_FMT_YUV420P || f == AV_PIX_FMT_NV12 || f == AV_PIX_FMT_P010 || f == AV_PIX_FMT_P016 ||
           f == AV_PIX_FMT_0RGB32 || f == AV_PIX_FMT_0BGR32;
}

static av_cold void scale_uninit(AVFilterContext *ctx)
{
    ScaleCudaContext *s = ctx->priv;
    if (s->sws) {
        gmatb_sws_free(s->sws);
        s->sws = NULL;
    }
    gmatb_uninit(ctx);
}

static int scale_config_props(AVFilterLink *outlink)
{
    AVFilterContext *ctx = outlink->src;
    AVFilterLink *inlink = ctx->inputs[0];
    ScaleCudaContext *s = ctx->priv;
    AVHWFramesContext *in_frames, *out_frames;
    AVCUDADeviceContext *hw;
    AVBufferRef *out_ref;
    CUcontext dummy;
    double prm[2];
    int w, h, ret, flags;

    if (!inlink->hw_frames_ctx) {
        av_log(ctx, AV_LOG_ERROR, "No hw context provided on input\n");
        return AVERROR(EINVAL);
    }
    in_frames = (AVHWFramesContext *)inlink->hw_frames_ctx->data;
    hw = in_frames->device_ctx->hwctx;
    if ((ret = ff_scale_eval_dimensions(s, s->w_expr, s->h_expr, inlink, outlink, &w, &h)) < 0)
        return ret;
    ff_scale_adjust_dimensions(inlink, &w, &h, s->force_original_aspect_ratio, s->force_divisible_by);
    outlink->w = w;
    outlink->h = h;
    s->in_w = inlink->w;
    s->in_h = inlink->h;
    s->in_fmt = in_frames->sw_format;
    s->out_fmt = s->format == AV_PIX_FMT_NONE ? s->in_fmt : s->format;
    if (!scale_format_supported(s->in_fmt) || !scale_format_supported(s->out_fmt)) {
        av_log(ctx, AV_LOG_ERROR, "Unsupported format: %s -> %s\n", av_get_pix_fmt_name(s->in_fmt), av_get_pix_fmt_name(s->out_fmt));
        return AVERROR(ENOSYS);
    }
    if (inlink->sample_aspect_ratio.num)
        outlink->sample_aspect_ratio = av_mul_q((AVRational){ outlink->h * inlink->w, outlink->w * inlink->h }, inlink->sample_aspect_ratio);
    else
        outlink->sample_aspect_ratio = inlink->sample_aspect_ratio;

    if (s->passthrough && inlink->w == w && inlink->h == h && s->in_fmt == s->out_fmt) {
        outlink->hw_frames_ctx = av_buffer_ref(inlink->hw_frames_ctx);
        return outlink->hw_frames_ctx ? 0 : AVERROR(ENOMEM);
    }
    s->passthrough = 0;
    if (inlink->w == w && inlink->h == h && s->in_fmt == s->out_fmt && s->interp_algo == INTERP_ALGO_DEFAULT)
        s->interp_algo = INTERP_ALGO_NEAREST;

    out_ref = av_hwframe_ctx_alloc(in_frames->device_ref);
    if (!out_ref)
        return AVERROR(ENOMEM);
    out_frames = (AVHWFramesContext *)out_ref->data;
    out_frames->format = AV_PIX_FMT_CUDA;
    out_frames->sw_format = s->base.sw_fmt = s->out_fmt;
    out_frames->width = FFALIGN(w, 32);
    out_frames->height = FFALIGN(h, 32);
    ret = av_hwframe_ctx_init(out_ref);
    if (ret < 0) {
        av_buffer_unref(&out_ref);
        return ret;
    }
    av_buffer_unref(&s->base.hw_frames_ctx);
    s->base.hw_frames_ctx = out_ref;
    outlink->hw_frames_ctx = av_buffer_ref(out_ref);
    if (!outlink->hw_frames_ctx)
        return AVERROR(ENOMEM);

    flags = s->interp_algo == INTERP_ALGO_NEAREST ? GMATB_SWS_POINT : s->interp_algo == INTERP_ALGO_BILINEAR ? GMATB_SWS_BILINEAR :
            s->interp_algo == INTERP_ALGO_LANCZOS ? GMATB_SWS_LANCZOS : GMATB_SWS_BICUBIC;
    prm[0] = s->param == SCALE_CUDA_PARAM_DEFAULT ? GMATB_SWS_PARAM_DEFAULT : (double)s->param;
    prm[1] = GMATB_SWS_PARAM_DEFAULT;
    if (cuCtxPushCurrent(hw->cuda_ctx) != CUDA_SUCCESS)
        return AVERROR_EXTERNAL;
    if (s->sws)
        gmatb_sws_free(s->sws);
    s->sws = gmatb_sws_create(inlink->w, inlink->h, (int)s->in_fmt, w, h, (int)s->out_fmt, flags | GMATB_SWS_HWACCEL_CUDA, prm, GMATB_SPC_DEFAULT);
    if (s->sws)
        gmatb_sws_set_stream(s->sws, (void *)hw->stream);
    cuCtxPopCurrent(&dummy);
    if (!s->sws) {
        av_log(ctx, AV_LOG_ERROR, "Unsupported conversion: %s %dx%d -> %s %dx%d\n", av_get_pix_fmt_name(s->in_fmt),
               inlink->w, inlink->h, av_get_pix_fmt_name(s->out_fmt), w, h);
        return AVERROR(ENOSYS);
    }
    av_log(ctx, AV_LOG_VERBOSE, "w:%d h:%d fmt:%s -> w:%d h:%d fmt:%s\n", inlink->w, inlink->h, av_get_pix_fmt_name(s->in_fmt),
           w, h, av_get_pix_fmt_name(s->out_fmt));
    return 0;
}

static int scale_filter_frame(AVFilterLink *inlink, AVFrame *in)
{
    AVFilterContext *ctx = inlink->dst;
    ScaleCudaContext *s = ctx->priv;
    AVFilterLink *outlink = ctx->outputs[0];
    AVHWFramesContext *frames = (AVHWFramesContext *)inlink->hw_frames_ctx->data;
    AVCUDADeviceContext *hw = frames->device_ctx->hwctx;
    AVFrame *out = NULL;
    CUcontext dummy;
    int ret, pushed = 0;

    if (s->passthrough)
        return ff_filter_frame(outlink, in);
    out = av_frame_alloc();
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
    out->width = outlink->w;
    out->height = outlink->h;
    ret = gmatb_sws_scale(s->sws, (const uint8_t *const *)in->data, in->linesize, out->data, out->linesize);
    if (ret < 0) {
        av_log(ctx, AV_LOG_ERROR, "gmat_b200 scale failed: %d (cuda %d: %s)\n", ret, gmatb_last_cuda_error(), gmatb_last_cuda_error_string());
        ret = ret == GMATB_ERR_INVAL ? AVERROR(EINVAL) : AVERROR_EXTERNAL;
        goto fail;
    }
    cuCtxPopCurrent(&dummy);
    pushed = 0;
    ret = av_frame_copy_props(out, in);
    if (ret < 0)
        goto fail;
    av_reduce(&out->sample_aspect_ratio.num, &out->sample_aspect_ratio.den,
              (int64_t)in->sample_aspect_ratio.num * outlink->h * inlink->w,
              (int64_t)in->sample_aspect_ratio.den * outlink->w * inlink->h, INT_MAX);
    av_frame_free(&in);
    return ff_filter_frame(outlink, out);
fail:
    if (pushed)
        cuCtxPopCurrent(&dummy);
    av_frame_free(&in);
    av_frame_free(&out);
    return ret;
}

static AVFrame *scale_get_video_buffer(AVFilterLink *inlink, int w, int h)
{
    ScaleCudaContext *s = inlink->dst->priv;
    return s->passthrough ? ff_null_get_video_buffer(inlink, w, h) : ff_default_get_video_buffer(inlink, w, h);
}

#define OFFSET(x) offsetof(ScaleCudaContext, x)
static const AVOption scale_cuda_options[] = {
    { "w", "Output video width",  OFFSET(w_expr), AV_OPT_TYPE_STRING, { .str = "iw" }, .flags = GMATB_FLAGS },
    { "h", "Output video height", OFFSET(h_expr), AV_OPT_TYPE_STRING, { .str = "ih" }, .flags = GMATB_FLAGS },
    { "interp_algo", "Interpolation algorithm used for resizing", OFFSET(interp_algo), AV_OPT_TYPE_INT, { .i64 = INTERP_ALGO_DEFAULT }, 0, INTERP_ALGO_COUNT - 1, GMATB_FLAGS, "interp_algo" },
        { "nearest",  "nearest neighbour", 0, AV_OPT_TYPE_CONST, { .i64 = INTERP_ALGO_NEAREST }, 0, 0, GMATB_FLAGS, "interp_algo" },
        { "bilinear", "bilinear", 0, AV_OPT_TYPE_CONST, { .i64 = INTERP_ALGO_BILINEAR }, 0, 0, GMATB_FLAGS, "interp_algo" },
        { "bicubic",  "bicubic",  0, AV_OPT_TYPE_CONST, { .i64 = INTERP_ALGO_BICUBIC  }, 0, 0, GMATB_FLAGS, "interp_algo" },
        { "lanczos",  "lanczos",  0, AV_OPT_TYPE_CONST, { .i64 = INTERP_ALGO_LANCZOS  }, 0, 0, GMATB_FLAGS, "interp_algo" },
    { "format", "Output video pixel format", OFFSET(format), AV_OPT_TYPE_PIXEL_FMT, { .i64 = AV_PIX_FMT_NONE }, INT_MIN, INT_MAX, .flags = GMATB_FLAGS },
    { "passthrough", "Do not process frames at all if parameters match", OFFSET(passthrough), AV_OPT_TYPE_BOOL, { .i64 = 1 }, 0, 1, GMATB_FLAGS },
    { "param", "Algorithm-Specific parameter", OFFSET(param), AV_OPT_TYPE_FLOAT, { .dbl = SCALE_CUDA_PARAM_DEFAULT }, -FLT_MAX, FLT_MAX, GMATB_FLAGS },
    { "force_original_aspect_ratio", "decrease or increase w/h if necessary to keep the original AR", OFFSET(force_original_aspect_ratio), AV_OPT_TYPE_INT, { .i64 = 0 }, 0, 2, GMATB_FLAGS, "force_oar" },
        { "disable",  NULL, 0, AV_OPT_TYPE_CONST, { .i64 = 0 }, 0, 0, GMATB_FLAGS, "force_oar" },
        { "decrease", NULL, 0, AV_OPT_TYPE_CONST, { .i64 = 1 }, 0, 0, GMATB_FLAGS, "force_oar" },
        { "increase", NULL, 0, AV_OPT_TYPE_CONST, { .i64 = 2 }, 0, 0, GMATB_FLAGS, "force_oar" },
    { "force_divisible_by", "enforce that the output resolution is divisible by a defined integer when force_original_aspect_ratio is used", OFFSET(force_divisible_by), AV_OPT_TYPE_INT, { .i64 = 1 }, 1, 256, GMATB_FLAGS },
    { NULL },
};
AVFILTER_DEFINE_CLASS(scale_cuda);

static const AVFilterPad scale_cuda_inputs[] = {
    { .name = "default", .type = AVMEDIA_TYPE_VIDEO, .filter_frame = scale_filter_frame, .get_buffer.video = scale_get_video_buffer },
};
static const AVFilterPad scale_cuda_outputs[] = {
    { .name = "default", .type = AVMEDIA_TYPE_VIDEO, .config_props = scale_config_props },
};

const AVFilter ff_vf_scale_cuda = {
    .name           = "scale_cuda",
    .description    = NULL_IF_CONFIG_SMALL("GPU accelerated video resizer (gmat_b200 kernels)"),
    FILTER_INPUTS(scale_cuda_inputs),
    FILTER_OUTPUTS(scale_cuda_outputs),
    .priv_class     = &scale_cuda_class,
    .priv_size      = sizeof(ScaleCudaContext),
    .uninit         = scale_uninit,
    FILTER_SINGLE_PIXFMT(AV_PIX_FMT_CUDA),
    .flags_internal = FF_FILTER_FLAG_HWFRAME_AWARE,
};
