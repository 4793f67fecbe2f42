/*
 * crop_cuda -- replaces crop_nvcv (libavfilter/vf_crop_nvcv.c).  Options (:80-86): w, h
 * (required, non-zero), x, y (default -1 = centred, :149-150).  The output link and pool take
 * the cropped size (:156-157,171-172).
 */
#include "gmatb_filter_common.h"

typedef struct CropCudaContext {
    GmatbFilterBase base;
    int w, h, x, y;
    int rx, ry;        /* resolved position */
} CropCudaContext;

#define OFFSET(x) offsetof(CropCudaContext, x)
static const AVOption crop_cuda_options[] = {
    { "w", "Width of the cropped area",  OFFSET(w), AV_OPT_TYPE_INT, { .i64 = 0 },  0, INT_MAX, .flags = GMATB_FLAGS },
    { "h", "Height of the cropped area", OFFSET(h), AV_OPT_TYPE_INT, { .i64 = 0 },  0, INT_MAX, .flags = GMATB_FLAGS },
    { "x", "Left edge of the cropped area (-1: centred)", OFFSET(x), AV_OPT_TYPE_INT, { .i64 = -1 }, -1, INT_MAX, .flags = GMATB_FLAGS },
    { "y", "Top edge of the cropped area (-1: centred)",  OFFSET(y), AV_OPT_TYPE_INT, { .i64 = -1 }, -1, INT_MAX, .flags = GMATB_FLAGS },
    { NULL }
};
AVFILTER_DEFINE_CLASS(crop_cuda);

static av_cold int crop_init(AVFilterContext *ctx)
{
    CropCudaContext *s = ctx->priv;
    if (s->w <= 0 || s->h <= 0) {
        av_log(ctx, AV_LOG_ERROR, "w and h are required and must be positive\n");
        return AVERROR(EINVAL);
    }
    return 0;
}

static int crop_config_props(AVFilterLink *outlink)
{
    AVFilterContext *ctx = outlink->src;
    AVFilterLink *inlink = ctx->inputs[0];
    CropCudaContext *s = ctx->priv;
    s->rx = s->x < 0 ? (inlink->w - s->w) / 2 : s->x;
    s->ry = s->y < 0 ? (inlink->h - s->h) / 2 : s->y;
    if (s->rx < 0 || s->ry < 0 || s->rx + s->w > inlink->w || s->ry + s->h > inlink->h) {
        av_log(ctx, AV_LOG_ERROR, "crop window %dx%d@%d,%d exceeds the %dx%d input\n", s->w, s->h, s->rx, s->ry, inlink->w, inlink->h);
        return AVERROR(EINVAL);
    }
    return gmatb_config_output(outlink, &s->base, s->w, s->h);
}
static int crop_launch(AVFilterContext *ctx, const GmatbImage *src, const GmatbImage *dst, void *stream)
{
    CropCudaContext *s = ctx->priv;
    return gmatb_crop(src, dst, s->rx, s->ry, stream);
}
static int crop_filter_frame(AVFilterLink *inlink, AVFrame *in)
{
    return gmatb_filter_frame(inlink, in, crop_launch);
}

static const AVFilterPad crop_cuda_inputs[] = {
    { .name = "default", .type = AVMEDIA_TYPE_VIDEO, .filter_frame = crop_filter_frame },
};
static const AVFilterPad crop_cuda_outputs[] = {
    { .name = "default", .type = AVMEDIA_TYPE_VIDEO, .config_props = crop_config_props },
};

const AVFilter ff_vf_crop_cuda = {
    .name           = "crop_cuda",
    .description    = NULL_IF_CONFIG_SMALL("Crop CUDA frames (gmat_b200 kernels)"),
    FILTER_INPUTS(crop_cuda_inputs),
    FILTER_OUTPUTS(crop_cuda_outputs),
    .priv_class     = &crop_cuda_class,
    .priv_size      = sizeof(CropCudaContext),
    .init           = crop_init,
    .uninit         = gmatb_uninit,
    FILTER_QUERY_FUNC(gmatb_query_formats),
    .flags_internal = FF_FILTER_FLAG_HWFRAME_AWARE,
};
