/*
 * smooth_cuda -- replaces smooth_nvcv (libavfilter/vf_smooth_nvcv.c).  Options (:88-103): type
 * (default|gaussian|median), kw / kh (3), border_type (constant|replicate|reflect|warp|reflect101),
 * sigmaX / sigmaY (0 = derived from the kernel size).  The reference's switch statements fall
 * through (:130-138, :288-296) so that it effectively always runs the median; here `type` means
 * what it says and "default" is gaussian.
 */
#include "gmatb_filter_common.h"

enum { SMOOTH_DEFAULT = 0, SMOOTH_GAUSSIAN = 1, SMOOTH_MEDIAN = 2 };

typedef struct SmoothCudaContext {
    GmatbFilterBase base;
    int type, kw, kh, border_type;
    double sigma_x, sigma_y;
} SmoothCudaContext;

#define OFFSET(x) offsetof(SmoothCudaContext, x)
static const AVOption smooth_cuda_options[] = {
    { "type", "Type of smooth filter", OFFSET(type), AV_OPT_TYPE_INT, { .i64 = SMOOTH_DEFAULT }, 0, 2, .flags = GMATB_FLAGS, "type" },
        { "default",  "gaussian", 0, AV_OPT_TYPE_CONST, { .i64 = SMOOTH_DEFAULT },  0, 0, .flags = GMATB_FLAGS, "type" },
        { "gaussian", "",         0, AV_OPT_TYPE_CONST, { .i64 = SMOOTH_GAUSSIAN }, 0, 0, .flags = GMATB_FLAGS, "type" },
        { "median",   "",         0, AV_OPT_TYPE_CONST, { .i64 = SMOOTH_MEDIAN },   0, 0, .flags = GMATB_FLAGS, "type" },
    { "kw", "Kernel width",  OFFSET(kw), AV_OPT_TYPE_INT, { .i64 = 3 }, 1, GMATB_GAUSS_MAXK, .flags = GMATB_FLAGS },
    { "kh", "Kernel height", OFFSET(kh), AV_OPT_TYPE_INT, { .i64 = 3 }, 1, GMATB_GAUSS_MAXK, .flags = GMATB_FLAGS },
    { "border_type", "Border mode", OFFSET(border_type), AV_OPT_TYPE_INT, { .i64 = GMATB_BORDER_CONSTANT }, 0, 4, .flags = GMATB_FLAGS, "border" },
        { "constant",   "", 0, AV_OPT_TYPE_CONST, { .i64 = GMATB_BORDER_CONSTANT },   0, 0, .flags = GMATB_FLAGS, "border" },
        { "replicate",  "", 0, AV_OPT_TYPE_CONST, { .i64 = GMATB_BORDER_REPLICATE },  0, 0, .flags = GMATB_FLAGS, "border" },
        { "reflect",    "", 0, AV_OPT_TYPE_CONST, { .i64 = GMATB_BORDER_REFLECT },    0, 0, .flags = GMATB_FLAGS, "border" },
        { "warp",       "", 0, AV_OPT_TYPE_CONST, { .i64 = GMATB_BORDER_WRAP },       0, 0, .flags = GMATB_FLAGS, "border" },
        { "reflect101", "", 0, AV_OPT_TYPE_CONST, { .i64 = GMATB_BORDER_REFLECT101 }, 0, 0, .flags = GMATB_FLAGS, "border" },
    { "sigmaX", "Gaussian sigma in x (0: from kw)", OFFSET(sigma_x), AV_OPT_TYPE_DOUBLE, { .dbl = 0.0 }, 0, 1e3, .flags = GMATB_FLAGS },
    { "sigmaY", "Gaussian sigma in y (0: sigmaX)",  OFFSET(sigma_y), AV_OPT_TYPE_DOUBLE, { .dbl = 0.0 }, 0, 1e3, .flags = GMATB_FLAGS },
    { NULL }
};
AVFILTER_DEFINE_CLASS(smooth_cuda);

static int smooth_config_props(AVFilterLink *outlink)
{
    AVFilterContext *ctx = outlink->src;
    AVFilterLink *inlink = ctx->inputs[0];
    SmoothCudaContext *s = ctx->priv;
    if (s->kw > inlink->w || s->kh > inlink->h) {      /* vf_smooth_nvcv.c:172-175 */
        av_log(ctx, AV_LOG_ERROR, "kernel %dx%d larger than the %dx%d frame\n", s->kw, s->kh, inlink->w, inlink->h);
        return AVERROR(EINVAL);
    }
    /* fail at graph configuration, not on the first frame: both kernels take odd windows, the median up to
     * GMATB_MEDIAN_MAXK, the gaussian up to GMATB_GAUSS_MAXK (include/gmat_b200.h) */
    if (!(s->kw & 1) || !(s->kh & 1)) {
        av_log(ctx, AV_LOG_ERROR, "kernel sizes must be odd (got %dx%d)\n", s->kw, s->kh);
        return AVERROR(EINVAL);
    }
    if (s->type == SMOOTH_MEDIAN && (s->kw > GMATB_MEDIAN_MAXK || s->kh > GMATB_MEDIAN_MAXK)) {
        av_log(ctx, AV_LOG_ERROR, "median kernel %dx%d larger than %dx%d\n", s->kw, s->kh, GMATB_MEDIAN_MAXK, GMATB_MEDIAN_MAXK);
        return AVERROR(EINVAL);
    }
    return gmatb_config_output(outlink, &s->base, 0, 0);
}
static int smooth_launch(AVFilterContext *ctx, const GmatbImage *src, const GmatbImage *dst, void *stream)
{
    SmoothCudaContext *s = ctx->priv;
    if (s->type == SMOOTH_MEDIAN)
        return gmatb_median(src, dst, s->kw, s->kh, stream);
    return gmatb_gaussian(src, dst, s->kw, s->kh, s->sigma_x, s->sigma_y, s->border_type, stream);
}
static int smooth_filter_frame(AVFilterLink *inlink, AVFrame *in)
{
    return gmatb_filter_frame(inlink, in, smooth_launch);
}

static const AVFilterPad smooth_cuda_inputs[] = {
    { .name = "default", .type = AVMEDIA_TYPE_VIDEO, .filter_frame = smooth_filter_frame },
};
static const AVFilterPad smooth_cuda_outputs[] = {
    { .name = "default", .type = AVMEDIA_TYPE_VIDEO, .config_props = smooth_config_props },
};

const AVFilter ff_vf_smooth_cuda = {
    .name           = "smooth_cuda",
    .description    = NULL_IF_CONFIG_SMALL("Gaussian / median smoothing of CUDA frames (gmat_b200 kernels)"),
    FILTER_INPUTS(smooth_cuda_inputs),
    FILTER_OUTPUTS(smooth_cuda_outputs),
    .priv_class     = &smooth_cuda_class,
    .priv_size      = sizeof(SmoothCudaContext),
    .uninit         = gmatb_uninit,
    FILTER_QUERY_FUNC(gmatb_query_formats),
    .flags_internal = FF_FILTER_FLAG_HWFRAME_AWARE,
};
