/*
 * flip_cuda -- replaces flip_nvcv (libavfilter/vf_flip_nvcv.c).  Option `code` (:77-80):
 * 0 = vertical, 1 = horizontal, -1 = both; default 0.
 */
#include "gmatb_filter_common.h"

typedef struct FlipCudaContext {
    GmatbFilterBase base;
    int code;
} FlipCudaContext;

#define OFFSET(x) offsetof(FlipCudaContext, x)
static const AVOption flip_cuda_options[] = {
    { "code", "Flip code: 0 vertical, 1 horizontal, -1 both", OFFSET(code), AV_OPT_TYPE_INT, { .i64 = 0 }, -1, 1, .flags = GMATB_FLAGS },
    { NULL }
};
AVFILTER_DEFINE_CLASS(flip_cuda);

static int flip_config_props(AVFilterLink *outlink)
{
    return gmatb_config_output(outlink, outlink->src->priv, 0, 0);
}
static int flip_launch(AVFilterContext *ctx, const GmatbImage *src, const GmatbImage *dst, void *stream)
{
    FlipCudaContext *s = ctx->priv;
    return gmatb_flip(src, dst, s->code, stream);
}
static int flip_filter_frame(AVFilterLink *inlink, AVFrame *in)
{
    return gmatb_filter_frame(inlink, in, flip_launch);
}

static const AVFilterPad flip_cuda_inputs[] = {
    { .name = "default", .type = AVMEDIA_TYPE_VIDEO, .filter_frame = flip_filter_frame },
};
static const AVFilterPad flip_cuda_outputs[] = {
    { .name = "default", .type = AVMEDIA_TYPE_VIDEO, .config_props = flip_config_props },
};

const AVFilter ff_vf_flip_cuda = {
    .name           = "flip_cuda",
    .description    = NULL_IF_CONFIG_SMALL("Flip CUDA frames (gmat_b200 kernels)"),
    FILTER_INPUTS(flip_cuda_inputs),
    FILTER_OUTPUTS(flip_cuda_outputs),
    .priv_class     = &flip_cuda_class,
    .priv_size      = sizeof(FlipCudaContext),
    .uninit         = gmatb_uninit,
    FILTER_QUERY_FUNC(gmatb_query_formats),
    .flags_internal = FF_FILTER_FLAG_HWFRAME_AWARE,
};
