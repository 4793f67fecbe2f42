/*
 * rotate_cuda -- replaces rotate_nvcv (libavfilter/vf_rotate_nvcv.c).  Same options, same
 * defaults (:79-88): angle (degrees, -360..360, 0), interp (linear|nearest|cubic|area,
 * "linear"), shift_x / shift_y (0).  Same-size output, pixels that map outside the source are 0.
 */
#include "gmatb_filter_common.h"

typedef struct RotateCudaContext {
    GmatbFilterBase base;
    char *interp_opt;
    int interp;
    double angle_deg, shift_x, shift_y;
} RotateCudaContext;

#define OFFSET(x) offsetof(RotateCudaContext, x)
static const AVOption rotate_cuda_options[] = {
    { "angle",   "Rotation angle in degree", OFFSET(angle_deg), AV_OPT_TYPE_DOUBLE, { .dbl = 0.0 }, -360, 360, .flags = GMATB_FLAGS },
    { "interp",  "Interpolation algorithm (linear, nearest, cubic, area)", OFFSET(interp_opt), AV_OPT_TYPE_STRING, { .str = "linear" }, 0, 0, .flags = GMATB_FLAGS },
    { "shift_x", "Shift in x directions to move the center at the same coord after rotation", OFFSET(shift_x), AV_OPT_TYPE_DOUBLE, { .dbl = 0.0 }, -1e9, 1e9, .flags = GMATB_FLAGS },
    { "shift_y", "Shift in y directions to move the center at the same coord after rotation", OFFSET(shift_y), AV_OPT_TYPE_DOUBLE, { .dbl = 0.0 }, -1e9, 1e9, .flags = GMATB_FLAGS },
    { NULL }
};
AVFILTER_DEFINE_CLASS(rotate_cuda);

static av_cold int rotate_init(AVFilterContext *ctx)
{
    RotateCudaContext *s = ctx->priv;
    if (!strcmp(s->interp_opt, "linear"))       s->interp = GMATB_INTERP_LINEAR;
    else if (!strcmp(s->interp_opt, "nearest")) s->interp = GMATB_INTERP_NEAREST;
    else if (!strcmp(s->interp_opt, "cubic"))   s->interp = GMATB_INTERP_CUBIC;
    else if (!strcmp(s->interp_opt, "area"))    s->interp = GMATB_INTERP_AREA;
    else {
        av_log(ctx, AV_LOG_ERROR, "Interpolation '%s' not supported.\n", s->interp_opt);
        return AVERROR(EINVAL);
    }
    return 0;
}

static int rotate_config_props(AVFilterLink *outlink)
{
    return gmatb_config_output(outlink, outlink->src->priv, 0, 0);
}

static int rotate_launch(AVFilterContext *ctx, const GmatbImage *src, const GmatbImage *dst, void *stream)
{
    RotateCudaContext *s = ctx->priv;
    return gmatb_rotate(src, dst, s->angle_deg, s->shift_x, s->shift_y, s->interp, stream);
}

static int rotate_filter_frame(AVFilterLink *inlink, AVFrame *in)
{
    return gmatb_filter_frame(inlink, in, rotate_launch);
}

static const AVFilterPad rotate_cuda_inputs[] = {
    { .name = "default", .type = AVMEDIA_TYPE_VIDEO, .filter_frame = rotate_filter_frame },
};
static const AVFilterPad rotate_cuda_outputs[] = {
    { .name = "default", .type = AVMEDIA_TYPE_VIDEO, .config_props = rotate_config_props },
};

const AVFilter ff_vf_rotate_cuda = {
    .name           = "rotate_cuda",
    .description    = NULL_IF_CONFIG_SMALL("Rotate CUDA frames (gmat_b200 kernels)"),
    FILTER_INPUTS(rotate_cuda_inputs),
    FILTER_OUTPUTS(rotate_cuda_outputs),
    .priv_class     = &rotate_cuda_class,
    .priv_size      = sizeof(RotateCudaContext),
    .init           = rotate_init,
    .uninit         = gmatb_uninit,
    FILTER_QUERY_FUNC(gmatb_query_formats),
    .flags_internal = FF_FILTER_FLAG_HWFRAME_AWARE,
};
