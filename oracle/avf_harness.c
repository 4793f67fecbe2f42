/* oracle/avf_harness.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Runs a libavfilter graph description on CUDA frames inside the reference's OWN libavfilter / libavutil
 * (avfilter.c, avfiltergraph.c, buffersrc.c, buffersink.c, formats.c, ..., hwcontext.c, hwcontext_cuda.c compiled
 * from /root/reference by oracle/refbuild `avf`), with our filter objects (gmat_b200/csrc/avfilter/vf_*_cuda.c)
 * registered through the generated filter_list.c exactly as INTEGRATION.md section 2 describes:
 *     buffer (hw_frames_ctx = a real CUDA frame pool) -> <filters> -> buffersink
 * Frames are uploaded / downloaded with av_hwframe_transfer_data, i.e. they live in the reference's frame pool
 * (hwcontext_cuda.c:96-205: one allocation per frame, pitch aligned to the texture alignment).
 * tests/test_gpu_avfilter.py drives this through ctypes and compares with the kernel layer called directly. */
#include <stdio.h>
#include <string.h>

#include <cuda.h>

#include "libavfilter/avfilter.h"
#include "libavfilter/buffersink.h"
#include "libavfilter/buffersrc.h"
#include "libavutil/frame.h"
#include "libavutil/hwcontext.h"
#include "libavutil/hwcontext_cuda.h"
#include "libavutil/imgutils.h"
#include "libavutil/mem.h"
#include "libavutil/opt.h"
#include "libavutil/pixdesc.h"

typedef struct AvfResult {
    int out_w, out_h, out_fmt;          /* sw format of the output pool */
    int in_pitch, out_pitch;            /* linesize[0] of the pool frames the filters saw / produced */
    int frames_out;
    long long out_bytes_per_frame;
    char error[256];
} AvfResult;

static int fail(AvfResult *r, const char *what, int err)
{
    char b[128];
    av_strerror(err, b, sizeof(b));
    snprintf(r->error, sizeof(r->error), "%s: %s (%d)", what, b, err);
    return err < 0 ? err : -1;
}

const char *avf_list_filters(void)
{
    static char buf[4096];
    void *it = NULL;
    const AVFilter *f;
    buf[0] = 0;
    while ((f = av_filter_iterate(&it))) { strncat(buf, f->name, sizeof(buf) - strlen(buf) - 2); strcat(buf, " "); }
    return buf;
}

/* in: n frames of sw_fmt, w x h, each plane tightly packed (av_image_copy_to_buffer layout, align 1).
 * out: room for out_cap bytes, frames written in the same layout. */
int avf_run(const char *sw_fmt_name, int w, int h, const char *filters, const uint8_t *in, int n,
            uint8_t *out, long long out_cap, AvfResult *r)
{
    AVBufferRef *dev = NULL, *frames_ref = NULL;
    AVFilterGraph *graph = NULL;
    AVFilterContext *src = NULL, *sink = NULL;
    AVFilterInOut *inputs = NULL, *outputs = NULL;
    AVBufferSrcParameters *par = NULL;
    AVFrame *sw = NULL, *hw = NULL, *got = NULL, *dl = NULL;
    AVHWFramesContext *fc;
    const enum AVPixelFormat sw_fmt = av_get_pix_fmt(sw_fmt_name);
    const int in_frame_bytes = av_image_get_buffer_size(sw_fmt, w, h, 1);
    int ret, i;

    memset(r, 0, sizeof(*r));
    if (sw_fmt == AV_PIX_FMT_NONE) return fail(r, "unknown pixel format", AVERROR(EINVAL));
    /* the application (the test process: torch) already owns the device's primary context: hand it to libavutil the
     * documented way for user-supplied contexts -- av_hwdevice_ctx_alloc, fill AVCUDADeviceContext, av_hwdevice_ctx_init
     * (hwcontext_cuda.c cuda_device_init) -- instead of av_hwdevice_ctx_create, which insists on its own flags */
    {
        CUdevice cudev;
        AVCUDADeviceContext *hc;
        if (!(dev = av_hwdevice_ctx_alloc(AV_HWDEVICE_TYPE_CUDA))) { ret = fail(r, "av_hwdevice_ctx_alloc", AVERROR(ENOMEM)); goto end; }
        hc = ((AVHWDeviceContext *)dev->data)->hwctx;
        if (cuInit(0) != CUDA_SUCCESS || cuDeviceGet(&cudev, 0) != CUDA_SUCCESS || cuDevicePrimaryCtxRetain(&hc->cuda_ctx, cudev) != CUDA_SUCCESS) {
            ret = fail(r, "cuDevicePrimaryCtxRetain", AVERROR_EXTERNAL); goto end;
        }
        hc->stream = NULL;
        if ((ret = av_hwdevice_ctx_init(dev)) < 0) { ret = fail(r, "av_hwdevice_ctx_init", ret); goto end; }
    }
    if (!(frames_ref = av_hwframe_ctx_alloc(dev))) { ret = fail(r, "av_hwframe_ctx_alloc", AVERROR(ENOMEM)); goto end; }
    fc = (AVHWFramesContext *)frames_ref->data;
    fc->format = AV_PIX_FMT_CUDA; fc->sw_format = sw_fmt; fc->width = w; fc->height = h;
    if ((ret = av_hwframe_ctx_init(frames_ref)) < 0) { ret = fail(r, "av_hwframe_ctx_init (is the sw_format in the CUDA pool's list?)", ret); goto end; }

    graph = avfilter_graph_alloc();
    src = avfilter_graph_alloc_filter(graph, avfilter_get_by_name("buffer"), "in");
    sink = avfilter_graph_alloc_filter(graph, avfilter_get_by_name("buffersink"), "out");
    if (!graph || !src || !sink) { ret = fail(r, "graph / buffer / buffersink allocation", AVERROR(ENOMEM)); goto end; }
    par = av_buffersrc_parameters_alloc();
    par->format = AV_PIX_FMT_CUDA; par->width = w; par->height = h;
    par->time_base = (AVRational){1, 25}; par->hw_frames_ctx = frames_ref;
    if ((ret = av_buffersrc_parameters_set(src, par)) < 0) { ret = fail(r, "av_buffersrc_parameters_set", ret); goto end; }
    if ((ret = avfilter_init_str(src, NULL)) < 0) { ret = fail(r, "init buffer", ret); goto end; }
    if ((ret = avfilter_init_str(sink, NULL)) < 0) { ret = fail(r, "init buffersink", ret); goto end; }
    outputs = avfilter_inout_alloc(); inputs = avfilter_inout_alloc();
    outputs->name = av_strdup("in");  outputs->filter_ctx = src;  outputs->pad_idx = 0; outputs->next = NULL;
    inputs->name  = av_strdup("out"); inputs->filter_ctx  = sink; inputs->pad_idx  = 0; inputs->next  = NULL;
    if ((ret = avfilter_graph_parse_ptr(graph, filters, &inputs, &outputs, NULL)) < 0) { ret = fail(r, "avfilter_graph_parse_ptr", ret); goto end; }
    if ((ret = avfilter_graph_config(graph, NULL)) < 0) { ret = fail(r, "avfilter_graph_config", ret); goto end; }

    sw = av_frame_alloc(); dl = av_frame_alloc(); got = av_frame_alloc();
    for (i = 0; i < n; i++) {
        hw = av_frame_alloc();
        sw->format = sw_fmt; sw->width = w; sw->height = h;
        if ((ret = av_image_fill_arrays(sw->data, sw->linesize, in + (size_t)i * in_frame_bytes, sw_fmt, w, h, 1)) < 0) { ret = fail(r, "av_image_fill_arrays", ret); goto end; }
        if ((ret = av_hwframe_get_buffer(frames_ref, hw, 0)) < 0) { ret = fail(r, "av_hwframe_get_buffer", ret); goto end; }
        if ((ret = av_hwframe_transfer_data(hw, sw, 0)) < 0) { ret = fail(r, "upload (av_hwframe_transfer_data)", ret); goto end; }
        hw->pts = i;
        r->in_pitch = hw->linesize[0];
        if ((ret = av_buffersrc_add_frame(src, hw)) < 0) { ret = fail(r, "av_buffersrc_add_frame", ret); goto end; }
        av_frame_free(&hw);
        while ((ret = av_buffersink_get_frame(sink, got)) >= 0) {
            AVHWFramesContext *oc = (AVHWFramesContext *)got->hw_frames_ctx->data;
            int sz;
            r->out_w = got->width; r->out_h = got->height; r->out_fmt = oc->sw_format; r->out_pitch = got->linesize[0];
            dl->format = oc->sw_format;
            if ((ret = av_hwframe_transfer_data(dl, got, 0)) < 0) { ret = fail(r, "download (av_hwframe_transfer_data)", ret); goto end; }
            sz = av_image_get_buffer_size(dl->format, dl->width, dl->height, 1);
            r->out_bytes_per_frame = sz;
            if ((long long)(r->frames_out + 1) * sz > out_cap) { ret = fail(r, "output buffer too small", AVERROR(ENOSPC)); goto end; }
            av_image_copy_to_buffer(out + (size_t)r->frames_out * sz, sz, (const uint8_t *const *)dl->data, dl->linesize, dl->format, dl->width, dl->height, 1);
            r->frames_out++;
            av_frame_unref(dl); av_frame_unref(got);
        }
        if (ret != AVERROR(EAGAIN) && ret != AVERROR_EOF) { ret = fail(r, "av_buffersink_get_frame", ret); goto end; }
    }
    ret = 0;
end:
    av_frame_free(&hw); av_frame_free(&got); av_frame_free(&dl);
    if (sw) { memset(sw->data, 0, sizeof(sw->data)); av_frame_free(&sw); }
    av_freep(&par);
    avfilter_inout_free(&inputs); avfilter_inout_free(&outputs);
    avfilter_graph_free(&graph);
    av_buffer_unref(&frames_ref);
    av_buffer_unref(&dev);
    return ret;
}
