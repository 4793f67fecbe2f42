/*
 * gmat_b200_sws.h -- the libswscale drop-in boundary.
 *
 * The reference's libswscale.so, built without its libswscale/cuda/ objects, has
 * exactly these nine undefined symbols (SURVEY 8b "B1"; nm -D --undefined-only).
 * libgmat_b200.so exports the five that take plain pointers; libgmat_b200_sws.so
 * (gmat_b200/csrc/sws_shim.c, compiled against the reference's own
 * libswscale/swscale_internal.h because it reads SwsContext fields) exports the four
 * that take a SwsContext.  Linking ffmpeg-gpu's libswscale against them replaces
 * libswscale/cuda/*.o + CV-CUDA with no C source change.
 *
 * All pointers are device pointers, strides are bytes, formats are enum
 * AVPixelFormat values, `stream` is the CUstream stored by sws_setCudaStream
 * (libswscale/swscale.c:1249).  Work is enqueued and never synchronised.
 */
#ifndef GMAT_B200_SWS_H
#define GMAT_B200_SWS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

struct SwsContext;

/* ---- exported by libgmat_b200.so ------------------------------------------- */

/* replaces yuv2rgb_cuda.cu:862-907; called by yuvToRgbWrapperCuda
 * (swscale_unscaled.c:1992-1997) with the full-frame geometry.  0 / -1. */
int yuv2rgb_cuda(const uint8_t *src[], int srcStride[], uint8_t *dst[], int dstStride[],
                 int width, int height, int srcFormat, int dstFormat, void *stream);
/* replaces yuv2rgb_cuda.cu:909-947; RgbToYuvWrapperCuda (swscale_unscaled.c:1999-2004) */
int rgb2yuv_cuda(const uint8_t *src[], int srcStride[], uint8_t *dst[], int dstStride[],
                 int width, int height, int srcFormat, int dstFormat, void *stream);
/* replaces yuv2yuv_cuda.cu:320-371; YuvToYuvWrapperCuda (swscale_unscaled.c:2006-2011).
 * NB the reference returns -1 unconditionally (:370); we return 0 on success. */
int yuv2yuv_cuda(const uint8_t *src[], int srcStride[], uint8_t *dst[], int dstStride[],
                 int width, int height, int srcFormat, int dstFormat, void *stream);
/* replaces rgb2rgb_cuda_kernel.cu:37-42; rgbToRgbWrapperCuda (swscale_unscaled.c:1972-1977) */
void rgb24tobgr24_cuda(const uint8_t *src[], uint8_t *dst[], int srcStride[], int dstStride[],
                       int width, int height, void *stream);
/* replaces rgb2rgb_cuda_kernel.cu:44-47; ff_sws_rgb2rgb_init_hw (rgb2rgb.c:147-150) */
void rgb2rgb_init_cuda(void);

/* The reference keeps its CSC matrices in process-global __constant__ memory written
 * by set_mat_*_cuda (yuv2rgb_cuda.cu:816-848).  The unscaled entry points above carry
 * no colourspace argument, so the process-global selection survives here as a host
 * variable (default 0 = BT.601 limited, what every reference context gets); the
 * matrices themselves travel as kernel arguments. */
void gmatb_set_process_colorspace(int av_color_space);
int  gmatb_get_process_colorspace(void);

/* ---- exported by libgmat_b200_sws.so ----------------------------------------- */

/* swscale_cuda.c:112-271; from sws_init_context_cuda (utils.c:2057) when sizes differ.
 * <0 makes sws_getContext return NULL. */
int  ff_sws_init_swscale_cuda(struct SwsContext *c);
/* swscale_cuda.c:273-479; from scale_internal (swscale.c:1042-1045).  Returns 0 on
 * success like the reference (its value becomes sws_scale()'s), negative on error. */
int  ff_swscale_cuda(struct SwsContext *c, const uint8_t *src[], int srcStride[], int srcSliceY, int srcSliceH,
                     uint8_t *dst[], int dstStride[], int dstSliceY, int dstSliceH);
/* swscale_cuda.c:86-110; from sws_freeContext_cuda (utils.c:2507-2510); does not free c */
int  ff_sws_free_swscale_cuda(struct SwsContext *c);
/* swscale_cuda.c:76-84; from ff_get_unscaled_swscale_cuda (swscale_unscaled.c:2053) */
void ff_yuv2rgb_init_tables_cuda(struct SwsContext *c);

#ifdef __cplusplus
}
#endif
#endif
