/*
 * gmat_b200.h -- C ABI of the B200-native pixel-transform kernel layer.
 *
 * Plain C: pointers, ints, doubles; no CUDA / torch / FFmpeg types.  `stream`
 * arguments are CUstream / cudaStream_t handles passed as `void *` (NULL = the
 * legacy default stream).  All image pointers are DEVICE pointers unless a
 * function name ends in `_host`.  Every function enqueues work on the given
 * stream and returns without synchronising (same contract as the reference,
 * libswscale/cuda/swscale_cuda.c never synchronises), returning 0 on
 * success or a negative GMATB_ERR_* code.  There is no CPU fallback anywhere: a
 * missing device or failed launch is an error.
 *
 * What each group replaces in the reference (NVIDIA/GMAT, ffmpeg-gpu/):
 *   gmatb_yuv2rgb / rgb2yuv / yuv2yuv / rgb24tobgr24
 *        -> libswscale/cuda/yuv2rgb_cuda.cu:862-947, yuv2yuv_cuda.cu:320-371,
 *           rgb2rgb_cuda_kernel.cu:37-42 (the unscaled converters)
 *   gmatb_sws_*    -> libswscale/cuda/swscale_cuda.c:76-479 (init / scale / free;
 *                     CSC + CV-CUDA resize), driven by SwsContext fields
 *                     (swscale_internal.h:682-695)
 *   gmatb_crop     -> libavfilter/vf_crop_nvcv.c:277   (cvcudaCustomCropSubmit)
 *   gmatb_flip     -> libavfilter/vf_flip_nvcv.c:251   (cvcudaFlipSubmit)
 *   gmatb_rotate   -> libavfilter/vf_rotate_nvcv.c:275 (cvcudaRotateSubmit)
 *   gmatb_gaussian -> libavfilter/vf_smooth_nvcv.c:290 (cvcudaGaussianSubmit)
 *   gmatb_median   -> libavfilter/vf_smooth_nvcv.c:294 (cvcudaMedianBlurSubmit)
 *   gmatb_format_* -> libavfilter/format_cuda_kernel.cu:583-632 (format_cuda filter)
 * The nine libswscale-internal symbols built on top of this layer are declared
 * in gmat_b200_sws.h.
 */
#ifndef GMAT_B200_H
#define GMAT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GMATB_VERSION 0x000100

/* ---- error codes ------------------------------------------------------- */
#define GMATB_OK              0
#define GMATB_ERR_INVAL     (-22)   /* bad argument (AVERROR(EINVAL) value)   */
#define GMATB_ERR_NOMEM     (-12)   /* device allocation failed               */
#define GMATB_ERR_UNSUPPORTED (-38) /* format pair / mode not implemented     */
#define GMATB_ERR_CUDA      (-5)    /* CUDA runtime error (see gmatb_last_cuda_error) */

/* ---- pixel formats: numerically equal to enum AVPixelFormat of the
 *      reference tree (libavutil/pixfmt.h) so ints pass through unchanged --- */
#define GMATB_FMT_YUV420P      0
#define GMATB_FMT_RGB24        2
#define GMATB_FMT_BGR24        3
#define GMATB_FMT_NV12        23
#define GMATB_FMT_RGBA        26
#define GMATB_FMT_BGRA        28
#define GMATB_FMT_RGB48LE     35
#define GMATB_FMT_YUV420P16LE 45
#define GMATB_FMT_BGR48LE     58
#define GMATB_FMT_YUV420P10LE 62
#define GMATB_FMT_RGBA64LE   105
#define GMATB_FMT_BGRA64LE   107
#define GMATB_FMT_0RGB       118
#define GMATB_FMT_RGB0       119
#define GMATB_FMT_0BGR       120
#define GMATB_FMT_BGR0       121
#define GMATB_FMT_P010LE     159
#define GMATB_FMT_P016LE     170
#define GMATB_FMT_RGBPF32LE  179   /* GMAT addition, pixfmt.h:315 */
#define GMATB_FMT_RGBAPF32LE 180   /* GMAT addition, pixfmt.h:316 */

/* ---- colour spaces: enum AVColorSpace values (pixfmt.h) ------------------ */
#define GMATB_SPC_DEFAULT      0    /* what every reference context gets: BT.601 limited */
#define GMATB_SPC_BT709        1
#define GMATB_SPC_FCC          4
#define GMATB_SPC_BT470BG      5
#define GMATB_SPC_SMPTE170M    6
#define GMATB_SPC_SMPTE240M    7
#define GMATB_SPC_BT2020_NCL   9
#define GMATB_SPC_BT2020_CL   10

/* ---- libswscale flag bits (libswscale/swscale.h:60-95) ------------------ */
#define GMATB_SWS_FAST_BILINEAR 0x1
#define GMATB_SWS_BILINEAR      0x2
#define GMATB_SWS_BICUBIC       0x4
#define GMATB_SWS_POINT         0x10
#define GMATB_SWS_AREA          0x20
#define GMATB_SWS_LANCZOS       0x200
#define GMATB_SWS_HWACCEL_CUDA  0x1000000
#define GMATB_SWS_PARAM_DEFAULT 123456.0   /* SWS_PARAM_DEFAULT */
/* gmat_b200 extension bit (unused by libswscale): reproduce the reference
 * scale_cuda kernels' missing upper clamp (values >= 256 wrap modulo 256,
 * vf_scale_cuda.cu:1057-1071 + cvt.rzi) instead of saturating. */
#define GMATB_SWS_PARITY_WRAP   0x40000000
/* gmat_b200 extension bit: run the exact-integer form of the fused 2:1 kernel (scale_fused4i.cuh) where it
 * applies (8-bit yuv sources, dyadic bicubic weights: param0 = 0.75, 0.5, 1.0).  Same output bytes as the default
 * float-chain kernel (tests/test_gpu_scale_int.py); measured slower on B200, hence opt-in. */
#define GMATB_SWS_INT_CHAIN     0x20000000
/* gmat_b200 extension bit: the same exact-integer chain with its horizontal pass on the tensor pipe (IMMA u8 x s8,
 * scale_fused5m.cuh); NV12 sources, 8-bit packed rgb destinations, same dyadic weights, same output bytes. */
#define GMATB_SWS_MMA_CHAIN     0x10000000
/* gmat_b200 extension bit: keep yuv -> rgb scaling at ratios other than 2:1 on the shared-memory tile kernel
 * (scale_generic.cuh) instead of the streaming kernel (scale_stream.cuh); same output bytes (A/B measurements, tests). */
#define GMATB_SWS_TILE_KERNEL   0x08000000

/* ---- interpolation / border codes of the filter layer (NVCV numbering:
 *      NVCV_INTERP_* / NVCV_BORDER_* as used by vf_rotate_nvcv.c:115-135 and
 *      vf_smooth_nvcv.c:96-101) ---------------------------------------------- */
#define GMATB_INTERP_NEAREST 0
#define GMATB_INTERP_LINEAR  1
#define GMATB_INTERP_CUBIC   2
#define GMATB_INTERP_AREA    3

#define GMATB_MEDIAN_MAXK 15   /* gmatb_median: odd kw, kh <= this */
#define GMATB_GAUSS_MAXK  31   /* gmatb_gaussian: odd kw, kh <= this */

#define GMATB_BORDER_CONSTANT   0
#define GMATB_BORDER_REPLICATE  1
#define GMATB_BORDER_REFLECT    2
#define GMATB_BORDER_WRAP       3
#define GMATB_BORDER_REFLECT101 4

/* ---- image descriptor ---------------------------------------------------- */
/* One frame (or a uniform batch of frames) in device memory, FFmpeg style:
 * up to 4 plane pointers + byte strides.  `batch` frames (0 or 1 = single) are
 * laid out at data[p] + i*batch_stride[p]; kernels put the frame index on
 * blockIdx.z so a whole batch is ONE launch (frames shard by batch index). */
typedef struct GmatbImage {
    void     *data[4];
    int       linesize[4];
    int       width, height;
    int       format;           /* GMATB_FMT_*                                  */
    int       batch;
    long long batch_stride[4];  /* bytes between consecutive frames, per plane  */
} GmatbImage;

/* ---- library / device ----------------------------------------------------- */
int         gmatb_version(void);
int         gmatb_device_count(void);           /* <0 on CUDA error              */
int         gmatb_last_cuda_error(void);        /* cudaError_t of the last failure */
const char *gmatb_last_cuda_error_string(void);
long long   gmatb_launch_count(void);           /* kernels launched by this library so far */
int         gmatb_device_sync(void);
/* Sink for the library's diagnostics (e.g. an interpolation mode that falls back to another one); NULL = stderr.
 * The FFmpeg-side glue passes a function that forwards to av_log. */
void        gmatb_set_log(void (*cb)(const char *msg));

/* 3x3 matrices exactly as the reference computes them (yuv2rgb_cuda.cu:782-848):
 * float arithmetic for the entries, double for the range scale, cast to float. */
void gmatb_csc_matrix_yuv2rgb(int colorspace, float out9[9]);
void gmatb_csc_matrix_rgb2yuv(int colorspace, float out9[9]);

/* ---- unscaled converters (colorspace = GMATB_SPC_*) ------------------------ */
/* src.width/height is the luma geometry; dst must have the same geometry. */
int gmatb_yuv2rgb(const GmatbImage *src, const GmatbImage *dst, int colorspace, void *stream);
int gmatb_rgb2yuv(const GmatbImage *src, const GmatbImage *dst, int colorspace, void *stream);
int gmatb_yuv2yuv(const GmatbImage *src, const GmatbImage *dst, void *stream);
int gmatb_rgb24tobgr24(const GmatbImage *src, const GmatbImage *dst, void *stream);
/* NV12/YUV420P -> planar float RGB with (x - shift[c]) / norm
 * (yuv2rgb_cuda.cu:381-389; the reference launches it with norm=255, shift=0) */
int gmatb_yuv2rgb_planar_f32(const GmatbImage *src, const GmatbImage *dst, int colorspace,
                             float norm, const float shift_rgb[3], void *stream);

/* ---- format_cuda filter kernels -------------------------------
 * Replace libavfilter/format_cuda_kernel.cu:583-632 as called by vf_format_cuda.c:185-203.
 * `av_colorspace` is the frame's enum AVColorSpace; the filter's own matrix selection
 * (GetConstants, format_cuda_kernel.cu:32-63: BT.709 is the DEFAULT branch, BT470BG = BT.601,
 * SMPTE170M falls into the default) is applied by gmatb_format_colorspace.
 *   nv12_to_rgbpf32: planes 0,1,2 = R,G,B (or B,G,R when bgr_planes), float (c - shift[c]) / norm
 *                    with c the truncated 8-bit result; nv12_to_rgbpf32 = norm 255, shift NULL;
 *                    nv12_to_rgbpf32_shift / nv12_to_bgrpf32_shift = the general form.
 *   rgbpf32_to_nv12: planar float in [0,1] x 255 -> NV12, 2x2 float mean chroma; width % 4 == 0,
 *                    height even, planes 16-byte aligned. */
int gmatb_format_colorspace(int av_colorspace);
int gmatb_format_nv12_to_rgbpf32(const GmatbImage *src, const GmatbImage *dst, int av_colorspace,
                                 float norm, const float shift_rgb[3], int bgr_planes, void *stream);
int gmatb_format_rgbpf32_to_nv12(const GmatbImage *src, const GmatbImage *dst, int av_colorspace, void *stream);

/* ---- scaling context (mirror of SwsContext's CUDA path) -------------------- */
typedef struct GmatbSws GmatbSws;

/* Same argument meaning as sws_getContext (libswscale/utils.c:2087): flags carries
 * the SWS_* algorithm bit, param[0..1] the algorithm parameters (NULL or
 * SWS_PARAM_DEFAULT = default).  Returns NULL on unsupported formats / sizes /
 * allocation failure (sws_getContext returns NULL when ff_sws_init_swscale_cuda
 * fails, utils.c:2102-2105). */
GmatbSws *gmatb_sws_create(int srcW, int srcH, int srcFormat,
                           int dstW, int dstH, int dstFormat,
                           int flags, const double *param, int colorspace);
void      gmatb_sws_free(GmatbSws *c);
void      gmatb_sws_set_stream(GmatbSws *c, void *stream);     /* sws_setCudaStream, swscale.c:1249 */
/* One frame; FFmpeg-style pointer/stride arrays (device pointers). */
int       gmatb_sws_scale(GmatbSws *c, const uint8_t *const src[4], const int srcStride[4],
                          uint8_t *const dst[4], const int dstStride[4]);
/* A uniform batch in one launch sequence. */
int       gmatb_sws_scale_batch(GmatbSws *c, const GmatbImage *src, const GmatbImage *dst);
/* End-to-end convenience: HOST frames in, HOST frames out.  Copies `n` source
 * frames host->device, converts, copies results device->host, on the context's
 * stream, then synchronises the stream.  Buffers are tightly described by the
 * GmatbImage (host pointers); pinned memory makes the copies asynchronous. */
int       gmatb_sws_scale_host(GmatbSws *c, const GmatbImage *src_host, const GmatbImage *dst_host);
/* Introspection used by the parity tests: copy the per-column / per-row
 * resample tables (4 taps each; device-computed) to host arrays.
 * axis 0 = horizontal (dstW entries), 1 = vertical (dstH entries).
 * coeffs: 4 floats per entry, pos: first source index (p-1) per entry. */
int       gmatb_sws_get_filter(GmatbSws *c, int axis, float *coeffs, int *pos);
/* which kernel family the context selected: 0 unscaled, 1 fused 2:1, 2 generic */
int       gmatb_sws_path(const GmatbSws *c);

/* ---- filter kernels ---------------------------------------------------------- */
/* Packed 8-bit images with 3 or 4 bytes per pixel (rgb24/bgr24/rgba/bgra/0rgb...),
 * which is what the reference filters accept (vf_rotate_nvcv.c:92-101). */
int gmatb_crop(const GmatbImage *src, const GmatbImage *dst, int x, int y, void *stream);
int gmatb_flip(const GmatbImage *src, const GmatbImage *dst, int flip_code, void *stream);
int gmatb_rotate(const GmatbImage *src, const GmatbImage *dst, double angle_deg,
                 double shift_x, double shift_y, int interp, void *stream);
int gmatb_gaussian(const GmatbImage *src, const GmatbImage *dst, int kw, int kh,
                   double sigma_x, double sigma_y, int border, void *stream);
int gmatb_median(const GmatbImage *src, const GmatbImage *dst, int kw, int kh, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GMAT_B200_H */
