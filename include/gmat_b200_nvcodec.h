/* gmat_b200_nvcodec.h -- SURVEY 8f N3: the scaling / colour-conversion helpers of GMAT's metrans toolkit
 * (metrans/include/NvCodec/Resize.cu, ColorSpace.cu; declared in NvCodec/NvCommon.h:232-255), the step right after
 * NVDEC in its transcoding apps, on the gmat_b200 kernels.  Same argument meaning as the reference functions:
 * device pointers, byte pitches, NV12 / P016 as ONE buffer with the interleaved chroma at base + height * pitch,
 * planar outputs as three stacked planes at base + i * pitch * height, iMatrix = ColorSpaceStandard (1 BT.709 --
 * also the default for unknown values --, 4 FCC, 5 BT.470, 6 BT.601, 7 SMPTE 240M, 9 / 10 BT.2020).  Only whole 2x2
 * blocks are converted (the reference's kernels skip a trailing odd row / column).  Results are bit-identical to
 * the reference's kernels compiled for sm_100a (tests/test_gpu_nvcodec.py, oracle O4).
 *
 * libgmat_b200.so exports the C ABI below; libgmat_b200_nvcodec.so adds the reference's own C++ signatures
 * (Nv12ToBgra32(uint8_t*, int, uint8_t*, int, int, int, int, cudaStream_t) ...) so that metrans objects link
 * against it unchanged.
 *
 * NOT provided: ScaleNv12 / ScaleP016 (Resize.cu:15-81).  They sample through the texture unit's bilinear filter
 * (cudaFilterModeLinear: 9-bit fixed-point weights in hardware), which plain arithmetic cannot reproduce bit for
 * bit and which has no place in a kernel that reads each byte once; use gmatb_sws (SWS_BILINEAR) for that job. */
#ifndef GMAT_B200_NVCODEC_H
#define GMAT_B200_NVCODEC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* which conversion of ColorSpace.cu (:219-348) */
enum {
    GMATB_NVC_NV12_TO_BGRA32 = 0,      /* Nv12ToBgra32          :219 */
    GMATB_NVC_NV12_TO_RGBA32,          /* Nv12ToRgba32          :226 */
    GMATB_NVC_NV12_TO_BGRA64,          /* Nv12ToBgra64          :232 */
    GMATB_NVC_P016_TO_BGRA32,          /* P016ToBgra32          :239 */
    GMATB_NVC_P016_TO_BGRA64,          /* P016ToBgra64          :246 */
    GMATB_NVC_NV12_TO_BGR_PLANAR,      /* Nv12ToBgrPlanar       :253 */
    GMATB_NVC_NV12_TO_RGB_PLANAR,      /* Nv12ToRgbPlanar       :260 */
    GMATB_NVC_P016_TO_BGR_PLANAR,      /* P016ToBgrPlanar       :266 */
    GMATB_NVC_NV12_TO_BGR_FLOAT_PLANAR,/* Nv12ToBgrFloatPlanar  :273 */
    GMATB_NVC_NV12_TO_RGB_FLOAT_PLANAR,/* Nv12ToRgbFloatPlanar  :280 */
    GMATB_NVC_P016_TO_BGR_FLOAT_PLANAR,/* P016ToBgrFloatPlanar  :287 */
    GMATB_NVC_BGRA64_TO_P016,          /* Bgra64ToP016          :343 */
    GMATB_NVC_COUNT
};

/* returns GMATB_OK (0) or a GMATB_ERR_* code (gmat_b200.h) */
int gmatb_nvcodec_convert(int kind, const uint8_t *src, int src_pitch, uint8_t *dst, int dst_pitch,
                          int width, int height, int iMatrix, void *stream);

/* ScaleNv12_Bicubic (Resize.cu:83-160): Catmull-Rom (a = -0.5) 4x4 on NV12, source position x * (srcW/dstW)
 * clamped to [2, srcW - 2] (no half-pixel offset), truncating store */
int gmatb_nvcodec_scale_nv12_bicubic(const uint8_t *src_nv12, int src_pitch, int src_width, int src_height,
                                     uint8_t *dst_nv12, int dst_pitch, int dst_width, int dst_height, void *stream);

#ifdef __cplusplus
}
#endif
#endif
