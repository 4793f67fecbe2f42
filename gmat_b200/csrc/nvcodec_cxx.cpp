// nvcodec_cxx.cpp -> libgmat_b200_nvcodec.so: the reference's own C++ signatures (metrans/include/NvCodec/
// NvCommon.h:232-255) over the C ABI of include/gmat_b200_nvcodec.h, so that metrans objects (NvDecoderImageProvider.h,
// the AppNvTrans tests) link against gmat_b200 unchanged.  The reference's functions return void and take their
// launch errors to cudaGetLastError; ours report through gmatb_last_cuda_error as well.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/gmat_b200_nvcodec.h"

#define CONV(Name, Kind, DstT) \
    void Name(uint8_t *dpSrc, int nSrcPitch, DstT *dpDst, int nDstPitch, int nWidth, int nHeight, int iMatrix, cudaStream_t stream) { \
        gmatb_nvcodec_convert(Kind, dpSrc, nSrcPitch, (uint8_t *)dpDst, nDstPitch, nWidth, nHeight, iMatrix, (void *)stream); }

CONV(Nv12ToBgra32, GMATB_NVC_NV12_TO_BGRA32, uint8_t)
CONV(Nv12ToRgba32, GMATB_NVC_NV12_TO_RGBA32, uint8_t)
CONV(Nv12ToBgra64, GMATB_NVC_NV12_TO_BGRA64, uint8_t)
CONV(P016ToBgra32, GMATB_NVC_P016_TO_BGRA32, uint8_t)
CONV(P016ToBgra64, GMATB_NVC_P016_TO_BGRA64, uint8_t)
CONV(Nv12ToBgrPlanar, GMATB_NVC_NV12_TO_BGR_PLANAR, uint8_t)
CONV(Nv12ToRgbPlanar, GMATB_NVC_NV12_TO_RGB_PLANAR, uint8_t)
CONV(P016ToBgrPlanar, GMATB_NVC_P016_TO_BGR_PLANAR, uint8_t)
CONV(Nv12ToBgrFloatPlanar, GMATB_NVC_NV12_TO_BGR_FLOAT_PLANAR, float)
CONV(Nv12ToRgbFloatPlanar, GMATB_NVC_NV12_TO_RGB_FLOAT_PLANAR, float)
CONV(P016ToBgrFloatPlanar, GMATB_NVC_P016_TO_BGR_FLOAT_PLANAR, float)
CONV(Bgra64ToP016, GMATB_NVC_BGRA64_TO_P016, uint8_t)

void ScaleNv12_Bicubic(unsigned char *dpSrcNv12, int nSrcPitch, int nSrcWidth, int nSrcHeight, unsigned char *dpDstNv12, int nDstPitch,
                       int nDstWidth, int nDstHeight) {
    gmatb_nvcodec_scale_nv12_bicubic(dpSrcNv12, nSrcPitch, nSrcWidth, nSrcHeight, dpDstNv12, nDstPitch, nDstWidth, nDstHeight, nullptr);
}
