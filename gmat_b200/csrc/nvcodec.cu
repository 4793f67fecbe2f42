// nvcodec.cu -- SURVEY 8f N3: metrans' NvCodec helpers (Resize.cu / ColorSpace.cu) on the gmat_b200 kernels.
// C ABI in include/gmat_b200_nvcodec.h.  The colour conversions are the libgpuscale kernels of csc.cu in the FMA form
// (every kernel of ColorSpace.cu compiles to FFMA(fv, mC, FFMA(fy, mA, FMUL(fu, mB))): SASS of oracle/_ref/
// libref_nvcodec.so) with BT.709 as the default matrix (ColorSpace.cu:36-42); the bicubic scaler is its own kernel.
#include <cstring>
#include "common.cuh"
#include "../../include/gmat_b200_nvcodec.h"

namespace gmatb {
int yuv2rgb_launch(const GmatbImage *, const GmatbImage *, const Mat9 &, cudaStream_t);
int yuv2rgb_nv12_fma_launch(const GmatbImage *, const GmatbImage *, const Mat9 &, cudaStream_t);
int yuv2rgb_planar8_launch(const GmatbImage *, uint8_t *, int, long long, bool, bool, const Mat9 &, cudaStream_t);
int rgb2yuv_launch(const GmatbImage *, const GmatbImage *, const Mat9 &, cudaStream_t);

// ColorSpaceStandard -> the colourspace code of gmatb_csc_matrix_* (AVColorSpace numbering); BT.709 is the default
static int nvc_colorspace(int iMatrix) {
    switch (iMatrix) {
    case 4: return GMATB_SPC_FCC;
    case 5: return GMATB_SPC_BT470BG;
    case 6: return GMATB_SPC_SMPTE170M;
    case 7: return GMATB_SPC_SMPTE240M;
    case 9: return GMATB_SPC_BT2020_NCL;
    case 10: return GMATB_SPC_BT2020_CL;
    default: return GMATB_SPC_BT709;
    }
}

// Catmull-Rom weight of a tap at distance d (Resize.cu:75-79) in the operation order nvcc gives the reference
// (SASS of ScaleNv12_Bicubic_Kernel):  |d| > 1:  FADD(FFMA(d, -4, FFMA(d, (d * -0.5) * d, (d * 2.5) * d)), 2)
//                                      |d| <= 1: FADD(FFMA(d, (d * 1.5) * d, (d * -2.5) * d), 1);   |d| > 2: 0
__device__ __forceinline__ float catmull_rom(float d) {
    d = fabsf(d);
    if (d > 2.0f) return 0.0f;
    if (d > 1.0f) {
        const float t1 = __fmul_rn(d, __fmul_rn(d, -0.5f)), t2 = __fmul_rn(d, __fmul_rn(d, 2.5f));
        return __fadd_rn(__fmaf_rn(d, -4.0f, __fmaf_rn(d, t1, t2)), 2.0f);
    }
    const float u1 = __fmul_rn(d, __fmul_rn(d, 1.5f)), u2 = __fmul_rn(d, __fmul_rn(d, -2.5f));
    return __fadd_rn(__fmaf_rn(d, u1, u2), 1.0f);
}

// One output sample of a plane of CH interleaved 8-bit components: 4x4 taps from (int)f - 1, rows then columns
// accumulated as FMA chains from zero (Resize.cu:81-99, :101-123), clamp, truncate.  The tap one past the clamped
// position has weight exactly 0 (d = 2); its address is clamped into the plane instead of read past it.
template <int CH>
__device__ __forceinline__ void bicubic_sample(const uint8_t *p, int pitch, int w, int h, float fx, float fy, uint8_t (&out)[CH]) {
    const int sx0 = (int)fx - 1, sy0 = (int)fy - 1;
    float cx[4], cy[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        cx[i] = catmull_rom(__fadd_rn((float)(sx0 + i), -fx));
        cy[i] = catmull_rom(__fadd_rn((float)(sy0 + i), -fy));
    }
    float r[CH];
#pragma unroll
    for (int c = 0; c < CH; c++) r[c] = 0.0f;
#pragma unroll
    for (int y = 0; y < 4; y++) {
        const uint8_t *row = p + (size_t)min(max(sy0 + y, 0), h - 1) * pitch;
        float rx[CH];
#pragma unroll
        for (int c = 0; c < CH; c++) rx[c] = 0.0f;
#pragma unroll
        for (int x = 0; x < 4; x++) {
            const uint8_t *q = row + (size_t)min(max(sx0 + x, 0), w - 1) * CH;
#pragma unroll
            for (int c = 0; c < CH; c++) rx[c] = __fmaf_rn((float)q[c], cx[x], rx[c]);
        }
#pragma unroll
        for (int c = 0; c < CH; c++) r[c] = __fmaf_rn(rx[c], cy[y], r[c]);
    }
#pragma unroll
    for (int c = 0; c < CH; c++) out[c] = (uint8_t)fmaxf(fminf(r[c], 255.0f), 0.0f);
}

// a thread = 2 x 2 luma outputs + their chroma pair (only whole blocks, like the reference: Resize.cu:129-131)
__global__ void __launch_bounds__(256) nv12_bicubic_kernel(const uint8_t *src, int spitch, int sw, int sh, uint8_t *dst, int dpitch, int dw, int dh,
                                                           float fxs, float fys) {
    const int ix = blockIdx.x * blockDim.x + threadIdx.x, iy = blockIdx.y * blockDim.y + threadIdx.y;
    if (ix >= dw / 2 || iy >= dh / 2) return;
    const float xmax = (float)(sw - 2), ymax = (float)(sh - 2);
    uint8_t l[2][2][1];
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
        for (int i = 0; i < 2; i++)
            bicubic_sample<1>(src, spitch, sw, sh, fminf(fmaxf(__fmul_rn((float)(2 * ix + i), fxs), 2.0f), xmax),
                              fminf(fmaxf(__fmul_rn((float)(2 * iy + j), fys), 2.0f), ymax), l[j][i]);
    *reinterpret_cast<uchar2 *>(dst + (size_t)(2 * iy) * dpitch + 2 * ix) = make_uchar2(l[0][0][0], l[0][1][0]);
    *reinterpret_cast<uchar2 *>(dst + (size_t)(2 * iy + 1) * dpitch + 2 * ix) = make_uchar2(l[1][0][0], l[1][1][0]);
    uint8_t c[2];
    bicubic_sample<2>(src + (size_t)sh * spitch, spitch, sw / 2, sh / 2, fminf(fmaxf(__fmul_rn((float)ix, fxs), 2.0f), (float)(sw / 2 - 2)),
                      fminf(fmaxf(__fmul_rn((float)iy, fys), 2.0f), (float)(sh / 2 - 2)), c);
    *reinterpret_cast<uchar2 *>(dst + (size_t)(dh + iy) * dpitch + 2 * ix) = make_uchar2(c[0], c[1]);
}

}  // namespace gmatb

using namespace gmatb;

extern "C" int gmatb_nvcodec_scale_nv12_bicubic(const uint8_t *src, int spitch, int sw, int sh, uint8_t *dst, int dpitch, int dw, int dh, void *stream) {
    if (!src || !dst || sw < 4 || sh < 4 || dw < 2 || dh < 2 || spitch < sw || dpitch < dw || (dpitch & 1)) return GMATB_ERR_INVAL;
    // (float)nSrcWidth / nDstWidth evaluated in binary32, as the kernel of the reference does per thread (Resize.cu:135)
    const float fxs = (float)sw / (float)dw, fys = (float)sh / (float)dh;
    dim3 b(32, 8), g((dw / 2 + 31) / 32, (dh / 2 + 7) / 8);
    nv12_bicubic_kernel<<<g, b, 0, (cudaStream_t)stream>>>(src, spitch, sw, sh, dst, dpitch, dw, dh, fxs, fys);
    count_launch();
    return set_cuda_error(cudaGetLastError());
}

extern "C" int gmatb_nvcodec_convert(int kind, const uint8_t *src, int spitch, uint8_t *dst, int dpitch, int width, int height,
                                     int iMatrix, void *stream) {
    if (!src || !dst || kind < 0 || kind >= GMATB_NVC_COUNT || width < 2 || height < 2) return GMATB_ERR_INVAL;
    const int w = width & ~1, h = height & ~1;                  // whole 2x2 blocks only (ColorSpace.cu:137-139)
    const int cs = nvc_colorspace(iMatrix);
    cudaStream_t st = (cudaStream_t)stream;
    Mat9 M;
    GmatbImage yuv, rgb;
    memset(&yuv, 0, sizeof(yuv)); memset(&rgb, 0, sizeof(rgb));
    yuv.width = rgb.width = w; yuv.height = rgb.height = h; yuv.batch = rgb.batch = 1;
    const bool to_yuv = kind == GMATB_NVC_BGRA64_TO_P016;
    const bool p016 = kind == GMATB_NVC_P016_TO_BGRA32 || kind == GMATB_NVC_P016_TO_BGRA64 || kind == GMATB_NVC_P016_TO_BGR_PLANAR ||
                      kind == GMATB_NVC_P016_TO_BGR_FLOAT_PLANAR || to_yuv;
    uint8_t *py = (uint8_t *)(to_yuv ? dst : src);
    const int ypitch = to_yuv ? dpitch : spitch;
    yuv.format = p016 ? GMATB_FMT_P016LE : GMATB_FMT_NV12;
    yuv.data[0] = py; yuv.data[1] = py + (size_t)height * ypitch;       // chroma after the FULL height (ColorSpace.cu:146)
    yuv.linesize[0] = yuv.linesize[1] = ypitch;
    rgb.data[0] = (void *)(to_yuv ? src : dst); rgb.linesize[0] = to_yuv ? spitch : dpitch;
    if (to_yuv) {
        gmatb_csc_matrix_rgb2yuv(cs, M.m);
        rgb.format = GMATB_FMT_BGRA64LE;
        return rgb2yuv_launch(&rgb, &yuv, M, st);
    }
    gmatb_csc_matrix_yuv2rgb(cs, M.m);
    switch (kind) {
    case GMATB_NVC_NV12_TO_BGRA32: rgb.format = GMATB_FMT_BGRA; return yuv2rgb_nv12_fma_launch(&yuv, &rgb, M, st);
    case GMATB_NVC_NV12_TO_RGBA32: rgb.format = GMATB_FMT_RGBA; return yuv2rgb_nv12_fma_launch(&yuv, &rgb, M, st);
    case GMATB_NVC_NV12_TO_BGRA64: rgb.format = GMATB_FMT_BGRA64LE; return yuv2rgb_nv12_fma_launch(&yuv, &rgb, M, st);
    case GMATB_NVC_P016_TO_BGRA32: rgb.format = GMATB_FMT_BGRA; return yuv2rgb_launch(&yuv, &rgb, M, st);
    case GMATB_NVC_P016_TO_BGRA64: rgb.format = GMATB_FMT_BGRA64LE; return yuv2rgb_launch(&yuv, &rgb, M, st);
    default: break;
    }
    const bool outf = kind >= GMATB_NVC_NV12_TO_BGR_FLOAT_PLANAR;
    const bool swap = kind != GMATB_NVC_NV12_TO_RGB_PLANAR && kind != GMATB_NVC_NV12_TO_RGB_FLOAT_PLANAR;
    return yuv2rgb_planar8_launch(&yuv, dst, dpitch, (long long)dpitch * height, outf, swap, M, st);
}
