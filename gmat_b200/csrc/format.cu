// format.cu -- the format_cuda filter's kernels: NV12 <-> planar float RGB.
//
// Reference: libavfilter/vf_format_cuda.c:185-203 (format_convert) calling
// libavfilter/format_cuda_kernel.cu:583-632 (nv12_to_rgbpf32[_shift], nv12_to_bgrpf32_shift,
// rgbpf32_to_nv12).  Parity is pinned by "O3" = that file compiled unmodified for sm_100a
// (oracle/refbuild/Makefile), arithmetic read from its SASS:
//
//   NV12 -> RGBPF32 (YuvToRgbPlanarKernel :257-297): exactly libgpuscale's planar kernel --
//     c = FFMA(fv, mC, FFMA(fy, mA, FMUL(fu, mB))), clamp to [0,255], truncate to u8, then the IEEE
//     division (c - shift) / norm -- so it runs on csc.cu's yuv2rgb_planar_f32_kernel.  The only
//     differences are the matrix selection (GetConstants :32-63: BT.709 is the default branch and
//     SMPTE170M falls into it) and that plane pointers / pitches are honoured.
//   RGBPF32 -> NV12 (RgbpToYuvKernel :516-570), per 2x2 block, with x' = FMUL(x, 255):
//     Y   = trunc(FADD(FFMA(b', m02, FFMA(r', m00, FMUL(g', m01))), 16))
//     r_m = FMUL(0.25, ((r'00 + r'01) + r'10) + r'11), g_m likewise,
//     b_m = FMUL(0.25, FFMA(b11, 255, (b'00 + b'01) + b'10))    (nvcc contracts the last product)
//     U   = trunc(FADD(FFMA(b_m, m12, FFMA(r_m, m10, FMUL(g_m, m11))), 128)), V with row 2
//     stored as the low byte of F2I.U32.TRUNC (negative -> 0, >= 256 wraps).
//     Reference defect reproduced because O3 is the parity definition: the bottom-right luma of
//     each block uses g' in place of b' (:560 `RgbToY(int2bR.y, int2bG.y, int2bG.y)`).
//     Not reproduced: the reference ignores data[1..2] of both frames (planes assumed contiguous
//     at height*pitch); we take every plane pointer -- identical for pool frames.
#include "csc_core.cuh"

namespace gmatb {

int yuv2rgb_planar_launch(const GmatbImage *, const GmatbImage *, const Mat9 &, float, const float *, cudaStream_t);

// a thread converts 4 columns x 2 rows: 6 x 16-byte loads, 2 x 4 luma bytes + 4 chroma bytes out
__global__ void __launch_bounds__(256) rgbpf32_to_nv12_kernel(Img src, Img dst, Mat9 M) {
    const int x0 = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int y0 = (blockIdx.y * 8 + threadIdx.y) * 2;
    if (x0 >= src.w || y0 >= src.h) return;
    const long long fz = blockIdx.z;
    float v[3][2][4];
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const float4 q = __ldcs(reinterpret_cast<const float4 *>(src.pl[c].p + fz * src.pl[c].bstride +
                                                                     (size_t)(y0 + r) * src.pl[c].pitch + (size_t)x0 * 4));
            v[c][r][0] = q.x; v[c][r][1] = q.y; v[c][r][2] = q.z; v[c][r][3] = q.w;
        }
    uint32_t yw[2] = {0u, 0u}, cw = 0u;
#pragma unroll
    for (int j = 0; j < 2; j++) {          // two 2x2 blocks
        float s[3][2][2];                  // x' = x * 255
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
            for (int r = 0; r < 2; r++)
#pragma unroll
                for (int h = 0; h < 2; h++) s[c][r][h] = __fmul_rn(v[c][r][2 * j + h], 255.0f);
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const float b = (r == 1 && h == 1) ? s[1][1][1] : s[2][r][h];      // format_cuda_kernel.cu:560
                const float t = __fmaf_rn(b, M.m[2], __fmaf_rn(s[0][r][h], M.m[0], __fmul_rn(s[1][r][h], M.m[1])));
                const uint32_t yv = (uint32_t)trunc_i(__fadd_rn(t, 16.0f));
                yw[r] |= (max((int)yv, 0) & 0xFFu) << (8 * (2 * j + h));
            }
        const float rm = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(s[0][0][0], s[0][0][1]), s[0][1][0]), s[0][1][1]), 0.25f);
        const float gm = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(s[1][0][0], s[1][0][1]), s[1][1][0]), s[1][1][1]), 0.25f);
        const float bm = __fmul_rn(__fmaf_rn(v[2][1][2 * j + 1], 255.0f, __fadd_rn(__fadd_rn(s[2][0][0], s[2][0][1]), s[2][1][0])), 0.25f);
        const float u = __fadd_rn(__fmaf_rn(bm, M.m[5], __fmaf_rn(rm, M.m[3], __fmul_rn(gm, M.m[4]))), 128.0f);
        const float w = __fadd_rn(__fmaf_rn(bm, M.m[8], __fmaf_rn(rm, M.m[6], __fmul_rn(gm, M.m[7]))), 128.0f);
        cw |= ((uint32_t)max(trunc_i(u), 0) & 0xFFu) << (16 * j);
        cw |= ((uint32_t)max(trunc_i(w), 0) & 0xFFu) << (16 * j + 8);
    }
    uint8_t *py = dst.pl[0].p + fz * dst.pl[0].bstride + (size_t)y0 * dst.pl[0].pitch + x0;
    stg32(py, yw[0]);
    stg32(py + dst.pl[0].pitch, yw[1]);
    stg32(dst.pl[1].p + fz * dst.pl[1].bstride + (size_t)(y0 >> 1) * dst.pl[1].pitch + x0, cw);
}

}  // namespace gmatb

using namespace gmatb;

extern "C" {

// GetConstants (format_cuda_kernel.cu:32-63) as a map onto the matrices of gmatb_csc_matrix_*:
// BT.709 is the default branch (SMPTE170M and "unspecified" land there), BT470BG is BT.601.
int gmatb_format_colorspace(int av_colorspace) {
    switch (av_colorspace) {
    case GMATB_SPC_FCC:        return GMATB_SPC_FCC;
    case GMATB_SPC_BT470BG:    return GMATB_SPC_BT470BG;
    case GMATB_SPC_SMPTE240M:  return GMATB_SPC_SMPTE240M;
    case GMATB_SPC_BT2020_NCL:
    case GMATB_SPC_BT2020_CL:  return GMATB_SPC_BT2020_NCL;
    default:                   return GMATB_SPC_BT709;
    }
}

int gmatb_format_nv12_to_rgbpf32(const GmatbImage *src, const GmatbImage *dst, int av_colorspace,
                                 float norm, const float shift_rgb[3], int bgr_planes, void *stream) {
    if (!src || !dst || src->format != GMATB_FMT_NV12) return GMATB_ERR_INVAL;
    Mat9 M; gmatb_csc_matrix_yuv2rgb(gmatb_format_colorspace(av_colorspace), M.m);
    GmatbImage d = *dst;
    if (bgr_planes) {        // BGRAF32: plane 0 = B, 1 = G, 2 = R (format_cuda_kernel.cu:601-609)
        d.data[0] = dst->data[2]; d.linesize[0] = dst->linesize[2]; d.batch_stride[0] = dst->batch_stride[2];
        d.data[2] = dst->data[0]; d.linesize[2] = dst->linesize[0]; d.batch_stride[2] = dst->batch_stride[0];
    }
    return yuv2rgb_planar_launch(src, &d, M, norm, shift_rgb, (cudaStream_t)stream);
}

int gmatb_format_rgbpf32_to_nv12(const GmatbImage *src, const GmatbImage *dst, int av_colorspace, void *stream) {
    if (!src || !dst || src->width != dst->width || src->height != dst->height || src->width <= 0 || src->height <= 0)
        return GMATB_ERR_INVAL;
    if (src->format != GMATB_FMT_RGBPF32LE || dst->format != GMATB_FMT_NV12) return GMATB_ERR_UNSUPPORTED;
    // the reference skips the last column / row of odd sizes (:519-521); 4-column vectors need width % 4 == 0
    if ((src->width & 3) || (src->height & 1)) return GMATB_ERR_UNSUPPORTED;
    Img s, d;
    if (!to_img(src, &s, 3) || !to_img(dst, &d, 2)) return GMATB_ERR_INVAL;
    for (int i = 0; i < 3; i++)
        if (((uintptr_t)s.pl[i].p | (uintptr_t)s.pl[i].pitch | (uintptr_t)s.pl[i].bstride) & 15) return GMATB_ERR_INVAL;
    for (int i = 0; i < 2; i++)
        if (((uintptr_t)d.pl[i].p | (uintptr_t)d.pl[i].pitch | (uintptr_t)d.pl[i].bstride) & 3) return GMATB_ERR_INVAL;
    Mat9 M; gmatb_csc_matrix_rgb2yuv(gmatb_format_colorspace(av_colorspace), M.m);
    dim3 g((s.w / 4 + 31) / 32, (s.h / 2 + 7) / 8, src->batch > 1 ? src->batch : 1), b(32, 8);
    rgbpf32_to_nv12_kernel<<<g, b, 0, (cudaStream_t)stream>>>(s, d, M);
    count_launch();
    return set_cuda_error(cudaGetLastError());
}

}  // extern "C"
