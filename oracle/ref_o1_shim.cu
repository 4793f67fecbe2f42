/* oracle/ref_o1_shim.cu -- TEST INFRASTRUCTURE ONLY (oracle "O1").
 *
 * Compiles the reference's own libgpuscale CSC kernels, unmodified, for sm_100a
 * by #including the file where it lies under /root/reference (nothing is copied
 * into this repo), and adds extern "C" entry points for the launchers that the
 * reference instantiates but never dispatches (yuv2rgb_cuda.cu:604-618, :741-746,
 * :564-570), so that the parity tests can pin the 16-bit and planar-float
 * arithmetic against the reference's own code on the same GPU.
 *
 * The dispatched entry points (yuv2rgb_cuda, rgb2yuv_cuda, set_mat_*_cuda) are
 * already extern "C" in the included file (yuv2rgb_cuda.cu:777-948).
 */
#include "libswscale/cuda/yuv2rgb_cuda.cu"

extern "C" {

/* P010/P016 (single buffer, UV at src + h*pitch) -> RGBA64 / BGRA64.
 * order: 0 = RGBA64, 1 = BGRA64.  Even width/height only (yuv2rgb_kernel :186). */
int ref_p016_to_color64(uint8_t *src, int src_pitch, uint8_t *dst, int dst_pitch,
                        int w, int h, int order, CUstream s)
{
    if (order == 0) p0162color64<RGBA64>(src, src_pitch, dst, dst_pitch, w, h, s);
    else            p0162color64<BGRA64>(src, src_pitch, dst, dst_pitch, w, h, s);
    return (int)cudaGetLastError();
}

/* NV12 -> planar float BGR (instantiated at :650, never dispatched) */
int ref_nv12_to_bgrpf32(const uint8_t *src, int src_pitch, uint8_t *dst, int dst_pitch,
                        int w, int h, CUstream s)
{
    nv122color_planar<BGRF32, float2>(src, src_pitch, dst, dst_pitch, w, h, s);
    return (int)cudaGetLastError();
}

/* BGRA64 -> P016 (yuv2rgb_cuda.cu:741-746) */
int ref_bgra64_to_p016(const uint8_t *src, int src_pitch, uint8_t *dst, int dst_pitch,
                       int w, int h, CUstream s)
{
    Bgra64ToP016(src, src_pitch, dst, dst_pitch, w, h, s);
    return (int)cudaGetLastError();
}

int ref_sync(void) { return (int)cudaDeviceSynchronize(); }

/* read back the 9 floats the reference uploaded (KAT for SURVEY 8a table) */
int ref_get_mat(int which, float *out9)
{
    return (int)(which == 0 ? cudaMemcpyFromSymbol(out9, matYuv2Rgb, 36)
                            : cudaMemcpyFromSymbol(out9, matRgb2Yuv, 36));
}
}
