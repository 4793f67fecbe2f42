/* oracle/ref_o2_driver.c -- TEST INFRASTRUCTURE ONLY (oracle "O2").
 *
 * Launches the reference's scale_cuda kernels (libavfilter/vf_scale_cuda.cu,
 * compiled unmodified to a cubin by oracle/refbuild/Makefile) the way
 * libavfilter/vf_scale_cuda.c does it: one pitch2D texture object per input
 * plane (vf_scale_cuda.c:442-471; point filter + normalised-float reads for
 * Bicubic/Lanczos, linear filter + integer reads for Bilinear, :294-313),
 * grid (ceil(W/32), ceil(H/16)) x block (32,16) (:57-58, :423-425), args in the
 * order of KERNEL_ARGS (vf_scale_cuda.cu:1086-1091).
 */
#include <cuda.h>
#include <stdio.h>
#include <string.h>

#define CK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char *n_ = 0; \
    cuGetErrorName(r_, &n_); fprintf(stderr, "ref_o2: %s -> %s\n", #x, n_ ? n_ : "?"); return -(int)r_; } } while (0)

static CUmodule g_mod;
static CUcontext g_ctx;

int ref_o2_load(const char *cubin_path)
{
    CUdevice dev;
    CK(cuInit(0));
    CK(cuCtxGetCurrent(&g_ctx));
    if (!g_ctx) {
        CK(cuDeviceGet(&dev, 0));
        CK(cuDevicePrimaryCtxRetain(&g_ctx, dev));
        CK(cuCtxSetCurrent(g_ctx));
    }
    if (!g_mod) CK(cuModuleLoad(&g_mod, cubin_path));
    return 0;
}

/* one launch == one call_resize_kernel() of the reference */
int ref_o2_launch(const char *func_name, int n_planes,
                  const void *src[4], const int src_pitch[4],
                  const int plane_w[4], const int plane_h[4],
                  const int plane_depth[4], const int plane_channels[4],
                  void *dst[4], int dst_w, int dst_h, int dst_pitch,
                  int src_w, int src_h, float param,
                  int use_linear, int as_integer)
{
    CUfunction fn;
    CUtexObject tex[4] = {0, 0, 0, 0};
    CUdeviceptr d[4];
    int i;
    CK(cuCtxSetCurrent(g_ctx));
    CK(cuModuleGetFunction(&fn, g_mod, func_name));
    for (i = 0; i < n_planes; i++) {
        CUDA_TEXTURE_DESC td;
        CUDA_RESOURCE_DESC rd;
        memset(&td, 0, sizeof(td));
        memset(&rd, 0, sizeof(rd));
        td.filterMode = use_linear ? CU_TR_FILTER_MODE_LINEAR : CU_TR_FILTER_MODE_POINT;
        td.flags = as_integer ? CU_TRSF_READ_AS_INTEGER : 0;
        rd.resType = CU_RESOURCE_TYPE_PITCH2D;
        rd.res.pitch2D.format = plane_depth[i] <= 8 ? CU_AD_FORMAT_UNSIGNED_INT8 : CU_AD_FORMAT_UNSIGNED_INT16;
        rd.res.pitch2D.numChannels = plane_channels[i];
        rd.res.pitch2D.pitchInBytes = src_pitch[i];
        rd.res.pitch2D.devPtr = (CUdeviceptr)src[i];
        rd.res.pitch2D.width = plane_w[i];
        rd.res.pitch2D.height = plane_h[i];
        CK(cuTexObjectCreate(&tex[i], &rd, &td, NULL));
    }
    for (i = 0; i < 4; i++) d[i] = (CUdeviceptr)dst[i];
    {
        void *args[] = { &tex[0], &tex[1], &tex[2], &tex[3], &d[0], &d[1], &d[2], &d[3],
                         &dst_w, &dst_h, &dst_pitch, &src_w, &src_h, &param };
        CK(cuLaunchKernel(fn, (dst_w + 31) / 32, (dst_h + 15) / 16, 1, 32, 16, 1, 0, 0, args, NULL));
    }
    CK(cuCtxSynchronize());
    for (i = 0; i < n_planes; i++) cuTexObjectDestroy(tex[i]);
    return 0;
}
