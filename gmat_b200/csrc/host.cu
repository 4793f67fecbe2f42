// host.cu -- host-side plumbing of the C ABI: error state, descriptors, CSC matrices,
// and the unscaled-converter entry points of include/gmat_b200.h.
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include "common.cuh"
#include "tma.cuh"

namespace gmatb {

static std::atomic<int> g_last_err{0};
static std::atomic<long long> g_launches{0};

int set_cuda_error(cudaError_t e) {
    if (e == cudaSuccess) return GMATB_OK;
    g_last_err.store((int)e);
    return GMATB_ERR_CUDA;
}
void count_launch(int n) { g_launches.fetch_add(n); }

// Diagnostics that must not be silent (an option that degrades to another algorithm): stderr by default, or the
// caller's sink -- the libswscale shim and the AVFilter glue route it to av_log.
static std::atomic<void (*)(const char *)> g_log{nullptr};
void gmatb_log(const char *msg) {
    void (*cb)(const char *) = g_log.load();
    if (cb) cb(msg); else fprintf(stderr, "%s\n", msg);
}

int fmt_planes(int fmt) {
    switch (fmt) {
    case GMATB_FMT_NV12: case GMATB_FMT_P010LE: case GMATB_FMT_P016LE: return 2;
    case GMATB_FMT_YUV420P: case GMATB_FMT_YUV420P10LE: case GMATB_FMT_YUV420P16LE: return 3;
    case GMATB_FMT_RGBPF32LE: return 3;
    case GMATB_FMT_RGBAPF32LE: return 4;
    default: return 1;
    }
}

bool to_img(const GmatbImage *g, Img *out, int nplanes) {
    if (!g || g->width <= 0 || g->height <= 0) return false;
    memset(out, 0, sizeof(*out));
    out->w = g->width; out->h = g->height;
    for (int i = 0; i < nplanes && i < 4; i++) {
        if (!g->data[i]) return false;
        out->pl[i].p = (uint8_t *)g->data[i];
        out->pl[i].pitch = g->linesize[i];
        out->pl[i].bstride = g->batch > 1 ? g->batch_stride[i] : 0;
    }
    return true;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point table (no link-time dependency on libcuda)
bool make_tensor_map_3d(CUtensorMap *out, CUtensorMapDataType dtype, int elem_bytes, const void *base, unsigned long long dim_x,
                        unsigned long long dim_y, unsigned long long dim_z, unsigned long long pitch_bytes, unsigned long long frame_bytes,
                        unsigned box_x, unsigned box_y) {
    typedef CUresult (*Encode)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                               const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                               CUtensorMapFloatOOBfill);
    static Encode enc = nullptr;
    static std::atomic<int> state{0};          // 0 unknown, 1 ok, 2 unavailable
    if (state.load() == 0) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && fn && q == cudaDriverEntryPointSuccess) {
            enc = (Encode)fn; state.store(1);
        } else state.store(2);
    }
    if (state.load() != 1) return false;
    if (((uintptr_t)base & 15) || (pitch_bytes & 15) || (frame_bytes & 15) || box_x > 256 || box_y > 256 || ((box_x * elem_bytes) & 15)) return false;
    if (dim_z < 1) dim_z = 1;
    if (frame_bytes == 0) frame_bytes = pitch_bytes * dim_y;      // single frame: any legal stride
    const cuuint64_t dims[3] = {dim_x, dim_y, dim_z};
    const cuuint64_t strides[2] = {pitch_bytes, frame_bytes};
    const cuuint32_t box[3] = {box_x, box_y, 1};
    const cuuint32_t es[3] = {1, 1, 1};
    return enc(out, dtype, 3, const_cast<void *>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Same expressions, same types, same evaluation order as the reference
// (yuv2rgb_cuda.cu:782-848): entries in float, range scale in double, cast to float.
static void get_constants(int cspace, float *wr, float *wb, int *black, int *white, int *max) {
    *black = 16; *white = 235; *max = 255;
    switch (cspace) {
    case GMATB_SPC_BT709:     *wr = 0.2126f; *wb = 0.0722f; break;
    case GMATB_SPC_FCC:       *wr = 0.30f;   *wb = 0.11f;   break;
    case GMATB_SPC_SMPTE240M: *wr = 0.212f;  *wb = 0.087f;  break;
    case GMATB_SPC_BT2020_NCL:
    case GMATB_SPC_BT2020_CL:
        *wr = 0.2627f; *wb = 0.0593f;
        *black = 64 << 6; *white = 940 << 6; *max = (1 << 16) - 1;
        break;
    default:                  *wr = 0.2990f; *wb = 0.1140f; break;   // BT470BG / SMPTE170M / everything else
    }
}

}  // namespace gmatb

extern "C" void gmatb_set_log(void (*cb)(const char *)) { gmatb::g_log.store(cb); }

using namespace gmatb;

extern "C" {

int gmatb_version(void) { return GMATB_VERSION; }
int gmatb_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { set_cuda_error(e); return GMATB_ERR_CUDA; }
    return n;
}
int gmatb_last_cuda_error(void) { return g_last_err.load(); }
const char *gmatb_last_cuda_error_string(void) { return cudaGetErrorString((cudaError_t)g_last_err.load()); }
long long gmatb_launch_count(void) { return g_launches.load(); }
int gmatb_device_sync(void) { return set_cuda_error(cudaDeviceSynchronize()); }

void gmatb_csc_matrix_yuv2rgb(int cspace, float out9[9]) {
    float wr, wb; int black, white, max;
    get_constants(cspace, &wr, &wb, &black, &white, &max);
    volatile float mat[3][3] = {
        {1.0f, 0.0f, (1.0f - wr) / 0.5f},
        {1.0f, -wb * (1.0f - wb) / 0.5f / (1 - wb - wr), -wr * (1 - wr) / 0.5f / (1 - wb - wr)},
        {1.0f, (1.0f - wb) / 0.5f, 0.0f},
    };
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            out9[i * 3 + j] = (float)(1.0 * max / (white - black) * mat[i][j]);
}
void gmatb_csc_matrix_rgb2yuv(int cspace, float out9[9]) {
    float wr, wb; int black, white, max;
    get_constants(cspace, &wr, &wb, &black, &white, &max);
    volatile float mat[3][3] = {
        {wr, 1.0f - wb - wr, wb},
        {-0.5f * wr / (1.0f - wb), -0.5f * (1 - wb - wr) / (1.0f - wb), 0.5f},
        {0.5f, -0.5f * (1.0f - wb - wr) / (1.0f - wr), -0.5f * wb / (1.0f - wr)},
    };
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            out9[i * 3 + j] = (float)(1.0 * (white - black) / max * mat[i][j]);
}

}  // extern "C"

namespace gmatb {
int yuv2rgb_launch(const GmatbImage *, const GmatbImage *, const Mat9 &, cudaStream_t);
int yuv2rgb_planar_launch(const GmatbImage *, const GmatbImage *, const Mat9 &, float, const float *, cudaStream_t);
int rgb2yuv_launch(const GmatbImage *, const GmatbImage *, const Mat9 &, cudaStream_t);
int yuv2yuv_launch(const GmatbImage *, const GmatbImage *, cudaStream_t);
int rgb24swap_launch(const GmatbImage *, const GmatbImage *, cudaStream_t);
}

extern "C" {

int gmatb_yuv2rgb(const GmatbImage *src, const GmatbImage *dst, int cspace, void *stream) {
    Mat9 M; gmatb_csc_matrix_yuv2rgb(cspace, M.m);
    if (dst && (dst->format == GMATB_FMT_RGBPF32LE || dst->format == GMATB_FMT_RGBAPF32LE))
        return yuv2rgb_planar_launch(src, dst, M, 255.0f, nullptr, (cudaStream_t)stream);
    return yuv2rgb_launch(src, dst, M, (cudaStream_t)stream);
}
int gmatb_yuv2rgb_planar_f32(const GmatbImage *src, const GmatbImage *dst, int cspace, float norm,
                             const float shift_rgb[3], void *stream) {
    Mat9 M; gmatb_csc_matrix_yuv2rgb(cspace, M.m);
    return yuv2rgb_planar_launch(src, dst, M, norm, shift_rgb, (cudaStream_t)stream);
}
int gmatb_rgb2yuv(const GmatbImage *src, const GmatbImage *dst, int cspace, void *stream) {
    Mat9 M; gmatb_csc_matrix_rgb2yuv(cspace, M.m);
    return rgb2yuv_launch(src, dst, M, (cudaStream_t)stream);
}
int gmatb_yuv2yuv(const GmatbImage *src, const GmatbImage *dst, void *stream) {
    return yuv2yuv_launch(src, dst, (cudaStream_t)stream);
}
int gmatb_rgb24tobgr24(const GmatbImage *src, const GmatbImage *dst, void *stream) {
    return rgb24swap_launch(src, dst, (cudaStream_t)stream);
}

}  // extern "C"
