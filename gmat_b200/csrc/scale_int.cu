// scale_int.cu -- instantiations and launcher of the exact-integer fused 2:1 kernel (scale_fused4i.cuh);
// its own translation unit so that it builds in parallel with scale.cu.
#include "scale_fused5m.cuh"

#ifndef GMATB_INT_MINB
#define GMATB_INT_MINB 16
#endif
#ifndef GMATB_MMA_MINB
#define GMATB_MMA_MINB 12
#endif
namespace gmatb {

template <int L, int DST>
static void launch_int_t(int iw, bool wrap, dim3 g, cudaStream_t st, const Fused3Params &P) {
#define K(W, A, B, S) fused_csc_scale2_int_kernel<L, DST, W, A, B, S, GMATB_INT_MINB><<<g, 32, 0, st>>>(P)
    if (wrap) { if (iw == 1) K(true, -3, 19, 5); else if (iw == 2) K(true, -1, 9, 4); else K(true, -1, 5, 3); }
    else      { if (iw == 1) K(false, -3, 19, 5); else if (iw == 2) K(false, -1, 9, 4); else K(false, -1, 5, 3); }
#undef K
}
template <int L>
static int launch_int_d(int dc, int iw, bool wrap, dim3 g, cudaStream_t st, const Fused3Params &P) {
    switch (dc) {
    case D_RGB24: launch_int_t<L, D_RGB24>(iw, wrap, g, st, P); break;
    case D_BGR24: launch_int_t<L, D_BGR24>(iw, wrap, g, st, P); break;
    case D_RGBA:  launch_int_t<L, D_RGBA>(iw, wrap, g, st, P); break;
    case D_BGRA:  launch_int_t<L, D_BGRA>(iw, wrap, g, st, P); break;
    default: return GMATB_ERR_UNSUPPORTED;
    }
    count_launch();
    return set_cuda_error(cudaGetLastError());
}

// tensor-pipe form (scale_fused5m.cuh): NV12 sources only
template <int DST>
static void launch_mma_t(int iw, bool wrap, dim3 g, cudaStream_t st, const Fused3Params &P) {
#define K(W, A, B, S) fused_csc_scale2_mma_kernel<DST, W, A, B, S, GMATB_MMA_MINB><<<g, 32, 0, st>>>(P)
    if (wrap) { if (iw == 1) K(true, -3, 19, 5); else if (iw == 2) K(true, -1, 9, 4); else K(true, -1, 5, 3); }
    else      { if (iw == 1) K(false, -3, 19, 5); else if (iw == 2) K(false, -1, 9, 4); else K(false, -1, 5, 3); }
#undef K
}
int fused_mma_launch(int dc, int iw, bool wrap, dim3 g, cudaStream_t st, const Fused3Params &P) {
    switch (dc) {
    case D_RGB24: launch_mma_t<D_RGB24>(iw, wrap, g, st, P); break;
    case D_BGR24: launch_mma_t<D_BGR24>(iw, wrap, g, st, P); break;
    case D_RGBA:  launch_mma_t<D_RGBA>(iw, wrap, g, st, P); break;
    case D_BGRA:  launch_mma_t<D_BGRA>(iw, wrap, g, st, P); break;
    default: return GMATB_ERR_UNSUPPORTED;
    }
    count_launch();
    return set_cuda_error(cudaGetLastError());
}

int fused_int_launch(bool semi, int dc, int iw, bool wrap, dim3 g, cudaStream_t st, const Fused3Params &P) {
    return semi ? launch_int_d<L_NV12>(dc, iw, wrap, g, st, P) : launch_int_d<L_I420>(dc, iw, wrap, g, st, P);
}

}  // namespace gmatb
