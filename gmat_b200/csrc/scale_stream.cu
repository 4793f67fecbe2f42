// scale_stream.cu -- instantiations and launcher of the any-ratio streaming kernel (scale_stream.cuh); its own
// translation unit so that it builds in parallel with scale.cu.
#include "scale_generic.cuh"
#include "scale_stream.cuh"

namespace gmatb {

template <int L, int DST>
static void launch_stream_t(int nout, int deal, int ra, dim3 g, cudaStream_t st, const StreamParams &P) {
    // NOUT = outputs per lane: 3 (ratios >= 2.5: at most 96 outputs per 240-column strip) or 5; DEAL: see the kernel;
    // RA: bilinear / nearest arithmetic (integer-valued samples, rint + saturate)
    if (ra) {
        if (nout <= 3)  fused_csc_scale_stream_kernel<L, DST, 3, 0, 1, 16><<<g, 32, 0, st>>>(P);
        else if (!deal) fused_csc_scale_stream_kernel<L, DST, 5, 0, 1, 12><<<g, 32, 0, st>>>(P);
        else            fused_csc_scale_stream_kernel<L, DST, 5, 1, 1, 12><<<g, 32, 0, st>>>(P);
        return;
    }
    if (nout <= 3)  fused_csc_scale_stream_kernel<L, DST, 3, 0, 0, 16><<<g, 32, 0, st>>>(P);
    else if (!deal) fused_csc_scale_stream_kernel<L, DST, 5, 0, 0, 12><<<g, 32, 0, st>>>(P);
    else            fused_csc_scale_stream_kernel<L, DST, 5, 1, 0, 12><<<g, 32, 0, st>>>(P);
}
template <int L>
static int launch_stream_d(int dc, int nout, int deal, int ra, dim3 g, cudaStream_t st, const StreamParams &P) {
    switch (dc) {
    case D_RGB24: launch_stream_t<L, D_RGB24>(nout, deal, ra, g, st, P); break;
    case D_BGR24: launch_stream_t<L, D_BGR24>(nout, deal, ra, g, st, P); break;
    case D_RGBA:  launch_stream_t<L, D_RGBA>(nout, deal, ra, g, st, P); break;
    case D_BGRA:  launch_stream_t<L, D_BGRA>(nout, deal, ra, g, st, P); break;
    default: return GMATB_ERR_UNSUPPORTED;
    }
    count_launch();
    return set_cuda_error(cudaGetLastError());
}

int stream_launch(bool semi, int dc, int nout, int deal, int ra, dim3 g, cudaStream_t st, const StreamParams &P) {
    return semi ? launch_stream_d<L_NV12>(dc, nout, deal, ra, g, st, P) : launch_stream_d<L_I420>(dc, nout, deal, ra, g, st, P);
}

template <int CH, int SBITS>
static void launch_plane_t(int nout, int deal, int ra, dim3 g, cudaStream_t st, const PlaneStreamParams &P) {
#define PK(N_, D_, R_) plane_scale_stream_kernel<CH, SBITS, N_, D_, R_, (CH >= 3 ? 12 : 16)><<<g, 32, 0, st>>>(P)
    if (ra) { if (nout <= 3) PK(3, 0, 1); else if (deal) PK(5, 1, 1); else PK(5, 0, 1); }
    else    { if (nout <= 3) PK(3, 0, 0); else if (deal) PK(5, 1, 0); else PK(5, 0, 0); }
#undef PK
}
int plane_stream_launch(int ch, int bits, int nout, int deal, int ra, dim3 g, cudaStream_t st, const PlaneStreamParams &P) {
    if (bits == 8) { if (ch == 1) launch_plane_t<1, 8>(nout, deal, ra, g, st, P); else if (ch == 2) launch_plane_t<2, 8>(nout, deal, ra, g, st, P); else if (ch == 3) launch_plane_t<3, 8>(nout, deal, ra, g, st, P); else launch_plane_t<4, 8>(nout, deal, ra, g, st, P); }
    else           { if (ch == 1) launch_plane_t<1, 16>(nout, deal, ra, g, st, P); else launch_plane_t<2, 16>(nout, deal, ra, g, st, P); }
    count_launch();
    return set_cuda_error(cudaGetLastError());
}

}  // namespace gmatb
