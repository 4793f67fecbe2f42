// tma_probe2.cu -- the CUDA programming guide's TMA example (libcu++ barrier + cp_async_bulk_tensor), to tell a
// problem of gmat_b200/csrc/tma.cuh from a problem of the box
#include <cstdio>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;

template <int BX, int BY>
__global__ void k(const __grid_constant__ CUtensorMap tensor_map, int x, int y, unsigned char *dbg) {
    __shared__ alignas(128) unsigned char smem_buffer[BY][BX];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
    } else {
        token = bar.arrive();
    }
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < BX * BY; i += blockDim.x) dbg[i] = (&smem_buffer[0][0])[i];
}

template <int BX, int BY>
int run(const char *name, int XO = 32) {
    const int W = 1920, H = 360, P = 2048;
    std::vector<unsigned char> h(P * H);
    for (size_t i = 0; i < h.size(); i++) h[i] = (unsigned char)(i * 7 + (i >> 11));
    unsigned char *d, *dbg;
    cudaMalloc(&d, P * H); cudaMalloc(&dbg, BX * BY);
    cudaMemcpy(d, h.data(), P * H, cudaMemcpyHostToDevice);
    typedef CUresult (*Encode)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                               const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    CUtensorMap m;
    const cuuint64_t dims[2] = {W, H}, strides[1] = {P};
    const cuuint32_t box[2] = {BX, BY}, es[2] = {1, 1};
    CUresult r = ((Encode)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    k<BX, BY><<<1, 128>>>(m, XO, 21, dbg);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<unsigned char> t(BX * BY);
    int bad = -1;
    if (e == cudaSuccess) {
        cudaMemcpy(t.data(), dbg, BX * BY, cudaMemcpyDeviceToHost); bad = 0;
        for (int yy = 0; yy < BY; yy++) for (int xx = 0; xx < BX; xx++) bad += t[yy * BX + xx] != h[(21 + yy) * P + XO + xx];
    }
    printf("%s box %dx%d: encode %d kernel: %s mismatches %d\n", name, BX, BY, (int)r, cudaGetErrorString(e), bad);
    return e != cudaSuccess;
}

int main(int argc, char **argv) {
    int v = argc > 1 ? atoi(argv[1]) : 0;
    if (v == 0) return run<64, 32>("libcu++");
    if (v == 1) return run<160, 48>("libcu++");
    if (v == 2) return run<128, 48>("libcu++");
    if (v == 3) return run<256, 16>("libcu++");
    if (v == 4) return run<160, 48>("libcu++ x=37", 37);
    if (v == 5) return run<160, 48>("libcu++ x=48", 48);
    return run<160, 48>("libcu++ x=4", 4);
}
