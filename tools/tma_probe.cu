// tma_probe.cu -- check of gmat_b200/csrc/tma.cuh on the box: one box load + one box store.
// usage: tma_probe <variant>   0: rank-3 map as __grid_constant__ parameter   1: rank-3 map in global memory
//                              2: rank-2 map as parameter (2d instructions)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/tma_probe tools/tma_probe.cu -Lgmat_b200 -lgmat_b200
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../gmat_b200/csrc/tma.cuh"
using namespace gmatb;

template <int V>
__global__ void k(const __grid_constant__ CUtensorMap in, const __grid_constant__ CUtensorMap out, const CUtensorMap *gin, const CUtensorMap *gout,
                  int x, int y, int bx, int by, uint8_t *dbg) {
    extern __shared__ __align__(128) uint8_t tile[];
    __shared__ uint64_t bar;
    const CUtensorMap *mi = V == 1 ? gin : &in, *mo = V == 1 ? gout : &out;
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, bx * by);
        if (V == 2) tma_load_2d(tile, mi, &bar, x, y); else tma_load_3d(tile, mi, &bar, x, y, 0);
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < bx * by; i += blockDim.x) { dbg[i] = tile[i]; tile[i] = tile[i] ^ 0xFF; }
    fence_async_smem();
    __syncthreads();
    if (threadIdx.x == 0 && V != 2) { tma_store_3d(mo, tile, x, y, 0); tma_store_commit_wait(); }
}

int main(int argc, char **argv) {
    const int V = argc > 1 ? atoi(argv[1]) : 0; const int XO = argc > 2 ? atoi(argv[2]) : 37;
    const int W = 1920, H = 360, P = 2048, BX = 160, BY = 48;
    std::vector<uint8_t> h(P * H);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 7 + (i >> 11));
    uint8_t *d, *o, *dbg;
    cudaMalloc(&d, P * H); cudaMalloc(&o, P * H); cudaMalloc(&dbg, BX * BY);
    cudaMemcpy(d, h.data(), P * H, cudaMemcpyHostToDevice); cudaMemset(o, 0, P * H);
    CUtensorMap mi, mo;
    bool a, b;
    if (V == 2) {
        typedef CUresult (*Encode)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                   const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void *fn = nullptr; cudaDriverEntryPointQueryResult q;
        cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        const cuuint64_t dims[2] = {W, H}, strides[1] = {P};
        const cuuint32_t box[2] = {BX, BY}, es[2] = {1, 1};
        CUresult r = ((Encode)fn)(&mi, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode2d: %d\n", (int)r); a = b = r == CUDA_SUCCESS; mo = mi;
    } else {
        a = make_tensor_map_3d(&mi, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, d, W, H, 1, P, 0, BX, BY);
        b = make_tensor_map_3d(&mo, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, o, W, H, 1, P, 0, BX, BY);
    }
    printf("variant %d encode: %d %d  sizeof(CUtensorMap)=%zu align=%zu\n", V, a, b, sizeof(CUtensorMap), alignof(CUtensorMap));
    CUtensorMap *gm; cudaMalloc(&gm, 2 * sizeof(CUtensorMap));
    cudaMemcpy(gm, &mi, sizeof(mi), cudaMemcpyHostToDevice); cudaMemcpy(gm + 1, &mo, sizeof(mo), cudaMemcpyHostToDevice);
    if (V == 0) k<0><<<1, 128, BX * BY>>>(mi, mo, gm, gm + 1, XO, 21, BX, BY, dbg);
    if (V == 1) k<1><<<1, 128, BX * BY>>>(mi, mo, gm, gm + 1, 37, 21, BX, BY, dbg);
    if (V == 2) k<2><<<1, 128, BX * BY>>>(mi, mo, gm, gm + 1, 37, 21, BX, BY, dbg);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<uint8_t> t(BX * BY), r(P * H);
    cudaMemcpy(t.data(), dbg, BX * BY, cudaMemcpyDeviceToHost); cudaMemcpy(r.data(), o, P * H, cudaMemcpyDeviceToHost);
    int bad = 0, bad2 = 0;
    for (int yy = 0; yy < BY; yy++) for (int xx = 0; xx < BX; xx++) {
        bad += t[yy * BX + xx] != h[(21 + yy) * P + XO + xx];
        bad2 += r[(21 + yy) * P + XO + xx] != (uint8_t)(h[(21 + yy) * P + XO + xx] ^ 0xFF);
    }
    printf("load mismatches %d, store mismatches %d\n", bad, bad2);
    return 0;
}
