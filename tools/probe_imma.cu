// probe_imma.cu -- legacy mma.sync throughput on sm_100a: m16n8k32 u8 x s8 -> s32 (and m16n8k16 f16 -> f32)
// per SM, as a function of resident warps, alone and mixed with IDP4A / PRMT streams.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_imma tools/probe_imma.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void imma(int (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void hmma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int MODE>
__global__ void k(int iters, uint32_t seed, int *out, long long *clk) {
    uint32_t a0 = seed + threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = 0x01FF0102u, b1 = 0x03FE0201u;
    int d[8][4]; float f[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) { d[i][j] = i + j; f[i][j] = (float)(i + j); }
    int x0 = seed, x1 = seed * 3, x2 = seed * 5, x3 = seed * 7;
    uint32_t p0 = seed, p1 = seed ^ 0x55, p2 = seed ^ 0x77, p3 = seed ^ 0x99;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {          // 8 independent IMMA
#pragma unroll
            for (int i = 0; i < 8; i++) imma(d[i], a0, a1, a2, a3, b0, b1);
        } else if (MODE == 1) {   // 8 independent HMMA
#pragma unroll
            for (int i = 0; i < 8; i++) hmma(f[i], a0, a1, a2, a3, b0, b1);
        } else if (MODE == 2) {   // 2 IMMA + 16 IDP4A + 16 PRMT
#pragma unroll
            for (int i = 0; i < 2; i++) imma(d[i], a0, a1, a2, a3, b0, b1);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                x0 = __dp4a((int)p0, (int)b0, x0); x1 = __dp4a((int)p1, (int)b0, x1); x2 = __dp4a((int)p2, (int)b1, x2); x3 = __dp4a((int)p3, (int)b1, x3);
                p0 = __byte_perm(p0, p1, 0x4321); p1 = __byte_perm(p1, p2, 0x6543); p2 = __byte_perm(p2, p3, 0x4321); p3 = __byte_perm(p3, p0, 0x6543);
            }
        } else if (MODE == 3) {   // 16 IDP4A + 16 PRMT only
#pragma unroll
            for (int i = 0; i < 4; i++) {
                x0 = __dp4a((int)p0, (int)b0, x0); x1 = __dp4a((int)p1, (int)b0, x1); x2 = __dp4a((int)p2, (int)b1, x2); x3 = __dp4a((int)p3, (int)b1, x3);
                p0 = __byte_perm(p0, p1, 0x4321); p1 = __byte_perm(p1, p2, 0x6543); p2 = __byte_perm(p2, p3, 0x4321); p3 = __byte_perm(p3, p0, 0x6543);
            }
        } else if (MODE == 4) {   // 1 dependent IMMA chain (latency)
            imma(d[0], a0, a1, a2, a3, b0, b1);
        }
    }
    const long long t1 = clock64();
    int s = x0 + x1 + x2 + x3 + p0;
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) s += d[i][j] + (int)f[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run(const char *name, int per_iter_mma, int per_iter_other) {
    int *out; long long *clk;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 148 * 8);
    const int iters = 4096;
    for (int warps = 4; warps <= 32; warps *= 2) {
        k<MODE><<<148, warps * 32>>>(16, 1, out, clk);
        k<MODE><<<148, warps * 32>>>(iters, 1, out, clk);
        cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
        double c = 0; for (int i = 0; i < 148; i++) c += h[i]; c /= 148;
        printf("%-40s warps/SM %2d : %7.3f mma/clk/SM  %7.3f other-instr/clk/SM  (%.1f clk/iter/warp-set)\n", name, warps,
               (double)per_iter_mma * iters * warps / c, (double)per_iter_other * iters * warps / c, c / iters);
    }
    cudaFree(out); cudaFree(clk);
}

int main() {
    run<0>("IMMA m16n8k32 u8*s8 (8 indep)", 8, 0);
    run<1>("HMMA m16n8k16 f16 f32acc (8 indep)", 8, 0);
    run<3>("16 IDP4A + 16 PRMT", 0, 32);
    run<2>("2 IMMA + 16 IDP4A + 16 PRMT", 2, 32);
    run<4>("IMMA dependent chain", 1, 0);
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
