// probe_minmax.cu -- which pipe runs the packed 16-bit min/max forms on sm_100a, and do they overlap?
// VIMNMX.U16x2 / VIMNMX3.U16x2 (alu pipe), HMNMX2 (packed half min/max) and the HFMA2.RELU compare-exchange
// (fma pipe), alone and interleaved, per SM as a function of resident warps.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_minmax tools/probe_minmax.cu
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned hmin(unsigned a, unsigned b) { unsigned r; asm volatile("min.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ unsigned hmax(unsigned a, unsigned b) { unsigned r; asm volatile("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ unsigned imin(unsigned a, unsigned b) { unsigned r; asm volatile("min.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ unsigned imax(unsigned a, unsigned b) { unsigned r; asm volatile("max.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ unsigned frelu(unsigned a, unsigned b, unsigned c) { unsigned r; asm volatile("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ unsigned hadd(unsigned a, unsigned b) { unsigned r; asm volatile("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ unsigned hsub(unsigned a, unsigned b) { unsigned r; asm volatile("sub.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }

template <int MODE>
__global__ void k(int iters, uint32_t seed, unsigned *out, long long *clk) {
    unsigned x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = 0x64006400u + ((seed * (i + 3) + threadIdx.x * 7) & 0x00ff00ffu);
    const unsigned one = 0x3c003c00u;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {                 // 16 VIMNMX.U16x2 (8 compare-exchanges)
#pragma unroll
            for (int i = 0; i < 16; i += 2) { unsigned a = imin(x[i], x[i + 1]), b = imax(x[i], x[i + 1]); x[i] = b; x[i + 1] = a; }
        } else if (MODE == 1) {          // 16 HMNMX2
#pragma unroll
            for (int i = 0; i < 16; i += 2) { unsigned a = hmin(x[i], x[i + 1]), b = hmax(x[i], x[i + 1]); x[i] = b; x[i + 1] = a; }
        } else if (MODE == 2) {          // 8 VIMNMX + 8 HMNMX2
#pragma unroll
            for (int i = 0; i < 8; i += 2) { unsigned a = imin(x[i], x[i + 1]), b = imax(x[i], x[i + 1]); x[i] = b; x[i + 1] = a; }
#pragma unroll
            for (int i = 8; i < 16; i += 2) { unsigned a = hmin(x[i], x[i + 1]), b = hmax(x[i], x[i + 1]); x[i] = b; x[i + 1] = a; }
        } else if (MODE == 3) {          // 8 compare-exchanges as HFMA2.RELU + HADD2 + HADD2 (24 fma-pipe instructions)
#pragma unroll
            for (int i = 0; i < 16; i += 2) { unsigned t = frelu(x[i + 1], one, x[i] ^ 0x80008000u); unsigned b = hadd(x[i], t), a = hsub(x[i + 1], t); x[i] = b; x[i + 1] = a; }
        } else if (MODE == 4) {          // 4 CE on VIMNMX (8 instr) + 4 CE on the fma pipe (12 + 4 LOP)
#pragma unroll
            for (int i = 0; i < 8; i += 2) { unsigned a = imin(x[i], x[i + 1]), b = imax(x[i], x[i + 1]); x[i] = b; x[i + 1] = a; }
#pragma unroll
            for (int i = 8; i < 16; i += 2) { unsigned t = frelu(x[i + 1], one, x[i] ^ 0x80008000u); unsigned b = hadd(x[i], t), a = hsub(x[i + 1], t); x[i] = b; x[i + 1] = a; }
        } else if (MODE == 5) {          // 16 VIMNMX3
#pragma unroll
            for (int i = 0; i < 16; i += 2) { unsigned a = __vimin3_u16x2(x[i], x[i + 1], x[(i + 2) & 15]), b = __vimax3_u16x2(x[i], x[i + 1], x[(i + 3) & 15]); x[i] = b; x[i + 1] = a; }
        }
    }
    const long long t1 = clock64();
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run(const char *name, int per_iter) {
    unsigned *out; long long *clk;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 148 * 8);
    const int iters = 4096;
    for (int warps = 4; warps <= 32; warps *= 2) {
        k<MODE><<<148, warps * 32>>>(16, 1, out, clk);
        k<MODE><<<148, warps * 32>>>(iters, 1, out, clk);
        cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
        double c = 0; for (int i = 0; i < 148; i++) c += h[i]; c /= 148;
        printf("%-44s warps/SM %2d : %6.3f compare-exchange/clk/SM (x2 samples)  %6.3f warp-instr/clk/SM\n", name, warps,
               8.0 * iters * warps / c, (double)per_iter * iters * warps / c);
    }
    cudaFree(out); cudaFree(clk);
}

int main() {
    run<0>("VIMNMX.U16x2", 16);
    run<1>("HMNMX2", 16);
    run<2>("VIMNMX.U16x2 + HMNMX2 (half each)", 16);
    run<3>("HFMA2.RELU + 2 HADD2 per CE", 32);
    run<4>("half VIMNMX, half HFMA2.RELU form", 28);
    run<5>("VIMNMX3.U16x2", 16);
    return 0;
}
