// tools/probe2.cu -- issue-rate probes by OPERAND FORM (register / uniform-register / immediate)
// and for the integer instructions an exact-integer resample would use (IMAD, IDP4A, IADD3,
// LOP3, VIMNMX) plus mixes across the fma and alu pipes.  Not product code.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe2 probe2.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)
typedef unsigned long long f2;
#define ITER 512
#define REP 8
#define R8(M) M(0) M(1) M(2) M(3) M(4) M(5) M(6) M(7)

template <int OP>
__global__ void __launch_bounds__(1024) tput(float *out, float a, float b, int ia, int ib, unsigned long long *cyc) {
    float x[8]; f2 p[8]; unsigned u[8];
    for (int i = 0; i < 8; i++) {
        x[i] = threadIdx.x * 1e-3f + a + i;
        float lo = x[i], hi = x[i] + 0.5f;
        asm("mov.b64 %0, {%1,%2};" : "=l"(p[i]) : "f"(lo), "f"(hi));
        u[i] = __float_as_uint(x[i]) * 2654435761u + i;
    }
    f2 pa, pb; asm("mov.b64 %0, {%1,%1};" : "=l"(pa) : "f"(a)); asm("mov.b64 %0, {%1,%1};" : "=l"(pb) : "f"(b));
    // per-thread (non-uniform) copies so that the compiler cannot use uniform registers
    float ta = a + threadIdx.x * 1e-9f, tb = b + threadIdx.x * 1e-12f;
    f2 qa, qb; asm("mov.b64 %0, {%1,%1};" : "=l"(qa) : "f"(ta)); asm("mov.b64 %0, {%1,%1};" : "=l"(qb) : "f"(tb));
    f2 qc, qd; { float c0 = ta, c1 = ta * 1.0000002f, d0 = tb, d1 = tb * 3.f; asm("mov.b64 %0, {%1,%2};" : "=l"(qc) : "f"(c0), "f"(c1)); asm("mov.b64 %0, {%1,%2};" : "=l"(qd) : "f"(d0), "f"(d1)); }
    __shared__ unsigned sm[8192];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i * 2654435761u;
    __syncthreads();
    unsigned tia = ia + (threadIdx.x & 1), tib = ib + (threadIdx.x & 2);
    unsigned long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
      for (int rep = 0; rep < REP; rep++) {
        if (OP == 0) {          // FFMA R,R,R,R  (all per-thread registers)
#define M(i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(ta), "f"(tb));
            R8(M)
#undef M
        } else if (OP == 1) {   // FFMA R,R,UR,UR (uniform multiplier and addend)
#define M(i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
            R8(M)
#undef M
        } else if (OP == 2) {   // FFMA R,R,imm,R
#define M(i) asm volatile("fma.rn.f32 %0, %0, 0f3F800001, %1;" : "+f"(x[i]) : "f"(tb));
            R8(M)
#undef M
        } else if (OP == 3) {   // FFMA2 R,R,R,R all 64-bit per-thread registers
#define M(i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(qa), "l"(qb));
            R8(M)
#undef M
        } else if (OP == 4) {   // FFMA2 R,R,U,U (uniform)
#define M(i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
            R8(M)
#undef M
        } else if (OP == 5) {   // FFMA2 chain form: acc = fma2(w_uniform, p_reg, acc)  (h-pass shape)
#define M(i) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(pa), "l"(p[(i + 1) & 7]));
            R8(M)
#undef M
        } else if (OP == 6) {   // FADD2 R,R,imm-ish uniform
#define M(i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
            R8(M)
#undef M
        } else if (OP == 7) {   // FFMA.SAT R,R,UR,R(same)
#define M(i) asm volatile("fma.rn.sat.f32 %0, %0, %1, %0;" : "+f"(x[i]) : "f"(b));
            R8(M)
#undef M
        } else if (OP == 8) {   // IMAD R,R,R,R
#define M(i) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(tia), "r"(tib));
            R8(M)
#undef M
        } else if (OP == 9) {   // IMAD R,R,U,R  (acc += h * w_uniform)
#define M(i) asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]), "r"(ia));
            R8(M)
#undef M
        } else if (OP == 10) {  // IDP4A R,R,U,R
#define M(i) asm volatile("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]), "r"(ia));
            R8(M)
#undef M
        } else if (OP == 11) {  // IDP4A R,R,R,R
#define M(i) asm volatile("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]), "r"(tia));
            R8(M)
#undef M
        } else if (OP == 12) {  // IADD3 R,R,R,R
#define M(i) asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(u[i]) : "r"(u[(i + 1) & 7]), "r"(tia));
            R8(M)
#undef M
        } else if (OP == 13) {  // LOP3
#define M(i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]), "r"(tia));
            R8(M)
#undef M
        } else if (OP == 14) {  // PRMT R,R,imm,R
#define M(i) asm volatile("prmt.b32 %0, %0, %1, 0x4321;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
            R8(M)
#undef M
        } else if (OP == 15) {  // VIMNMX relu clamp: min.s32.relu
#define M(i) asm volatile("min.relu.s32 %0, %0, %1;" : "+r"(u[i]) : "r"(tia));
            R8(M)
#undef M
        } else if (OP == 16) {  // mix: 4 FFMA2(U,U) + 4 PRMT
#define MA(i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
#define MB(i) asm volatile("prmt.b32 %0, %0, %1, 0x4321;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
            MA(0) MB(0) MA(1) MB(1) MA(2) MB(2) MA(3) MB(3)
        } else if (OP == 17) {  // mix: 4 FFMA2 chain-form + 4 PRMT
#define MC(i) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(pa), "l"(p[(i + 1) & 7]));
            MC(0) MB(0) MC(1) MB(1) MC(2) MB(2) MC(3) MB(3)
        } else if (OP == 18) {  // mix: 4 IDP4A(U) + 4 PRMT
#define MD(i) asm volatile("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(u[i]) : "r"(u[(i + 1) & 3]), "r"(ia));
#define ME(i) asm volatile("prmt.b32 %0, %0, %1, 0x4321;" : "+r"(u[i]) : "r"(u[4 + ((i + 1) & 3)]));
            MD(0) ME(4) MD(1) ME(5) MD(2) ME(6) MD(3) ME(7)
        } else if (OP == 19) {  // mix: 4 FFMA2(U,U) + 4 IMAD(U)
#define MF(i) asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]), "r"(ia));
            MA(0) MF(0) MA(1) MF(1) MA(2) MF(2) MA(3) MF(3)
        } else if (OP == 20) {  // mix: 4 FFMA(U,U) scalar + 4 PRMT
#define MG(i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
            MG(0) MB(0) MG(1) MB(1) MG(2) MB(2) MG(3) MB(3)
        } else if (OP == 21) {  // mix: 2 FFMA2(U,U) + 2 PRMT + 2 IADD3 + 2 FFMA2 chain
#define MH(i) asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(u[i]) : "r"(u[(i + 1) & 7]), "r"(tia));
            MA(0) MB(0) MH(1) MC(1) MA(2) MB(2) MH(3) MC(3)
        } else if (OP == 22) {  // FMUL2.RZ R,R,U
#define M(i) asm volatile("mul.rz.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pa));
            R8(M)
#undef M
        } else if (OP == 23) {  // I2IP R,R,R,R
#define M(i) asm volatile("cvt.pack.sat.u8.s32.b32 %0, %1, %0, %2;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]), "r"(u[(i + 2) & 7]));
            R8(M)
#undef M
        } else if (OP == 24) {  // SHF (funnel shift) R,R,imm,R
#define M(i) asm volatile("shf.r.wrap.b32 %0, %0, %1, 8;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
            R8(M)
#undef M
        } else if (OP == 26) {  // FFMA2 with three true 64-bit register operands
#define M(i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(qc), "l"(qd));
            R8(M)
#undef M
        } else if (OP == 27) {  // mix 4 FFMA2(U,U) + 4 FFMA.SAT
#define MS(i) asm volatile("fma.rn.sat.f32 %0, %0, %1, %0;" : "+f"(x[i]) : "f"(b));
            MA(0) MS(0) MA(1) MS(1) MA(2) MS(2) MA(3) MS(3)
        } else if (OP == 28) {  // LDS.32 conflict-free (lane-private column), address from data
#define ML(i) { u[i] = sm[((u[i] & 0xff) << 5) | lane]; }
            ML(0) ML(1) ML(2) ML(3) ML(4) ML(5) ML(6) ML(7)
        } else if (OP == 29) {  // LDS.32 random addresses (bank conflicts)
#define MR(i) { u[i] = sm[u[i] & 0x1fff]; }
            MR(0) MR(1) MR(2) MR(3) MR(4) MR(5) MR(6) MR(7)
        } else if (OP == 30) {  // mix 6 FFMA2(U,U) + 2 LDS conflict-free
            MA(0) MA(1) MA(2) ML(0) MA(3) MA(4) MA(5) ML(1)
        } else if (OP == 31) {  // alternating FFMA2(U,U) / FFMA(U,U)
            MA(0) MG(0) MA(1) MG(1) MA(2) MG(2) MA(3) MG(3)
        } else if (OP == 32) {  // grouped 4 FFMA2 then 4 FFMA
            MA(0) MA(1) MA(2) MA(3) MG(0) MG(1) MG(2) MG(3)
        } else if (OP == 33) {  // grouped 8 FFMA2 then 8 FFMA (16 instr per rep)
            MA(0) MA(1) MA(2) MA(3) MA(4) MA(5) MA(6) MA(7) MG(0) MG(1) MG(2) MG(3) MG(4) MG(5) MG(6) MG(7)
        } else if (OP == 34) {  // 6 FFMA2 + 2 FFMA
            MA(0) MA(1) MA(2) MG(0) MA(3) MA(4) MA(5) MG(1)
        } else if (OP == 35) {  // 2 FFMA2 + 6 FFMA
            MA(0) MG(0) MG(1) MG(2) MA(1) MG(3) MG(4) MG(5)
        } else if (OP == 25) {  // mix: 4 IMAD(U) + 4 IADD3
            MF(0) MH(4) MF(1) MH(5) MF(2) MH(6) MF(3) MH(7)
        }
      }
    }
    unsigned long long t1 = clock64();
    float s = 0; unsigned long long q = 0; unsigned v = 0;
    for (int i = 0; i < 8; i++) { s += x[i]; q ^= p[i]; v ^= u[i]; }
    if (s + (float)q + (float)v == 12345.678f) out[0] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
int run(const char *name) {
    float *out; unsigned long long *cyc;
    CK(cudaMalloc(&out, 4)); CK(cudaMalloc(&cyc, 8 * 148));
    printf("%-44s", name);
    for (int warps = 8; warps <= 32; warps *= 2) {
        tput<OP><<<148, warps * 32>>>(out, 1.0000001f, 1e-9f, 19, -3, cyc);
        CK(cudaDeviceSynchronize());
        std::vector<unsigned long long> h(148);
        CK(cudaMemcpy(h.data(), cyc, 8 * 148, cudaMemcpyDeviceToHost));
        double avg = 0; for (auto v : h) avg += v; avg /= 148;
        printf("  w%-2d %.3f", warps, (double)ITER * REP * 8 * warps / avg);
    }
    printf("   warp-instr/clk/SM\n");
    cudaFree(out); cudaFree(cyc);
    return 0;
}

int main() {
    run<0>("FFMA R,R,R,R");
    run<1>("FFMA R,R,U,U");
    run<2>("FFMA R,R,imm,R");
    run<3>("FFMA2 R,R,R,R");
    run<4>("FFMA2 R,R,U,U");
    run<5>("FFMA2 acc=fma2(U,R,acc)");
    run<6>("FADD2 R,R,U");
    run<22>("FMUL2.RZ R,R,U");
    run<7>("FFMA.SAT R,R,U,R");
    run<8>("IMAD R,R,R,R");
    run<9>("IMAD acc=R*U+acc");
    run<10>("IDP4A acc=dp4a(R,U,acc)");
    run<11>("IDP4A acc=dp4a(R,R,acc)");
    run<12>("IADD3 R,R,R");
    run<13>("LOP3 R,R,R");
    run<14>("PRMT R,R,imm,R");
    run<24>("SHF.R R,R,imm,R");
    run<23>("I2IP R,R,R,R");
    run<15>("VIMNMX.RELU");
    run<16>("mix 4 FFMA2(U,U) + 4 PRMT");
    run<17>("mix 4 FFMA2 chain + 4 PRMT");
    run<18>("mix 4 IDP4A(U) + 4 PRMT");
    run<19>("mix 4 FFMA2(U,U) + 4 IMAD(U)");
    run<20>("mix 4 FFMA(U,U) + 4 PRMT");
    run<21>("mix 2 FFMA2 + 2 PRMT + 2 IADD3 + 2 FFMA2ch");
    run<25>("mix 4 IMAD(U) + 4 IADD3");
    run<31>("alt 4 FFMA2 + 4 FFMA (12 units)");
    run<32>("grp 4 FFMA2 + 4 FFMA (12 units)");
    run<33>("grp 8 FFMA2 + 8 FFMA (24 units; x2 instr)");
    run<34>("6 FFMA2 + 2 FFMA (14 units)");
    run<35>("2 FFMA2 + 6 FFMA (10 units)");
    run<26>("FFMA2 R,R64,R64,R64");
    run<27>("mix 4 FFMA2(U,U) + 4 FFMA.SAT");
    run<28>("LDS.32 conflict-free data-dependent");
    run<29>("LDS.32 random (conflicts)");
    run<30>("mix 6 FFMA2(U,U) + 2 LDS cf");
    return 0;
}
