/* oracle/ref_o2_coeffs.cu -- TEST INFRASTRUCTURE ONLY (part of oracle "O2").
 *
 * The reference's own coefficient functions, libavfilter/vf_scale_cuda.cu:948-981 (lanczos_coeffs with its
 * fast-math __sinf, bicubic_coeffs), called from a two-line kernel: the reference file is #included where it
 * lies under /root/reference (nothing is copied), so the numbers are what its Subsample_* kernels use.
 * tests/golden/make_golden_lanczos.py runs this on a B200 to produce the committed Lanczos tables that pin the
 * CPU oracle; tests/test_gpu_scale.py compares our device tables against it live. */
#include "vf_scale_cuda.cu"

extern "C" __global__ void ref_coeffs_kernel(int lanczos, const float *x, float param, float4 *out, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = lanczos ? lanczos_coeffs(x[i], param) : bicubic_coeffs(x[i], param);
}

/* host arrays in, host arrays out; returns 0 or a cudaError_t */
extern "C" int ref_o2_coeffs(int lanczos, const float *x, float param, float *out4, int n)
{
    float *dx = 0; float4 *dout = 0;
    cudaError_t e = cudaMalloc(&dx, sizeof(float) * n);
    if (e == cudaSuccess) e = cudaMalloc(&dout, sizeof(float4) * n);
    if (e == cudaSuccess) e = cudaMemcpy(dx, x, sizeof(float) * n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        ref_coeffs_kernel<<<(n + 127) / 128, 128>>>(lanczos, dx, param, dout, n);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out4, dout, sizeof(float4) * n, cudaMemcpyDeviceToHost);
    cudaFree(dx); cudaFree(dout);
    return (int)e;
}
