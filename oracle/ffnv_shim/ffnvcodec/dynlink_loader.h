/* oracle/ffnv_shim/ffnvcodec/dynlink_loader.h -- TEST INFRASTRUCTURE ONLY.
 *
 * A stand-in for nv-codec-headers' <ffnvcodec/dynlink_loader.h> (not in the reference tree, not in this image;
 * compat/cuda/dynlink_loader.h:33 includes it) that lets the reference's own libavutil/hwcontext_cuda.c compile
 * and run in the AVFilter harness (oracle/refbuild `avf`): `CudaFunctions` holds the driver entry points
 * hwcontext_cuda.c calls through `cu->`, filled from the libcuda we link instead of dlopen()ing it.
 * Member names go through <cuda.h>'s versioning macros on both sides (cu->cuMemAlloc is cu->cuMemAlloc_v2 in the
 * struct and at the call site alike). */
#ifndef GMATB_FFNV_SHIM_DYNLINK_LOADER_H
#define GMATB_FFNV_SHIM_DYNLINK_LOADER_H

#include <cuda.h>
#include <stdlib.h>

#ifndef CUDAAPI
#define CUDAAPI
#endif

#define GMATB_FFNV_FUNCS(X) \
    X(cuInit) X(cuDeviceGet) X(cuDeviceGetCount) X(cuDeviceGetAttribute) X(cuDeviceGetUuid) X(cuDeviceGetName) \
    X(cuDevicePrimaryCtxGetState) X(cuDevicePrimaryCtxSetFlags) X(cuDevicePrimaryCtxRetain) X(cuDevicePrimaryCtxRelease) \
    X(cuCtxCreate) X(cuCtxDestroy) X(cuCtxPushCurrent) X(cuCtxPopCurrent) X(cuCtxSetLimit) \
    X(cuMemAlloc) X(cuMemAllocPitch) X(cuMemFree) X(cuMemcpy) X(cuMemcpyAsync) X(cuMemcpy2D) X(cuMemcpy2DAsync) \
    X(cuMemsetD8Async) X(cuStreamCreate) X(cuStreamDestroy) X(cuStreamSynchronize) X(cuStreamQuery) \
    X(cuEventCreate) X(cuEventDestroy) X(cuEventRecord) X(cuEventSynchronize) X(cuEventQuery) X(cuStreamWaitEvent) \
    X(cuGetErrorName) X(cuGetErrorString) \
    X(cuModuleLoadData) X(cuModuleUnload) X(cuModuleGetFunction) X(cuModuleGetGlobal) X(cuLaunchKernel) \
    X(cuTexObjectCreate) X(cuTexObjectDestroy)

typedef struct CudaFunctions {
#define GMATB_FFNV_MEMBER(f) __typeof__(f) *f;
    GMATB_FFNV_FUNCS(GMATB_FFNV_MEMBER)
#undef GMATB_FFNV_MEMBER
} CudaFunctions;

static inline void cuda_free_functions(CudaFunctions **pf)
{
    if (pf && *pf) { free(*pf); *pf = NULL; }
}

static inline int cuda_load_functions(CudaFunctions **pf, void *logctx)
{
    CudaFunctions *f = (CudaFunctions *)calloc(1, sizeof(*f));
    (void)logctx;
    if (!f) return -1;
#define GMATB_FFNV_ASSIGN(fn) f->fn = &fn;
    GMATB_FFNV_FUNCS(GMATB_FFNV_ASSIGN)
#undef GMATB_FFNV_ASSIGN
    *pf = f;
    return 0;
}

#endif
