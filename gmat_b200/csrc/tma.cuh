// tma.cuh -- Tensor Memory Accelerator plumbing (sm_90+; here sm_100a): cp.async.bulk.tensor tile loads into
// shared memory completing on an mbarrier, tile stores from shared memory, and the host-side tensor-map encoder
// (cuTensorMapEncodeTiled, reached through cudaGetDriverEntryPoint so that the library keeps linking against the
// runtime only).  SASS: UTMALDG / UTMASTG / SYNCS.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace gmatb {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");      // visible to the async proxy before a TMA names it
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// tile load: box (fixed in the tensor map) whose first element is at coordinates (x, y[, z]); out-of-range
// elements arrive as zeros; completion is signalled on `bar` as transaction bytes
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int x, int y, int z) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}
// tile store (the part of the box that lies inside the tensor): call after the writers' fence_async_smem() + barrier
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, const void *smem_src, int x, int y, int z) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(map), "r"(smem_u32(smem_src)), "r"(x), "r"(y), "r"(z) : "memory");
}
__device__ __forceinline__ void tma_store_commit_wait() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// generic-proxy writes to shared memory -> visible to the async proxy (the TMA engine)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// host: a rank-3 map (x bytes / elements, rows, frames) over a pitched plane; returns false when the plane cannot
// be described (alignment) so that the caller can take its non-TMA path
bool make_tensor_map_3d(CUtensorMap *out, CUtensorMapDataType dtype, int elem_bytes, const void *base, unsigned long long dim_x,
                        unsigned long long dim_y, unsigned long long dim_z, unsigned long long pitch_bytes, unsigned long long frame_bytes,
                        unsigned box_x, unsigned box_y);

}  // namespace gmatb
