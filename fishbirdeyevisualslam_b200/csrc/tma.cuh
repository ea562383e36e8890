// TMA (cp.async.bulk.tensor) + mbarrier helpers for the stencil kernels: one elected thread asks the copy engine for a
// 3-D box (x, y, image slot) of a padded pyramid level; the tile lands densely in shared memory and completes on an
// mbarrier.  Out-of-range box parts (negative x on the first tile column, right/bottom overhang) are zero-filled by
// the hardware, so the kernels need no edge cases.  SASS: UTMALDG / SYNCS.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace fbe {

struct __align__(64) TmaMaps { CUtensorMap m[FBE_MAX_LEVELS]; };   // one map per pyramid level (passed __grid_constant__)

// host: u8 tensor [nslots][rows][pitch] at `base` (slot stride `slot_bytes`), box = box_w x box_h x 1
int tma_encode_level(CUtensorMap* out, const void* base, int pitch, int rows, int nslots, size_t slot_bytes, int box_w, int box_h);

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}
#endif

}  // namespace fbe
