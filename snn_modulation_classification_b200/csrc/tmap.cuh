// TMA tensor-map descriptors (tmap.cu) and the PTX of the tile loads that use them.
#pragma once
#include <stdint.h>

namespace dcll {

struct alignas(64) TmapDesc {
    unsigned char bytes[128];   // an opaque CUtensorMap; passed to kernels as a __grid_constant__ parameter
};
bool tmap_bf16(TmapDesc *out, const void *base, int rank, const uint64_t *dims, const uint64_t *strides, const uint32_t *box);

#ifdef __CUDACC__
namespace tc {
// cp.async.bulk.tensor (UTMALDG): one box of the tensor -> shared memory, bytes counted on the mbarrier; coordinates are signed,
// everything outside the tensor arrives as zeros.
__device__ __forceinline__ void tma_load_5d(uint32_t dst_smem, const TmapDesc *map, uint32_t bar_smem, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst_smem),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_smem), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
// CTA-pair form: the bytes are counted on the barrier at `bar_smem` in the pair's LEADER CTA (tc::PEER_BIT_MASK), whichever CTA issues
__device__ __forceinline__ void tma_load_5d_2cta(uint32_t dst_smem, const TmapDesc *map, uint32_t bar_smem, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst_smem),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_smem & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst_smem, const TmapDesc *map, uint32_t bar_smem, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst_smem),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_smem), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const TmapDesc *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
}  // namespace tc
#endif

}  // namespace dcll
