// Tensor-map (TMA) descriptors for the tile loads of the tensor-core kernels.  The library links cudart only: the driver's
// cuTensorMapEncodeTiled is fetched through cudaGetDriverEntryPoint.  A descriptor depends only on the buffer and its
// geometry, so the few a network needs are kept in a small cache (encoding costs microseconds of host time per call).
#include <cuda.h>
#include <string.h>

#include "common.cuh"
#include "tmap.cuh"

namespace dcll {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        (void)cudaGetLastError();
    }
    return fn;
}

struct TmapKey {
    const void *base;
    int rank;
    uint64_t dims[5], strides[4];
    uint32_t box[5];
};
static constexpr int TMAP_CACHE = 32;
static TmapKey g_keys[TMAP_CACHE];
static TmapDesc g_maps[TMAP_CACHE];
static int g_n = 0, g_next = 0;

// bf16 tensor, dims[0] contiguous; strides[i] = byte stride of dims[i+1] (any order: the dimension order only fixes the order of
// the box in shared memory); zero fill outside the tensor; no swizzle.  Returns false when the driver refuses (the callers then
// fall back to their cp.async loaders).
bool tmap_bf16(TmapDesc *out, const void *base, int rank, const uint64_t *dims, const uint64_t *strides, const uint32_t *box) {
    TmapKey k;
    memset(&k, 0, sizeof(k));
    k.base = base, k.rank = rank;
    for (int i = 0; i < rank; ++i) k.dims[i] = dims[i], k.box[i] = box[i];
    for (int i = 0; i + 1 < rank; ++i) k.strides[i] = strides[i];
    for (int i = 0; i < g_n; ++i)
        if (memcmp(&g_keys[i], &k, sizeof(k)) == 0) {
            *out = g_maps[i];
            return true;
        }
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    alignas(64) CUtensorMap m;
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    cuuint64_t gd[5], gs[4];
    cuuint32_t bx[5];
    for (int i = 0; i < rank; ++i) gd[i] = dims[i], bx[i] = box[i];
    for (int i = 0; i + 1 < rank; ++i) gs[i] = strides[i];
    CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void *>(base), gd, gs, bx, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
    static_assert(sizeof(TmapDesc) == sizeof(CUtensorMap), "descriptor size");
    memcpy(out, &m, sizeof(m));
    const int slot = g_n < TMAP_CACHE ? g_n++ : (g_next++ % TMAP_CACHE);
    g_keys[slot] = k, g_maps[slot] = *out;
    return true;
}

}  // namespace dcll
