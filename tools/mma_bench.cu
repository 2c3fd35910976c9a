// tcgen05 issue-rate microbenchmark for the short MMAs of the DCLL conv / wgrad kernels (M=128, K=16, N=64/32, SWIZZLE_NONE).
// One CTA per SM; lane 0 of `issuers` warps issues `n` MMAs each, then commits; cycles are taken around issue+completion.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I snn_modulation_classification_b200/csrc -I include tools/mma_bench.cu -o tools/_build/mma_bench
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"

using namespace dcll::tc;

struct Cfg {
    int n;          // MMAs per issuer
    int n_acc;      // accumulators cycled through per issuer
    int pair;       // 1: alternate N=64 / N=32 (as the conv kernel), 0: all N=nsize
    int nsize;      // N of the MMAs when !pair
    int shift;      // vary the A start address per MMA (implicit im2col) or not
    int issuers;    // issuing warps (each its own accumulators)
    int commit_every;   // tcgen05.commit + wait every k MMAs (0 = only at the end)
    int swz;            // 0: SWIZZLE_NONE canonical layout, 2: SWIZZLE_128B, 4: 64B, 6: 32B (descriptor bits 61..63), K-major
    int a_mn;           // 1: A is MN-major (the weight-gradient kernels: M = (kernel row, channel) contiguous), SWIZZLE_NONE
    int ts;             // 1: A comes from tensor memory (TS mode: tcgen05.mma [d], [a_tmem], b_desc), B from shared memory
    int f16;            // 1: fp16 x fp16 operands instead of bf16 x bf16
    int b256;           // 1: B as the weight-gradient kernels stage it (N groups 256 B apart, the two K chunks 128 B apart)
};

template <int TS>
__global__ void __launch_bounds__(256, 1) bench(Cfg c, long long *out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar[8];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u + i;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) mbar_init(bar + i, 1);
        mbar_fence_init();
    }
    if (warp == 7) tmem_alloc(&slot, 512);
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = slot;
    long long t0 = 0, t1 = 0;
    if (warp < c.issuers) {
        const uint32_t elected = elect_one();
        const uint32_t a_base = desc_lo(smem_u32(smem), c.swz ? 16 : 7744), b_base = desc_lo(smem_u32(smem + 128 * 1024), c.swz ? 16 : 1024);
        const uint32_t swz = (uint32_t)c.swz << 29;
        const uint32_t sbo_sw = c.swz == 2 ? 1024 : (c.swz == 4 ? 512 : 256);
        const uint32_t A_HI = c.swz ? (desc_hi(sbo_sw) | swz) : desc_hi(352), B_HI = c.swz ? (desc_hi(sbo_sw) | swz) : desc_hi(128);
        const uint32_t i64 = idesc_bf16(128, 64, c.a_mn != 0, false), i32 = idesc_bf16(128, 32, c.a_mn != 0, false);
        const uint32_t iN = c.f16 ? idesc_f16a(128, c.nsize, c.a_mn != 0, false) : idesc_bf16(128, c.nsize, c.a_mn != 0, false);
        const uint32_t a_base_mn = desc_lo(smem_u32(smem), 128);      // MN-major: LBO = next 8 K rows, SBO (352) = next 8 M elements
        const uint32_t b_base256 = desc_lo(smem_u32(smem + 128 * 1024), 128);
        const uint32_t B_HI256 = desc_hi(256);
        __syncwarp();
        t0 = clock64();
        if (elected) {
            int ph = 0;
            const uint32_t dbase = tmem + warp * (512 / c.issuers);
            const uint32_t acc_mask = c.n_acc - 1, acc_cols = 64 * (c.nsize > 64 ? c.nsize / 64 : 1);
            const uint32_t sh_mask = c.shift ? 31u : 0u;
            const uint32_t lo_off = c.pair ? (32768 >> 4) : 0;
            const int ce = c.commit_every ? c.commit_every : (1 << 30);
            int since = 0;
#pragma unroll 1
            for (int k = 0; k < c.n; k += 8) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t kk = k + u;
                    const uint32_t d = dbase + (((kk >> 1) & acc_mask) * acc_cols);
                    const uint32_t sh = c.swz ? ((kk >> 1) & sh_mask & 3) * 2 : ((kk >> 1) & sh_mask);
                    const uint64_t a = desc(A_HI, (c.a_mn ? a_base_mn : a_base) + sh + ((u & 1) ? lo_off : 0));
                    const uint64_t b = c.b256 ? desc(B_HI256, b_base256 + (sh & 3) * 256) : desc(B_HI, b_base + (c.swz ? (sh & 3) * 2 : (sh & 7) * 256));
                    const uint32_t id = c.pair ? ((u & 1) ? i32 : i64) : iN;
                    if (TS) {
                        // A = 128 lanes x 8 columns (16 bf16) of tensor memory, columns 448.. (never written: timing only)
                        asm volatile(
                            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
                            "r"(tmem + 448 + 8 * (kk & 7)), "l"(b), "r"(id), "r"(1u)
                            : "memory");
                    } else {
                        mma_bf16(d, a, b, id, 1);
                    }
                }
                since += 8;
                if (since >= ce) {
                    since = 0;
                    commit(bar + warp);
                    mbar_wait(bar + warp, ph);
                    ph ^= 1;
                }
            }
            commit(bar + 4 + warp);
            mbar_wait(bar + 4 + warp, 0);
        }
        __syncwarp();
        t1 = clock64();
    }
    if (lane == 0 && warp < c.issuers && blockIdx.x == 0) out[warp] = t1 - t0;
    fence_before();
    __syncthreads();
    if (warp == 7) tmem_dealloc(tmem, 512);
}

// ---- CTA-pair variant: tcgen05.mma.cta_group::2, M = 256 (128 rows of A from each CTA's shared memory), each CTA holds half of B.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1) bench2(Cfg c, long long *out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar[2];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u + i;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1), mbar_init(bar + 1, 1);
        mbar_fence_init();
    }
    if (warp == 7) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    fence_before();
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    fence_after();
    const uint32_t tmem = slot;
    long long t0 = 0, t1 = 0;
    if (warp == 0) {
        const uint32_t elected = elect_one();
        const uint32_t a_base = desc_lo(smem_u32(smem), 7744), b_base = desc_lo(smem_u32(smem + 128 * 1024), 1024);
        constexpr uint32_t A_HI = desc_hi(352), B_HI = desc_hi(128);
        const uint32_t i64 = idesc_bf16(256, 64, c.a_mn != 0, false), i32 = idesc_bf16(256, 32, c.a_mn != 0, false);
        const uint32_t iN = idesc_bf16(256, c.nsize, c.a_mn != 0, false);
        const uint32_t a_base_mn = desc_lo(smem_u32(smem), 128);
        __syncwarp();
        t0 = clock64();
        if (elected && rank == 0) {
            const uint32_t sh_mask = c.shift ? 31u : 0u;
            const uint32_t lo_off = c.pair ? (32768 >> 4) : 0;
#pragma unroll 1
            for (int k = 0; k < c.n; k += 8) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t kk = k + u;
                    const uint32_t sh = (kk >> 1) & sh_mask;
                    const uint64_t a = desc(A_HI, (c.a_mn ? a_base_mn : a_base) + sh + ((u & 1) ? lo_off : 0));
                    const uint64_t b = desc(B_HI, b_base + (sh & 7) * 256);
                    const uint32_t id = c.pair ? ((u & 1) ? i32 : i64) : iN;
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
                        "l"(a), "l"(b), "r"(id), "r"(1u)
                        : "memory");
                }
            }
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                         "h"((uint16_t)3)
                         : "memory");
        }
        __syncwarp();
        mbar_wait(bar, 0);          // both CTAs: the multicast commit arrives on each CTA's barrier
        t1 = clock64();
    }
    if (lane == 0 && warp == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    fence_before();
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 7) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
    long long *out;
    cudaMallocManaged(&out, 64);
    cudaFuncSetAttribute(bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const Cfg cfgs[] = {
        // n, n_acc, pair, nsize, shift, issuers, commit_every, swz
        {4096, 1, 0, 64, 1, 1, 0, 0}, {4096, 1, 0, 32, 1, 1, 0, 0}, {4096, 1, 0, 128, 1, 1, 0, 0}, {4096, 1, 0, 256, 1, 1, 0, 0},
        {4096, 1, 1, 64, 1, 1, 0, 0}, {4096, 1, 1, 64, 1, 2, 0, 0},
        {4096, 1, 0, 64, 0, 1, 0, 2}, {4096, 1, 0, 32, 0, 1, 0, 2}, {4096, 1, 0, 128, 0, 1, 0, 2}, {4096, 1, 0, 64, 1, 1, 0, 2},
        {4096, 1, 0, 32, 1, 1, 0, 2}, {4096, 1, 1, 64, 1, 1, 0, 2},
        {4096, 1, 0, 64, 0, 1, 0, 4}, {4096, 1, 0, 32, 0, 1, 0, 4}, {4096, 1, 0, 64, 1, 1, 0, 4}, {4096, 1, 1, 64, 1, 1, 0, 4},
        {4096, 1, 0, 64, 0, 1, 0, 6}, {4096, 1, 0, 32, 0, 1, 0, 6},
    };
    printf("%6s %5s %4s %5s %5s %7s %6s %3s | cycles/MMA (issuer 0)\n", "n", "n_acc", "pair", "N", "shift", "issuers", "commit", "swz");
    for (const Cfg &c : cfgs) {
        for (int rep = 0; rep < 2; ++rep) {
            bench<0><<<148, 256, 200 * 1024>>>(c, out);
            if (cudaDeviceSynchronize() != cudaSuccess) {
                printf("CUDA error: %s\n", cudaGetErrorString(cudaGetLastError()));
                return 1;
            }
        }
        printf("%6d %5d %4d %5d %5d %7d %6d %3d | %.1f\n", c.n, c.n_acc, c.pair, c.nsize, c.shift, c.issuers, c.commit_every, c.swz,
               (double)out[0] / c.n);
    }
    printf("\nA MN-major (weight gradient: M = (kernel row, channel) contiguous, K = positions), SWIZZLE_NONE\n");
    const Cfg cmn[] = {{4096, 1, 0, 32, 1, 1, 0, 0, 1, 0}, {4096, 1, 0, 64, 1, 1, 0, 0, 1, 0}, {4096, 1, 0, 128, 1, 1, 0, 0, 1, 0},
                       {4096, 1, 0, 256, 1, 1, 0, 0, 1, 0}, {4096, 1, 0, 64, 1, 2, 0, 0, 1, 0}, {4096, 1, 0, 128, 1, 2, 0, 0, 1, 0}};
    for (const Cfg &c : cmn) {
        for (int rep = 0; rep < 2; ++rep) {
            bench<0><<<148, 256, 200 * 1024>>>(c, out);
            if (cudaDeviceSynchronize() != cudaSuccess) {
                printf("CUDA error: %s\n", cudaGetErrorString(cudaGetLastError()));
                return 1;
            }
        }
        printf("A MN-major N=%3d issuers=%d | %.1f cycles per MMA (issuer 0)\n", c.nsize, c.issuers, (double)out[0] / c.n);
    }
    printf("\nAccumulator rotation: consecutive MMA pairs go to n_acc different accumulators (N = 128: 128 columns each)\n");
    const Cfg cacc[] = {{4096, 1, 0, 128, 1, 1, 0, 0, 0, 0}, {4096, 2, 0, 128, 1, 1, 0, 0, 0, 0}, {4096, 4, 0, 128, 1, 1, 0, 0, 0, 0},
                        {4096, 4, 0, 128, 1, 1, 0, 0, 1, 0}, {4096, 2, 0, 64, 1, 1, 0, 0, 0, 0}, {4096, 4, 0, 64, 1, 1, 0, 0, 0, 0},
                        {4096, 8, 0, 64, 1, 1, 0, 0, 0, 0}};
    for (const Cfg &c : cacc) {
        for (int rep = 0; rep < 2; ++rep) {
            bench<0><<<148, 256, 200 * 1024>>>(c, out);
            if (cudaDeviceSynchronize() != cudaSuccess) {
                printf("CUDA error: %s\n", cudaGetErrorString(cudaGetLastError()));
                return 1;
            }
        }
        printf("n_acc=%d N=%3d A %s-major | %.1f cycles per MMA\n", c.n_acc, c.nsize, c.a_mn ? "MN" : "K", (double)out[0] / c.n);
    }
    printf("\nfp16 operands / weight-gradient B layout (N = 128, A MN-major)\n");
    const Cfg cf[] = {{4096, 4, 0, 128, 1, 1, 0, 0, 1, 0, 0, 0}, {4096, 4, 0, 128, 1, 1, 0, 0, 1, 0, 1, 0}, {4096, 4, 0, 128, 1, 1, 0, 0, 1, 0, 0, 1},
                      {4096, 4, 0, 128, 1, 1, 0, 0, 1, 0, 1, 1}, {4096, 2, 0, 128, 1, 2, 0, 0, 1, 0, 1, 1}, {4096, 4, 0, 128, 1, 1, 0, 0, 0, 0, 1, 0}};
    for (const Cfg &c : cf) {
        for (int rep = 0; rep < 2; ++rep) {
            bench<0><<<148, 256, 200 * 1024>>>(c, out);
            if (cudaDeviceSynchronize() != cudaSuccess) {
                printf("CUDA error: %s\n", cudaGetErrorString(cudaGetLastError()));
                return 1;
            }
        }
        printf("f16=%d b256=%d A %s-major issuers=%d | %.1f cycles per MMA\n", c.f16, c.b256, c.a_mn ? "MN" : "K", c.issuers, (double)out[0] / c.n);
    }
    printf("\nTS mode (A in tensor memory, B = N x 16 bf16 from shared memory)\n");
    const Cfg cts[] = {{4096, 1, 0, 32, 1, 1, 0, 0, 0, 1}, {4096, 1, 0, 64, 1, 1, 0, 0, 0, 1}, {4096, 1, 0, 128, 1, 1, 0, 0, 0, 1},
                       {4096, 1, 0, 256, 1, 1, 0, 0, 0, 1}};
    for (const Cfg &c : cts) {
        for (int rep = 0; rep < 2; ++rep) {
            bench<1><<<148, 256, 200 * 1024>>>(c, out);
            if (cudaDeviceSynchronize() != cudaSuccess) {
                printf("CUDA error: %s\n", cudaGetErrorString(cudaGetLastError()));
                return 1;
            }
        }
        printf("TS N=%3d | %.1f cycles per MMA\n", c.nsize, (double)out[0] / c.n);
    }
    cudaFuncSetAttribute(bench2, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    printf("\nCTA pairs (cta_group::2, M = 256 = 128 rows per CTA, half of B per CTA), 148 CTAs = 74 pairs\n");
    const Cfg c2[] = {{4096, 1, 0, 64, 1, 1, 0, 0}, {4096, 1, 0, 32, 1, 1, 0, 0}, {4096, 1, 0, 128, 1, 1, 0, 0}, {4096, 1, 1, 64, 1, 1, 0, 0},
                      {4096, 1, 0, 256, 1, 1, 0, 0}, {4096, 1, 0, 64, 1, 1, 0, 0, 1, 0}, {4096, 1, 0, 128, 1, 1, 0, 0, 1, 0}};
    for (const Cfg &c : c2) {
        for (int rep = 0; rep < 2; ++rep) {
            bench2<<<148, 256, 200 * 1024>>>(c, out);
            if (cudaDeviceSynchronize() != cudaSuccess) {
                printf("CUDA error: %s\n", cudaGetErrorString(cudaGetLastError()));
                return 1;
            }
        }
        printf("pair=%d N=%3d A %s-major | %.1f cycles per M=256 MMA\n", c.pair, c.nsize, c.a_mn ? "MN" : "K", (double)out[0] / c.n);
    }
    return 0;
}
