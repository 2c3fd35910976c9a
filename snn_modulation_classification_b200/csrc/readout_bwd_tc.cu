// Backward local read-out on tcgen05 (F16X2 layers whose g_u leaves as the fp16 {hi,lo} operand image of wgrad_tc2p):
//
//     g_u[b,f] = (sum_k g_o[b,k] Wo[k,f]) * (1 - pv[b,f]) * pv[b,f]          (reference dcll/pytorch_libdcll.py:690-704)
//
// readout_bwd2_kernel (readout.cu) spends 24 FFMA2 + 6 LDS.128 per sample and feature pair on the K = 24 sum and is
// issue / latency bound at 0.57 of the HBM roofline.  Here the sum is ONE skinny GEMM per CTA,
//     D[f, b] = sum_k Wo[k,f] g_o[b,k]      M = 128 features, N = 64 samples, K = 32 (padded),
// in split bf16 with the three products concatenated along K (K = 96 = [hi|hi|lo] x [hi|lo|hi], six MMAs), so that the
// threads only stream: TMEM -> registers, pv in, image out.  A CTA's 128 features are 4 adjacent channels x 32 consecutive
// positions (warp = channel, lane = position): pv loads are 128-byte rows, and the image words of two adjacent channels are
// neighbours, so the CTA's 4-byte stores fill whole sectors.  The two lanes of a position pair exchange their values; the
// even lane stores the hi word of the pair, the odd lane the lo word.
// Numerics: the K = 24 sum carries ~2^-16 relative rounding (bf16 x3) instead of fp32's 2^-24 -- F16X2 only, whose g_u image
// is fp16 {hi,lo} anyway and whose tests bound the first Adam step statistically (tests/util_build.py).
// EXPERIMENTAL, off unless DCLL_RB_TC=1: correct (tests/test_gpu_f16x2.py runs the f16x2 bounds with the switch on) but
// measured 0.102 ms against 0.085 ms for readout_bwd2_kernel at 128x128, B = 64 -- 4096 short-lived CTAs pay TMEM allocation,
// a Wo round trip, the MMA round trip and the pv round trips in series, and store 16-byte runs.  DESIGN.md section 4.4 (g).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace dcll {

struct RbTcP {
    const float *pv, *wo, *g_o;
    float *g_u;
    int B, F, K, hw;
    float g_scale;
};

namespace rbtc {
constexpr int NT = 128, KP = 32, KG = 3 * KP / 8;          // 12 groups of 8 along the concatenated K
constexpr int A_F8 = 128 * 16, B_F8 = 64 * 16;             // bytes per K group (rows x 16 B)
constexpr int A_BYTES = KG * A_F8, B_BYTES = KG * B_F8;

__device__ __forceinline__ uint32_t bf16x2_hi_lo(float a, float b, uint32_t &lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    const uint32_t hb = *reinterpret_cast<const uint32_t *>(&h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(a - __uint_as_float(hb << 16), b - __uint_as_float(hb & 0xffff0000u));
    lo = *reinterpret_cast<const uint32_t *>(&l);
    return hb;
}
// one row (32 values along K) -> its 16-byte pieces in the three K blocks; blk1 / blk2 select {hi, lo} for the second / third block
__device__ __forceinline__ void store_row(unsigned char *base, int f8, int row, const float (&v)[KP], bool second_lo) {
#pragma unroll
    for (int g = 0; g < KP / 8; ++g) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) hi[j] = bf16x2_hi_lo(v[8 * g + 2 * j], v[8 * g + 2 * j + 1], lo[j]);
        const uint4 h4 = make_uint4(hi[0], hi[1], hi[2], hi[3]), l4 = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        unsigned char *d = base + row * 16;
        *reinterpret_cast<uint4 *>(d + g * f8) = h4;                                       // block 0: hi
        *reinterpret_cast<uint4 *>(d + (KP / 8 + g) * f8) = second_lo ? l4 : h4;           // block 1: A hi, B lo
        *reinterpret_cast<uint4 *>(d + (2 * KP / 8 + g) * f8) = second_lo ? h4 : l4;       // block 2: A lo, B hi
    }
}
}  // namespace rbtc

__global__ void __launch_bounds__(rbtc::NT, 5) readout_bwd_tc_kernel(const RbTcP p) {
    using namespace rbtc;
    using namespace tc;
    __shared__ __align__(128) unsigned char sA[A_BYTES];
    __shared__ __align__(128) unsigned char sB[B_BYTES];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // tile -> (channel group of 8, block of 32 positions, half of the group): 4 adjacent channels x 32 positions
    const int pblks = p.hw >> 5, chunks = p.hw >> 3;
    const int chalf = blockIdx.x & 1, t = blockIdx.x >> 1;
    const int pblk = t % pblks, cog = t / pblks;
    const int co8 = chalf * 4 + warp, pos = pblk * 32 + lane;
    const int f = (cog * 8 + co8) * p.hw + pos;
    const int img_word = ((cog * chunks + (pos >> 3)) * 8 + co8) * 4 + ((pos & 7) >> 1);   // 32-bit word of the position pair in one part
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        tmem_alloc(&tmem_slot, 64u);
        // several CTAs share an SM here: hand the allocation permit back at once, or the next CTA's tcgen05.alloc waits until
        // this CTA has exited (measured without it: the CTAs of an SM ran one after the other, 0.178 ms per launch)
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = tmem_slot;
    pdl_entry();

    // ---- A: this thread's Wo column (row m = tid of the tile), K-major [k/8][m][8].  The g_o rows of the first batch block are
    //      requested together with it (one round trip instead of two); loads are unconditional (clamped) so that all fly at once.
    constexpr uint32_t IDESC = idesc_bf16(128, 64, false, false);
    constexpr uint32_t SBO128 = desc_hi(128);
    const bool even = (lane & 1) == 0;
    auto load_g = [&](int b0, int nb, float (&g)[KP]) {
#pragma unroll
        for (int k = 0; k < KP; ++k) g[k] = __ldg(p.g_o + (size_t)(b0 + min(tid & 63, nb - 1)) * p.K + min(k, p.K - 1));
    };
    // the image's power-of-two scale rides in the B operand (exact), so the epilogue has no multiply for it
    auto store_g = [&](int nb, float (&g)[KP]) {
#pragma unroll
        for (int k = 0; k < KP; ++k) g[k] = (tid < nb && k < p.K) ? __fmul_rn(g[k], p.g_scale) : 0.f;
        store_row(sB, B_F8, tid, g, true);
    };
    {
        float w[KP], g[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) w[k] = __ldg(p.wo + (size_t)min(k, p.K - 1) * p.F + f);
        if (tid < 64) load_g(0, min(64, p.B), g);
#pragma unroll
        for (int k = 0; k < KP; ++k) w[k] = k < p.K ? w[k] : 0.f;
        store_row(sA, A_F8, tid, w, false);
        if (tid < 64) store_g(min(64, p.B), g);
    }
    const float *pvf = p.pv + f;
    uint32_t phase = 0;
    for (int b0 = 0; b0 < p.B; b0 += 64) {
        const int nb = min(64, p.B - b0);
        if (b0 != 0 && tid < 64) {                                        // later batch blocks: g_o rows, K-major [k/8][n][8]
            float g[KP];
            load_g(b0, nb, g);
            store_g(nb, g);
        }
        fence_async_smem();
        __syncthreads();
        if (warp == 0) {
            if (elect_one()) {
                const uint32_t a_base = desc_lo(smem_u32(sA), A_F8), b_base = desc_lo(smem_u32(sB), B_F8);   // LBO: next K group
#pragma unroll
                for (int j = 0; j < KG / 2; ++j)
                    mma_bf16(tmem_base, desc(SBO128, a_base + ((2 * j * A_F8) >> 4)), desc(SBO128, b_base + ((2 * j * B_F8) >> 4)), IDESC, j != 0);
                commit(&bar);
            }
            __syncwarp();
        }
        // ---- epilogue: thread = feature f, columns = samples.  Rows go in batches of 16 with the NEXT batch's pv values already
        //      requested (the first before the MMAs have finished).  Within a pair of rows the even lane converts and stores the
        //      position pair of the even row, the odd lane that of the odd row: one shuffle, one conversion per row and lane.
        auto load_pv = [&](int r0, float (&v)[16]) {
#pragma unroll
            for (int u = 0; u < 16; ++u) v[u] = __ldg(pvf + (size_t)(b0 + min(r0 + u, nb - 1)) * p.F);
        };
        uint32_t *img = reinterpret_cast<uint32_t *>(p.g_u) + img_word;
        const int half_f = p.F >> 1;
        float pa[16], pb[16];
        load_pv(0, pa);
        mbar_wait(&bar, phase);
        phase ^= 1u;
        fence_after();
        auto batch = [&](const uint32_t (&d)[32], int dj, int r0, const float (&v)[16]) {
#pragma unroll
            for (int u = 0; u < 16; u += 2) {
                // s * (1 - pv) * pv in the FP32 kernel's order (s already carries the image scale)
                const float x0 = __fmul_rn(__fmul_rn(__uint_as_float(d[dj + u]), __fmaf_rn(v[u], -1.f, 1.f)), v[u]);
                const float x1 = __fmul_rn(__fmul_rn(__uint_as_float(d[dj + u + 1]), __fmaf_rn(v[u + 1], -1.f, 1.f)), v[u + 1]);
                const float got = __shfl_xor_sync(0xffffffffu, even ? x1 : x0, 1);   // the partner's value of MY row
                const float mine = even ? x0 : x1;
                const float a = even ? mine : got, b = even ? got : mine;            // (even position, odd position)
                uint32_t hh;
                asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hh) : "f"(b), "f"(a));   // low half = even position
                const float2 hf = __half22float2(*reinterpret_cast<const __half2 *>(&hh));
                uint32_t ll;
                asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(ll) : "f"(b - hf.y), "f"(a - hf.x));
                const int row = r0 + u + (even ? 0 : 1);
                if (row < nb) {
                    uint32_t *q = img + (size_t)(b0 + row) * p.F;
                    q[0] = hh;
                    q[half_f] = ll;
                }
            }
        };
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            if (32 * h >= nb) break;
            uint32_t d[32];
            ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + 32 * h, d);
            load_pv(32 * h + 16, pb);
            batch(d, 0, 32 * h, pa);
            if (h == 0) load_pv(32, pa);
            batch(d, 16, 32 * h + 16, pb);
        }
        fence_before();
        __syncthreads();   // accumulator and sB are free for the next batch block
        fence_after();
    }
    if (warp == 0) tmem_dealloc(tmem_base, 64u);
}

bool readout_bwd_tc_enabled() {
    static int on = -1;
    if (on < 0) {
        const char *e = getenv("DCLL_RB_TC");
        on = (e && atoi(e) != 0) ? 1 : 0;
    }
    return on == 1;
}

// F16X2 layer with the image-form g_u (wgrad_tc2 consumes it), planes of a multiple of 32 positions, 8 | channels, K <= 32
bool readout_bwd_tc_supported(const dcll_conv_layer *L, int hw, int F) {
    return readout_bwd_tc_enabled() && prec_f16(L) && L->K <= rbtc::KP && hw % 32 == 0 && F % (8 * hw) == 0;
}

int launch_readout_bwd_tc(const dcll_conv_layer *L, const float *g_o, int hw, int F, float g_scale, cudaStream_t st) {
    RbTcP p;
    p.pv = L->pv, p.wo = L->wo, p.g_o = g_o, p.g_u = L->g_u;
    p.B = L->B, p.F = F, p.K = L->K, p.hw = hw, p.g_scale = g_scale;
    const unsigned grid = (unsigned)(F / 128);                       // (F / (8 hw)) groups x (hw / 32) blocks x 2 halves
    launch_k(readout_bwd_tc_kernel, grid, rbtc::NT, 0, st, p);
    DCLL_LAUNCH_OK("readout_bwd_tc_kernel");
    return DCLL_OK;
}

}  // namespace dcll
