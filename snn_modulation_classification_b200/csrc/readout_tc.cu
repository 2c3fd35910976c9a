// Local read-outs on tcgen05 (split-bf16 x3): partial[fs][row][kt] = sum_{f in range fs} pv[row,f] * Wcat[kt,f].
//
// Replaces readout_fwd_kernel in the bf16x3 precision mode (reference: dcll/pytorch_libdcll.py:602-606, i2o / output_).
// The GEMM is skinny (N = Ktot <= 64) and each pv element is used Ktot times, so the kernel is HBM-bound once the
// contraction leaves the FMA pipe (the FP32 kernel is shared-memory bound at ~16 % of FMA peak).
//
//   * M = 128 rows (samples, or sample-timesteps after dcll_infer_stack16), N = Ktot padded to 16, K = 16 features per MMA;
//   * both operands K-major, no swizzle: a 32-feature chunk of pv is staged as [f/8][row][8] (rows 16 B apart) and the
//     matching chunk of the read-out matrices as [f/8][{hi,lo}][k][8]; [W_hi | W_lo] is ONE N = 2*NPAD operand (A_hi is
//     read once for two of the three products), A_lo x W_hi the second MMA.  Feature-group pitches carry 32 B of padding
//     so that the converters' 8-byte stores are bank-conflict free;
//   * CTA = (feature range, row tile): it walks its range chunk by chunk, accumulating in TMEM, so partial blocks are
//     few (<= 148; ONE when there are enough row tiles to fill the GPU) -- the read-out matrices stream from L2, pv
//     from HBM exactly once;
//   * warp-specialised: three groups of 5 warps own alternate chunks and one operand-ring stage each; a thread
//     cp.async's (LDGSTS, coalesced 16-byte pieces) its own items one own-chunk ahead into thread-private raw fp32
//     slots (no registers held, no barrier), then converts them to bf16 hi/lo into its group's stage; one elected lane
//     issues the MMAs; warps 0-3 finally drain TMEM into the partials.  (A bulk copy per 128-byte row segment was tried: ~50 cycles per copy
//     through the TMA unit, 4x slower.)
// Rows and output columns of the MMA are independent, so rows >= `rows` and columns >= Ktot are simply never written
// (their accumulator lanes hold garbage that the epilogue does not read); only the feature tail is zero-filled.
// Partials are reduced by readout_finish_kernel exactly as in the FP32 path (fixed order, deterministic).
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace dcll {

struct RoTcP {
    const float *pv, *wo, *wout;
    float *partial;
    int rows, F, K, Ktot, n_chunks, n_fs;   // n_chunks = ceil(F/32), n_fs = feature ranges (grid.x)
};

namespace rotc {
constexpr int NT = 512, CH = 32;                                 // features per pipeline chunk
constexpr int CS = 3;                                            // converted (bf16 operand) ring depth
constexpr int MMA_WARP = 4, NGROUP = 3, GROUP_WARPS = 5, GT = GROUP_WARPS * 32;   // 3 loader/converter groups of 5 warps
constexpr int ROWS_PER_PASS = GT / 8;                            // a group covers 20 rows x 8 float4 per pass
// MROWS = rows a CTA stages per chunk: 128, or 64 when the whole batch is <= 64 rows (training at B = 64: one half-empty row
// tile).  The MMA stays M = 128 -- rows >= nrow are garbage lanes nobody reads -- but the raw ring shrinks with the rows, which
// pays for FOUR raw stages per group instead of two: with one chunk ahead per group only ~33 KB per SM were in flight and the
// kernel sat at 3.5 TB/s (latency x bandwidth of HBM wants ~70 KB per SM); three chunks ahead keep ~100 KB in flight.
template <int NPAD, int MROWS>
struct Lay {
    static constexpr int A_F8 = MROWS * 16 + 32;                 // pitch of one 8-feature group of pv rows (padded)
    static constexpr int A_PART = (CH / 8) * A_F8;               // one of {hi,lo}
    static constexpr int W_F8 = 2 * NPAD * 16 + 32;              // pitch of one feature group: [{hi,lo}][k][8] (padded)
    static constexpr int OFF_W = 2 * A_PART;
    static constexpr int CONV = OFF_W + (CH / 8) * W_F8;         // converted stage (one per group)
    static constexpr int U = (MROWS + NPAD + ROWS_PER_PASS - 1) / ROWS_PER_PASS;   // float4 items per thread and chunk
    static constexpr int RAW = U * GT * 16;                      // raw stage: [u][thread] float4, thread-private slots
    // raw (fp32, cp.async) stages per group: as many as fit beside the converted ring, 2..4
    static constexpr int RS_FIT = (227 * 1024 - 128 - CS * CONV) / (NGROUP * RAW);
    static constexpr int RS = RS_FIT >= 4 ? 4 : (RS_FIT >= 3 ? 3 : 2);
    static constexpr int OFF_CONV = NGROUP * RS * RAW;
    static constexpr int OFF_BAR = OFF_CONV + CS * CONV;
    static constexpr int SMEM = OFF_BAR + 128;
    static constexpr int TMEM_COLS = 2 * NPAD <= 32 ? 32 : (2 * NPAD <= 64 ? 64 : 128);
    // an M = 128 MMA reads 2 KB from the base of every feature group: with MROWS = 64 the tail lies in the following groups /
    // the W block of the same stage (garbage rows), which must still be inside the CTA's allocation
    static_assert(SMEM <= 227 * 1024, "shared memory");
    static_assert(A_PART + (CH / 8 - 1) * A_F8 + 128 * 16 <= CONV, "garbage rows of the last lo group stay inside the stage");
};
static_assert(CS == NGROUP, "one operand-ring stage per converter group");

// split-bf16 of four values with the PACKED conversion (cvt.rn.bf16x2.f32 = one F2FP for two values on the ALU pipe; the scalar
// F2F.BF16.F32 is a half-rate instruction and needs a byte permute per pair on top): 12 instructions per float4 instead of ~22
__device__ __forceinline__ uint32_t bf16x2_hi_lo(float a, float b, uint32_t &lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);                   // low half = a
    const uint32_t hb = *reinterpret_cast<const uint32_t *>(&h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(a - __uint_as_float(hb << 16), b - __uint_as_float(hb & 0xffff0000u));
    lo = *reinterpret_cast<const uint32_t *>(&l);
    return hb;
}
__device__ __forceinline__ uint2 pack_hi_lo(const float4 v, uint2 &lo) {
    uint2 hi;
    hi.x = bf16x2_hi_lo(v.x, v.y, lo.x);
    hi.y = bf16x2_hi_lo(v.z, v.w, lo.y);
    return hi;
}
// 16-byte asynchronous global -> shared copy (LDGSTS), L2-only; src_bytes = 0 zero-fills
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
}  // namespace rotc

template <int NPAD, int MROWS>
__global__ void __launch_bounds__(rotc::NT, 1) readout_tc_kernel(const RoTcP p) {
    using namespace rotc;
    using namespace tc;
    using G = Lay<NPAD, MROWS>;
    constexpr int RS = G::RS, A_F8 = G::A_F8, A_PART = G::A_PART;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + G::OFF_BAR);
    uint64_t *conv_full = bars, *conv_empty = bars + CS, *acc_full = bars + 2 * CS;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * CS + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = blockIdx.y * 128;
    const int nrow = min(MROWS, p.rows - row0);
    // feature range of this CTA, in 32-feature chunks (balanced split)
    const int c0 = (int)(((long long)blockIdx.x * p.n_chunks) / p.n_fs), c1 = (int)(((long long)(blockIdx.x + 1) * p.n_chunks) / p.n_fs);
    if (tid == 0) {
        for (int i = 0; i < CS; ++i) mbar_init(conv_full + i, GROUP_WARPS), mbar_init(conv_empty + i, 1);
        mbar_init(acc_full, 1);
        mbar_fence_init();
    }
    if (warp == MMA_WARP) tmem_alloc(tmem_slot, G::TMEM_COLS);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_entry();   // barrier init and the TMEM allocation above overlap the previous grid's tail; no global access before here

    if (warp == MMA_WARP) {
        // ================= MMA issuer =================
        constexpr uint32_t IDESC_N2 = idesc_bf16(128, 2 * NPAD, false, false), IDESC_N1 = idesc_bf16(128, NPAD, false, false);
        constexpr uint32_t SBO128 = desc_hi(128);                             // next 8 rows / next 8 outputs
        const uint32_t elected = elect_one();
        int it = 0;
        for (int c = c0; c < c1; ++c, ++it) {
            const int s = it % CS;
            mbar_wait(conv_full + s, (it / CS) & 1);
            fence_after();
            if (elected) {
                const uint32_t sb = smem_u32(smem + G::OFF_CONV + s * G::CONV);
                const uint32_t a_base = desc_lo(sb, A_F8), w_base = desc_lo(sb + G::OFF_W, G::W_F8);   // LBO: next feature group
#pragma unroll
                for (int j = 0; j < CH / 16; ++j) {
                    const uint64_t bd = desc(SBO128, w_base + ((2 * j * G::W_F8) >> 4));
                    mma_bf16(tmem_base, desc(SBO128, a_base + ((2 * j * A_F8) >> 4)), bd, IDESC_N2, (it | j) != 0);
                    mma_bf16(tmem_base, desc(SBO128, a_base + ((A_PART + 2 * j * A_F8) >> 4)), bd, IDESC_N1, 1);
                }
                commit(conv_empty + s);
                if (c == c1 - 1) commit(acc_full);
            }
            __syncwarp();
        }
    } else {
        // ================= loaders / converters =================
        // Group g (5 warps) owns chunks c0+g, c0+g+3, ... and operand stage g.  A thread owns float4 q of rows
        // n0, n0+20, ... (pv rows first, then read-out rows) for the whole range: it cp.async's them one own-chunk ahead
        // into private raw slots and converts its own slots -- no barrier on the raw ring, and the three groups' chains
        // (wait -> convert -> store -> fence -> arrive) overlap.
        const int wl = warp < MMA_WARP ? warp : warp - 1;        // 0..14
        const int g = wl / GROUP_WARPS, lg = (wl - g * GROUP_WARPS) * 32 + lane;
        const int q = lg & 7, n0 = lg >> 3, ntot = nrow + p.Ktot;
        const float *src[G::U];
        int dst[G::U];
#pragma unroll
        for (int u = 0; u < G::U; ++u) {
            const int n = n0 + u * ROWS_PER_PASS;
            src[u] = nullptr, dst[u] = 0;
            if (n < nrow) {
                src[u] = p.pv + (size_t)(row0 + n) * p.F + q * 4;
                dst[u] = (q >> 1) * A_F8 + n * 16 + (q & 1) * 8;
            } else if (n < ntot) {
                const int k = n - nrow;
                src[u] = (k < p.K ? p.wo + (size_t)k * p.F : p.wout + (size_t)(k - p.K) * p.F) + q * 4;
                dst[u] = G::OFF_W + (q >> 1) * G::W_F8 + k * 16 + (q & 1) * 8;
            }
        }
        const int raw_off = g * RS * G::RAW + lg * 16;
        auto issue = [&](int c, int stage) {
            const bool in = c * CH + q * 4 < p.F;                                         // feature tail -> zero fill
            const uint32_t d = smem_u32(smem) + raw_off + stage * G::RAW;
#pragma unroll
            for (int u = 0; u < G::U; ++u)
                if (src[u]) cp_async16(d + u * (GT * 16), in ? src[u] + (size_t)c * CH : src[u], in ? 16u : 0u);
        };
        unsigned char *cv = smem + G::OFF_CONV + g * G::CONV;
        // RS - 1 own chunks ahead, one commit group per chunk (empty groups keep the count uniform at the tail)
#pragma unroll
        for (int a = 0; a < RS - 1; ++a) {
            if (c0 + g + a * NGROUP < c1) issue(c0 + g + a * NGROUP, a);
            cp_async_commit();
        }
        int k = 0;
        for (int c = c0 + g; c < c1; c += NGROUP, ++k) {
            // slot (k-1) % RS was converted by this very thread in the previous iteration: refill it with chunk k + RS - 1
            if (c + (RS - 1) * NGROUP < c1) issue(c + (RS - 1) * NGROUP, (k + RS - 1) % RS);
            cp_async_commit();
            cp_async_wait<RS - 1>();
            const unsigned char *raw = smem + raw_off + (k % RS) * G::RAW;
            if (k >= 1) mbar_wait(conv_empty + g, (k - 1) & 1);
#pragma unroll
            for (int u = 0; u < G::U; ++u) {
                if (!src[u]) continue;
                uint2 lo;
                const uint2 hi = pack_hi_lo(*reinterpret_cast<const float4 *>(raw + u * (GT * 16)), lo);
                *reinterpret_cast<uint2 *>(cv + dst[u]) = hi;
                *reinterpret_cast<uint2 *>(cv + dst[u] + (n0 + u * ROWS_PER_PASS < nrow ? A_PART : NPAD * 16)) = lo;
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(conv_full + g);
        }
        if (warp < MMA_WARP) {
            // ================= epilogue: warps 0..3 = TMEM lane quarters, thread = one row =================
            const int r = warp * 32 + lane;
            float *out = p.partial + ((size_t)blockIdx.x * p.rows + row0 + (r < nrow ? r : 0)) * p.Ktot;
            if (c0 < c1) {
                mbar_wait(acc_full, 0);
                fence_after();
            }
#pragma unroll
            for (int n0 = 0; n0 < NPAD; n0 += 16) {
                uint32_t v[16], v2[16];
                if (c0 < c1) {
                    const uint32_t ta = tmem_base + ((uint32_t)(warp * 32) << 16) + n0;
                    ld16(ta, v);
                    ld16(ta + NPAD, v2);
                } else {
#pragma unroll
                    for (int k = 0; k < 16; ++k) v[k] = v2[k] = 0u;
                }
                if (r < nrow) {
#pragma unroll
                    for (int k = 0; k < 16; ++k)
                        if (n0 + k < p.Ktot) out[n0 + k] = __uint_as_float(v[k]) + __uint_as_float(v2[k]);
                }
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc(tmem_base, G::TMEM_COLS);
}

bool readout_tc_supported(const dcll_conv_layer *L) {
    Geo g = geo_of(L);
    return prec_tc(L) && g.Ktot <= 64 && (g.F % 4) == 0 &&
           (((uintptr_t)L->pv | (uintptr_t)L->wo | (uintptr_t)L->wout) % 16) == 0;
}

// Feature ranges (= partial blocks): fill the 148 SMs when the row tiles alone do not, else pick the split with the
// least idle tail in the last wave.
int readout_tc_blocks(const dcll_conv_layer *L) {
    const int n_chunks = ceil_div(geo_of(L).F, rotc::CH), n_rt = ceil_div(L->B, 128);
    if (n_rt < 148) return max(1, min(n_chunks, sm_budget() / n_rt));
    int best = 1;
    double best_eff = 0.0;
    for (int s = 1; s <= 4 && s <= n_chunks; ++s) {
        const int ctas = n_rt * s;
        const double eff = (double)ctas / (148.0 * ceil_div(ctas, 148));
        if (eff > best_eff + 0.02) best_eff = eff, best = s;
    }
    return best;
}

template <int NPAD, int MROWS>
static int launch_rotc_m(const RoTcP &p, dim3 grid, cudaStream_t st) {
    DCLL_SMEM_ATTR((readout_tc_kernel<NPAD, MROWS>), (rotc::Lay<NPAD, MROWS>::SMEM));
    launch_k(readout_tc_kernel<NPAD, MROWS>, grid, rotc::NT, rotc::Lay<NPAD, MROWS>::SMEM, st, p);
    DCLL_LAUNCH_OK("readout_tc_kernel");
    return DCLL_OK;
}
template <int NPAD>
static int launch_rotc(const RoTcP &p, dim3 grid, cudaStream_t st) {
    return p.rows <= 64 ? launch_rotc_m<NPAD, 64>(p, grid, st) : launch_rotc_m<NPAD, 128>(p, grid, st);
}

int launch_readout_tc(const dcll_conv_layer *L, float *partial, cudaStream_t st) {
    Geo g = geo_of(L);
    RoTcP p;
    p.pv = L->pv, p.wo = L->wo, p.wout = L->wout, p.partial = partial;
    p.rows = L->B, p.F = g.F, p.K = L->K, p.Ktot = g.Ktot;
    p.n_chunks = ceil_div(g.F, rotc::CH), p.n_fs = readout_tc_blocks(L);
    dim3 grid(p.n_fs, ceil_div(L->B, 128));
    DCLL_REQUIRE(grid.y <= 65535, DCLL_EUNSUPPORTED, "read-out over %d rows: more than 65535 row tiles", L->B);
    const int npad = ceil_div(g.Ktot, 16) * 16;
    switch (npad) {
        case 16: return launch_rotc<16>(p, grid, st);
        case 32: return launch_rotc<32>(p, grid, st);
        case 48: return launch_rotc<48>(p, grid, st);
        default: return launch_rotc<64>(p, grid, st);
    }
}

}  // namespace dcll
