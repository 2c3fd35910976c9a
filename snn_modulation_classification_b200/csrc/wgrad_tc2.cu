// Local weight gradient of a 7x7, 32 -> 32 Conv2dDCLLlayer on tcgen05 with ROW-PAIR N-CONCATENATION (split-bf16 x3).
//
//     gW[co,ci,kh,kw] = sum_{b,h,w} g_u[b,co,h,w] * eps1[b,ci,h+kh-pad,w+kw-pad]      (dcll/pytorch_libdcll.py:704, see wgrad.cu)
//
// wgrad_tc_kernel (wgrad_tc.cu) issues one N = 64 / N = 32 MMA pair per (output row, kernel column): an SS-mode MMA with
// M = 128, K = 16 costs max(N/2, 32 + N/4) cycles (4 KB of A + 32 N bytes of B through the 128 B/cycle shared-memory operand
// path; tools/mma_bench.cu), so those MMAs run at 49 + 45 cycles for 32 + 16 cycles of math.  Here N is doubled without more
// output channels: the B operand holds TWO adjacent output rows,
//
//     B_main = [ G_hi(r) | G_hi(r+1) | G_lo(r) | G_lo(r+1) ]   N = 128        B_lo = its first half, N = 64
//     A      = X[halo rows r + 4g + dy, dy = 0..3][cols c + kw][ci]            M = 128 = (dy, ci), read ONCE for both rows
//
// Column block j = 0 (row r) meets kernel row kh = 4g + dy, block j = 1 (row r+1) meets kh = 4g + dy - 1; with g in {0,1} both
// row parities see every kh = 0..6 exactly once (kh = -1 and 7 are the padding each group already paid).  Stepping r by two
// halves the MMAs over positions: 8 row pairs x (64 + 49) cycles per kernel column and tile instead of 16 x (49 + 45).
// The accumulator of one kernel column is now 128 x 128 fp32, so a CTA holds at most four of them (512 TMEM columns):
// FOUR CTA roles -- kernel-row group g in {0,1} x kernel columns {0..3} (role A) or {4..6} (role B).  Role B has a spare
// accumulator, which takes the bias gradient as one more MMA per row pair: ones(128 x 16) x B_main = sum over positions of
// G_hi and G_lo (alternate pairs on the g = 0 / g = 1 CTA).  nA : nB CTA pairs = 40 : 34 balances the two roles.
//
// Operands arrive without any conversion in this kernel:
//   * eps1: the bf16 {hi,lo} operand image the forward wrote ([b][part][ci/8][H][W][8 ci]), staged [row][ci/8][col][8 ci] exactly as
//     in wgrad_tc_kernel (MN-major A; (dy, ci/8) uniformly strided; kw = 16-byte start shift);
//   * g_u: the packed backward read-out (readout_bwd2_kernel<.., IMG>) leaves it as bf16 {hi,lo} NCHW planes
//     [b][part][co][Hc][Wc] in the g_u buffer (same bytes as fp32).  16-byte pieces (8 columns of one channel) are staged as the
//     K-MAJOR B operand: core matrix = [co % 8][8 columns], N-group order (part, row parity, co/8), two K chunks per group.
// Tiles are moved by TMA: one thread issues three cp.async.bulk.tensor boxes per unit of 8 rows x 16 columns (tensor maps whose
// DIMENSION ORDER is chosen so that the box lands in shared memory already in operand order; zero fill outside the picture)
// through a FOUR-stage ring while two issuer warps consume; mbarriers both ways.  (History, B200 measurements: the first version
// staged 16-row tiles with 5 216 16-byte cp.async per tile into a double buffer -- 0.39 ms per launch, 0.34 ms with the MMAs
// removed and 0.27 ms with the loads removed: bound by the load path.  One TMA box instead of the cp.async changed nothing
// (0.35 ms loads-only): it is the LATENCY of a tile load with a single tile in flight per SM, 12 B/cycle/SM, hence the deeper ring.)  Each CTA leaves its accumulators as one compact partial block
// [co][ci][kernel column a][slot = kh - (4g-1)] (+ 32 bias sums); reduce_adam_rp_kernel (wgrad.cu) adds the blocks that hold a
// given element in a fixed order -- deterministic, no float atomics.
#include <cuda_bf16.h>
#include <string.h>

#include "common.cuh"
#include "tc_ptx.cuh"
#include "tmap.cuh"

namespace dcll {

struct Wg2P {
    float *partial;      // [n_cta][Wg2Geo::BLK]
    int B, H, W, padH, padW, Hc, Wc;
    int tiles_h, tiles_w, n_units;
    int nA, nB;          // CTA pairs of role A (kernel columns 0..3) and role B (4..6 + bias)
    int dbg;             // DCLL_WG2_DEBUG (timing experiments only, results are garbage): bit 0 skip eps1 loads, bit 1 skip g_u loads,
                         // bit 2 skip the MMAs
    unsigned long long *tl;   // in-kernel stopwatch block (common.cuh TL_*), null when off
    int f16;                  // DCLL_PREC_F16X2: the eps1 image holds ONE fp16 part (scaled), g_u fp16 {hi,lo} (scaled); no X_lo product
    float unscale, unscale_g; // F16X2: 2^-(a_exp + g_exp) for the weight-gradient sums, 2^-g_exp for the bias sums; else 1
    int g16;                  // the g_u tensor map merges the two 8-column chunks of a unit row into one 256-byte box row
};

template <int TH_, int NSTAGE_>
struct Wg2GeoT {
    static constexpr int KH = 7, KW = 7, CIN = 32, COUT = 32;
    static constexpr int TH = TH_, TW = 16, PAIRS = TH / 2;  // a unit = TH output rows x 16 columns = TH/2 row pairs
    static constexpr int CGR = 4, DY = 4;
    static constexpr int XROWS = TH + DY - 2;               // halo rows 2*pair + dy of one kernel-row group: 10
    static constexpr int XCOLS = TW + KW - 1;               // 22
    static constexpr int X_CP = XCOLS * 16, X_RP = CGR * X_CP, X_PART = XROWS * X_RP, X_BYTES = 2 * X_PART;
    // g_u tile: [pair 4][plane 8 = (part, co/8)][row parity 2][k chunk 2][co % 8][8 columns] bf16 -- ONE tensor-map box per unit out
    // of the image the backward read-out wrote (128-byte core matrices).  N-group (8 channels) = plane*2 + parity, 256 bytes
    // apart; the two K chunks of a group 128 bytes apart.
    static constexpr int G_GRP = 256;
    static constexpr int G_PAIR = 16 * G_GRP;                // 4 KB
    static constexpr int G_BYTES = PAIRS * G_PAIR;           // 16 KB
    static constexpr int BUF = X_BYTES + G_BYTES;            // 44 544 B per pipeline stage
    // FOUR stages: a tile load takes ~2.7 us from request to landing (measured: loads alone 0.34 ms per launch with one tile in
    // flight per SM -- 12 B/cycle/SM, a third of what L2 can deliver), the MMAs of a unit ~1 us, so three loads must be in flight
    static constexpr int NSTAGE = NSTAGE_;
    static constexpr int OFF_ONES = NSTAGE * BUF, OFF_BAR = OFF_ONES + 4096, SMEM = OFF_BAR + 128;
    static constexpr int NT = 512;
    static constexpr int ACC_COLS = 128, TMEM_COLS = 512;
    static constexpr int SLOTS = 5, NACC = 4, PER_CC = NACC * SLOTS;   // floats per (co, ci) in a partial block
    static constexpr int PITCH = PER_CC + 1;                 // staging pitch (odd: conflict-free across ci)
    static constexpr int NW_BLK = COUT * CIN * PER_CC;       // 20 480
    static constexpr int BLK = NW_BLK + COUT;                // + bias sums
    static_assert(COUT * CIN * PITCH * 4 <= NSTAGE * BUF, "staging fits the tile buffers");
    static_assert(X_PART % 16 == 0 && BUF % 128 == 0 && X_BYTES % 128 == 0, "alignment");
};
using Wg2Geo = Wg2GeoT<8, 4>;          // default: 8-row units, four stages (DCLL_WG2_TILE=16: 16-row units, two stages)

size_t wgrad_tc2_partial_floats() { return (size_t)148 * Wg2Geo::BLK; }
static bool wg2_pair();

template <int TH_, int NSTAGE_>
__global__ void __launch_bounds__(512, 1) wgrad_tc2_kernel(const Wg2P p, const __grid_constant__ TmapDesc tmx,
                                                           const __grid_constant__ TmapDesc tmg) {
    using G = Wg2GeoT<TH_, NSTAGE_>;
    static_assert(G::SMEM <= 227 * 1024, "shared memory");
    using namespace tc;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + G::OFF_BAR);
    uint64_t *full = bars, *empty = bars + G::NSTAGE, *done = bars + 2 * G::NSTAGE;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * G::NSTAGE + 1);
    uint32_t *ones_used = tmem_slot + 1;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool tl_on = p.tl != nullptr;
    unsigned long long *tl = tl_on ? p.tl + (size_t)blockIdx.x * TL_SLOTS : nullptr;
    const long long tl_entry = tl_on ? clock64() : 0;
    if (tl_on && threadIdx.x == 0) tl[TL_T_ENTRY] = gtimer_ns();
    if (tid == 0) {
        for (int i = 0; i < G::NSTAGE; ++i) mbar_init(full + i, 1), mbar_init(empty + i, 1);
        mbar_init(done, 1);
        *ones_used = 0u;
        mbar_fence_init();
    }
    // the all-ones A tile of the bias-gradient MMA (any layout: every element is 1.0)
    for (int i = tid; i < 4096 / 4; i += G::NT) reinterpret_cast<uint32_t *>(smem + G::OFF_ONES)[i] = p.f16 ? 0x3c003c00u : 0x3f803f80u;   // 1.0
    if (warp == 2) tmem_alloc(tmem_slot, (uint32_t)G::TMEM_COLS);
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_entry();   // nothing global is touched before here
    if (tl_on && tid == 0) tl[TL_PROLOGUE] = clock64() - tl_entry;

    // role / group / unit walk of this CTA: CTAs (2k, 2k+1) form a pair (g = 0, 1) that walks the same units
    const int pair_id = blockIdx.x >> 1, grp = blockIdx.x & 1;
    const bool roleB = pair_id >= p.nA;
    const int u_first = roleB ? pair_id - p.nA : pair_id, u_step = roleB ? p.nB : p.nA;
    const int kw_base = roleB ? 4 : 0;
    const int tiles = p.tiles_h * p.tiles_w;
    const int row_off = G::DY * grp;

    if (warp == 4) {
        // ================= tile producer: one thread, three tensor-map boxes per unit (UTMALDG), zero fill outside the picture:
        //   eps1 image  dims (8 ci, W, ci/8, H, (b, part))          box (8, 22, 4, 10, 2)  -> [part][row][ci/8][col][8 ci]
        //   g_u image   dims (64 = co%8 x 8 w, w/8, row parity, (b, part, co/8), row/2)  box (64, 2, 2, 8, 4)
        //                                                           -> [pair][part, co/8][parity][k chunk][co % 8][8 w]
        // It runs up to NSTAGE units ahead of the issuers.
        if (lane == 0) {
            tma_prefetch_desc(&tmx);
            tma_prefetch_desc(&tmg);
            int i = 0;
            long long tl_wait = 0;
            for (int u = u_first; u < p.n_units; u += u_step, ++i) {
                const int sg = i % G::NSTAGE;
                const uint32_t sX = smem_u32(smem + sg * G::BUF), sG = sX + G::X_BYTES, bar = smem_u32(full + sg);
                const int b = u / tiles, tile = u - b * tiles;
                const int th_i = tile / p.tiles_w, tw_i = tile - th_i * p.tiles_w;
                const int h0 = th_i * G::TH, w0 = tw_i * G::TW;
                if (i >= G::NSTAGE) TL_TIMED(tl_on, tl_wait, mbar_wait_relaxed(empty + sg, ((i / G::NSTAGE) - 1) & 1));   // MMAs of unit i-NSTAGE have read this stage
                mbar_expect_tx(full + sg, ((p.dbg & 1) ? 0 : (p.f16 ? G::X_PART : G::X_BYTES)) + ((p.dbg & 2) ? 0 : G::G_BYTES));
                if (!(p.dbg & 1)) tma_load_5d(sX, &tmx, bar, 0, w0 - p.padW, 0, h0 - p.padH + row_off, 2 * b);
                if (!(p.dbg & 2)) tma_load_5d(sG, &tmg, bar, 0, p.g16 ? w0 >> 4 : w0 >> 3, 0, 8 * b, h0 >> 1);
            }
            if (tl_on) tl[TL_APROD_EMPTY] = tl_wait;
        }
    } else if (warp == 0) {
        // ================= MMA issuer: ONE warp.  (Two issuer warps -- kernel columns split between them -- were right for the
        // N = 64 / 32 MMAs of wgrad_tc_kernel, which a single thread cannot issue fast enough; with N = 128 and the MN-major A
        // operand two concurrent streams cost 92 cycles per MMA in aggregate against 64 for one, tools/mma_bench.cu.) ==========
        const uint32_t IDESC_N128 = p.f16 ? idesc_f16a(128, 128, true, false) : idesc_bf16(128, 128, true, false);   // A MN-major, B K-major
        constexpr uint32_t IDESC_N64 = idesc_bf16(128, 64, true, false);
        constexpr uint32_t A_HI = desc_hi(G::X_CP);          // SBO: next 8 rows of M = next (dy, cg) group
        constexpr uint32_t B_HI = desc_hi(G::G_GRP);         // SBO: next 8 rows of N = next (part, co/8, parity) group
        constexpr uint32_t ONES_HI = desc_hi(128);
        const uint32_t elected = elect_one();
        const int kw0 = kw_base, kw1 = roleB ? 7 : 4;
        const bool do_ones = roleB;
        const uint64_t ones_desc = desc(ONES_HI, desc_lo(smem_u32(smem + G::OFF_ONES), 2048));
        uint32_t ones_acc = 0;
        int i = 0;
        long long tl_wait = 0;
        const long long tl_loop0 = tl_on ? clock64() : 0;
        for (int u = u_first; u < p.n_units; u += u_step, ++i) {
            const int sg = i % G::NSTAGE;
            const int tile = u % tiles;
            const int h0 = (tile / p.tiles_w) * G::TH;
            const int npair = (min(G::TH, p.Hc - h0) + 1) >> 1;
            const uint32_t a_base = desc_lo(smem_u32(smem + sg * G::BUF), 128);                      // LBO: next 8 positions (K)
            const uint32_t b_base = desc_lo(smem_u32(smem + sg * G::BUF + G::X_BYTES), 128);         // LBO: next 8 columns (K)
            TL_TIMED(tl_on, tl_wait, mbar_wait(full + sg, (i / G::NSTAGE) & 1));
            fence_after();
            if (tl_on && i == 0 && warp == 0 && elected) tl[TL_FIRST_MMA] = clock64() - tl_entry;
            if (elected) {
                for (int pr = 0; pr < ((p.dbg & 4) ? (i == 0 ? 1 : 0) : npair); ++pr) {
                    const uint64_t b = desc(B_HI, b_base + pr * (G::G_PAIR >> 4));
                    const uint32_t acc = (i == 0 && pr == 0) ? 0u : 1u;
                    const uint32_t a_row = a_base + ((2 * pr * G::X_RP) >> 4);
#pragma unroll 2
                    for (int kw = kw0; kw < kw1; ++kw) {
                        const uint32_t d = tmem_base + (kw - kw_base) * G::ACC_COLS;
                        mma_bf16(d, desc(A_HI, a_row + kw), b, IDESC_N128, acc);
                        if (!p.f16) mma_bf16(d, desc(A_HI, a_row + kw + (G::X_PART >> 4)), b, IDESC_N64, 1);
                    }
                    if (do_ones && (pr & 1) == grp) {
                        mma_bf16(tmem_base + 3 * G::ACC_COLS, ones_desc, b, IDESC_N128, ones_acc);
                        ones_acc = 1;
                    }
                }
                commit(empty + sg);
            }
            __syncwarp();
        }
        if (elected) {
            if (do_ones) *ones_used = ones_acc;
            commit(done);
            if (tl_on) tl[warp == 0 ? TL_ISS_A_FULL : TL_ISS_W_FULL] = tl_wait, tl[warp == 0 ? TL_ISS_LOOP : TL_EPI_LOOP] = clock64() - tl_loop0;
        }
        __syncwarp();
    }
    // ---- drain.  Lane m = (dy, ci) of accumulator a; its 128 columns are N-groups (part, co/8, row parity) x 8 channels:
    //        column = part*64 + (co/8)*16 + parity*8 + co%8;   part 0 = X_hi G_hi + X_lo G_hi, part 1 = X_hi G_lo
    //        parity 0 (even rows): kernel row 4g + dy      -> slot dy + 1
    //        parity 1 (odd rows) : kernel row 4g + dy - 1  -> slot dy
    //      (slot s <-> kh = 4g - 1 + s).  Two passes through the idle tile buffers: parity 0 stores slots 1..4, then parity 1 stores
    //      slot 0 and adds to slots 1..3; each (co, ci, a, slot) is touched by one thread per pass.
    float *out = p.partial + (size_t)blockIdx.x * G::BLK;
    mbar_wait_relaxed(done, 0, 1000);   // 13+ idle warps: sleep between polls (see tc_ptx.cuh)
    fence_after();
    __syncthreads();
    const long long tl_drain0 = tl_on ? clock64() : 0;
    {
        const int q = warp & 3;                                          // TMEM lane quarter of this warp = dy
        const int ci = lane;
        float *stg = reinterpret_cast<float *>(smem);
        const int nacc = roleB ? 3 : 4;
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            for (int a = (warp >> 2); a < nacc; a += 4) {
                const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + a * G::ACC_COLS;
#pragma unroll
                for (int hc = 0; hc < 2; ++hc) {                         // columns [32 hc, 32 hc + 32) of both parts: co/8 = 2 hc, 2 hc + 1
                    uint32_t v[32], v2[32];
                    ld32(ta + 32 * hc, v);
                    ld32(ta + 64 + 32 * hc, v2);
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        if (((c >> 3) & 1) != pass) continue;
                        const int co = (2 * hc + (c >> 4)) * 8 + (c & 7);
                        float val = __uint_as_float(v[c]) + __uint_as_float(v2[c]);
                        if (p.f16) val *= p.unscale;
                        float *sp = stg + (co * G::CIN + ci) * G::PITCH + a * G::SLOTS + q + 1 - pass;
                        *sp = (pass == 0 || q == 0) ? val : *sp + val;
                    }
                }
            }
            __syncthreads();
        }
        // bias gradient (role B): every lane of the spare accumulator holds the column sums of B_main over this CTA's pairs
        if (warp == 0) {
            float bsum = 0.f;
            if (roleB && *ones_used) {
                const uint32_t ta = tmem_base + 3 * G::ACC_COLS;
                const int idx = ((lane >> 3) & 1) * 16 + (lane & 7);     // column of (co/8 parity-0 block) inside a 32-column chunk
#pragma unroll
                for (int hc = 0; hc < 2; ++hc) {
                    uint32_t c0[32], c1[32];
                    ld32(ta + 32 * hc, c0);
                    ld32(ta + 64 + 32 * hc, c1);
                    if ((lane >> 4) == hc) {
#pragma unroll
                        for (int c = 0; c < 24; ++c)
                            if (c == idx) bsum = (__uint_as_float(c0[c]) + __uint_as_float(c1[c])) + (__uint_as_float(c0[c + 8]) + __uint_as_float(c1[c + 8]));
                    }
                }
            }
            out[G::NW_BLK + lane] = p.f16 ? bsum * p.unscale_g : bsum;
        }
        fence_before();
        __syncthreads();
        for (int e = tid; e < G::NW_BLK; e += G::NT) {
            const int cc = e / G::PER_CC, r = e - cc * G::PER_CC;
            out[e] = (r < nacc * G::SLOTS) ? stg[cc * G::PITCH + r] : 0.f;
        }
    }
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, (uint32_t)G::TMEM_COLS);
    if (tl_on && tid == 0) tl[TL_DRAIN] = clock64() - tl_drain0, tl[TL_TOTAL] = clock64() - tl_entry, tl[TL_T_EXIT] = gtimer_ns();
}

// ---------------------------------------------------------------------------------------------------------------------
// CTA-PAIR form (cta_group::2): the two CTAs of a pair (kernel-row groups g = 0, 1 of the same units) become one cluster and the
// leader issues ONE M = 256 MMA for both.  Why: the single-CTA kernel is bound by the 128 B/cycle shared-memory path (in-kernel
// stopwatch: the issuers never wait for tiles; per unit 224 KB of operand reads + 44.5 KB of tile writes = 2 100 cycles, which
// is the measured rate).  In a pair every CTA keeps only HALF of the N columns of B:
//     CTA 0 (g = 0): [ G_hi(r)   | G_lo(r)   ]      CTA 1 (g = 1): [ G_hi(r+1) | G_lo(r+1) ]       64 columns each
// so the g_u tile a CTA stages halves (16 -> 8 KB per unit) and so does the B read of every MMA (4 -> 2 KB, 2 -> 1 KB):
//     main : X_hi x B,  N = 128 -> D[  0..127] = [ hi.hi(r) | hi.lo(r) | hi.hi(r+1) | hi.lo(r+1) ]
//     lo   : X_lo x (first half of each CTA's B = G_hi(r) | G_hi(r+1)),  N = 64 -> D[ 32.. 95]
// The lo product lands on the hi.lo(r) and hi.hi(r+1) columns -- any column of the right (row parity, channel) will do, the
// drain adds the two 32-column blocks of a parity anyway.  A (the eps1 halo rows of the CTA's own kernel-row group) stays
// per CTA.  Both producers count their bytes on the LEADER's full barrier; the leader's two issuer warps release the stage in
// both CTAs with multicast commits.  The bias-gradient MMA (ones x B) now serves both CTAs at once and runs on every row pair.
template <int TH_, int NSTAGE_>
struct Wg2PairGeoT : Wg2GeoT<TH_, NSTAGE_> {
    using Base = Wg2GeoT<TH_, NSTAGE_>;
    // g_u half tile: [pair][plane 8 = (part, co/8)][k chunk 2][co % 8][8 columns] bf16 of ONE row parity
    static constexpr int G_GRP = 256, G_PAIR = 8 * G_GRP, G_BYTES = Base::PAIRS * G_PAIR;
    static constexpr int BUF = Base::X_BYTES + G_BYTES;
    static constexpr int OFF_ONES = NSTAGE_ * BUF, OFF_BAR = OFF_ONES + 4096, SMEM = OFF_BAR + 128;
    static_assert(Base::COUT * Base::CIN * Base::PITCH * 4 <= NSTAGE_ * BUF, "staging fits the tile buffers");
    static_assert(BUF % 128 == 0, "alignment");
};

// The MMAs of one unit, straight-line: with the pair / kernel-column loops rolled the issuing thread spent ~110 cycles of
// uniform-datapath work per iteration (ncu: tensor pipe 59 % active, the issuer never waiting on a barrier), i.e. the 64-cycle
// N = 128 MMAs were issue-bound.  Unrolled, an MMA is a descriptor add and the instruction.
template <class G, bool ROLEB, bool F16>
__device__ __forceinline__ void wg2p_issue_unit(uint32_t tmem_base, uint32_t a_base, uint32_t b_base, uint64_t ones_desc, uint32_t acc0,
                                                int npair) {
    using namespace tc;
    constexpr int NKW = ROLEB ? 3 : 4, KW0 = ROLEB ? 4 : 0;
    constexpr uint32_t IDESC_N128 = F16 ? idesc_f16a(256, 128, true, false) : idesc_bf16(256, 128, true, false);   // M = 256 over the pair
    constexpr uint32_t IDESC_N64 = idesc_bf16(256, 64, true, false);
    constexpr uint32_t A_HI = desc_hi(G::X_CP);          // SBO: next 8 rows of M = next (dy, cg) group
    constexpr uint32_t B_HI = desc_hi(G::G_GRP);         // SBO: next 8 columns of N = next (part, co/8) plane
    auto pair_row = [&](int pr, uint32_t acc) {
        const uint64_t b = desc(B_HI, b_base + pr * (G::G_PAIR >> 4));
        const uint32_t a_row = a_base + ((2 * pr * G::X_RP) >> 4) + KW0;
#pragma unroll
        for (int k = 0; k < NKW; ++k) {
            const uint32_t d = tmem_base + k * G::ACC_COLS;
            mma_bf16_2cta(d, desc(A_HI, a_row + k), b, IDESC_N128, acc);
            if (!F16) mma_bf16_2cta(d + 32, desc(A_HI, a_row + k + (G::X_PART >> 4)), b, IDESC_N64, 1);
        }
        if (ROLEB) mma_bf16_2cta(tmem_base + 3 * G::ACC_COLS, ones_desc, b, IDESC_N128, acc);
    };
    if (npair == G::PAIRS) {
        pair_row(0, acc0);
#pragma unroll
        for (int pr = 1; pr < G::PAIRS; ++pr) pair_row(pr, 1u);
    } else {
#pragma unroll 1
        for (int pr = 0; pr < npair; ++pr) pair_row(pr, pr == 0 ? acc0 : 1u);
    }
}

template <int TH_, int NSTAGE_>
__global__ void __launch_bounds__(512, 1) wgrad_tc2p_kernel(const Wg2P p, const __grid_constant__ TmapDesc tmx,
                                                            const __grid_constant__ TmapDesc tmg) {
    using G = Wg2PairGeoT<TH_, NSTAGE_>;
    static_assert(G::SMEM <= 227 * 1024, "shared memory");
    using namespace tc;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + G::OFF_BAR);
    uint64_t *full = bars, *empty = bars + G::NSTAGE, *done = bars + 2 * G::NSTAGE;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * G::NSTAGE + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();                    // 0: leader, kernel-row group 0; 1: group 1
    const bool tl_on = p.tl != nullptr;
    unsigned long long *tl = tl_on ? p.tl + (size_t)blockIdx.x * TL_SLOTS : nullptr;
    const long long tl_entry = tl_on ? clock64() : 0;
    if (tl_on && threadIdx.x == 0) tl[TL_T_ENTRY] = gtimer_ns();
    if (tid == 0) {
        // full: ONE arrival (the leader's producer, with the byte count of both CTAs); only the leader's copy is ever waited on
        for (int i = 0; i < G::NSTAGE; ++i) mbar_init(full + i, 1), mbar_init(empty + i, 1);
        mbar_init(done, 1);
        mbar_fence_init();
    }
    for (int i = tid; i < 4096 / 4; i += G::NT) reinterpret_cast<uint32_t *>(smem + G::OFF_ONES)[i] = p.f16 ? 0x3c003c00u : 0x3f803f80u;   // 1.0
    if (warp == 2) tmem_alloc2(tmem_slot, (uint32_t)G::TMEM_COLS);
    fence_async_smem();
    fence_before();
    __syncthreads();
    cluster_sync();                                              // both CTAs' barriers exist before anybody signals across the pair
    fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_entry();   // nothing global is touched before here (except the stopwatch stamp)
    if (tl_on && tid == 0) tl[TL_PROLOGUE] = clock64() - tl_entry;

    const int pair_id = blockIdx.x >> 1, grp = (int)rank;
    const bool roleB = pair_id >= p.nA;
    const int u_first = roleB ? pair_id - p.nA : pair_id, u_step = roleB ? p.nB : p.nA;
    const int kw_base = roleB ? 4 : 0;
    const int tiles = p.tiles_h * p.tiles_w;
    const int row_off = G::DY * grp;

    if (warp == 4) {
        // ================= tile producer (one thread per CTA): its own eps1 halo rows and its own row parity of g_u =================
        //   g_u image dims (64 = co%8 x 8 w, w/8, row parity, (b, part, co/8), row/2), box (64, 2, 1, 8, PAIRS) at parity = rank
        if (lane == 0) {
            tma_prefetch_desc(&tmx);
            tma_prefetch_desc(&tmg);
            int i = 0;
            long long tl_wait = 0;
            for (int u = u_first; u < p.n_units; u += u_step, ++i) {
                const int sg = i % G::NSTAGE;
                const uint32_t sX = smem_u32(smem + sg * G::BUF), sG = sX + G::X_BYTES, bar = smem_u32(full + sg);
                const int b = u / tiles, tile = u - b * tiles;
                const int th_i = tile / p.tiles_w, tw_i = tile - th_i * p.tiles_w;
                const int h0 = th_i * G::TH, w0 = tw_i * G::TW;
                if (i >= G::NSTAGE) TL_TIMED(tl_on, tl_wait, mbar_wait_relaxed(empty + sg, ((i / G::NSTAGE) - 1) & 1));   // the pair's MMAs have read this stage
                if (rank == 0) mbar_expect_tx(full + sg, 2 * ((p.f16 ? G::X_PART : G::X_BYTES) + G::G_BYTES));
                tma_load_5d_2cta(sX, &tmx, bar, 0, w0 - p.padW, 0, h0 - p.padH + row_off, 2 * b);
                tma_load_5d_2cta(sG, &tmg, bar, 0, p.g16 ? w0 >> 4 : w0 >> 3, (int)rank, 8 * b, h0 >> 1);
            }
            if (tl_on) tl[TL_APROD_EMPTY] = tl_wait;
        }
    } else if (warp == 0 && rank == 0) {
        // ================= MMA issuer (leader only, one warp: see wgrad_tc2_kernel) =========
        constexpr uint32_t ONES_HI = desc_hi(128);
        const uint32_t elected = elect_one();
        const uint64_t ones_desc = desc(ONES_HI, desc_lo(smem_u32(smem + G::OFF_ONES), 2048));
        int i = 0;
        long long tl_wait = 0;
        const long long tl_loop0 = tl_on ? clock64() : 0;
        for (int u = u_first; u < p.n_units; u += u_step, ++i) {
            const int sg = i % G::NSTAGE;
            const int tile = u % tiles;
            const int h0 = (tile / p.tiles_w) * G::TH;
            const int npair = (min(G::TH, p.Hc - h0) + 1) >> 1;
            const uint32_t a_base = desc_lo(smem_u32(smem + sg * G::BUF), 128);                      // LBO: next 8 positions (K)
            const uint32_t b_base = desc_lo(smem_u32(smem + sg * G::BUF + G::X_BYTES), 128);         // LBO: next 8 columns (K)
            TL_TIMED(tl_on, tl_wait, mbar_wait(full + sg, (i / G::NSTAGE) & 1));
            fence_after();
            if (tl_on && i == 0 && warp == 0 && elected) tl[TL_FIRST_MMA] = clock64() - tl_entry;
            if (elected) {
                const uint32_t acc0 = i == 0 ? 0u : 1u;
                if (roleB) {
                    if (p.f16) wg2p_issue_unit<G, true, true>(tmem_base, a_base, b_base, ones_desc, acc0, npair);
                    else wg2p_issue_unit<G, true, false>(tmem_base, a_base, b_base, ones_desc, acc0, npair);
                } else {
                    if (p.f16) wg2p_issue_unit<G, false, true>(tmem_base, a_base, b_base, ones_desc, acc0, npair);
                    else wg2p_issue_unit<G, false, false>(tmem_base, a_base, b_base, ones_desc, acc0, npair);
                }
                commit_2cta(empty + sg);
            }
            __syncwarp();
        }
        if (elected) {
            commit_2cta(done);
            if (tl_on) tl[warp == 0 ? TL_ISS_A_FULL : TL_ISS_W_FULL] = tl_wait, tl[warp == 0 ? TL_ISS_LOOP : TL_EPI_LOOP] = clock64() - tl_loop0;
        }
        __syncwarp();
    }
    // ---- drain.  Lane m = (dy, ci) of accumulator a; its 128 columns are [parity 2][part 2][co 32]:
    //        parity 0 (even rows): kernel row 4g + dy      -> slot dy + 1
    //        parity 1 (odd rows) : kernel row 4g + dy - 1  -> slot dy
    //      (slot s <-> kh = 4g - 1 + s).  Two passes through the idle tile buffers: parity 0 stores slots 1..4, then parity 1 stores
    //      slot 0 and adds to slots 1..3; each (co, ci, a, slot) is touched by one thread per pass.
    float *out = p.partial + (size_t)blockIdx.x * G::BLK;
    mbar_wait_relaxed(done, 0, 1000);   // 13+ idle warps: sleep between polls (see tc_ptx.cuh)
    fence_after();
    __syncthreads();
    const long long tl_drain0 = tl_on ? clock64() : 0;
    {
        const int q = warp & 3;                                          // TMEM lane quarter of this warp = dy
        const int ci = lane;
        float *stg = reinterpret_cast<float *>(smem);
        const int nacc = roleB ? 3 : 4;
        const bool any = u_first < p.n_units;                            // a CTA without units holds no accumulators
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            for (int a = (warp >> 2); a < nacc; a += 4) {
                const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + a * G::ACC_COLS + 64 * pass;
                uint32_t v[32], v2[32];
                ld32(ta, v);
                ld32(ta + 32, v2);
#pragma unroll
                for (int co = 0; co < 32; ++co) {
                    float val = any ? __uint_as_float(v[co]) + __uint_as_float(v2[co]) : 0.f;
                    if (p.f16) val *= p.unscale;
                    float *sp = stg + (co * G::CIN + ci) * G::PITCH + a * G::SLOTS + q + 1 - pass;
                    *sp = (pass == 0 || q == 0) ? val : *sp + val;
                }
            }
            __syncthreads();
        }
        // bias gradient (role B, leader): every lane of the spare accumulator holds the column sums of B over all row pairs
        if (warp == 0) {
            float bsum = 0.f;
            if (roleB && rank == 0 && any) {
                const uint32_t ta = tmem_base + 3 * G::ACC_COLS;
#pragma unroll
                for (int blk = 0; blk < 4; ++blk) {
                    uint32_t c[32];
                    ld32(ta + 32 * blk, c);
#pragma unroll
                    for (int k = 0; k < 32; ++k)
                        if (k == lane) bsum += __uint_as_float(c[k]);
                }
            }
            out[G::NW_BLK + lane] = p.f16 ? bsum * p.unscale_g : bsum;
        }
        fence_before();
        __syncthreads();
        for (int e = tid; e < G::NW_BLK; e += G::NT) {
            const int cc = e / G::PER_CC, r = e - cc * G::PER_CC;
            out[e] = (r < nacc * G::SLOTS) ? stg[cc * G::PITCH + r] : 0.f;
        }
    }
    __syncthreads();
    cluster_sync();                                              // the peer has drained too: nobody reads this TMEM any more
    if (warp == 2) tmem_dealloc2(tmem_base, (uint32_t)G::TMEM_COLS);
    if (tl_on && tid == 0) tl[TL_DRAIN] = clock64() - tl_drain0, tl[TL_TOTAL] = clock64() - tl_entry, tl[TL_T_EXIT] = gtimer_ns();
}

// Tensor maps of the two operand images (cached per buffer and geometry, tmap.cu).  Dimension ORDER = order of the box in
// shared memory; the row-parity / row-pair split of the g_u rows needs an even Hc, global strides multiples of 16 bytes.
static int wg2_tile() {
    static int th = -1;
    if (th < 0) {
        const char *e = getenv("DCLL_WG2_TILE");
        th = (e && atoi(e) == 16) ? 16 : 8;
    }
    return th;
}

// DCLL_WG2_PAIR=0: the single-CTA kernel (A/B measurements); default: CTA pairs (cta_group::2) with 8-row units
static bool wg2_pair() {
    static int on = -1;
    if (on < 0) {
        const char *e = getenv("DCLL_WG2_PAIR");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on != 0 && wg2_tile() == 8;
}

static bool wg2_g16(const dcll_conv_layer *L) { return (geo_of(L).Wc % 16) == 0; }

template <class G>
static bool wg2_tmaps_g(const dcll_conv_layer *L, TmapDesc *tmx, TmapDesc *tmg, bool pair = false) {
    Geo g = geo_of(L);
    const uint64_t hw16 = (uint64_t)L->H * L->W * 16, plane = (uint64_t)g.Hc * g.Wc * 2;
    const uint64_t xd[5] = {8, (uint64_t)L->W, 4, (uint64_t)L->H, (uint64_t)2 * L->B};
    const uint64_t xs[4] = {16, hw16, (uint64_t)L->W * 16, 4 * hw16};
    const uint32_t xb[5] = {8, (uint32_t)G::XCOLS, 4, (uint32_t)G::XROWS, prec_f16(L) ? 1u : 2u};   // F16X2: the fp16 part only
    // g_u image [b][part][co/8][position/8][co % 8][8 positions]: 128 bytes per (channel group, 8 positions).  The two chunks of
    // a 16-column unit row are adjacent in memory, so with Wc % 16 == 0 they are ONE 256-byte box row (the TMA unit works row by
    // row: 32 instead of 64 rows per half tile; with fp16 traces the producer, not the MMAs, bounded the kernel).
    const bool g16 = wg2_g16(L);
    const uint64_t gd[5] = {g16 ? 128u : 64u, (uint64_t)g.Wc / (g16 ? 16 : 8), 2, (uint64_t)8 * L->B, (uint64_t)g.Hc / 2};
    const uint64_t gs[4] = {g16 ? 256u : 128u, (uint64_t)g.Wc * 16, plane * 8, (uint64_t)g.Wc * 32};
    const uint32_t gb[5] = {g16 ? 128u : 64u, g16 ? 1u : 2u, pair ? 1u : 2u, 8, (uint32_t)G::PAIRS};   // pair: one row parity per CTA
    return tmap_bf16(tmx, L->eps1_mma, 5, xd, xs, xb) && tmap_bf16(tmg, L->g_u, 5, gd, gs, gb);
}
static bool wg2_tmaps(const dcll_conv_layer *L, TmapDesc *tmx, TmapDesc *tmg) {
    return wg2_tile() == 16 ? wg2_tmaps_g<Wg2GeoT<16, 2>>(L, tmx, tmg) : wg2_tmaps_g<Wg2GeoT<8, 4>>(L, tmx, tmg, wg2_pair());
}

// The row-pair kernel takes the layer when its operands exist in image form -- the tensor-core forward wrote eps1_mma, and the
// packed backward read-out can write g_u as bf16 planes (even F, rows of 8 columns 16-byte aligned, even Hc) -- and the driver
// accepts the tensor maps.  The writer of g_u (launch_readout_bwd) and its reader (launch_wgrad) both ask this function.
bool wgrad_tc2_supported(const dcll_conv_layer *L) {
    static int on = -1;                                     // DCLL_WGRAD_TC2=0: keep wgrad_tc_kernel (A/B measurements)
    if (on < 0) {
        const char *e = getenv("DCLL_WGRAD_TC2");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    if (!on) return false;
    Geo g = geo_of(L);
    if (!(prec_tc(L) && tc_supported(L) && L->Cin == 32 && L->eps1_mma && L->g_u && (g.Wc % 8) == 0 && (g.Hc % 2) == 0 &&
          L->K <= 32 && (((uintptr_t)L->pv | (uintptr_t)L->wo | (uintptr_t)L->g_u | (uintptr_t)L->eps1_mma) % 16) == 0))
        return false;
    TmapDesc tmx, tmg;
    return wg2_tmaps(L, &tmx, &tmg);
}

void wgrad_tc2_roles(const dcll_conv_layer *L, int *nA, int *nB) {
    Geo g = geo_of(L);
    const int n_units = L->B * ceil_div(g.Hc, wg2_tile()) * ceil_div(g.Wc, Wg2Geo::TW);
    // 74 CTA pairs: role A does 4 kernel columns per unit (4 x 4 x 113 cycles), role B 3 + half of the bias MMAs (4 x 3 x 113 + 2 x 64)
    const int pairs = sm_budget() / 2;                       // 74 unless the data-parallel driver holds SMs back for NCCL
    // (CTA pairs: 4 x 4 x 104 against 4 x 3 x 104 + 4 x 64 cycles plus the tile writes -> 39 : 35; DCLL_WG2_NA overrides, for tuning)
    static int na_env = -1;
    if (na_env < 0) {
        const char *e = getenv("DCLL_WG2_NA");
        na_env = e ? atoi(e) : 0;
    }
    // (F16X2: no lo products -- 16 N = 128 MMAs per unit in both roles, the bias MMAs included: 37 : 37)
    const int na74 = na_env > 0 && na_env < 74 ? na_env : (prec_f16(L) ? 37 : (wg2_pair() ? 39 : 40));
    if (n_units >= 40) *nA = (pairs * na74 + 37) / 74, *nB = pairs - *nA;
    else *nA = *nB = n_units < pairs / 2 ? n_units : pairs / 2;
}

template <class G>
static int launch_wgrad_tc2_g(const dcll_conv_layer *L, float *partial, int *nA_out, int *nB_out, cudaStream_t st) {
    Geo g = geo_of(L);
    Wg2P p;
    p.partial = partial;
    p.B = L->B, p.H = L->H, p.W = L->W, p.padH = L->padH, p.padW = L->padW, p.Hc = g.Hc, p.Wc = g.Wc;
    p.tiles_h = ceil_div(g.Hc, G::TH), p.tiles_w = ceil_div(g.Wc, G::TW);
    p.n_units = L->B * p.tiles_h * p.tiles_w;
    wgrad_tc2_roles(L, &p.nA, &p.nB);
    *nA_out = p.nA, *nB_out = p.nB;
    static int dbg = -1;
    if (dbg < 0) {
        const char *e = getenv("DCLL_WG2_DEBUG");
        dbg = e ? atoi(e) : 0;
    }
    p.dbg = dbg;
    p.f16 = prec_f16(L) ? 1 : 0, p.g16 = wg2_g16(L) ? 1 : 0;
    p.unscale = p.f16 ? pow2i(-(L->a_exp + L->g_exp)) : 1.f, p.unscale_g = p.f16 ? pow2i(-L->g_exp) : 1.f;
    DCLL_REQUIRE(!p.f16 || (abs(L->a_exp + L->g_exp) <= 120 && abs(L->g_exp) <= 120), DCLL_EINVAL, "f16x2: operand exponents out of range");
    p.tl = timeline_buf(TL_WGRAD2);
    TmapDesc tmx, tmg;
    DCLL_REQUIRE(wg2_tmaps(L, &tmx, &tmg), DCLL_ECUDA, "wgrad_tc2: cuTensorMapEncodeTiled failed");
    DCLL_SMEM_ATTR((wgrad_tc2_kernel<G::TH, G::NSTAGE>), G::SMEM);
    launch_k(wgrad_tc2_kernel<G::TH, G::NSTAGE>, 2 * (p.nA + p.nB), G::NT, G::SMEM, st, p, tmx, tmg);
    DCLL_LAUNCH_OK("wgrad_tc2_kernel");
    return DCLL_OK;
}

static int launch_wgrad_tc2_pair(const dcll_conv_layer *L, float *partial, int *nA_out, int *nB_out, cudaStream_t st) {
    using G = Wg2PairGeoT<8, 6>;             // six stages fit once the g_u half tile is 8 KB (36 KB per stage)
    Geo g = geo_of(L);
    Wg2P p;
    p.partial = partial;
    p.B = L->B, p.H = L->H, p.W = L->W, p.padH = L->padH, p.padW = L->padW, p.Hc = g.Hc, p.Wc = g.Wc;
    p.tiles_h = ceil_div(g.Hc, G::TH), p.tiles_w = ceil_div(g.Wc, G::TW);
    p.n_units = L->B * p.tiles_h * p.tiles_w;
    wgrad_tc2_roles(L, &p.nA, &p.nB);
    *nA_out = p.nA, *nB_out = p.nB;
    p.dbg = 0;
    p.f16 = prec_f16(L) ? 1 : 0, p.g16 = wg2_g16(L) ? 1 : 0;
    p.unscale = p.f16 ? pow2i(-(L->a_exp + L->g_exp)) : 1.f, p.unscale_g = p.f16 ? pow2i(-L->g_exp) : 1.f;
    DCLL_REQUIRE(!p.f16 || (abs(L->a_exp + L->g_exp) <= 120 && abs(L->g_exp) <= 120), DCLL_EINVAL, "f16x2: operand exponents out of range");
    p.tl = timeline_buf(TL_WGRAD2);
    TmapDesc tmx, tmg;
    DCLL_REQUIRE(wg2_tmaps(L, &tmx, &tmg), DCLL_ECUDA, "wgrad_tc2: cuTensorMapEncodeTiled failed");
    DCLL_SMEM_ATTR((wgrad_tc2p_kernel<8, 6>), G::SMEM);
    launch_k_pair(wgrad_tc2p_kernel<8, 6>, 2 * (p.nA + p.nB), G::NT, G::SMEM, st, p, tmx, tmg);
    DCLL_LAUNCH_OK("wgrad_tc2p_kernel");
    return DCLL_OK;
}

int launch_wgrad_tc2(const dcll_conv_layer *L, float *partial, int *nA_out, int *nB_out, cudaStream_t st) {
    if (wg2_pair()) return launch_wgrad_tc2_pair(L, partial, nA_out, nB_out, st);
    return wg2_tile() == 16 ? launch_wgrad_tc2_g<Wg2GeoT<16, 2>>(L, partial, nA_out, nB_out, st)
                            : launch_wgrad_tc2_g<Wg2GeoT<8, 4>>(L, partial, nA_out, nB_out, st);
}

}  // namespace dcll
