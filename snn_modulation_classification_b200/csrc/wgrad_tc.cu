// Local weight gradient of a Conv2dDCLLlayer on tcgen05 (split-bf16 x3), 7x7 kernels, 32 -> 32 channels.
//
//     gW[co,ci,kh,kw] = sum_{b,h,w} g_u[b,co,h,w] * eps1[b,ci,h+kh-pad,w+kw-pad]           (see wgrad.cu)
//
// As a GEMM the reduction (K) runs over positions and the output is tiny, so ALL of gW stays resident in tensor memory
// while a persistent CTA streams its share of the positions through shared memory:
//
//   D_{g,kw}[(dy,ci), co] += sum_{16 columns} X[row+4g+dy][col+kw][ci] * G[row][col][co]       g in {0,1}, dy in 0..3
//
// * M = 128 = 4 kernel rows x 32 input channels.  With the eps1 halo tile staged as [halo row][ci/8][halo col][8 ci]
//   (row pitch = 4 x channel-group pitch) the M index (dy, ci/8) has ONE uniform stride, so a single MN-major
//   no-swizzle descriptor addresses four kernel rows at once; the kernel column kw is a 16-byte start-address shift.
// * N = 32 output channels, K = 16 consecutive columns of one output row per MMA; G staged as [co/8][row][col][8 co].
// * bf16 hi/lo split of both operands, products hi*hi + lo*hi + hi*lo, FP32 accumulation.  As in the forward kernel an
//   N = 32 MMA is bound by re-reading its A tile from shared memory, so [G_hi | G_lo] are ONE operand with N = 64
//   (X_hi is read once for two products) and X_lo x G_hi is the second MMA.  64-column accumulators x 7 kernel columns
//   = 448 of the 512 TMEM columns per kernel-row group, hence a CTA owns ONE group (g = blockIdx & 1; the 8th kernel
//   row of group 1 is padding) and CTA pairs walk the same tiles.
// * Persistent and warp-specialised: 12 loader warps fill one half of a double buffer with the next (eps1, g_u) tile
//   (eps1 straight from the bf16 operand image the forward wrote, by cp.async; g_u converted fp32 -> bf16 hi/lo) while
//   two issuer warps (kernel columns 0..3 / 4..6) issue the MMAs of the current tile; mbarriers (full: loader
//   arrivals, empty: tcgen05.commit of both issuers).
// Partials (one per CTA pair) are reduced in fixed order by reduce_adam_kernel, exactly like the FP32 path.
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace dcll {

struct WgTcP {
    const float *g_u;   // [B,32,Hc,Wc]
    const float *eps1;  // [B,32,H,W] (state after the forward step)
    const uint4 *img;   // bf16 {hi,lo} operand image of eps1 written by the tensor-core forward ([b][part][ci/8][H][W][8]), or null
    float *partial;     // [S][nW + Cout]
    int B, H, W, padH, padW, Hc, Wc;
    int tiles_h, tiles_w, n_units, n_tot, nW;
};

// CIN = 32: M = 4 kernel rows x 32 channels, two row groups (g = 0,1), one accumulator per kernel column.
// CIN = 1 (layer 0): the 8 slots of a 16-byte piece hold the 8 column shifts eps1[row][col .. col+7] instead of 8
// channels, so M = 16 kernel rows x 8 kernel columns (rows 7..15 and column 7 are padding) and ONE MMA pair per
// output row covers every tap: 32 MMAs per tile instead of 224; the kernel is bound by streaming g_u.
template <int CIN_>
struct WgTcGeoT {
    static constexpr int KH = 7, KW = 7, CIN = CIN_, COUT = 32;
    static constexpr int TH = 16, TW = 16;                   // one 16-column K chunk per output row
    static constexpr int CGR = CIN == 32 ? 4 : 1;            // channel groups per halo row
    static constexpr int NG = CIN == 32 ? 2 : 1;             // kernel-row groups; a CTA handles ONE (g = blockIdx & 1)
    static constexpr int DY = 128 / (CGR * 8);               // kernel rows per group (incl. padding)
    static constexpr int NACC = CIN == 32 ? KW : 1;          // accumulators per CTA, each 128 x ACC_COLS fp32
    static constexpr int ACC_COLS = 2 * COUT;                // [X_hi*G_hi + X_lo*G_hi | X_hi*G_lo], summed when draining
    static constexpr int XROWS = TH + DY - 1;
    static constexpr int XCOLS = CIN == 32 ? TW + KW - 1 : TW;   // 22; CIN = 1: the column shifts live inside the pieces
    static constexpr int X_CP = XCOLS * 16;                  // channel-group pitch (bytes)
    static constexpr int X_RP = CGR * X_CP;                  // halo-row pitch
    static constexpr int X_PART = XROWS * X_RP;              // one of {hi,lo}
    static constexpr int G_ROW = TW * 16, G_PLANE = TH * G_ROW, G_PART = 4 * G_PLANE;
    static constexpr int X_BYTES = 2 * X_PART, G_BYTES = 2 * G_PART;
    static constexpr int BUF = X_BYTES + G_BYTES;            // one pipeline stage
    static constexpr int NT = 512;
    static constexpr int LOADER_WARPS = 12;                  // warps 4..15
    // CIN = 1 with the forward's operand image (the shipped case): g_u is STREAMED -- 134 MB per launch against 32 MMAs per unit --
    // so the loaders keep two units of raw fp32 g_u in flight in a thread-private cp.async ring (RAW_UNIT bytes per unit: 24
    // values per loader thread) and the small X tiles in a ring of 4 slots, requested two units ahead (slots 0, 1 are the X
    // areas of the two stage buffers, slots 2, 3 extra).  With the
    // values prefetched into REGISTERS (first version) the six scoreboards of a warp alias: the first conversion of a unit
    // waited for the loads of the NEXT unit issued just before it (ncu: 28 % of the samples on that F2F, 1.6 TB/s).
    static constexpr int XRING = CIN == 1 ? 2 : 0, RAW_UNIT = CIN == 1 ? 24 * LOADER_WARPS * 32 * 4 : 0, RAW_SLOTS = 2;
    static constexpr int XRING_OFF = 2 * BUF + 128 + 16 * 8 * 4, RAW_OFF = XRING_OFF + XRING * X_BYTES;
    static constexpr int SMEM = RAW_OFF + RAW_SLOTS * RAW_UNIT;
    static constexpr int NISS_ACC = CIN == 1 ? 2 : 1;        // CIN = 1: the two issuer warps take alternate rows into accumulators of their own
    static constexpr int TMEM_COLS = NACC * NISS_ACC * ACC_COLS <= 256 ? 256 : 512;
    static_assert(SMEM <= 227 * 1024, "shared memory");
    __host__ __device__ static constexpr int xslot(int i) { return (i & 3) < 2 ? (i & 3) * BUF : XRING_OFF + ((i & 3) - 2) * X_BYTES; }
    static_assert(G_PART == 4 * G_PLANE, "{hi,lo} x channel-group must be uniformly strided for the N = 2*Cout operand");
    static_assert(CIN == 32 || CIN == 1, "instantiated for 32 and 1 input channels");
};

// Persistent, warp-specialised: warps 4..15 stage (eps1, g_u) tiles of unit i+1 into the free half of a double buffer
// while the elected lane of warp 0 issues the MMAs of unit i; tcgen05.commit hands buffers back to the loaders.
template <int CIN_>
__global__ void __launch_bounds__(512, 1) wgrad_tc_kernel(const WgTcP p) {
    using G = WgTcGeoT<CIN_>;
    using namespace tc;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 2 * G::BUF);
    uint64_t *full = bars, *empty = bars + 2, *done = bars + 4;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 6);
    float *bias_red = reinterpret_cast<float *>(smem + 2 * G::BUF + 128);   // [16 warps][8]
    uint32_t *iss_used = tmem_slot + 2;                                      // [2]: issuer warp w has written its accumulator
    const bool ring = G::CIN == 1 && p.img != nullptr;                       // streamed g_u (see WgTcGeoT)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) mbar_init(full + i, G::LOADER_WARPS), mbar_init(empty + i, 2);
        mbar_init(done, 2);
        iss_used[0] = iss_used[1] = 0;
        mbar_fence_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, (uint32_t)G::TMEM_COLS);
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_entry();   // barrier init and the TMEM allocation above overlap the previous grid's tail; no global access before here

    const int tiles = p.tiles_h * p.tiles_w;
    // CTA pairs (2k, 2k+1) walk the same units; each CTA accumulates one kernel-row group
    const int grp = G::NG == 2 ? (blockIdx.x & 1) : 0;
    const int u_first = G::NG == 2 ? (blockIdx.x >> 1) : blockIdx.x;
    const int u_step = G::NG == 2 ? (gridDim.x >> 1) : gridDim.x;
    const int row_off = G::DY * grp;                          // first halo row of this group relative to the tile's halo
    float gsum[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) gsum[k] = 0.f;

    if (warp >= 4) {
        // ================= loaders =================
        const float *__restrict__ gg = p.g_u;
        const float *__restrict__ ge = p.eps1;
        const int cgw = warp & 3;                 // channel group (of eps1 and of g_u) this warp stages
        const int l96 = ((warp >> 2) - 1) * 32 + lane;   // 0..95 within the group
        const size_t xcs = (size_t)p.H * p.W, gcs = (size_t)p.Hc * p.Wc;
        static_assert(G::TH * G::TW <= 3 * 96, "three positions per loader thread and channel group");
        // g_u values of one unit for this thread: positions l96, l96 + 96, l96 + 192 of the 16 x 16 tile x 8 channels
        auto load_g = [&](int u, float (&v)[3][8]) {
            const int b = u / tiles, tile = u - b * tiles;
            const int th_i = tile / p.tiles_w, tw_i = tile - th_i * p.tiles_w;
#pragma unroll
            for (int h = 0; h < 3; ++h) {
                const int itt = l96 + h * 96;
                const int r = itt / G::TW, c = itt - r * G::TW;
                const int oh = th_i * G::TH + r, ow = tw_i * G::TW + c;
                const bool ok = itt < G::TH * G::TW && oh < p.Hc && ow < p.Wc;
                const size_t off = ok ? ((size_t)(b * G::COUT + cgw * 8) * p.Hc + oh) * p.Wc + ow : 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) v[h][k] = ok ? __ldg(gg + off + k * gcs) : 0.f;
            }
        };
        if (G::CIN == 1 && ring) {
            // Lean instruction stream (the loaders, not DRAM, paced the first version: ~820 instructions per warp and unit, most of
            // them address arithmetic): everything that does not depend on the unit is computed once per thread.
            const int lt = tid - 128;                                        // 0..383
            float *raw = reinterpret_cast<float *>(smem + G::RAW_OFF);
            const uint32_t hw = (uint32_t)(p.H * p.W), gcs32 = (uint32_t)(p.Hc * p.Wc);
            // X tile: only the halo rows 0..21 feed real kernel rows (dy <= 6); piece (row xr, column xc) of both parts
            constexpr int XR_USED = G::TH + G::KH - 1;
            const int xr = lt >> 4, xc = lt & 15;
            const bool x_mine = xr < XR_USED;
            const uint32_t x_dst = (uint32_t)(xr * G::X_RP + xc * 16);
            // g_u: positions l96 + 96 h of the tile, 8 channels each (element offsets relative to the tile origin fit 32 bits)
            int g_r[3], g_c[3];
            uint32_t g_off[3], g_dst[3];
#pragma unroll
            for (int h = 0; h < 3; ++h) {
                const int itt = l96 + h * 96;
                g_r[h] = itt >> 4, g_c[h] = itt & 15;
                g_off[h] = (uint32_t)((cgw * 8 * p.Hc + g_r[h]) * p.Wc + g_c[h]);
                g_dst[h] = (uint32_t)(cgw * G::G_PLANE + g_r[h] * G::G_ROW + g_c[h] * 16);
            }
            const bool h2_mine = l96 + 192 < G::TH * G::TW;                  // the third position exists for l96 < 64 only
            // group C_i = { X tile of the CTA's i-th unit -> X slot i & 3,  raw g_u of that unit -> raw slot i & 1 }
            auto issue = [&](int u, int i) {
                if (u < p.n_units) {
                    const unsigned b = (unsigned)u / (unsigned)tiles, tile = (unsigned)u - b * (unsigned)tiles;
                    const unsigned th_i = tile / (unsigned)p.tiles_w, tw_i = tile - th_i * (unsigned)p.tiles_w;
                    const int h0 = (int)th_i * G::TH, w0 = (int)tw_i * G::TW;
                    if (x_mine) {
                        const int gh = h0 - p.padH + xr, gw = w0 + xc;
                        const bool in = gh >= 0 && gh < p.H && gw < p.W;
                        const uint4 *src = in ? p.img + (size_t)b * 2 * hw + (uint32_t)(gh * p.W + gw) : p.img;
                        const uint32_t dst = smem_u32(smem + G::xslot(i)) + x_dst;
                        const uint32_t sz = in ? 16u : 0u;
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + G::X_PART), "l"(in ? src + hw : src), "r"(sz) : "memory");
                    }
                    const uint32_t rs = smem_u32(raw + (size_t)(i & 1) * (G::RAW_UNIT / 4) + lt);
                    const float *tile0 = gg + ((size_t)b * G::COUT * p.Hc + h0) * p.Wc + w0;
#pragma unroll
                    for (int h = 0; h < 3; ++h) {
                        if (h == 2 && !h2_mine) break;
                        const bool ok = h0 + g_r[h] < p.Hc && w0 + g_c[h] < p.Wc;
                        const float *src = ok ? tile0 + g_off[h] : gg;
                        const uint32_t step = ok ? gcs32 : 0u, sz = ok ? 4u : 0u;
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(rs + (h * 8 + k) * (G::LOADER_WARPS * 32 * 4)),
                                         "l"(src + k * step), "r"(sz)
                                         : "memory");
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            };
            issue(u_first, 0);
            issue(u_first + u_step, 1);
            int i = 0;
            for (int u = u_first; u < p.n_units; u += u_step, ++i) {
                const int buf = i & 1;
                unsigned char *sG = smem + buf * G::BUF + G::X_BYTES;
                if (i >= 2) mbar_wait(empty + buf, ((i >> 1) - 1) & 1);   // MMAs of unit i-2 done: this G half and X slot (i+2) & 3 are free
                asm volatile("cp.async.wait_group 1;" ::: "memory");      // C_i has landed (C_{i+1} may still be in flight)
                const float *rv = raw + (size_t)(i & 1) * (G::RAW_UNIT / 4) + lt;
#pragma unroll
                for (int h = 0; h < 3; ++h) {
                    if (h == 2 && !h2_mine) break;
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int k = 0; k < 8; k += 2) {
                        // two channels per conversion instruction (cvt.rn.bf16x2.f32: low half = first value)
                        const float v0 = rv[(h * 8 + k) * (G::LOADER_WARPS * 32)], v1 = rv[(h * 8 + k + 1) * (G::LOADER_WARPS * 32)];
                        gsum[k] += v0, gsum[k + 1] += v1;
                        const __nv_bfloat162 h2 = __floats2bfloat162_rn(v0, v1);
                        const uint32_t hb = *reinterpret_cast<const uint32_t *>(&h2);
                        const __nv_bfloat162 l2 = __floats2bfloat162_rn(v0 - __uint_as_float(hb << 16), v1 - __uint_as_float(hb & 0xffff0000u));
                        hi[k >> 1] = hb, lo[k >> 1] = *reinterpret_cast<const uint32_t *>(&l2);
                    }
                    *reinterpret_cast<uint4 *>(sG + g_dst[h]) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint4 *>(sG + g_dst[h] + G::G_PART) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(full + buf);
                issue(u + 2 * u_step, i + 2);                              // refills the raw slot this thread has just read
            }
        } else {
        float gv[3][8];
        if (u_first < p.n_units) load_g(u_first, gv);
        int i = 0;
        for (int u = u_first; u < p.n_units; u += u_step, ++i) {
            const int buf = i & 1;
            unsigned char *sX = smem + buf * G::BUF, *sG = sX + G::X_BYTES;
            const int b = u / tiles;
            const int tile = u - b * tiles;
            const int th_i = tile / p.tiles_w, tw_i = tile - th_i * p.tiles_w;
            const int h0 = th_i * G::TH, w0 = tw_i * G::TW;
            if (i >= 2) mbar_wait(empty + buf, ((i >> 1) - 1) & 1);   // MMAs of unit i-2 have finished reading this half
            // ---- eps1 halo tile: [row][cg][col][8 ci], bf16 hi | lo
            if (G::CIN == 32 && p.img) {
                // the forward already wrote eps1 as a bf16 {hi,lo} image in 16-byte (position, channel group) pieces:
                // asynchronous copies (zero fill outside the picture) that land while g_u is converted below
                const int l384 = (warp - 4) * 32 + lane;
                const size_t hw = (size_t)p.H * p.W;
                const uint4 *src0 = p.img + (size_t)b * 2 * G::CGR * hw;
                const uint32_t dst0 = smem_u32(sX);
                for (int idx = l384; idx < 2 * G::XROWS * G::CGR * G::XCOLS; idx += G::LOADER_WARPS * 32) {
                    const int c = idx % G::XCOLS;
                    int t = idx / G::XCOLS;
                    const int cg = t % G::CGR;
                    t /= G::CGR;
                    const int r = t % G::XROWS, part = t / G::XROWS;
                    const int gh = h0 - p.padH + row_off + r, gw = w0 - p.padW + c;
                    const bool in = gh >= 0 && gh < p.H && gw >= 0 && gw < p.W;
                    const uint4 *src = in ? src0 + (size_t)(part * G::CGR + cg) * hw + (size_t)gh * p.W + gw : src0;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + part * G::X_PART + r * G::X_RP + cg * G::X_CP + c * 16),
                                 "l"(src), "r"(in ? 16u : 0u)
                                 : "memory");
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            } else if (G::CIN == 1 && p.img) {
                // layer 0: the forward's image already holds the pieces eps1[y][x-3 .. x+4] = our piece (r, c) at x = w0 + c
                const int l384 = (warp - 4) * 32 + lane;
                const size_t hw = (size_t)p.H * p.W;
                const uint4 *src0 = p.img + (size_t)b * 2 * hw;
                const uint32_t dst0 = smem_u32(sX);
                for (int idx = l384; idx < 2 * G::XROWS * G::XCOLS; idx += G::LOADER_WARPS * 32) {
                    const int c = idx % G::XCOLS;
                    const int t = idx / G::XCOLS;
                    const int r = t % G::XROWS, part = t / G::XROWS;
                    const int gh = h0 - p.padH + r, gw = w0 + c;
                    const bool in = gh >= 0 && gh < p.H && gw < p.W;
                    const uint4 *src = in ? src0 + (size_t)part * hw + (size_t)gh * p.W + gw : src0;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + part * G::X_PART + r * G::X_RP + c * 16), "l"(src),
                                 "r"(in ? 16u : 0u)
                                 : "memory");
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            } else if (G::CIN == 32) {
                // two positions per iteration (16 loads in flight)
                for (int it = l96; it < G::XROWS * G::XCOLS; it += 2 * 96) {
                    float v[2][8];
                    int r_[2], c_[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int itt = it + h * 96;
                        r_[h] = itt / G::XCOLS, c_[h] = itt - r_[h] * G::XCOLS;
                        const int gh = h0 - p.padH + row_off + r_[h], gw = w0 - p.padW + c_[h];
                        const bool ok = itt < G::XROWS * G::XCOLS && gh >= 0 && gh < p.H && gw >= 0 && gw < p.W;
                        const size_t off = ok ? ((size_t)(b * G::CIN + cgw * 8) * p.H + gh) * p.W + gw : 0;
#pragma unroll
                        for (int k = 0; k < 8; ++k) v[h][k] = ok ? __ldg(ge + off + k * xcs) : 0.f;
                    }
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if (it + h * 96 >= G::XROWS * G::XCOLS) continue;
                        __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            hi[k] = __float2bfloat16_rn(v[h][k]);
                            lo[k] = __float2bfloat16_rn(v[h][k] - __bfloat162float(hi[k]));
                        }
                        unsigned char *dst = sX + r_[h] * G::X_RP + cgw * G::X_CP + c_[h] * 16;
                        *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(hi);
                        *reinterpret_cast<uint4 *>(dst + G::X_PART) = *reinterpret_cast<const uint4 *>(lo);
                    }
                }
            } else {
                // single input channel: piece (r, c) = the 8 column shifts eps1[r][c + 0..7] (slot 7 is padding)
                for (int it = (warp - 4) * 32 + lane; it < G::XROWS * G::XCOLS; it += G::LOADER_WARPS * 32) {
                    const int r = it / G::XCOLS, c = it - r * G::XCOLS;
                    const int gh = h0 - p.padH + r, gw0 = w0 - p.padW + c;
                    __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int gw = gw0 + k;
                        float v = 0.f;
                        if (k < G::KW && gh >= 0 && gh < p.H && gw >= 0 && gw < p.W) v = __ldg(ge + ((size_t)b * p.H + gh) * p.W + gw);
                        hi[k] = __float2bfloat16_rn(v);
                        lo[k] = __float2bfloat16_rn(v - __bfloat162float(hi[k]));
                    }
                    unsigned char *dst = sX + r * G::X_RP + c * 16;
                    *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(hi);
                    *reinterpret_cast<uint4 *>(dst + G::X_PART) = *reinterpret_cast<const uint4 *>(lo);
                }
            }
            // ---- g_u tile: [cog][row][col][8 co], bf16 hi | lo (zero outside the image).  The values of THIS unit were requested
            //      while the previous one was converted (gv), so their DRAM latency is covered by that work and by the wait for
            //      the free buffer above; the next unit's are requested now.  (Loading and converting in the same iteration left
            //      the loaders waiting on every load: ncu long-scoreboard stalls at the first conversion, 1.4 TB/s.)
            float gn[3][8];
            {
                const int un = u + u_step;
                if (un < p.n_units) load_g(un, gn);
            }
#pragma unroll
            for (int h = 0; h < 3; ++h) {
                const int itt = l96 + h * 96;
                if (itt >= G::TH * G::TW) continue;
                const int r = itt / G::TW, c = itt - r * G::TW;
                __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (grp == 0) gsum[k] += gv[h][k];
                    hi[k] = __float2bfloat16_rn(gv[h][k]);
                    lo[k] = __float2bfloat16_rn(gv[h][k] - __bfloat162float(hi[k]));
                }
                unsigned char *dst = sG + cgw * G::G_PLANE + r * G::G_ROW + c * 16;
                *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(hi);
                *reinterpret_cast<uint4 *>(dst + G::G_PART) = *reinterpret_cast<const uint4 *>(lo);
            }
#pragma unroll
            for (int h = 0; h < 3; ++h)
#pragma unroll
                for (int k = 0; k < 8; ++k) gv[h][k] = gn[h][k];
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            fence_async_smem();   // this thread's smem writes -> async proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(full + buf);
        }
        }
    } else if (warp < 2) {
        // ================= MMA issuer =================
        // a_major = b_major = MN (bits 15,16), bf16 x bf16 -> f32, N = 32, M = 128
        constexpr uint32_t IDESC_N2 = idesc_bf16(128, 2 * G::COUT, true, true);   // X_hi x [G_hi | G_lo]
        constexpr uint32_t IDESC_N1 = idesc_bf16(128, G::COUT, true, true);       // X_lo x G_hi
        constexpr uint32_t A_HI = desc_hi(G::X_CP);          // SBO: next 8 rows of M = next (dy, cg) group
        constexpr uint32_t B_HI = desc_hi(G::G_PLANE);       // SBO: next 8 columns of N = next ({hi,lo}, co/8) group
        const uint32_t elected = elect_one();
        // two issuer warps (a single issuing thread tops out at ~54 cycles per MMA, the tensor pipe at ~44 for these
        // short MMAs): warp 0 owns the accumulators of kernel columns 0..3, warp 1 those of 4..6
        const int kw0 = warp == 0 ? 0 : 4, kw1 = G::NACC == 1 ? (warp == 0 ? 1 : 0) : (warp == 0 ? 4 : G::KW);
        int i = 0;
        uint32_t iss_started = 0u;
        for (int u = u_first; u < p.n_units; u += u_step, ++i) {
            const int buf = i & 1;
            const int tile = u % tiles;
            const int h0 = (tile / p.tiles_w) * G::TH;
            const int rows = min(G::TH, p.Hc - h0);
            const uint32_t a_base = desc_lo(smem_u32(smem + (ring ? G::xslot(i) : buf * G::BUF)), 128);   // LBO: next 8 positions (K)
            const uint32_t b_base = desc_lo(smem_u32(smem + buf * G::BUF + G::X_BYTES), 128);
            mbar_wait(full + buf, (i >> 1) & 1);
            fence_after();
            if (G::CIN == 1) {
                // one MMA pair per output row: the two issuer warps take alternate rows into accumulators of their own (summed
                // when draining), so that the issue rate of one thread (~54 cycles per MMA) does not pace the stream
                if (elected) {
                    for (int r = warp; r < rows; r += 2) {
                        const uint64_t b = desc(B_HI, b_base + r * G::TW);
                        const uint32_t a_lo0 = a_base + ((r * G::X_RP) >> 4);
                        const uint32_t d = tmem_base + warp * G::ACC_COLS;
                        mma_bf16(d, desc(A_HI, a_lo0), b, IDESC_N2, iss_started);
                        mma_bf16(d, desc(A_HI, a_lo0 + (G::X_PART >> 4)), b, IDESC_N1, 1);
                        iss_started = 1u;
                    }
                    commit(empty + buf);
                }
            } else if (elected) {
                for (int r = 0; r < rows; ++r) {
                    const uint64_t b = desc(B_HI, b_base + r * G::TW);
                    const uint32_t acc = (i == 0 && r == 0) ? 0u : 1u;
#pragma unroll 4
                    for (int kw = kw0; kw < kw1; ++kw) {
                        const uint32_t a_lo0 = a_base + ((r * G::X_RP) >> 4) + kw;
                        const uint32_t d = tmem_base + kw * G::ACC_COLS;
                        mma_bf16(d, desc(A_HI, a_lo0), b, IDESC_N2, acc);
                        mma_bf16(d, desc(A_HI, a_lo0 + (G::X_PART >> 4)), b, IDESC_N1, 1);
                    }
                }
                commit(empty + buf);
            }
            __syncwarp();
        }
        if (elected) {
            if (G::CIN == 1) iss_used[warp] = iss_started;
            commit(done);
        }
        __syncwarp();
    }
    // ---- drain: once every MMA has completed (`done` flips), add the two accumulator halves and scatter into the
    //      [Cout,Cin,KH,KW] layout.  The two CTAs of a pair own complementary kernel rows and SHARE one partial block
    //      (every element is written by exactly one of them: no zero fill, half as many blocks to reduce).
    float *out = p.partial + (size_t)(G::NG == 2 ? blockIdx.x >> 1 : blockIdx.x) * p.n_tot;
    if (G::CIN == 1) __syncthreads();                                    // iss_used[] of both issuers is visible below
    mbar_wait(done, 0);
    fence_after();
    {
        const int q = warp & 3;                                          // TMEM lane quarter of this warp
        const int m = q * 32 + lane;                                     // (dy, ci) resp. (dy, kw)
        const int dy = G::CIN == 32 ? (m >> 5) : (m >> 3);
        const int ci = G::CIN == 32 ? (m & 31) : 0;
        const bool lane_ok = G::CIN == 32 ? true : ((m & 7) < G::KW);
        const int kh = G::DY * grp + dy;
        if (G::CIN == 32) {
            // A thread holds gW[co = 0..31][ci][kh][kw] for its (ci, kh): 196 bytes apart between lanes in the output.  The tile
            // buffers are idle now (all MMAs done), so the block is staged there as [co][ci][dy][kw] (pitch 29 floats per
            // (co,ci): conflict-free) and written out in runs of DY*KW contiguous floats.
            float *stg = reinterpret_cast<float *>(smem);
            constexpr int RUN = G::DY * G::KW, PITCH = RUN + 1;          // 28 values (+1 pad) per (co, ci)
            static_assert(G::COUT * G::CIN * PITCH * 4 <= 2 * G::BUF, "staging fits the tile buffers");
            for (int a = (warp >> 2); a < G::NACC; a += 4) {
                uint32_t v[32], v2[32];
                ld32(tmem_base + ((uint32_t)(q * 32) << 16) + a * G::ACC_COLS, v);
                ld32(tmem_base + ((uint32_t)(q * 32) << 16) + a * G::ACC_COLS + G::COUT, v2);
#pragma unroll
                for (int co = 0; co < 32; ++co)
                    stg[(co * G::CIN + ci) * PITCH + dy * G::KW + a] = __uint_as_float(v[co]) + __uint_as_float(v2[co]);
            }
            __syncthreads();
            const int rows_valid = min(G::DY, G::KH - G::DY * grp);      // kernel rows of this group that exist (4 or 3)
            const int run = rows_valid * G::KW;
            for (int i = tid; i < G::COUT * G::CIN * RUN; i += G::NT) {
                const int cc = i / RUN, e = i - cc * RUN;                // cc = co * CIN + ci
                if (e < run) out[(size_t)cc * (G::KH * G::KW) + G::DY * grp * G::KW + e] = stg[cc * PITCH + e];
            }
        } else {
            if (warp < 4) {
                // (row parity 0: hi + lo products) + (row parity 1: hi + lo products), in that fixed order
                float sum[32];
#pragma unroll
                for (int co = 0; co < 32; ++co) sum[co] = 0.f;
                for (int a = 0; a < G::NISS_ACC; ++a) {
                    if (!iss_used[a]) continue;                              // that issuer never had a row (single-row images)
                    uint32_t v[32], v2[32];
                    ld32(tmem_base + ((uint32_t)(q * 32) << 16) + a * G::ACC_COLS, v);
                    ld32(tmem_base + ((uint32_t)(q * 32) << 16) + a * G::ACC_COLS + G::COUT, v2);
#pragma unroll
                    for (int co = 0; co < 32; ++co) sum[co] += __uint_as_float(v[co]) + __uint_as_float(v2[co]);
                }
                const int kw = m & 7;
                if (kh < G::KH && lane_ok) {
#pragma unroll
                    for (int co = 0; co < 32; ++co) out[((size_t)(co * G::CIN + ci) * G::KH + kh) * G::KW + kw] = sum[co];
                }
            }
        }
    }
    // ---- bias gradient: sum of g_u over this CTA's positions (fixed-order in-CTA reduction)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float s = gsum[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) bias_red[warp * 8 + k] = s;
    }
    fence_before();
    __syncthreads();
    if (tid < 32) {
        const int cog = tid >> 3, k = tid & 7;                            // loader warps with (warp & 3) == cog staged this group
        float s = 0.f;
        for (int w = 4 + cog; w < 16; w += 4) s += bias_red[w * 8 + k];
        if (grp == 0) out[p.nW + tid] = s;
    }
    if (warp == 2) {
        tmem_dealloc(tmem_base, (uint32_t)G::TMEM_COLS);
    }
}

bool wgrad_tc_supported(const dcll_conv_layer *L) {
    // (F16X2 with 32 input channels: the fp16 trace image is only understood by wgrad_tc2; without it the FP32 kernel runs)
    return prec_tc(L) && !prec_f16(L) && L->KH == 7 && L->KW == 7 && (L->Cin == 32 || L->Cin == 1) && L->Cout == 32 &&
           L->poolH == 1 && L->poolW == 1;
}

int wgrad_tc_splits(const dcll_conv_layer *L) {
    Geo g = geo_of(L);
    int n_units = L->B * ceil_div(g.Hc, 16) * ceil_div(g.Wc, 16);
    const int sms = sm_budget();
    if (L->Cin == 32) return n_units < sms / 2 ? n_units : sms / 2;   // partial blocks = CTA pairs (one kernel-row group per CTA)
    return n_units < sms ? n_units : sms;
}

int launch_wgrad_tc(const dcll_conv_layer *L, float *partial, int S, cudaStream_t st) {
    Geo g = geo_of(L);
    WgTcP p;
    p.g_u = L->g_u, p.eps1 = L->eps1[L->cur & 1], p.partial = partial;
    p.img = tc_supported(L) ? reinterpret_cast<const uint4 *>(L->eps1_mma) : nullptr;
    p.B = L->B, p.H = L->H, p.W = L->W, p.padH = L->padH, p.padW = L->padW, p.Hc = g.Hc, p.Wc = g.Wc;
    p.tiles_h = ceil_div(g.Hc, 16), p.tiles_w = ceil_div(g.Wc, 16);
    p.n_units = L->B * p.tiles_h * p.tiles_w;
    p.nW = g.nW, p.n_tot = g.nW + L->Cout;
    DCLL_SMEM_ATTR(wgrad_tc_kernel<32>, WgTcGeoT<32>::SMEM);
    DCLL_SMEM_ATTR(wgrad_tc_kernel<1>, WgTcGeoT<1>::SMEM);
    if (L->Cin == 32) launch_k(wgrad_tc_kernel<32>, 2 * S, 512, WgTcGeoT<32>::SMEM, st, p);
    else launch_k(wgrad_tc_kernel<1>, S, 512, WgTcGeoT<1>::SMEM, st, p);
    DCLL_LAUNCH_OK("wgrad_tc_kernel");
    return DCLL_OK;
}

}  // namespace dcll
