// Local read-outs of a Conv2dDCLLlayer and the local-loss gradient.
//
// Replaces (reference, dcll/pytorch_libdcll.py):
//   :602-603  pvoutput = i2o(flatten(pool(pv)))                     -> readout_fwd_kernel (+ finish)
//   :605-606  output   = output_(flatten.detach())  (last layer)    -> same kernel, extra K rows
//   :694-697  SmoothL1Loss(pvoutput, target) [+ (output, target)]   -> readout_finish_kernel (gradient only)
//   :724-728  clout.append(argmax)                                  -> readout_finish_kernel (device side)
//   :704      autograd: d loss / d membrane through i2o, pool, sigmoid -> readout_bwd_kernel
//             d loss / d output_.{weight,bias} + optimizer2.step()  -> readout_bwd_kernel<.,true> (fused)
//
// The read-out is a skinny GEMM [B,F] x [F,Ktot] whose frozen matrix (4*K*F bytes, 50 MB per layer at
// 128x128) must be amortised over the batch, so it is a separate HBM-bound pass over pv rather than an
// epilogue of the convolution: every CTA keeps a 32-wide slice of Wo in shared memory and sweeps it
// over all samples.  Partials are reduced in a fixed order (no float atomics): results are
// deterministic run to run.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace dcll {

constexpr int RO_BM = 64;   // samples per tile
constexpr int RO_BK = 64;   // features per pipeline stage
constexpr int RO_PITCH = RO_BK + 4;   // floats; (pitch/4) odd -> conflict-free float4 reads across rows

__device__ __forceinline__ void ro_cp16(void *smem_dst, const void *gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void ro_zero16(void *smem_dst) { *reinterpret_cast<float4 *>(smem_dst) = make_float4(0.f, 0.f, 0.f, 0.f); }

// partial[blk][b][kt] = sum over this CTA's feature tiles of pv[b,f] * Wcat[kt,f],  Wcat = [wo ; wout]
// Two-stage cp.async pipeline: the (64 x 64) pv tile and the (Ktot x 64) read-out tile of stage s+1 stream into shared
// memory while stage s is multiplied; thread = 4 samples x KJ outputs, float4 along the feature axis.
// Requires F % 4 == 0 and 16-byte aligned rows (checked by the launcher; otherwise the scalar-load variant runs).
template <int KJ, bool VEC>
__global__ void __launch_bounds__(256, 3) readout_fwd_kernel(const float *__restrict__ pv, const float *__restrict__ wo,
                                                             const float *__restrict__ wout, int B, int F, int K, int Ktot,
                                                             float *__restrict__ partial) {
    pdl_entry();
    extern __shared__ __align__(16) float ro_smem[];
    constexpr int ROWS = RO_BM + 16 * KJ;                 // pv rows followed by read-out rows
    constexpr int STAGE = ROWS * RO_PITCH;
    const int tid = threadIdx.x;
    const int tb = tid & 15, tk = tid >> 4;
    const int n_ft = (F + RO_BK - 1) / RO_BK;

    auto issue = [&](int stage, int ft, int b0) {
        float *dst = ro_smem + stage * STAGE;
        const int f0 = ft * RO_BK;
        // ROWS x 16 chunks of 16 bytes
        for (int i = tid; i < ROWS * (RO_BK / 4); i += 256) {
            const int row = i / (RO_BK / 4), c4 = i - row * (RO_BK / 4);
            const int f = f0 + c4 * 4;
            float *d = dst + row * RO_PITCH + c4 * 4;
            const float *src = nullptr;
            if (row < RO_BM) {
                const int b = b0 + row;
                if (b < B) src = pv + (size_t)b * F + f;
            } else {
                const int k = row - RO_BM;
                if (k < Ktot) src = (k < K ? wo + (size_t)k * F : wout + (size_t)(k - K) * F) + f;
            }
            if (VEC) {
                if (src && f + 4 <= F) ro_cp16(d, src);
                else {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (src) {
                        if (f < F) v.x = src[0];
                        if (f + 1 < F) v.y = src[1];
                        if (f + 2 < F) v.z = src[2];
                    }
                    *reinterpret_cast<float4 *>(d) = v;
                }
            } else {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (src) {
                    if (f < F) v.x = __ldg(src);
                    if (f + 1 < F) v.y = __ldg(src + 1);
                    if (f + 2 < F) v.z = __ldg(src + 2);
                    if (f + 3 < F) v.w = __ldg(src + 3);
                }
                *reinterpret_cast<float4 *>(d) = v;
            }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };

    // grid.y strides over 64-sample tiles (small F / large B: not enough feature tiles to fill the machine)
    for (int b0 = blockIdx.y * RO_BM; b0 < B; b0 += gridDim.y * RO_BM) {
        float acc[4][KJ];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < KJ; ++j) acc[i][j] = 0.f;
        int ft = blockIdx.x, stage = 0;
        if (ft < n_ft) issue(0, ft, b0);
        for (; ft < n_ft; ft += gridDim.x, stage ^= 1) {
            const int nxt = ft + gridDim.x;
            if (nxt < n_ft) {
                issue(stage ^ 1, nxt, b0);
                asm volatile("cp.async.wait_group 1;\n" ::: "memory");
            } else {
                asm volatile("cp.async.wait_group 0;\n" ::: "memory");
            }
            __syncthreads();
            const float *pvs = ro_smem + stage * STAGE;
            const float *wos = pvs + RO_BM * RO_PITCH;
#pragma unroll 4
            for (int f4 = 0; f4 < RO_BK; f4 += 4) {
                float4 a[4], w[KJ];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4 *>(pvs + (tb + 16 * i) * RO_PITCH + f4);
#pragma unroll
                for (int j = 0; j < KJ; ++j) w[j] = *reinterpret_cast<const float4 *>(wos + (tk + 16 * j) * RO_PITCH + f4);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < KJ; ++j) {
                        acc[i][j] = fmaf(a[i].x, w[j].x, acc[i][j]);
                        acc[i][j] = fmaf(a[i].y, w[j].y, acc[i][j]);
                        acc[i][j] = fmaf(a[i].z, w[j].z, acc[i][j]);
                        acc[i][j] = fmaf(a[i].w, w[j].w, acc[i][j]);
                    }
            }
            __syncthreads();   // everyone is done with this stage before it is refilled two iterations later
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int b = b0 + tb + 16 * i;
#pragma unroll
            for (int j = 0; j < KJ; ++j) {
                int k = tk + 16 * j;
                if (b < B && k < Ktot) partial[((size_t)blockIdx.x * B + b) * Ktot + k] = acc[i][j];
            }
        }
    }
}

// One CTA per sample: fixed-order reduction of the partials, + bias, loss gradient, argmax.
// blockDim.x / 64 slices walk the partial blocks (4 after the FP32 kernel, 16 after the tcgen05 kernel's many blocks).
__global__ void __launch_bounds__(1024) readout_finish_kernel(const float *__restrict__ partial, int n_part, int B, int K,
                                                             int Ktot, const float *__restrict__ bo,
                                                             const float *__restrict__ bout, const float *__restrict__ target,
                                                             int loss_kind, float *__restrict__ pvoutput,
                                                             float *__restrict__ output, float *__restrict__ g_o,
                                                             float *__restrict__ g_o2, int32_t *__restrict__ clout,
                                                             float *__restrict__ loss_out) {
    pdl_entry();
    __shared__ float red[16][64];
    __shared__ float vals[64];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int kk = tid & 63, sl = tid >> 6, n_sl = blockDim.x >> 6;
    float s = 0.f;
    if (kk < Ktot) {
        // FIN_U loads in flight per thread, added in the same fixed order as a serial walk (a rolled loop paid one L2 round trip
        // per partial block: ~8 us per launch, three launches per timestep on the critical path)
        constexpr int FIN_U = 10;
        const float *src = partial + (size_t)b * Ktot + kk;
        const size_t pitch = (size_t)B * Ktot;
        // (unconditional loads -- past the end the slice's last block is read again -- so that they are issued back to back)
        const int c_hi = sl + n_sl * ((n_part - 1 - sl) / n_sl);
        for (int c0 = sl; c0 < n_part; c0 += n_sl * FIN_U) {
            float t[FIN_U];
#pragma unroll
            for (int u = 0; u < FIN_U; ++u) t[u] = src[(size_t)min(c0 + u * n_sl, c_hi) * pitch];
#pragma unroll
            for (int u = 0; u < FIN_U; ++u)
                if (c0 + u * n_sl < n_part) s += t[u];
        }
    }
    red[sl][kk] = s;
    __syncthreads();
    float lsum = 0.f;
    if (tid < Ktot) {
        const bool second = tid >= K;
        const int k = second ? tid - K : tid;
        float v = ((red[0][tid] + red[1][tid]) + red[2][tid]) + red[3][tid];
        for (int q = 4; q < n_sl; ++q) v += red[q][tid];
        v += second ? bout[k] : bo[k];
        vals[tid] = v;
        (second ? output : pvoutput)[(size_t)b * K + k] = v;
        if (target) {
            float d = v - target[(size_t)b * K + k];
            float g = loss_grad_elem(d, loss_kind, B * K);
            (second ? g_o2 : g_o)[(size_t)b * K + k] = g;
            lsum = loss_value_elem(d, loss_kind) / (float)(B * K);
        }
    }
    __syncthreads();
    if (tid == 0) {
        // DCLLClassification.forward :725-728 -- argmax of output on the output layer, else of pvoutput;
        // first maximum wins, as torch.argmax does.
        if (clout) {
            const int base = (Ktot > K) ? K : 0;
            int best = 0;
            float bv = vals[base];
            for (int k = 1; k < K; ++k)
                if (vals[base + k] > bv) bv = vals[base + k], best = k;
            clout[b] = best;
        }
    }
    if (loss_out && target) {
        // diagnostic only (train_dcll returns it, ConvNetwork.learn drops it): order-dependent atomics are fine
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
        if ((tid & 31) == 0 && tid < 64 && lsum != 0.f) atomicAdd(loss_out, lsum);
    }
}

// g_u[b,f] = (sum_k g_o[b,k] Wo[k,f]) * (1 - pv) * pv      (gradient w.r.t. the membrane at the pool argmax)
// Thread = one feature column of Wo held in registers, swept over a slice of the batch (4 samples in flight).
// WOUT variant (output layer, whole batch per CTA): the same sweep also accumulates
//   gWout[k,f] = sum_b g_o2[b,k] pv[b,f]   and applies optimizer2.step() (Adam, lr 1e-4, torch defaults;
//   dcll/pytorch_libdcll.py:636-638,713-714) thread-locally, so pv is read once for both purposes.
template <int KMAX, bool WOUT>
__global__ void __launch_bounds__(256) readout_bwd_kernel(const float *__restrict__ pv, const float *__restrict__ wo,
                                                          const float *__restrict__ g_o, const float *__restrict__ g_o2, int B,
                                                          int F, int K, int b_per_blk, float *__restrict__ g_u,
                                                          float *__restrict__ wout, float *__restrict__ bout,
                                                          float *__restrict__ m_w, float *__restrict__ v_w,
                                                          float *__restrict__ m_b, float *__restrict__ v_b,
                                                          float *__restrict__ grad_w, float *__restrict__ grad_b, int apply,
                                                          AdamScalars sc) {
    pdl_entry();
    // rows are read back as float4 (one broadcast LDS.128 per 4 FMAs; scalar LDS made the sweep LDS-issue bound)
    __shared__ __align__(16) float gs[64][KMAX];
    __shared__ __align__(16) float gs2[WOUT ? 64 : 1][KMAX];
    static_assert(KMAX % 4 == 0, "float4 rows");
    const int tid = threadIdx.x;
    const int f = blockIdx.x * 256 + tid;
    const bool fok = f < F;
    float w[KMAX], acc[WOUT ? KMAX : 1];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) w[k] = (k < K && fok) ? __ldg(wo + (size_t)k * F + f) : 0.f;
    if (WOUT) {
#pragma unroll
        for (int k = 0; k < KMAX; ++k) acc[k] = 0.f;
    }
    float bsum = 0.f;
    const int b_begin = blockIdx.y * b_per_blk, b_end = min(B, b_begin + b_per_blk);
    for (int b0 = b_begin; b0 < b_end; b0 += 64) {
        const int nb = min(64, b_end - b0);
        __syncthreads();
        for (int i = tid; i < 64 * KMAX; i += 256) {
            int bb = i / KMAX, k = i - bb * KMAX;
            const bool ok = bb < nb && k < K;
            gs[bb][k] = ok ? g_o[(size_t)(b0 + bb) * K + k] : 0.f;
            if (WOUT) gs2[bb][k] = ok ? g_o2[(size_t)(b0 + bb) * K + k] : 0.f;
        }
        __syncthreads();
        if (fok) {
            constexpr int UB = 8;   // samples in flight per thread (16 was slower: fewer resident CTAs)
            for (int bb = 0; bb < nb; bb += UB) {
                float pvv[UB];
#pragma unroll
                for (int u = 0; u < UB; ++u) pvv[u] = (bb + u < nb) ? __ldg(pv + (size_t)(b0 + bb + u) * F + f) : 0.f;
#pragma unroll
                for (int u = 0; u < UB; ++u) {
                    if (bb + u >= nb) break;
                    float s = 0.f;
                    const float4 *g4 = reinterpret_cast<const float4 *>(gs[bb + u]);
#pragma unroll
                    for (int k = 0; k < KMAX; k += 4) {
                        const float4 g = g4[k >> 2];
                        s = fmaf(g.x, w[k], s), s = fmaf(g.y, w[k + 1], s), s = fmaf(g.z, w[k + 2], s), s = fmaf(g.w, w[k + 3], s);
                    }
                    g_u[(size_t)(b0 + bb + u) * F + f] = s * (1.f - pvv[u]) * pvv[u];
                    if (WOUT) {
                        const float4 *h4 = reinterpret_cast<const float4 *>(gs2[bb + u]);
#pragma unroll
                        for (int k = 0; k < KMAX; k += 4) {
                            const float4 g = h4[k >> 2];
                            acc[k] = fmaf(g.x, pvv[u], acc[k]), acc[k + 1] = fmaf(g.y, pvv[u], acc[k + 1]);
                            acc[k + 2] = fmaf(g.z, pvv[u], acc[k + 2]), acc[k + 3] = fmaf(g.w, pvv[u], acc[k + 3]);
                        }
                    }
                }
            }
        }
        if (WOUT && blockIdx.x == 0 && tid < K)
            for (int bb = 0; bb < nb; ++bb) bsum += gs2[bb][tid];
    }
    if (WOUT) {
        if (fok) {
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
                if (k < K) {
                    size_t o = (size_t)k * F + f;
                    if (grad_w) grad_w[o] = acc[k];
                    if (apply) {
                        float wv = wout[o], m = m_w[o], v = v_w[o];
                        adam_elem(wv, acc[k], m, v, sc);
                        wout[o] = wv, m_w[o] = m, v_w[o] = v;
                    }
                }
            }
        }
        if (blockIdx.x == 0 && tid < K) {
            if (grad_b) grad_b[tid] = bsum;
            if (apply) {
                float wv = bout[tid], m = m_b[tid], v = v_b[tid];
                adam_elem(wv, bsum, m, v, sc);
                bout[tid] = wv, m_b[tid] = m, v_b[tid] = v;
            }
        }
    }
}

// Packed variant of the non-fused sweep above (g_u only): thread = TWO adjacent feature columns, every multiply-add is one
// FFMA2 (fma.rn.f32x2: two independent IEEE FMAs, so each lane-half is bit-identical to the scalar kernel).  The scalar
// kernel is instruction-issue bound (64 instructions per sample-feature, issue slots 67 % busy at 3.4 TB/s); here a
// sample costs 6 broadcast LDS.128 + 24 FFMA2 for two features.  Needs an even F and 8-byte aligned rows.  On the output
// layer it is paired with wout_grad_adam2_kernel (a fused packed sweep needs 179 registers -- one CTA per SM).
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b),
                       rc = *reinterpret_cast<unsigned long long *>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2 *>(&rd);
}

__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b), rd;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2 *>(&rd);
}

// IMG: g_u leaves the kernel as a bf16 {hi,lo} operand image in the same buffer -- sample b occupies the same F*4 bytes,
//     [part 2][co/8][position/8][co % 8][8 positions] bf16,
// i.e. 128-byte core matrices (8 channels x 8 consecutive positions) of the K-major B operand of wgrad_tc2_kernel, which moves
// them with one tensor-map box per unit (full 128-byte lines; the first, plain-NCHW form of the image made every 16-byte piece
// its own half-used L2 sector).  The block -> feature map follows the image: a CTA = 8 channels x 64 consecutive positions.
// q32 = 32-bit address of the bf16 pair in the hi part of the sample; the lo part starts F/2 words later.
// F16X2 (g_scale != 0): fp16 {hi,lo} of g_u * g_scale, saturated to the fp16 range (the scale comes from a bound on |g_u|)
// (packed conversions: one F2FP per pair on the ALU pipe instead of two half-rate F2F and a byte permute)
__device__ __forceinline__ void store_gu_img_f16(uint32_t *q32, int half_f, float2 v, float g_scale) {
    const float x = fminf(fmaxf(__fmul_rn(v.x, g_scale), -65504.f), 65504.f), y = fminf(fmaxf(__fmul_rn(v.y, g_scale), -65504.f), 65504.f);
    const __half2 h = __floats2half2_rn(x, y);                              // low half = x
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(x - hf.x, y - hf.y);
    q32[0] = *reinterpret_cast<const uint32_t *>(&h);
    q32[half_f] = *reinterpret_cast<const uint32_t *>(&l);
}
__device__ __forceinline__ void store_gu_img(uint32_t *q32, int half_f, float2 v) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(v.x, v.y);
    const uint32_t hb = *reinterpret_cast<const uint32_t *>(&h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(v.x - __uint_as_float(hb << 16), v.y - __uint_as_float(hb & 0xffff0000u));
    q32[0] = hb;
    q32[half_f] = *reinterpret_cast<const uint32_t *>(&l);
}

// pv rows reach the thread through a ring of thread-PRIVATE shared-memory slots filled by 8-byte cp.async (LDGSTS): NG groups of
// UB rows, requested NG - 1 groups ahead -- no registers held while the loads fly and no barrier (a thread only ever reads what it
// copied itself).  With the register double buffer of the first packed version a CTA had 16 KB in flight (32 KB per SM): latency-
// bound at ~3.3 TB/s of reads; at B = 64 the whole 32-row slice of a CTA is now requested before the first multiply (64 KB per
// CTA).  The CTAs that share a block of Wo columns (the batch slices) are adjacent in the 1-D grid, so the second read of those
// columns hits L2 instead of DRAM.
constexpr int RB2_NG = 4;
__device__ __forceinline__ void cp_async8(uint32_t dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
template <int KMAX, int UB, int MINB, bool IMG>
__global__ void __launch_bounds__(256, MINB) readout_bwd2_kernel(const float *__restrict__ pv, const float *__restrict__ wo,
                                                              const float *__restrict__ g_o, int B, int F, int K, int b_per_blk,
                                                              float *__restrict__ g_u, int hw, int slices, float g_scale) {
    extern __shared__ __align__(16) float2 rb2_ring[];                 // [RB2_NG][UB][256]
    const int fblk = blockIdx.x / slices, slice = blockIdx.x - fblk * slices;
    pdl_entry();
    // g rows are read back as broadcast LDS.128 (4 k per load) and enter the FFMA2 as a scalar-broadcast operand: ptxas folds
    // the {g,g} pair into `FFMA2 Rd, Rg.F32, Rw.F32x2, Rs.F32x2`.  (Storing the rows duplicated {g,g} doubled the LDS traffic
    // and made the sweep L1TEX-bound: ncu l1tex throughput 74 %, short-scoreboard stalls above the DRAM ones.)
    __shared__ __align__(16) float gd[64][KMAX];
    static_assert(KMAX % 4 == 0, "float4 rows");
    const int tid = threadIdx.x;
    int f = 2 * (fblk * 256 + tid);
    bool fok = f < F;
    int img_word = 0;                                      // IMG: 32-bit word of this thread's bf16 pair inside one part of the sample
    if (IMG) {
        const int chunks = hw >> 3, blk_per_cog = (chunks + 7) >> 3;          // 8-position chunks per plane, 64-position blocks
        const int cog = fblk / blk_per_cog, pb = fblk - cog * blk_per_cog;
        // warp = one channel of the group, lanes = 64 consecutive positions: the pv / Wo loads stay 256-byte contiguous per warp
        // (with lanes = 8 channels x 8 positions they were 8 scattered sectors: +0.02 ms per launch, measured); the 4-byte image
        // stores of a warp land as 16-byte runs in 8 lines, whose other runs come from the 7 sibling warps of this CTA
        // (round 2: a warp covers TWO adjacent channels x 32 positions instead of one x 64, so that the 4-byte image stores of a
        //  warp fill whole 32-byte sectors -- the 16-byte runs of two adjacent channels are neighbours in the image; measured with
        //  the stores removed they cost 0.023 ms per launch as half sectors, the conversions 0.012)
        const int wq = tid >> 5, ln = tid & 31;
        const int co8 = (wq & 3) * 2 + (ln >> 4), c8 = pb * 8 + (wq >> 2) * 4 + ((ln >> 2) & 3), pr = ln & 3;
        f = (cog * 8 + co8) * hw + c8 * 8 + 2 * pr;
        fok = c8 < chunks && f < F;
        img_word = ((cog * chunks + c8) * 8 + co8) * 4 + pr;
    }
    const int b_begin = slice * b_per_blk, b_end = min(B, b_begin + b_per_blk);
    const uint32_t ring0 = (uint32_t)__cvta_generic_to_shared(rb2_ring) + tid * 8;
    // group g of the rows b0.. : UB rows of this thread's feature pair -> ring stage g % NG
    auto request = [&](int b0, int g) {
        const float *src = pv + (size_t)(b0 + g * UB) * F + f;
        const uint32_t dst = ring0 + (g % RB2_NG) * (UB * 256 * 8);
#pragma unroll
        for (int u = 0; u < UB; ++u, src += F) cp_async8(dst + u * (256 * 8), src);
    };
    // the first groups are requested before anything else, so they fly together with the Wo columns and the g_o rows
    if (fok && b_begin < b_end) {
        const int ng0 = min(64, b_end - b_begin) / UB;
#pragma unroll
        for (int g = 0; g < RB2_NG - 1; ++g) {
            if (g < ng0) request(b_begin, g);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
    }
    float2 w[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) w[k] = (k < K && fok) ? __ldg(reinterpret_cast<const float2 *>(wo + (size_t)k * F + f)) : make_float2(0.f, 0.f);
    for (int b0 = b_begin; b0 < b_end; b0 += 64) {
        const int nb = min(64, b_end - b0);
        __syncthreads();
        for (int i = tid; i < 64 * KMAX; i += 256) {
            int bb = i / KMAX, k = i - bb * KMAX;
            gd[bb][k] = (bb < nb && k < K) ? g_o[(size_t)(b0 + bb) * K + k] : 0.f;
        }
        __syncthreads();
        if (fok) {
            const size_t rowF = (size_t)F;
            const float *pvp = pv + (size_t)b0 * F + f;
            float *gup = g_u + (size_t)b0 * F + f;
            const float2 one2 = make_float2(1.f, 1.f), neg2 = make_float2(-1.f, -1.f);
            auto sample = [&](int bb, float2 p) {
                float2 s = make_float2(0.f, 0.f);
                const float4 *g4 = reinterpret_cast<const float4 *>(gd[bb]);
#pragma unroll
                for (int k = 0; k < KMAX; k += 4) {
                    const float4 g = g4[k >> 2];
                    s = ffma2(make_float2(g.x, g.x), w[k], s);
                    s = ffma2(make_float2(g.y, g.y), w[k + 1], s);
                    s = ffma2(make_float2(g.z, g.z), w[k + 2], s);
                    s = ffma2(make_float2(g.w, g.w), w[k + 3], s);
                }
                // s * (1 - pv) * pv, each lane-half rounded exactly like the scalar expression (1 - pv == fma(pv, -1, 1))
                return fmul2(fmul2(s, ffma2(p, neg2, one2)), p);
            };
            int bb = 0;
            // full groups of UB rows through the cp.async ring (requested RB2_NG - 1 groups ahead; the first ones at kernel entry)
            {
                const int ng = nb / UB;
                if (b0 != b_begin) {
#pragma unroll
                    for (int g = 0; g < RB2_NG - 1; ++g) {
                        if (g < ng) request(b0, g);
                        asm volatile("cp.async.commit_group;" ::: "memory");
                    }
                }
                for (int g = 0; g < ng; ++g, bb += UB) {
                    if (g + RB2_NG - 1 < ng) request(b0, g + RB2_NG - 1);   // its stage was read by this thread one iteration ago
                    asm volatile("cp.async.commit_group;" ::: "memory");
                    asm volatile("cp.async.wait_group %0;" ::"n"(RB2_NG - 1) : "memory");
                    const float2 *stage = rb2_ring + (g % RB2_NG) * (UB * 256) + tid;
                    float2 cur[UB];
#pragma unroll
                    for (int u = 0; u < UB; ++u) cur[u] = stage[u * 256];
                    float *q = gup;
#pragma unroll
                    for (int u = 0; u < UB; ++u, q += rowF) {
                        const float2 v = sample(bb + u, cur[u]);
                        if (IMG && g_scale != 0.f) store_gu_img_f16(reinterpret_cast<uint32_t *>(q - f) + img_word, F >> 1, v, g_scale);
                        else if (IMG) store_gu_img(reinterpret_cast<uint32_t *>(q - f) + img_word, F >> 1, v);
                        else *reinterpret_cast<float2 *>(q) = v;
                    }
                    gup = q;
                }
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                pvp += (size_t)ng * UB * rowF;
            }
            for (; bb < nb; ++bb, pvp += rowF, gup += rowF) {
                const float2 v = sample(bb, __ldg(reinterpret_cast<const float2 *>(pvp)));
                if (IMG && g_scale != 0.f) store_gu_img_f16(reinterpret_cast<uint32_t *>(gup - f) + img_word, F >> 1, v, g_scale);
                else if (IMG) store_gu_img(reinterpret_cast<uint32_t *>(gup - f) + img_word, F >> 1, v);
                else *reinterpret_cast<float2 *>(gup) = v;
            }
        }
    }
}

// output_ gradient + Adam alone, for small feature counts (16x16 planes: F = 8192), where the fused sweep above would
// leave each thread with a serial loop over the whole batch.  CTA = 32 features x 8 batch slices (one warp each, so the
// pv loads are 128-byte coalesced and g_o2 is a warp-broadcast load) x KG outputs (grid.y); fixed-order reduction of the
// 8 slices through shared memory, then Adam.
template <int KG>
__global__ void __launch_bounds__(256) wout_grad_adam_kernel(const float *__restrict__ pv, const float *__restrict__ g_o2, int B,
                                                             int F, int K, float *__restrict__ wout, float *__restrict__ bout,
                                                             float *__restrict__ m_w, float *__restrict__ v_w,
                                                             float *__restrict__ m_b, float *__restrict__ v_b,
                                                             float *__restrict__ grad_w, float *__restrict__ grad_b, int apply,
                                                             AdamScalars sc) {
    pdl_entry();
    __shared__ float red[8][KG][33];
    __shared__ float bred[8][KG];
    const int fx = threadIdx.x & 31, bs = threadIdx.x >> 5;
    const int f = blockIdx.x * 32 + fx;
    const int k0 = blockIdx.y * KG;
    const bool fok = f < F;
    float acc[KG], bsum[KG];
#pragma unroll
    for (int k = 0; k < KG; ++k) acc[k] = 0.f, bsum[k] = 0.f;
    for (int b = bs; b < B; b += 8 * 4) {
        float pvv[4], g[4][KG];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int bb = b + 8 * u;
            pvv[u] = (bb < B && fok) ? __ldg(pv + (size_t)bb * F + f) : 0.f;
#pragma unroll
            for (int k = 0; k < KG; ++k) g[u][k] = (bb < B && k0 + k < K) ? __ldg(g_o2 + (size_t)bb * K + k0 + k) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < KG; ++k) acc[k] = fmaf(g[u][k], pvv[u], acc[k]), bsum[k] += g[u][k];
    }
#pragma unroll
    for (int k = 0; k < KG; ++k) red[bs][k][fx] = acc[k];
    if (fx == 0) {
#pragma unroll
        for (int k = 0; k < KG; ++k) bred[bs][k] = bsum[k];
    }
    __syncthreads();
    if (bs == 0 && fok) {
#pragma unroll
        for (int k = 0; k < KG; ++k) {
            if (k0 + k < K) {
                float gsum = 0.f;
#pragma unroll
                for (int s = 0; s < 8; ++s) gsum += red[s][k][fx];
                size_t o = (size_t)(k0 + k) * F + f;
                if (grad_w) grad_w[o] = gsum;
                if (apply) {
                    float wv = wout[o], m = m_w[o], v = v_w[o];
                    adam_elem(wv, gsum, m, v, sc);
                    wout[o] = wv, m_w[o] = m, v_w[o] = v;
                }
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < KG && k0 + threadIdx.x < K) {
        const int k = k0 + threadIdx.x;
        float gsum = 0.f;
        for (int s = 0; s < 8; ++s) gsum += bred[s][threadIdx.x];
        if (grad_b) grad_b[k] = gsum;
        if (apply) {
            float wv = bout[k], m = m_b[k], v = v_b[k];
            adam_elem(wv, gsum, m, v, sc);
            bout[k] = wv, m_b[k] = m, v_b[k] = v;
        }
    }
}

// output_ gradient + Adam for LARGE feature counts (128x128 planes: F = 524 288), the companion of readout_bwd2_kernel on
// the output layer.  The fused scalar sweep (readout_bwd_kernel<., true>) keeps the whole batch serial in one thread with
// 48 scalar FMAs per sample-feature and a 2.3-wave grid (0.27 ms against a 0.09 ms HBM floor); splitting it into the packed
// g_u sweep plus this kernel re-reads pv once (134 MB, 0.02 ms) but both halves run packed and with >= 4 waves.
// CTA = 128 feature PAIRS x 2 halves of the OUTPUT index k: a thread accumulates gWout[k][f, f+1] for its KMAX/2 outputs over
// the whole batch in order (FFMA2 with the g_o2 value as broadcast operand, rows prefetched one group ahead) and then applies
// Adam to exactly those elements -- no cross-thread reduction.  The two halves read the same pv rows (second read: L1/L2).
// (First version: halves of the BATCH, all KMAX outputs per thread -- 128 registers, 2 CTAs per SM, 0.104 ms.)
template <int KMAX>
__global__ void __launch_bounds__(256, 3) wout_grad_adam2_kernel(const float *__restrict__ pv, const float *__restrict__ g_o2, int B,
                                                                 int F, int K, float *__restrict__ wout, float *__restrict__ bout,
                                                                 float *__restrict__ m_w, float *__restrict__ v_w,
                                                                 float *__restrict__ m_b, float *__restrict__ v_b,
                                                                 float *__restrict__ grad_w, float *__restrict__ grad_b, int apply,
                                                                 AdamScalars sc) {
    pdl_entry();
    constexpr int KH = KMAX / 2, UB = 8;
    static_assert(KH % 4 == 0, "float4 rows per half");
    __shared__ __align__(16) float gd[64][KMAX];                  // broadcast LDS.128 rows, see readout_bwd2_kernel
    extern __shared__ __align__(16) float2 rb2_ring[];            // [RB2_NG][UB][256]: thread-private cp.async slots (readout_bwd2_kernel)
    const int tid = threadIdx.x, pair = tid & 127, half = tid >> 7;
    const int f = 2 * (blockIdx.x * 128 + pair);
    const bool fok = f < F;
    const int kbase = half * KH;
    const size_t rowF = (size_t)F;
    const uint32_t ring0 = (uint32_t)__cvta_generic_to_shared(rb2_ring) + tid * 8;
    auto request = [&](int b0, int g) {
        const float *src = pv + (size_t)(b0 + g * UB) * F + f;
        const uint32_t dst = ring0 + (g % RB2_NG) * (UB * 256 * 8);
#pragma unroll
        for (int u = 0; u < UB; ++u, src += F) cp_async8(dst + u * (256 * 8), src);
    };
    if (fok) {
        const int ng0 = min(64, B) / UB;
#pragma unroll
        for (int g = 0; g < RB2_NG - 1; ++g) {
            if (g < ng0) request(0, g);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
    }
    // the Adam tail reads w, m, v of this thread's outputs in dependent rounds: pull their lines into L2 now, while the batch streams
    // (no measurable effect at 128x128, B = 64 -- 0.102 ms either way -- kept because it is free)
    if (fok && apply) {
#pragma unroll
        for (int k = 0; k < KH; ++k) {
            if (kbase + k < K) {
                const size_t o = (size_t)(kbase + k) * F + f;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(wout + o));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(m_w + o));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(v_w + o));
            }
        }
    }
    // The Adam tail's w, m, v rows travel through the SAME ring as extra groups behind the last pv rows (TG = KH / 2 groups of
    // 2 outputs x {w, m, v}), requested RB2_NG - 1 groups ahead like everything else: their DRAM latency is covered by the last
    // multiplies and by the Adam arithmetic of the groups before them.  (Loading them in dependent rounds after the sweep left
    // every CTA waiting four times: 18 % of the samples on the first use of each round.)
    constexpr int TG = KH / 2;
    static_assert(KH % 2 == 0 && 6 <= UB, "tail groups of two outputs fit a ring group");
    const bool chain_ok = apply && K == KMAX;
    bool chained = false;
    int ng_last = 0;
    auto request_tail = [&](int tg) {
        const uint32_t dst = ring0 + ((ng_last + tg) % RB2_NG) * (UB * 256 * 8);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const size_t o = (size_t)(kbase + 2 * tg + j) * F + f;
            cp_async8(dst + (3 * j + 0) * (256 * 8), wout + o);
            cp_async8(dst + (3 * j + 1) * (256 * 8), m_w + o);
            cp_async8(dst + (3 * j + 2) * (256 * 8), v_w + o);
        }
    };
    float2 acc[KH];
#pragma unroll
    for (int k = 0; k < KH; ++k) acc[k] = make_float2(0.f, 0.f);
    for (int b0 = 0; b0 < B; b0 += 64) {
        const int nb = min(64, B - b0);
        __syncthreads();
        for (int i = tid; i < 64 * KMAX; i += 256) {
            int bb = i / KMAX, k = i - bb * KMAX;
            gd[bb][k] = (bb < nb && k < K) ? g_o2[(size_t)(b0 + bb) * K + k] : 0.f;
        }
        __syncthreads();
        if (fok) {
            const float *pvp = pv + (size_t)b0 * F + f;
            auto sample = [&](int bb, float2 p) {
                const float4 *g4 = reinterpret_cast<const float4 *>(&gd[bb][kbase]);
#pragma unroll
                for (int k = 0; k < KH; k += 4) {
                    const float4 g = g4[k >> 2];
                    acc[k] = ffma2(make_float2(g.x, g.x), p, acc[k]);
                    acc[k + 1] = ffma2(make_float2(g.y, g.y), p, acc[k + 1]);
                    acc[k + 2] = ffma2(make_float2(g.z, g.z), p, acc[k + 2]);
                    acc[k + 3] = ffma2(make_float2(g.w, g.w), p, acc[k + 3]);
                }
            };
            int bb = 0;
            {
                const int ng = nb / UB;
                const bool chain = chain_ok && b0 + 64 >= B && ng >= RB2_NG - 1;   // last batch block: the tail follows in the ring
                if (chain) chained = true, ng_last = ng;
                if (b0 != 0) {
#pragma unroll
                    for (int g = 0; g < RB2_NG - 1; ++g) {
                        if (g < ng) request(b0, g);
                        asm volatile("cp.async.commit_group;" ::: "memory");
                    }
                }
                for (int g = 0; g < ng; ++g, bb += UB) {
                    if (g + RB2_NG - 1 < ng) request(b0, g + RB2_NG - 1);
                    else if (chain && g + RB2_NG - 1 - ng < TG) request_tail(g + RB2_NG - 1 - ng);
                    asm volatile("cp.async.commit_group;" ::: "memory");
                    asm volatile("cp.async.wait_group %0;" ::"n"(RB2_NG - 1) : "memory");
                    const float2 *stage = rb2_ring + (g % RB2_NG) * (UB * 256) + tid;
                    float2 cur[UB];
#pragma unroll
                    for (int u = 0; u < UB; ++u) cur[u] = stage[u * 256];
#pragma unroll
                    for (int u = 0; u < UB; ++u) sample(bb + u, cur[u]);
                }
                if (!chain) asm volatile("cp.async.wait_group 0;" ::: "memory");
                pvp += (size_t)ng * UB * rowF;
            }
            for (; bb < nb; ++bb, pvp += rowF) sample(bb, __ldg(reinterpret_cast<const float2 *>(pvp)));
        }
    }
    if (fok && chained) {
#pragma unroll
        for (int tg = 0; tg < TG; ++tg) {
            if (tg + RB2_NG - 1 < TG) request_tail(tg + RB2_NG - 1);
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group %0;" ::"n"(RB2_NG - 1) : "memory");
            const float2 *stage = rb2_ring + ((ng_last + tg) % RB2_NG) * (UB * 256) + tid;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const size_t o = (size_t)(kbase + 2 * tg + j) * F + f;
                float2 w = stage[(3 * j + 0) * 256], m = stage[(3 * j + 1) * 256], v = stage[(3 * j + 2) * 256];
                const float2 g = acc[2 * tg + j];
                if (grad_w) *reinterpret_cast<float2 *>(grad_w + o) = g;
                adam_elem(w.x, g.x, m.x, v.x, sc);
                adam_elem(w.y, g.y, m.y, v.y, sc);
                *reinterpret_cast<float2 *>(wout + o) = w;
                *reinterpret_cast<float2 *>(m_w + o) = m;
                *reinterpret_cast<float2 *>(v_w + o) = v;
            }
        }
    } else if (fok) {
        // Adam tail on this thread's own outputs, TB rows at a time with every load issued before the first store
        constexpr int TB = (KH % 3 == 0) ? 3 : 4;
        static_assert(KH % TB == 0, "tail groups");
#pragma unroll
        for (int kk = 0; kk < KH; kk += TB) {
            float2 w[TB], m[TB], v[TB];
#pragma unroll
            for (int j = 0; j < TB; ++j) {
                const size_t o = (size_t)(kbase + kk + j) * F + f;
                if (apply && kbase + kk + j < K) {
                    w[j] = *reinterpret_cast<const float2 *>(wout + o);
                    m[j] = *reinterpret_cast<const float2 *>(m_w + o);
                    v[j] = *reinterpret_cast<const float2 *>(v_w + o);
                }
            }
#pragma unroll
            for (int j = 0; j < TB; ++j) {
                if (kbase + kk + j < K) {
                    const size_t o = (size_t)(kbase + kk + j) * F + f;
                    const float2 g = acc[kk + j];
                    if (grad_w) *reinterpret_cast<float2 *>(grad_w + o) = g;
                    if (apply) {
                        adam_elem(w[j].x, g.x, m[j].x, v[j].x, sc);
                        adam_elem(w[j].y, g.y, m[j].y, v[j].y, sc);
                        *reinterpret_cast<float2 *>(wout + o) = w[j];
                        *reinterpret_cast<float2 *>(m_w + o) = m[j];
                        *reinterpret_cast<float2 *>(v_w + o) = v[j];
                    }
                }
            }
        }
    }
    if (blockIdx.x == 0 && tid < K) {
        float bsum = 0.f;
        for (int b = 0; b < B; ++b) bsum += g_o2[(size_t)b * K + tid];
        if (grad_b) grad_b[tid] = bsum;
        if (apply) {
            float wv = bout[tid], m = m_b[tid], v = v_b[tid];
            adam_elem(wv, bsum, m, v, sc);
            bout[tid] = wv, m_b[tid] = m, v_b[tid] = v;
        }
    }
}

// g_o / g_o2 from stored read-outs (layer-level API: the target is only known after forward returned)
__global__ void loss_grad_kernel(const float *__restrict__ pvoutput, const float *__restrict__ output,
                                 const float *__restrict__ target, int B, int K, int loss_kind, float *__restrict__ g_o,
                                 float *__restrict__ g_o2, float *__restrict__ loss_out) {
    pdl_entry();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    float lsum = 0.f;
    if (i < B * K) {
        float t = target[i];
        float d = pvoutput[i] - t;
        g_o[i] = loss_grad_elem(d, loss_kind, B * K);
        lsum = loss_value_elem(d, loss_kind) / (float)(B * K);
        if (output) {
            float d2 = output[i] - t;
            g_o2[i] = loss_grad_elem(d2, loss_kind, B * K);
            lsum += loss_value_elem(d2, loss_kind) / (float)(B * K);
        }
    }
    if (loss_out) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
        if ((threadIdx.x & 31) == 0 && lsum != 0.f) atomicAdd(loss_out, lsum);
    }
}

// generic Adam over a flat parameter (data-parallel path: gradients arrive from the allreduce)
__global__ void adam_flat_kernel(float *__restrict__ w, const float *__restrict__ g, float *__restrict__ m,
                                 float *__restrict__ v, size_t n, AdamScalars sc) {
    pdl_entry();
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float wv = w[i], mv = m[i], vv = v[i];
    adam_elem(wv, g[i], mv, vv, sc);
    w[i] = wv, m[i] = mv, v[i] = vv;
}

int launch_readout_fwd(const dcll_conv_layer *L, const float *target, int loss_kind, int32_t *clout, float *loss_out,
                       cudaStream_t st) {
    Geo g = geo_of(L);
    WsLayout ws = ws_layout(L);
    char *base = (char *)L->workspace;
    float *partial = (float *)(base + ws.off_ro_part);
    float *g_o = (float *)(base + ws.off_go), *g_o2 = (float *)(base + ws.off_go2);
    if (readout_tc_supported(L)) {
        int rc = launch_readout_tc(L, partial, st);
        if (rc != DCLL_OK) return rc;
        launch_k(readout_finish_kernel, L->B, ws.n_ro_tc > 16 ? 1024 : 256, 0, st, partial, ws.n_ro_tc, L->B, L->K, g.Ktot, L->bo, L->bout, target, loss_kind,
                                                      L->pvoutput, L->output, g_o, g_o2, clout, loss_out);
        DCLL_LAUNCH_OK("readout_finish_kernel");
        return DCLL_OK;
    }
    const int kj = ceil_div(g.Ktot, 16);
    DCLL_REQUIRE(kj >= 1 && kj <= 4, DCLL_EUNSUPPORTED, "read-out width %d > 64 unsupported", g.Ktot);
    const bool vec = (g.F % 4 == 0) && (((uintptr_t)L->pv | (uintptr_t)L->wo | (uintptr_t)L->wout) % 16 == 0);
#define RO_CFG(J)                                                                                            \
    DCLL_SMEM_ATTR((readout_fwd_kernel<J, true>), 2 * (RO_BM + 16 * J) * RO_PITCH * sizeof(float));          \
    DCLL_SMEM_ATTR((readout_fwd_kernel<J, false>), 2 * (RO_BM + 16 * J) * RO_PITCH * sizeof(float));
    RO_CFG(1) RO_CFG(2) RO_CFG(3) RO_CFG(4)
#undef RO_CFG
    // batch tiles in grid.y until the grid holds ~4 CTAs per SM
    const int b_tiles = ceil_div(L->B, RO_BM);
    dim3 grid(ws.n_ro, max(1, min(b_tiles, ceil_div(4 * 148, ws.n_ro))));
#define RO_CASE(J)                                                                                                        \
    case J: {                                                                                                             \
        size_t sm = 2 * (RO_BM + 16 * J) * RO_PITCH * sizeof(float);                                                      \
        if (vec) launch_k(readout_fwd_kernel<J, true>, grid, 256, sm, st, L->pv, L->wo, L->wout, L->B, g.F, L->K, g.Ktot, partial); \
        else launch_k(readout_fwd_kernel<J, false>, grid, 256, sm, st, L->pv, L->wo, L->wout, L->B, g.F, L->K, g.Ktot, partial);    \
        break;                                                                                                            \
    }
    switch (kj) { RO_CASE(1) RO_CASE(2) RO_CASE(3) RO_CASE(4) }
#undef RO_CASE
    DCLL_LAUNCH_OK("readout_fwd_kernel");
    launch_k(readout_finish_kernel, L->B, 256, 0, st, partial, ws.n_ro, L->B, L->K, g.Ktot, L->bo, L->bout, target, loss_kind,
                                                 L->pvoutput, L->output, g_o, g_o2, clout, loss_out);
    DCLL_LAUNCH_OK("readout_finish_kernel");
    return DCLL_OK;
}

int launch_loss_grad(const dcll_conv_layer *L, const float *target, int loss_kind, float *loss_out, cudaStream_t st) {
    WsLayout ws = ws_layout(L);
    char *base = (char *)L->workspace;
    float *g_o = (float *)(base + ws.off_go), *g_o2 = (float *)(base + ws.off_go2);
    int n = L->B * L->K;
    launch_k(loss_grad_kernel, ceil_div(n, 256), 256, 0, st, L->pvoutput, L->output_layer ? L->output : nullptr, target, L->B, L->K,
                                                        loss_kind, g_o, g_o2, loss_out);
    DCLL_LAUNCH_OK("loss_grad_kernel");
    return DCLL_OK;
}

int launch_readout_bwd(const dcll_conv_layer *L, dcll_train_args *a, cudaStream_t st) {
    Geo g = geo_of(L);
    WsLayout ws = ws_layout(L);
    char *base = (char *)L->workspace;
    const float *g_o = (const float *)(base + ws.off_go), *g_o2 = (const float *)(base + ws.off_go2);
    DCLL_REQUIRE(L->K <= 32, DCLL_EUNSUPPORTED, "target_size %d > 32 unsupported in the backward read-out", L->K);
    // g_u alone: packed kernel (two features per thread, FFMA2) when rows are 8-byte aligned, else the scalar kernel.
    // Output layer with a large F: ONE scalar sweep over pv serves g_u and gWout (+ Adam); with a small F the output_
    // gradient gets its own kernel below.
    const bool packed = (g.F % 2 == 0) && (((uintptr_t)L->pv | (uintptr_t)L->wo | (uintptr_t)L->g_u) % 8 == 0);
    AdamScalars sc = {};
    const bool big_out = L->output_layer && ceil_div(g.F, 256) >= 2 * 148;
    const bool packed_out = big_out && packed && (!a->apply_update || (a->adam_out.m_w && a->adam_out.v_w)) &&
                            (((uintptr_t)L->wout | (uintptr_t)a->adam_out.m_w | (uintptr_t)a->adam_out.v_w | (uintptr_t)a->grad_wout) % 8 == 0);
    // odd F / unaligned rows: the scalar fused sweep -- never when the weight-gradient kernel expects g_u in image form (then the
    // packed g_u sweep below runs and the output_ gradient takes the generic kernel)
    const bool fused_out = big_out && !packed_out && !wgrad_tc2_supported(L);
    int fblk = ceil_div(g.F, (packed && !fused_out) ? 512 : 256);
    const int hw = g.Hp * g.Wp;
    if (wgrad_tc2_supported(L)) fblk = (L->Cout / 8) * ceil_div(hw / 8, 8);   // image-ordered blocks: 8 channels x 64 positions
    if (packed_out) {
        // large F: packed g_u sweep below + packed output_ gradient/Adam kernel (pv read twice, both at >= 4 waves)
        sc = adam_scalars(a->adam_out, a->adam_out.step + 1);
        dcll_adam &o = a->adam_out;
        const int grid = ceil_div(g.F, 256);
        constexpr int ring_bytes = RB2_NG * 8 * 256 * 8;
#define WG2(KM)                                                                                                           \
    DCLL_SMEM_ATTR(wout_grad_adam2_kernel<KM>, ring_bytes);                                                               \
    launch_k(wout_grad_adam2_kernel<KM>, grid, 256, ring_bytes, st, L->pv, g_o2, L->B, g.F, L->K, L->wout, L->bout, o.m_w, o.v_w, o.m_b, \
                                                     o.v_b, a->grad_wout, a->grad_bout, a->apply_update, sc)
        if (L->K <= 16) { WG2(16); }
        else if (L->K <= 24) { WG2(24); }
        else { WG2(32); }
#undef WG2
        DCLL_LAUNCH_OK("wout_grad_adam2_kernel");
    } else if (L->output_layer && !fused_out) {
        // small F: the output_ gradient gets its own kernel (batch split over warps inside the CTA)
        sc = adam_scalars(a->adam_out, a->adam_out.step + 1);
        dcll_adam &o = a->adam_out;
        dim3 grid(ceil_div(g.F, 32), ceil_div(L->K, 8));
        launch_k(wout_grad_adam_kernel<8>, grid, 256, 0, st, L->pv, g_o2, L->B, g.F, L->K, L->wout, L->bout, o.m_w, o.v_w, o.m_b, o.v_b,
                                                      a->grad_wout, a->grad_bout, a->apply_update, sc);
        DCLL_LAUNCH_OK("wout_grad_adam_kernel");
    }
#define RB_LAUNCH(KM, WO, GRID, BPER, ...) \
    launch_k(readout_bwd_kernel<KM, WO>, GRID, 256, 0, st, L->pv, L->wo, g_o, WO ? g_o2 : nullptr, L->B, g.F, L->K, BPER, L->g_u, __VA_ARGS__)
    if (fused_out) {
        // the output_ gradient reduces over the whole batch inside one thread: no batch slicing
        sc = adam_scalars(a->adam_out, a->adam_out.step + 1);
        dcll_adam &o = a->adam_out;
        dim3 grid(fblk, 1);
        if (L->K <= 16)
            RB_LAUNCH(16, true, grid, L->B, L->wout, L->bout, o.m_w, o.v_w, o.m_b, o.v_b, a->grad_wout, a->grad_bout, a->apply_update, sc);
        else if (L->K <= 24)
            RB_LAUNCH(24, true, grid, L->B, L->wout, L->bout, o.m_w, o.v_w, o.m_b, o.v_b, a->grad_wout, a->grad_bout, a->apply_update, sc);
        else
            RB_LAUNCH(32, true, grid, L->B, L->wout, L->bout, o.m_w, o.v_w, o.m_b, o.v_b, a->grad_wout, a->grad_bout, a->apply_update, sc);
    } else {
        // enough CTAs for >= 4 waves of 3 CTAs per SM (a 2.3-wave grid loses 23 % to the partial last wave): slice the batch
        // (packed kernel, round 2: with the pv rows in the cp.async ring a CTA's fixed cost -- its Wo columns and g_o rows, one DRAM
        //  round trip before the first multiply -- is what more slices multiply; two waves of 2 CTAs per SM are enough.  Measured at
        //  128x128, B = 64: one slice 0.066 / 0.083 ms against 0.074 / 0.089 with two.)
        int slices = max(1, min(ceil_div(L->B, 32), ceil_div((packed ? 2 * 2 : 4 * 3) * 148, fblk)));
        {
            static int sl_env = -1;                                       // DCLL_RB2_SLICES=n: force the number of batch slices (A/B)
            if (sl_env < 0) {
                const char *e = getenv("DCLL_RB2_SLICES");
                sl_env = e ? atoi(e) : 0;
            }
            if (sl_env > 0) slices = sl_env;
        }
        int b_per = ceil_div(ceil_div(L->B, slices), 32) * 32;
        slices = ceil_div(L->B, b_per);
        dim3 grid(fblk, slices);
        float *nf = nullptr;
        if (packed) {
            const unsigned grid1 = (unsigned)fblk * slices;                   // 1-D: the slices of a feature block are adjacent
            constexpr int ring_bytes = RB2_NG * 8 * 256 * 8;
            const float g_scale = prec_f16(L) ? pow2i(L->g_exp) : 0.f;   // F16X2: the image is fp16 {hi,lo} of g_u * 2^g_exp
            // DCLL_RB_TC=1: the K-sum as a skinny tcgen05 GEMM per 128 features (readout_bwd_tc.cu), F16X2 image form only
            if (g_scale != 0.f && wgrad_tc2_supported(L) && readout_bwd_tc_supported(L, hw, g.F))
                return launch_readout_bwd_tc(L, g_o, hw, g.F, g_scale, st);
            // 8 rows in flight per thread, next group prefetched: 128 registers, 2 CTAs per SM (4 rows at 3 CTAs per SM was slower)
            // image form of g_u (bf16 {hi,lo} planes in the same buffer) when the row-pair weight-gradient kernel consumes it
#define RB2_M(KM, MB)                                                                                                              \
    do {                                                                                                                           \
        if (wgrad_tc2_supported(L)) {                                                                                              \
            DCLL_SMEM_ATTR((readout_bwd2_kernel<KM, 8, MB, true>), ring_bytes);                                                    \
            launch_k(readout_bwd2_kernel<KM, 8, MB, true>, grid1, 256, ring_bytes, st, L->pv, L->wo, g_o, L->B, g.F, L->K, b_per, L->g_u, hw, slices, g_scale); \
        } else {                                                                                                                   \
            DCLL_SMEM_ATTR((readout_bwd2_kernel<KM, 8, MB, false>), ring_bytes);                                                   \
            launch_k(readout_bwd2_kernel<KM, 8, MB, false>, grid1, 256, ring_bytes, st, L->pv, L->wo, g_o, L->B, g.F, L->K, b_per, L->g_u, hw, slices, 0.f); \
        }                                                                                                                          \
    } while (0)
            // CTAs per SM: two.  With the rows in the cp.async ring instead of registers three fit for K <= 24 (80 registers, a few
            // bytes of spill) but run SLOWER (measured, 128x128 B = 64: 0.084 / 0.121 ms against 0.075 / 0.102); DCLL_RB2_MINB=3 selects it
            static int minb = -1;
            if (minb < 0) {
                const char *e = getenv("DCLL_RB2_MINB");
                minb = (e && atoi(e) == 3) ? 3 : 2;
            }
#define RB2(KM)                  \
    do {                         \
        if (minb == 3 && KM <= 24) RB2_M(KM, 3); \
        else RB2_M(KM, 2);       \
    } while (0)
            if (L->K <= 16) { RB2(16); }
            else if (L->K <= 24) { RB2(24); }
            else { RB2(32); }
#undef RB2
#undef RB2_M
        } else if (L->K <= 16)
            RB_LAUNCH(16, false, grid, b_per, nf, nf, nf, nf, nf, nf, nf, nf, 0, sc);
        else if (L->K <= 24)
            RB_LAUNCH(24, false, grid, b_per, nf, nf, nf, nf, nf, nf, nf, nf, 0, sc);
        else
            RB_LAUNCH(32, false, grid, b_per, nf, nf, nf, nf, nf, nf, nf, nf, 0, sc);
    }
#undef RB_LAUNCH
    DCLL_LAUNCH_OK("readout_bwd_kernel");
    return DCLL_OK;
}

int launch_adam_flat(float *w, const float *g, float *m, float *v, size_t n, const AdamScalars &sc, cudaStream_t st) {
    launch_k(adam_flat_kernel, (unsigned)((n + 255) / 256), 256, 0, st, w, g, m, v, n, sc);
    DCLL_LAUNCH_OK("adam_flat_kernel");
    return DCLL_OK;
}

}  // namespace dcll
