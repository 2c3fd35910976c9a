// Fused per-timestep forward of one Conv2dDCLLlayer, FP32 parity mode.
//
// Replaces (reference, dcll/pytorch_libdcll.py):
//   :415-416  eps0/eps1 synaptic-trace recurrences        -> prologue (while staging the conv input tile)
//   :417-418  F.conv2d(eps1, W, b)                        -> register-blocked direct convolution on FP32 FMA
//   :497-503  refractory trace (RRP variant)              -> epilogue
//   :419-420  sigmoid, (u > 0) threshold                  -> epilogue
//   :601      MaxPool2d on spikes and on pv               -> epilogue (registers + one warp shuffle)
//
// Layout: all tensors NCHW float32 exactly as the Python API exposes them.  The neuron state is a
// ping-pong pair: a CTA reads the OLD traces of its tile plus halo and writes the NEW traces of the
// part of the input plane it owns, so neighbouring CTAs never observe half-updated halos.
//
// Work decomposition: one CTA = one sample x one (TH x 8*SEGS) tile of conv outputs x up to 32 output
// channels.  A thread owns 8 consecutive output columns x 8 output channels (64 FP32 accumulators) and
// slides a 16-float register window over the input row, so one (ci, kh) step costs 4 LDS.128 of input +
// 2*KW broadcast LDS.128 of weights for 64*KW FMAs: the kernel is bound by the FP32 FMA pipe.
#include "common.cuh"

namespace dcll {

struct FwdP {
    const float *x;
    const int2 *cells;
    const float *e0_old, *e1_old;
    float *e0_new, *e1_new;
    const float *alpha, *alphas, *tau_m, *tau_s;
    const float *wt, *bias;
    float *arp, *spikes, *pv, *pvmem;
    uint8_t *pool_idx;
    float alpharp, wrp;
    int coef_mode;
    int B, Cin, H, W, Cout, CoutPad, padH, padW, Hc, Wc, Hp, Wp;
    int tiles_h, tiles_w;
};

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

__host__ __device__ constexpr int pitch_for(int tw, int kw) {
    // >= halo width and >= tw + 8 (16-float register window of the last segment), multiple of 4,
    // (pitch/4) odd so that 8 consecutive rows hit 8 different 16-byte bank groups.
    int need = (tw + kw - 1) > (tw + 8) ? (tw + kw - 1) : (tw + 8);
    int p = (need + 3) / 4 * 4;
    if (((p / 4) & 1) == 0) p += 4;
    return p;
}

template <int KH, int KW, int TH, int SEGS, int CI_T, int PH, int PW>
__global__ void __launch_bounds__(TH *SEGS * 4, 512 / (TH * SEGS * 4)) conv_fwd_kernel(const FwdP p) {
    pdl_entry();
    constexpr int TW = 8 * SEGS;
    constexpr int POS_T = TH * SEGS;
    constexpr int HALO_H = TH + KH - 1, HALO_W = TW + KW - 1;
    constexpr int PITCH = pitch_for(TW, KW);
    constexpr int XS_PLANE = HALO_H * PITCH;
    constexpr int KHKW = KH * KW;
    constexpr int NX4 = (8 + KW - 1 + 3) / 4;
    static_assert(KW <= 9, "register window holds 16 floats");
    static_assert(TH % 2 == 0, "row pairs for pooling");

    __shared__ __align__(16) float xs[CI_T * XS_PLANE];
    __shared__ __align__(16) float ws[CI_T * KHKW * 32];

    const int NT = blockDim.x;
    const int tid = threadIdx.x;
    const int tiles = p.tiles_h * p.tiles_w;
    const int b = blockIdx.x / tiles;
    const int tile = blockIdx.x - b * tiles;
    const int th_i = tile / p.tiles_w, tw_i = tile - th_i * p.tiles_w;
    const int h0 = th_i * TH, w0 = tw_i * TW;
    const int z = blockIdx.y;  // 32-wide output-channel chunk

    const int cog = tid / POS_T;
    const int pp = tid - cog * POS_T;
    const int seg = pp / TH;
    const int row = pp - seg * TH;

    // input positions whose NEW traces this CTA writes (channel chunk 0 only)
    const int own_h_end = (th_i == p.tiles_h - 1) ? p.H : h0 + TH;
    const int own_w_end = (tw_i == p.tiles_w - 1) ? p.W : w0 + TW;
    const bool writer = (z == 0);

    int cq = -1, cI = -1;
    if (p.cells) {
        int2 c = __ldg(p.cells + b);
        cq = c.x;
        cI = c.y;
    }

    float acc[8][8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[j][c] = 0.f;

    for (int ci0 = 0; ci0 < p.Cin; ci0 += CI_T) {
        const int cin_here = min(CI_T, p.Cin - ci0);
        __syncthreads();
        // ---- weights of this channel chunk: [ci][tap][32 co], straight 16-byte async copies
        {
            const int n16 = cin_here * KHKW * 8;
            const float *src = p.wt + (size_t)ci0 * KHKW * p.CoutPad + z * 32;
            for (int i = tid; i < n16; i += NT) {
                int r = i >> 3, q = i & 7;
                cp_async16(ws + r * 32 + q * 4, src + (size_t)r * p.CoutPad + q * 4);
            }
        }
        // ---- trace recurrences on the tile + halo (dcll/pytorch_libdcll.py:415-416), one rounding per
        //      reference operation so that the traces are bit-identical to the reference's.  Elements are
        //      processed in batches of U with all loads issued before the first store: a store to the
        //      ping-pong half would otherwise fence the next element's loads (one exposed latency each).
        {
            constexpr int U = 4;
            const float *__restrict__ ge0 = p.e0_old;
            const float *__restrict__ ge1 = p.e1_old;
            const float *__restrict__ gx = p.x;
            float *__restrict__ ne0 = p.e0_new;
            float *__restrict__ ne1 = p.e1_new;
            const int n_el = cin_here * HALO_H * HALO_W;
            for (int idx0 = tid; idx0 < n_el; idx0 += U * NT) {
                float e0[U], e1[U], xin[U], cts[U], cas[U], cal[U], ctm[U];
                size_t off[U];
                int sidx[U];
                bool inimg[U], own[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int idx = idx0 + u * NT;
                    inimg[u] = false, own[u] = false, sidx[u] = -1, off[u] = 0;
                    e0[u] = e1[u] = xin[u] = cts[u] = cas[u] = cal[u] = ctm[u] = 0.f;
                    if (idx < n_el) {
                        int ci_l = idx / (HALO_H * HALO_W);
                        int rem = idx - ci_l * (HALO_H * HALO_W);
                        int r = rem / HALO_W;
                        int c = rem - r * HALO_W;
                        int gh = h0 - p.padH + r, gw = w0 - p.padW + c;
                        int ci = ci0 + ci_l;
                        sidx[u] = ci_l * XS_PLANE + r * PITCH + c;
                        if (gh >= 0 && gh < p.H && gw >= 0 && gw < p.W) {
                            inimg[u] = true;
                            own[u] = writer && gh >= h0 && gh < own_h_end && gw >= w0 && gw < own_w_end;
                            off[u] = ((size_t)(b * p.Cin + ci) * p.H + gh) * p.W + gw;
                            e0[u] = __ldg(ge0 + off[u]);
                            e1[u] = __ldg(ge1 + off[u]);
                            xin[u] = p.cells ? ((gh == cq && gw == cI) ? 1.f : 0.f) : __ldg(gx + off[u]);
                            int k = p.coef_mode == DCLL_COEF_SCALAR ? 0
                                                                    : (p.coef_mode == DCLL_COEF_CHANNEL ? ci : (ci * p.H + gh) * p.W + gw);
                            cts[u] = __ldg(p.tau_s + k), cas[u] = __ldg(p.alphas + k);
                            cal[u] = __ldg(p.alpha + k), ctm[u] = __ldg(p.tau_m + k);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (sidx[u] < 0) continue;
                    float n1 = 0.f;
                    if (inimg[u]) {
                        float n0 = __fadd_rn(__fmul_rn(xin[u], cts[u]), __fmul_rn(cas[u], e0[u]));
                        n1 = __fadd_rn(__fmul_rn(cal[u], e1[u]), __fmul_rn(n0, ctm[u]));
                        if (own[u]) ne0[off[u]] = n0, ne1[off[u]] = n1;
                    }
                    xs[sidx[u]] = n1;
                }
            }
        }
        cp_async_wait_all();
        __syncthreads();
        // ---- direct convolution on the chunk
        for (int ci_l = 0; ci_l < cin_here; ++ci_l) {
#pragma unroll 1
            for (int kh = 0; kh < KH; ++kh) {
                const float *xrow = xs + ci_l * XS_PLANE + (row + kh) * PITCH + seg * 8;
                float xr[NX4 * 4];
#pragma unroll
                for (int q = 0; q < NX4; ++q) {
                    float4 v = *reinterpret_cast<const float4 *>(xrow + 4 * q);
                    xr[4 * q] = v.x, xr[4 * q + 1] = v.y, xr[4 * q + 2] = v.z, xr[4 * q + 3] = v.w;
                }
                const float *wrow = ws + (ci_l * KH + kh) * KW * 32 + cog * 8;
#pragma unroll
                for (int kw = 0; kw < KW; ++kw) {
                    float4 wa = *reinterpret_cast<const float4 *>(wrow + kw * 32);
                    float4 wb = *reinterpret_cast<const float4 *>(wrow + kw * 32 + 4);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float xv = xr[j + kw];
                        acc[j][0] = fmaf(xv, wa.x, acc[j][0]);
                        acc[j][1] = fmaf(xv, wa.y, acc[j][1]);
                        acc[j][2] = fmaf(xv, wa.z, acc[j][2]);
                        acc[j][3] = fmaf(xv, wa.w, acc[j][3]);
                        acc[j][4] = fmaf(xv, wb.x, acc[j][4]);
                        acc[j][5] = fmaf(xv, wb.y, acc[j][5]);
                        acc[j][6] = fmaf(xv, wb.z, acc[j][6]);
                        acc[j][7] = fmaf(xv, wb.w, acc[j][7]);
                    }
                }
            }
        }
    }

    // ---------------------------------------------------------------------------------------------
    // epilogue
    // ---------------------------------------------------------------------------------------------
    const int oh = h0 + row, ow0 = w0 + seg * 8;
    const bool row_ok = oh < p.Hc;
    const bool refr = p.wrp > 0.f;
    const bool vec_c = ((p.Wc & 3) == 0) && (ow0 + 8 <= p.Wc);  // conv-grid rows are float4 addressable
    const int ohp = oh / PH;                                    // pooled row
    const int owp0 = ow0 / PW;                                  // first pooled column
    constexpr int NPW = 8 / PW;                                 // pooled columns per thread
    const bool prow_ok = ohp < p.Hp;
    const bool vec_p = ((p.Wp & 3) == 0) && (owp0 + NPW <= p.Wp);

#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int co = z * 32 + cog * 8 + c;
        const bool co_ok = co < p.Cout;  // warp-uniform when POS_T >= 32
        const float bv = co_ok ? __ldg(p.bias + co) : 0.f;
        const size_t off = ((size_t)(b * p.Cout + (co_ok ? co : 0)) * p.Hc + (row_ok ? oh : 0)) * p.Wc + ow0;
        float u[8], a_old[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) u[j] = __fadd_rn(acc[j][c], bv), a_old[j] = 0.f;
        const bool any = row_ok && co_ok;
        if (refr && any) {  // arp = alpharp * state.arp ; u = pvmem + arp   (:497-498)
            if (vec_c) {
                float4 v0 = *reinterpret_cast<const float4 *>(p.arp + off);
                float4 v1 = *reinterpret_cast<const float4 *>(p.arp + off + 4);
                a_old[0] = v0.x, a_old[1] = v0.y, a_old[2] = v0.z, a_old[3] = v0.w;
                a_old[4] = v1.x, a_old[5] = v1.y, a_old[6] = v1.z, a_old[7] = v1.w;
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (ow0 + j < p.Wc) a_old[j] = p.arp[off + j];
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                a_old[j] = __fmul_rn(p.alpharp, a_old[j]);
                u[j] = __fadd_rn(u[j], a_old[j]);
            }
        }
        float sp[8], pvv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            sp[j] = u[j] > 0.f ? 1.f : 0.f;   // :420 / :499
            pvv[j] = sigmoidf_ref(u[j]);      // :419 / :500
        }
        if (any) {
            if (refr) {  // arp -= output * wrp  (:503)
#pragma unroll
                for (int j = 0; j < 8; ++j) a_old[j] = __fsub_rn(a_old[j], __fmul_rn(sp[j], p.wrp));
                if (vec_c) {
                    *reinterpret_cast<float4 *>(p.arp + off) = make_float4(a_old[0], a_old[1], a_old[2], a_old[3]);
                    *reinterpret_cast<float4 *>(p.arp + off + 4) = make_float4(a_old[4], a_old[5], a_old[6], a_old[7]);
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (ow0 + j < p.Wc) p.arp[off + j] = a_old[j];
                }
            }
            if (p.pvmem) {
                if (vec_c) {
                    *reinterpret_cast<float4 *>(p.pvmem + off) = make_float4(u[0], u[1], u[2], u[3]);
                    *reinterpret_cast<float4 *>(p.pvmem + off + 4) = make_float4(u[4], u[5], u[6], u[7]);
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (ow0 + j < p.Wc) p.pvmem[off + j] = u[j];
                }
            }
        }
        // ---- max-pool (k = stride, no padding for k <= 2): PyTorch keeps the FIRST maximum in
        //      row-major window order (strict '>' update), both for the value and for the index
        float ps[NPW], pp_[NPW];
        int pi[NPW];
#pragma unroll
        for (int q = 0; q < NPW; ++q) {
            if (PW == 2) {
                bool second = pvv[2 * q + 1] > pvv[2 * q];
                pp_[q] = second ? pvv[2 * q + 1] : pvv[2 * q];
                pi[q] = second ? 1 : 0;
                ps[q] = fmaxf(sp[2 * q], sp[2 * q + 1]);
            } else {
                pp_[q] = pvv[q], pi[q] = 0, ps[q] = sp[q];
            }
        }
        if (PH == 2) {
#pragma unroll
            for (int q = 0; q < NPW; ++q) {
                float o_p = __shfl_xor_sync(0xffffffffu, pp_[q], 1);
                int o_i = __shfl_xor_sync(0xffffffffu, pi[q], 1);
                float o_s = __shfl_xor_sync(0xffffffffu, ps[q], 1);
                // even row = top of the window; bottom wins only when strictly greater
                if (o_p > pp_[q]) pp_[q] = o_p, pi[q] = 2 + o_i;
                ps[q] = fmaxf(ps[q], o_s);
            }
        }
        const bool pool_writer = (PH == 1) || ((row & 1) == 0);
        if (pool_writer && prow_ok && co_ok) {
            const size_t poff = ((size_t)(b * p.Cout + co) * p.Hp + ohp) * p.Wp + owp0;
            if (vec_p) {
#pragma unroll
                for (int q = 0; q < NPW; q += 4) {
                    *reinterpret_cast<float4 *>(p.spikes + poff + q) = make_float4(ps[q], ps[q + 1], ps[q + 2], ps[q + 3]);
                    *reinterpret_cast<float4 *>(p.pv + poff + q) = make_float4(pp_[q], pp_[q + 1], pp_[q + 2], pp_[q + 3]);
                }
            } else {
#pragma unroll
                for (int q = 0; q < NPW; ++q)
                    if (owp0 + q < p.Wp) p.spikes[poff + q] = ps[q], p.pv[poff + q] = pp_[q];
            }
            if (PH * PW > 1 && p.pool_idx) {
#pragma unroll
                for (int q = 0; q < NPW; ++q)
                    if (owp0 + q < p.Wp) p.pool_idx[poff + q] = (uint8_t)pi[q];
            }
        }
    }
}

// [Cout,Cin,KH,KW] -> [Cin,KH*KW,CoutPad] (zero padded)
__global__ void weight_transpose_kernel(const float *__restrict__ w, float *__restrict__ wt, int Cout, int CoutPad,
                                        int CinKK) {
    pdl_entry();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= CinKK * CoutPad) return;
    int co = i % CoutPad, r = i / CoutPad;
    wt[i] = co < Cout ? w[(size_t)co * CinKK + r] : 0.f;
}

// Quantised-weight mode: one CTA per output channel computes s = max|w|/127 (1 if the row is zero) and writes the
// dequantised row clamp(rint(w/s),-127,127)*s into the kernel-side layout [Cin*KH*KW, CoutPad] (and, through
// `deq`, a dense [Cout, Cin*KH*KW] copy that feeds the tensor-core weight split).
__global__ void __launch_bounds__(256) weight_fakequant_transpose_kernel(const float *__restrict__ w, float *__restrict__ wt,
                                                                         float *__restrict__ deq, int CoutPad, int CinKK) {
    pdl_entry();
    __shared__ float red[8];
    __shared__ float s_scale;
    const int co = blockIdx.x;
    const float *row = w + (size_t)co * CinKK;
    float m = 0.f;
    for (int i = threadIdx.x; i < CinKK; i += 256) m = fmaxf(m, fabsf(row[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        float mm = 0.f;
        for (int i = 0; i < 8; ++i) mm = fmaxf(mm, red[i]);
        s_scale = mm > 0.f ? __fdiv_rn(mm, 127.f) : 1.f;
    }
    __syncthreads();
    const float s = s_scale;
    for (int i = threadIdx.x; i < CinKK; i += 256) {
        float q = fminf(fmaxf(rintf(__fdiv_rn(row[i], s)), -127.f), 127.f);
        float v = __fmul_rn(q, s);
        wt[(size_t)i * CoutPad + co] = v;
        if (deq) deq[(size_t)co * CinKK + i] = v;
    }
}

template <int KH, int KW, int TH, int SEGS, int CI_T, int PH, int PW>
static int launch_inst(const FwdP &p, int B, cudaStream_t st) {
    int ncog = min(4, ceil_div(p.Cout, 8));
    dim3 grid((unsigned)(p.tiles_h * p.tiles_w * B), ceil_div(p.Cout, 32));
    launch_k(conv_fwd_kernel<KH, KW, TH, SEGS, CI_T, PH, PW>, grid, TH * SEGS * ncog, 0, st, p);
    DCLL_LAUNCH_OK("conv_fwd_kernel");
    return DCLL_OK;
}

template <int KH, int KW, int CI_T, int PH, int PW>
static int launch_tile(FwdP &p, int B, cudaStream_t st) {
    if (p.Wc > 16) {
        p.tiles_h = ceil_div(p.Hc, 16), p.tiles_w = ceil_div(p.Wc, 32);
        return launch_inst<KH, KW, 16, 4, CI_T, PH, PW>(p, B, st);
    }
    p.tiles_h = ceil_div(p.Hc, 16), p.tiles_w = ceil_div(p.Wc, 16);
    return launch_inst<KH, KW, 16, 2, CI_T, PH, PW>(p, B, st);
}

// Square kernels ship with pooling (1,1) and (2,2); the 1 x KW kernels of radio_ml_conv_ref.yaml with
// (1,1) and (1,2).  Anything else is reported, not emulated.
template <int KH, int KW, int CI_T>
static int launch_pool(FwdP &p, int B, int PH, int PW, cudaStream_t st) {
    if (PH == 1 && PW == 1) return launch_tile<KH, KW, CI_T, 1, 1>(p, B, st);
    if constexpr (KH > 1) {
        if (PH == 2 && PW == 2) return launch_tile<KH, KW, CI_T, 2, 2>(p, B, st);
    } else {
        if (PH == 1 && PW == 2) return launch_tile<KH, KW, CI_T, 1, 2>(p, B, st);
    }
    set_error("pooling (%d,%d) with kernel (%d,%d) has no sm_100a instantiation", PH, PW, KH, KW);
    return DCLL_EUNSUPPORTED;
}

// weight -> weight_t (and weight_mma): plain transpose, or the int8 quantise->dequantise image in quantised mode.
// Called after every weight change (Adam step, load_state_dict).
int sync_kernel_weights(const dcll_conv_layer *L, cudaStream_t st) {
    Geo g = geo_of(L);
    const int cinkk = L->Cin * L->KH * L->KW;
    if (L->quantized) {
        // the dequantised dense copy for the tensor-core split lives at the tail of weight_t's allocation
        float *deq = L->weight_mma ? L->weight_t + (size_t)cinkk * g.CoutPad : nullptr;
        if (g.CoutPad > L->Cout) DCLL_CUDA_OK(cudaMemsetAsync(L->weight_t, 0, sizeof(float) * (size_t)cinkk * g.CoutPad, st));
        launch_k(weight_fakequant_transpose_kernel, L->Cout, 256, 0, st, L->weight, L->weight_t, deq, g.CoutPad, cinkk);
        DCLL_LAUNCH_OK("weight_fakequant_transpose_kernel");
        if (L->weight_mma) return launch_weight_mma(L, deq, st);
        return DCLL_OK;
    }
    const int n = cinkk * g.CoutPad;
    launch_k(weight_transpose_kernel, ceil_div(n, 256), 256, 0, st, L->weight, L->weight_t, L->Cout, g.CoutPad, cinkk);
    DCLL_LAUNCH_OK("weight_transpose_kernel");
    if (L->weight_mma) return launch_weight_mma(L, L->weight, st);
    return DCLL_OK;
}

int launch_conv_fwd(const dcll_conv_layer *L, const void *x, cudaStream_t st, const dcll_conv_layer *next, bool trace_done,
                    int spike_io) {
    // "bf16x3" = tensor cores wherever a shape is instantiated; the rest stays on the FMA pipe
    if (prec_tc(L) && tc_supported(L)) return launch_conv_fwd_tc(L, x, st, next, trace_done, spike_io);
    DCLL_REQUIRE(!next && !trace_done && !(spike_io & (SPK_PACKED | SPK_X_PACKED)), DCLL_EINVAL,
                 "fused next-layer trace / packed spikes need the tensor-core path");
    Geo g = geo_of(L);
    FwdP p;
    p.x = L->x_mode == DCLL_X_DENSE ? (const float *)x : nullptr;
    p.cells = L->x_mode == DCLL_X_CELLS ? (const int2 *)x : nullptr;
    int cur = L->cur & 1;
    p.e0_old = L->eps0[cur], p.e1_old = L->eps1[cur];
    p.e0_new = L->eps0[cur ^ 1], p.e1_new = L->eps1[cur ^ 1];
    p.alpha = L->alpha, p.alphas = L->alphas, p.tau_m = L->tau_m, p.tau_s = L->tau_s;
    p.wt = L->weight_t, p.bias = L->bias;
    p.arp = L->arp, p.spikes = L->spikes, p.pv = L->pv, p.pvmem = L->write_pvmem ? L->pvmem : nullptr;
    p.pool_idx = L->pool_idx;
    p.alpharp = L->alpharp, p.wrp = L->wrp, p.coef_mode = L->coef_mode;
    p.B = L->B, p.Cin = L->Cin, p.H = L->H, p.W = L->W, p.Cout = L->Cout, p.CoutPad = g.CoutPad;
    p.padH = L->padH, p.padW = L->padW, p.Hc = g.Hc, p.Wc = g.Wc, p.Hp = g.Hp, p.Wp = g.Wp;
    p.tiles_h = p.tiles_w = 0;
    const int kh = L->KH, kw = L->KW;
    if (kh == 7 && kw == 7) return launch_pool<7, 7, 4>(p, L->B, L->poolH, L->poolW, st);
    if (kh == 5 && kw == 5) return launch_pool<5, 5, 4>(p, L->B, L->poolH, L->poolW, st);
    if (kh == 3 && kw == 3) return launch_pool<3, 3, 8>(p, L->B, L->poolH, L->poolW, st);
    if (kh == 1 && kw == 3) return launch_pool<1, 3, 8>(p, L->B, L->poolH, L->poolW, st);
    set_error("kernel_size (%d,%d) has no sm_100a instantiation (have 7x7, 5x5, 3x3, 1x3)", kh, kw);
    return DCLL_EUNSUPPORTED;
}

}  // namespace dcll

using namespace dcll;

extern "C" __attribute__((visibility("default"))) int dcll_conv_sync_weights(const dcll_conv_layer *L, void *stream) {
    DCLL_REQUIRE(L && L->weight && L->weight_t, DCLL_EINVAL, "dcll_conv_sync_weights: null pointer");
    return sync_kernel_weights(L, (cudaStream_t)stream);
}
