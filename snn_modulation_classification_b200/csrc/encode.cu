// IQ -> spike-cell encoder and dense frame materialisation.
// Replaces data/utils.py:43-87 (iq2spiketrain) of the reference.
//
// HBM-bound streaming kernels: 8 B in / 8 B out per sample-timestep for the encoder.
#include "common.cuh"

namespace dcll {

// One I or Q value -> cell index.  Every reference operation (data/utils.py:65-79) is one
// float32 rounding; the gamma power is evaluated in double with the float32-rounded exponent
// float(1/1.2) = 0.8333333134651184 and rounded once, which reproduces torch-CPU pow
// cell-for-cell.
__device__ __forceinline__ int encode_axis(float v, float lo, float span, int n, int do_gamma) {
    float c = __fdiv_rn(__fsub_rn(v, lo), span);
    if (do_gamma) {
        c = __fsub_rn(__fmul_rn(c, 2.0f), 1.0f);
        float mag = (float)pow((double)fabsf(c), 0.8333333134651184);
        float sg = c > 0.f ? 1.f : (c < 0.f ? -1.f : 0.f);
        c = __fmul_rn(sg, mag);
        c = __fmul_rn(__fadd_rn(c, 1.0f), 0.5f);
    }
    c = fminf(fmaxf(c, 0.f), 1.f);
    c = __fmul_rn(c, (float)(n - 1));
    return __float2int_rz(c);
}

// x [B,2,N] (t contiguous) -> cells [T,B,2] (b contiguous): 32x32 transpose through shared memory
// so that both the reads (along t) and the int2 writes (along b) are coalesced.
__global__ void __launch_bounds__(256) iq_encode_kernel(const float *__restrict__ x, int B, int N, float lo_I,
                                                        float span_I, float lo_Q, float span_Q, int out_w,
                                                        int out_h, int t_start, int T, int do_gamma,
                                                        int2 *__restrict__ cells) {
    pdl_entry();
    __shared__ int2 tile[32][33];
    const int t0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        int b = b0 + ty + 8 * r, t = t0 + tx;
        if (b < B && t < T) {
            const float *row = x + (size_t)b * 2 * N + t_start + t;
            float vi = __ldg(row), vq = __ldg(row + N);
            int ci = encode_axis(vi, lo_I, span_I, out_w, do_gamma);
            int cq = encode_axis(vq, lo_Q, span_Q, out_h, do_gamma);
            tile[ty + 8 * r][tx] = make_int2(cq, ci);  // (row = Q, col = I), data/utils.py:82
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        int t = t0 + ty + 8 * r, b = b0 + tx;
        if (b < B && t < T) cells[(size_t)t * B + b] = tile[tx][ty + 8 * r];
    }
}

// frames [T*B][H*W] float32, exactly one 1.0 per frame (data/utils.py:57,81-82).
template <int VEC>
__global__ void __launch_bounds__(256) cells_to_frames_kernel(const int2 *__restrict__ cells, size_t n_frames, int HW,
                                                              int W, float *__restrict__ frames) {
    pdl_entry();
    const int per_frame = HW / VEC;
    size_t total = n_frames * (size_t)per_frame;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t fr = i / per_frame;
        int p = (int)(i - fr * per_frame) * VEC;
        int2 c = __ldg(cells + fr);
        int hot = c.x * W + c.y;
        if (VEC == 4) {
            float4 v = make_float4(p == hot, p + 1 == hot, p + 2 == hot, p + 3 == hot);
            reinterpret_cast<float4 *>(frames)[i] = v;
        } else {
            frames[i] = (p == hot) ? 1.f : 0.f;
        }
    }
}

// image2spiketrain (data/utils.py:15-40): frozen Poisson spike train of an image.
//     p[b,n] = (1000 - gain * x[b,n]) / 1000   (float32, as numpy evaluates it);  spike[t,b,n] = t < T_b && !(u[t,b,n] < p[b,n])
// u: the host-drawn uniforms of the reference's numpy stream (float64 [B][Tmax][Nin], compared in double: bit-exact with the
// reference for the same draws), or -- u == NULL -- a counter-based generator (one 64-bit mix per element, seeded; the same
// distribution, not the numpy stream).  HBM-bound: 4 B written per element (+ 8 B read in the parity mode).
__device__ __forceinline__ double mix_uniform(unsigned long long seed, unsigned long long idx) {
    unsigned long long z = seed + (idx + 1) * 0x9E3779B97F4A7C15ull;      // splitmix64
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);                // 53 random bits -> [0, 1)
}

__global__ void __launch_bounds__(256) image_encode_kernel(const float *__restrict__ x, const double *__restrict__ u,
                                                           const int32_t *__restrict__ t_len, int B, int Nin, int Tmax, float gain,
                                                           unsigned long long seed, float *__restrict__ out) {
    pdl_entry();
    const size_t total = (size_t)Tmax * B * Nin;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int n = (int)(i % Nin);
        const size_t tb = i / Nin;
        const int b = (int)(tb % B), t = (int)(tb / B);
        float v = 0.f;
        if (t < __ldg(t_len + b)) {
            const float rate = __fmul_rn(gain, __ldg(x + (size_t)b * Nin + n));
            const float p = __fdiv_rn(__fsub_rn(1000.0f, rate), 1000.0f);
            const double r = u ? __ldg(u + ((size_t)b * Tmax + t) * Nin + n) : mix_uniform(seed, i);
            v = (r < (double)p) ? 0.f : 1.f;
        }
        out[i] = v;
    }
}

}  // namespace dcll

using namespace dcll;

extern "C" __attribute__((visibility("default"))) int dcll_image_encode(const float *x, const double *u, const int32_t *t_len, int B, int Nin, int Tmax,
                                                                         double gain, uint64_t seed, float *out, void *stream) {
    DCLL_REQUIRE(x && t_len && out && B > 0 && Nin > 0 && Tmax > 0, DCLL_EINVAL, "dcll_image_encode: bad arguments");
    const size_t total = (size_t)Tmax * B * Nin;
    const int blocks = (int)((total + 255) / 256 < (size_t)148 * 16 ? (total + 255) / 256 : (size_t)148 * 16);
    launch_k(image_encode_kernel, blocks, 256, 0, (cudaStream_t)stream, x, u, t_len, B, Nin, Tmax, (float)gain, (unsigned long long)seed, out);
    DCLL_LAUNCH_OK("image_encode_kernel");
    return DCLL_OK;
}

extern "C" __attribute__((visibility("default"))) int dcll_iq_encode(const float *x, int B, int N, double min_I, double max_I, double min_Q, double max_Q,
                              int out_w, int out_h, int t_start, int T, int do_gamma, int32_t *cells, void *stream) {
    DCLL_REQUIRE(x && cells, DCLL_EINVAL, "dcll_iq_encode: null pointer");
    DCLL_REQUIRE(B > 0 && N > 0 && T > 0 && out_w > 0 && out_h > 0, DCLL_EINVAL, "dcll_iq_encode: bad sizes");
    // assert max_duration <= num_timesteps (data/utils.py:56)
    DCLL_REQUIRE(t_start >= 0 && t_start + T <= N, DCLL_EINVAL,
                 "dcll_iq_encode: window [%d,%d) exceeds the %d samples of the record", t_start, t_start + T, N);
    dim3 grid(ceil_div(T, 32), ceil_div(B, 32));
    launch_k(iq_encode_kernel, grid, 256, 0, (cudaStream_t)stream, x, B, N, (float)min_I, (float)(max_I - min_I), (float)min_Q,
                                                             (float)(max_Q - min_Q), out_w, out_h, t_start, T, do_gamma,
                                                             reinterpret_cast<int2 *>(cells));
    DCLL_LAUNCH_OK("iq_encode_kernel");
    return DCLL_OK;
}

extern "C" __attribute__((visibility("default"))) int dcll_cells_to_frames(const int32_t *cells, int T, int B, int H, int W, float *frames, void *stream) {
    DCLL_REQUIRE(cells && frames && T > 0 && B > 0 && H > 0 && W > 0, DCLL_EINVAL, "dcll_cells_to_frames: bad args");
    size_t n_frames = (size_t)T * B;
    int HW = H * W;
    bool vec = (HW % 4 == 0) && ((uintptr_t)frames % 16 == 0);
    size_t total = n_frames * (size_t)(vec ? HW / 4 : HW);
    int blocks = (int)((total + 255) / 256 < (size_t)148 * 16 ? (total + 255) / 256 : (size_t)148 * 16);
    if (vec)
        launch_k(cells_to_frames_kernel<4>, blocks, 256, 0, (cudaStream_t)stream, reinterpret_cast<const int2 *>(cells),
                                                                            n_frames, HW, W, frames);
    else
        launch_k(cells_to_frames_kernel<1>, blocks, 256, 0, (cudaStream_t)stream, reinterpret_cast<const int2 *>(cells),
                                                                            n_frames, HW, W, frames);
    DCLL_LAUNCH_OK("cells_to_frames_kernel");
    return DCLL_OK;
}
