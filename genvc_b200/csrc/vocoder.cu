// HiFi-GAN generator building blocks (the stage right after the codec-token path):
//   reference  layers/hifigan.py:118-153 (ResBlock2), :28-116 (ResBlock1), :156-232 (HiFiGAN.forward)
// Two stateless kernels behind the C ABI (include/genvc_b200.h: genvc_conv1d, genvc_conv_transpose1d); the host side
// (genvc_b200/vocoder.py) strings them together exactly as HiFiGAN.forward does.  Everything that surrounds a convolution in
// the reference is folded into it:
//   * leaky_relu in front of the conv  -> applied while the input tile is staged in shared memory (pre_slope; 1 = none)
//   * bias, the residual `xt + x`       -> epilogue
//   * `xs += resblock(x)`, `xs / num_kernels` -> epilogue accumulates into y and scales (accumulate, out_scale)
//   * tanh of conv_post                 -> epilogue
// so a ResBlock2 is two launches and no elementwise kernel exists.  fp32 FFMA out of shared memory: the whole generator is
// 4.5 GFLOP per second of audio (channels 256 -> 32), i.e. launch- and latency-bound, not a tensor-core problem at
// streaming chunk sizes (32 frames): the tiles are sized so that even the last stage (32 channels, 24 000 samples per
// second of audio) fills the 148 SMs.
// Weights are repacked by the host to [Cin][K][Cout] (Cout contiguous): a warp reads the four output channels of a thread
// as one broadcast LDS.128.
#include "common.cuh"
#include "../../include/genvc_b200.h"
#include <algorithm>

namespace gv {

constexpr int VC_THREADS = 256;
constexpr int VC_CO_T = 32;    // output channels per block (8 channel groups x 4)
constexpr int VC_T_T = 128;    // output samples per block (32 sample groups x 4)
constexpr int VC_CI_C = 16;    // input channels staged per round

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.0f ? v : v * slope; }

// y[b, co, t] = epi( bias[co] + sum_ci sum_j w[ci][j][co] * lrelu(x[b, ci, t + j * dil - pad]) )
__global__ void __launch_bounds__(VC_THREADS)
conv1d_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
              const float* __restrict__ res, float* __restrict__ y, int Cin, int Cout, int T, int To, int K, int dil, int pad,
              int stride, float pre_slope, int accumulate, float out_scale, int act_out, int ksplit, float* __restrict__ scratch) {
    extern __shared__ float sm[];
    const int halo = (K - 1) * dil;
    const int XW = (VC_T_T - 1) * stride + halo + 1;  // staged samples per input channel (T = input, To = output length)
    float* xin = sm;                          // [VC_CI_C][XW]
    float* ws = sm + VC_CI_C * XW;            // [VC_CI_C][K][VC_CO_T]
    const int tid = threadIdx.x;
    const int cg = tid >> 5, tg = tid & 31;   // channel group (warp: 4 channels), lane: outputs t0 + tg + {0, 32, 64, 96}
    // split over the input channels (small layers: too few output tiles to fill the GPU): slice ks of ksplit writes its
    // partial sums to scratch[ks][b][co][t]; splitk_epilogue_kernel adds the slices in a fixed order
    const int t0 = blockIdx.x * VC_T_T, co0 = blockIdx.y * VC_CO_T, b = blockIdx.z / ksplit, ks = blockIdx.z % ksplit;
    const int per = ((Cin + ksplit - 1) / ksplit + VC_CI_C - 1) / VC_CI_C * VC_CI_C;
    const int c_begin = ks * per, c_end = min(Cin, c_begin + per);
    const float* xb = x + (size_t)b * Cin * T;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[i][q] = 0.0f;

    // Software pipeline: the global loads of round r + 1 are issued into registers before round r is computed, and
    // written to shared memory after it -- a block has few rounds' worth of parallelism (one block per SM at streaming
    // sizes), so an exposed L2 / HBM latency per round would dominate the layer.
    // Staging layout: warp cg stages input channels cg and cg + 8 of the round (lane = sample, 32 apart) and their K weight
    // rows (lane = output channel) -- no index arithmetic per element.  Bounds (host-checked): XW <= 9 * 32, K <= 12.
    constexpr int XM = 9, KM = 12, CH = VC_CI_C / 8;
    float xreg[CH][XM], wreg[CH][KM];
    auto fetch = [&](int c0) {
#pragma unroll
        for (int h = 0; h < CH; ++h) {
            const int ci = c0 + cg + 8 * h;
            const bool cok = ci < c_end;
            const float* xrow = xb + (size_t)ci * T;
#pragma unroll
            for (int m = 0; m < XM; ++m) {
                const int sidx = tg + 32 * m;
                const int t = t0 * stride + sidx - pad;
                xreg[h][m] = (cok && sidx < XW && t >= 0 && t < T) ? lrelu(__ldg(xrow + t), pre_slope) : 0.0f;
            }
            const float* wrow = w + (size_t)ci * K * Cout + co0 + tg;
            const bool wok = cok && co0 + tg < Cout;
#pragma unroll
            for (int j = 0; j < KM; ++j) wreg[h][j] = (wok && j < K) ? __ldg(wrow + (size_t)j * Cout) : 0.0f;
        }
    };
    auto commit = [&]() {
#pragma unroll
        for (int h = 0; h < CH; ++h) {
            const int cl = cg + 8 * h;
#pragma unroll
            for (int m = 0; m < XM; ++m) {
                const int sidx = tg + 32 * m;
                if (sidx < XW) xin[cl * XW + sidx] = xreg[h][m];
            }
#pragma unroll
            for (int j = 0; j < KM; ++j)
                if (j < K) ws[(cl * K + j) * VC_CO_T + tg] = wreg[h][j];
        }
    };
    if (c_begin < c_end) fetch(c_begin);
    for (int c0 = c_begin; c0 < c_end; c0 += VC_CI_C) {
        commit();
        __syncthreads();
        if (c0 + VC_CI_C < c_end) fetch(c0 + VC_CI_C);
#pragma unroll 2
        for (int ci = 0; ci < VC_CI_C; ++ci) {
            // the thread's four outputs are t0 + tg + 32 q: for a fixed q the 32 lanes read consecutive words (no bank
            // conflicts; a 4-consecutive-outputs layout makes every one of these loads a 4-way conflict and the kernel
            // shared-memory bound)
            const float* xr = xin + ci * XW + tg * stride;
            const float* wr = ws + ci * K * VC_CO_T + cg * 4;
            for (int j = 0; j < K; ++j) {
                const float4 wv = *reinterpret_cast<const float4*>(wr + j * VC_CO_T);
                const float* xp = xr + j * dil;
                const float x0 = xp[0], x1 = xp[32 * stride], x2 = xp[64 * stride], x3 = xp[96 * stride];
                acc[0][0] = fmaf(wv.x, x0, acc[0][0]); acc[0][1] = fmaf(wv.x, x1, acc[0][1]);
                acc[0][2] = fmaf(wv.x, x2, acc[0][2]); acc[0][3] = fmaf(wv.x, x3, acc[0][3]);
                acc[1][0] = fmaf(wv.y, x0, acc[1][0]); acc[1][1] = fmaf(wv.y, x1, acc[1][1]);
                acc[1][2] = fmaf(wv.y, x2, acc[1][2]); acc[1][3] = fmaf(wv.y, x3, acc[1][3]);
                acc[2][0] = fmaf(wv.z, x0, acc[2][0]); acc[2][1] = fmaf(wv.z, x1, acc[2][1]);
                acc[2][2] = fmaf(wv.z, x2, acc[2][2]); acc[2][3] = fmaf(wv.z, x3, acc[2][3]);
                acc[3][0] = fmaf(wv.w, x0, acc[3][0]); acc[3][1] = fmaf(wv.w, x1, acc[3][1]);
                acc[3][2] = fmaf(wv.w, x2, acc[3][2]); acc[3][3] = fmaf(wv.w, x3, acc[3][3]);
            }
        }
        __syncthreads();
    }
    if (ksplit > 1) {
        const int Bn = gridDim.z / ksplit;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int co = co0 + cg * 4 + i;
            if (co >= Cout) continue;
            float* prow = scratch + (((size_t)ks * Bn + b) * Cout + co) * To;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int t = t0 + tg + 32 * q;
                if (t < To) prow[t] = acc[i][q];
            }
        }
        return;
    }
    // epilogue: bias, residual, accumulate / scale, output activation (1 tanh, 2 relu)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int co = co0 + cg * 4 + i;
        if (co >= Cout) continue;
        const float bv = bias ? __ldg(bias + co) : 0.0f;
        const size_t row = ((size_t)b * Cout + co) * To;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int t = t0 + tg + 32 * q;
            if (t >= To) continue;
            float v = acc[i][q] + bv;
            if (res) v += res[row + t];
            if (accumulate) v += y[row + t];
            v *= out_scale;
            if (act_out == 1) v = tanhf(v);
            else if (act_out == 2) v = fmaxf(v, 0.0f);
            y[row + t] = v;
        }
    }
}

// y[b, co, t] = bias[co] + sum_ci sum_{kk = (t + pad) % stride + m * stride < K} w[ci][kk][co] * lrelu(x[b, ci, (t + pad - kk) / stride])
__global__ void __launch_bounds__(VC_THREADS)
conv_transpose1d_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                        float* __restrict__ y, int Cin, int Cout, int Tin, int Tout, int K, int stride, int pad, float pre_slope,
                        int ksplit, float* __restrict__ scratch) {
    extern __shared__ float sm[];
    const int tid = threadIdx.x;
    const int cg = tid >> 5, tg = tid & 31;
    const int t0 = blockIdx.x * VC_T_T, co0 = blockIdx.y * VC_CO_T, b = blockIdx.z / ksplit, ks = blockIdx.z % ksplit;
    const int per = ((Cin + ksplit - 1) / ksplit + VC_CI_C - 1) / VC_CI_C * VC_CI_C;
    const int c_begin = ks * per, c_end = min(Cin, c_begin + per);
    // input samples that can reach outputs [t0, t0 + VC_T_T): s in [(t0 + pad - (K - 1)) / stride, (t0 + VC_T_T - 1 + pad) / stride]
    const int s_lo = (t0 + pad - (K - 1)) >= 0 ? (t0 + pad - (K - 1)) / stride : -(((K - 1) - t0 - pad + stride - 1) / stride);
    const int XW = (VC_T_T + K - 1) / stride + 2;
    float* xin = sm;                      // [VC_CI_C][XW]
    float* ws = sm + VC_CI_C * XW;        // [VC_CI_C][K][VC_CO_T]
    const float* xb = x + (size_t)b * Cin * Tin;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[i][q] = 0.0f;
    int kk0[4], sr0[4];  // first tap and its (staged) input index for each of the thread's four outputs
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int tp = t0 + tg + 32 * q + pad;
        kk0[q] = tp % stride;
        sr0[q] = tp / stride - s_lo;
    }
    for (int c0 = c_begin; c0 < c_end; c0 += VC_CI_C) {
        for (int e = tid; e < VC_CI_C * XW; e += VC_THREADS) {
            const int ci = e / XW, s = s_lo + (e - ci * XW);
            float v = 0.0f;
            if (c0 + ci < c_end && s >= 0 && s < Tin) v = lrelu(__ldg(xb + (size_t)(c0 + ci) * Tin + s), pre_slope);
            xin[e] = v;
        }
        for (int e = tid; e < VC_CI_C * K * VC_CO_T; e += VC_THREADS) {
            const int co = e % VC_CO_T, cj = e / VC_CO_T;
            const int ci = cj / K;
            float v = 0.0f;
            if (c0 + ci < c_end && co0 + co < Cout) v = __ldg(w + ((size_t)(c0 + ci) * K + (cj - ci * K)) * Cout + co0 + co);
            ws[e] = v;
        }
        __syncthreads();
        for (int ci = 0; ci < VC_CI_C; ++ci) {
            const float* xr = xin + ci * XW;
            const float* wr = ws + ci * K * VC_CO_T + cg * 4;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                int s = sr0[q];
                for (int kk = kk0[q]; kk < K; kk += stride, --s) {
                    const float xv = xr[s];  // s >= 0 by construction of s_lo; inputs outside [0, Tin) were staged as zeros
                    const float4 wv = *reinterpret_cast<const float4*>(wr + kk * VC_CO_T);
                    acc[0][q] = fmaf(wv.x, xv, acc[0][q]);
                    acc[1][q] = fmaf(wv.y, xv, acc[1][q]);
                    acc[2][q] = fmaf(wv.z, xv, acc[2][q]);
                    acc[3][q] = fmaf(wv.w, xv, acc[3][q]);
                }
            }
        }
        __syncthreads();
    }
    if (ksplit > 1) {
        const int Bn = gridDim.z / ksplit;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int co = co0 + cg * 4 + i;
            if (co >= Cout) continue;
            float* prow = scratch + (((size_t)ks * Bn + b) * Cout + co) * Tout;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int t = t0 + tg + 32 * q;
                if (t < Tout) prow[t] = acc[i][q];
            }
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int co = co0 + cg * 4 + i;
        if (co >= Cout) continue;
        const float bv = bias ? __ldg(bias + co) : 0.0f;
        const size_t row = ((size_t)b * Cout + co) * Tout;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int t = t0 + tg + 32 * q;
            if (t < Tout) y[row + t] = acc[i][q] + bv;
        }
    }
}

// Nearest codebook entry (layers/dvae.py:84-88, Quantize.forward at inference): for every (b, t)
//   dist[n] = sum_c f[c]^2 - 2 * sum_c f[c] * E[c][n] + sum_c E[c][n]^2 ;  code = argmax_n(-dist[n]), first index on ties.
// x is the encoder output as it leaves the last convolution, [B, dim, T] (the reference permutes to [B, T, dim] first).
// One block per (b, t); thread n owns code n (strided if n_embed > blockDim); E is [dim][n_embed], read coalesced.
__global__ void __launch_bounds__(256)
codebook_argmin_kernel(const float* __restrict__ x, const float* __restrict__ embed, long long* __restrict__ codes, int dim,
                       int n_embed, int T) {
    extern __shared__ float f[];  // [dim] feature vector of this position, then reduction scratch
    __shared__ float s_val[256];
    __shared__ int s_idx[256];
    const int b = blockIdx.x / T, t = blockIdx.x % T, tid = threadIdx.x;
    float part = 0.0f;
    for (int c = tid; c < dim; c += blockDim.x) {
        const float v = x[((size_t)b * dim + c) * T + t];
        f[c] = v;
        part += v * v;
    }
    s_val[tid] = part;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (tid < o) s_val[tid] += s_val[tid + o];
        __syncthreads();
    }
    const float f2 = s_val[0];
    __syncthreads();
    float best = -INFINITY;
    int best_n = 0x7fffffff;
    for (int n = tid; n < n_embed; n += blockDim.x) {
        float dot = 0.0f, e2 = 0.0f;
        for (int c = 0; c < dim; ++c) {
            const float e = __ldg(embed + (size_t)c * n_embed + n);
            dot = fmaf(f[c], e, dot);
            e2 = fmaf(e, e, e2);
        }
        const float neg = -((f2 - 2.0f * dot) + e2);
        if (neg > best) {  // n ascends: the first maximum is kept
            best = neg;
            best_n = n;
        }
    }
    s_val[tid] = best;
    s_idx[tid] = best_n;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (tid < o) {
            const float v = s_val[tid + o];
            const int i = s_idx[tid + o];
            if (v > s_val[tid] || (v == s_val[tid] && i < s_idx[tid])) {
                s_val[tid] = v;
                s_idx[tid] = i;
            }
        }
        __syncthreads();
    }
    if (tid == 0) codes[(size_t)b * T + t] = s_idx[0];
}

// sum of the ksplit partial slices in slice order, then the same epilogue as the unsplit kernels
__global__ void __launch_bounds__(256)
splitk_conv_epilogue_kernel(const float* __restrict__ scratch, const float* __restrict__ bias, const float* __restrict__ res,
                            float* __restrict__ y, int ksplit, int B, int Cout, int T, int accumulate, float out_scale, int act_tanh) {
    const size_t n = (size_t)B * Cout * T;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
        float v = 0.0f;
        for (int ks = 0; ks < ksplit; ++ks) v += scratch[(size_t)ks * n + e];
        const int co = (int)((e / T) % Cout);
        if (bias) v += __ldg(bias + co);
        if (res) v += res[e];
        if (accumulate) v += y[e];
        v *= out_scale;
        if (act_tanh == 1) v = tanhf(v);
        else if (act_tanh == 2) v = fmaxf(v, 0.0f);
        y[e] = v;
    }
}

// slices so that small layers still put ~2 blocks on every SM; each slice keeps at least one staging round
static int pick_ksplit(int tiles, int Cin, size_t out_elems, size_t scratch_floats) {
    if (scratch_floats == 0) return 1;
    int ks = 1;
    while (tiles * ks < 200 && ks * 2 * VC_CI_C <= Cin && (size_t)(ks * 2) * out_elems <= scratch_floats) ks *= 2;
    return ks;
}

}  // namespace gv

// ---------------------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------------------
extern "C" int genvc_conv1d(const float* x, const float* w, const float* bias, const float* residual, float* y, int B, int Cin,
                            int Cout, int T, int K, int dilation, int padding, int stride, float pre_slope, int accumulate,
                            float out_scale, int act_out, float* scratch, uint64_t scratch_floats, void* stream) {
    if (!x || !w || !y || B <= 0 || Cin <= 0 || Cout <= 0 || T <= 0 || K <= 0 || K > 12 || dilation <= 0 || padding < 0 ||
        stride <= 0 || stride > 2 || act_out < 0 || act_out > 2)
        return GENVC_E_INVALID;
    const int To = (T + 2 * padding - dilation * (K - 1) - 1) / stride + 1;  // torch.nn.Conv1d
    if (To <= 0 || (residual && To != T)) return GENVC_E_INVALID;
    const int XW = (gv::VC_T_T - 1) * stride + (K - 1) * dilation + 1;
    if (XW > 9 * 32 || K > 12) return GENVC_E_UNSUPPORTED;  // register staging of the kernel (conv1d_kernel: XM, KM)
    const size_t smem = ((size_t)gv::VC_CI_C * XW + (size_t)gv::VC_CI_C * K * gv::VC_CO_T) * sizeof(float);
    if (smem > 200 * 1024) return GENVC_E_INVALID;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(gv::conv1d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return GENVC_E_CUDA;
    dim3 grid((To + gv::VC_T_T - 1) / gv::VC_T_T, (Cout + gv::VC_CO_T - 1) / gv::VC_CO_T, B);
    const size_t out_elems = (size_t)B * Cout * To;
    const int ks = gv::pick_ksplit((int)(grid.x * grid.y * grid.z), Cin, out_elems, scratch ? (size_t)scratch_floats : 0);
    grid.z = B * ks;
    gv::conv1d_kernel<<<grid, gv::VC_THREADS, smem, (cudaStream_t)stream>>>(x, w, bias, residual, y, Cin, Cout, T, To, K, dilation,
                                                                             padding, stride, pre_slope, accumulate, out_scale,
                                                                             act_out, ks, scratch);
    if (ks > 1) {
        const int blocks = (int)std::min<size_t>((out_elems + 255) / 256, 1184);
        gv::splitk_conv_epilogue_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(scratch, bias, residual, y, ks, B, Cout, To, accumulate,
                                                                                  out_scale, act_out);
    }
    return cudaGetLastError() == cudaSuccess ? GENVC_OK : GENVC_E_CUDA;
}

extern "C" int genvc_conv_transpose1d(const float* x, const float* w, const float* bias, float* y, int B, int Cin, int Cout, int Tin,
                                      int K, int stride, int padding, float pre_slope, float* scratch, uint64_t scratch_floats,
                                      void* stream) {
    if (!x || !w || !y || B <= 0 || Cin <= 0 || Cout <= 0 || Tin <= 0 || K <= 0 || K > 64 || stride <= 0 || padding < 0)
        return GENVC_E_INVALID;
    const int Tout = (Tin - 1) * stride - 2 * padding + K;  // torch.nn.ConvTranspose1d, output_padding = 0, dilation = 1
    if (Tout <= 0) return GENVC_E_INVALID;
    const size_t smem = ((size_t)gv::VC_CI_C * ((gv::VC_T_T + K - 1) / stride + 2) + (size_t)gv::VC_CI_C * K * gv::VC_CO_T) * sizeof(float);
    if (smem > 200 * 1024) return GENVC_E_INVALID;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(gv::conv_transpose1d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return GENVC_E_CUDA;
    dim3 grid((Tout + gv::VC_T_T - 1) / gv::VC_T_T, (Cout + gv::VC_CO_T - 1) / gv::VC_CO_T, B);
    const size_t out_elems = (size_t)B * Cout * Tout;
    const int ks = gv::pick_ksplit((int)(grid.x * grid.y * grid.z), Cin, out_elems, scratch ? (size_t)scratch_floats : 0);
    grid.z = B * ks;
    gv::conv_transpose1d_kernel<<<grid, gv::VC_THREADS, smem, (cudaStream_t)stream>>>(x, w, bias, y, Cin, Cout, Tin, Tout, K, stride,
                                                                                       padding, pre_slope, ks, scratch);
    if (ks > 1) {
        const int blocks = (int)std::min<size_t>((out_elems + 255) / 256, 1184);
        gv::splitk_conv_epilogue_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(scratch, bias, nullptr, y, ks, B, Cout, Tout, 0, 1.0f, 0);
    }
    return cudaGetLastError() == cudaSuccess ? GENVC_OK : GENVC_E_CUDA;
}

extern "C" int genvc_codebook_argmin(const float* x, const float* embed, int64_t* codes, int B, int dim, int n_embed, int T, void* stream) {
    if (!x || !embed || !codes || B <= 0 || dim <= 0 || n_embed <= 0 || T <= 0 || dim > 8192) return GENVC_E_INVALID;
    gv::codebook_argmin_kernel<<<B * T, 256, (size_t)dim * sizeof(float), (cudaStream_t)stream>>>(x, embed, reinterpret_cast<long long*>(codes),
                                                                                                    dim, n_embed, T);
    return cudaGetLastError() == cudaSuccess ? GENVC_OK : GENVC_E_CUDA;
}

// ---------------------------------------------------------------------------------------------------------------------
// Mel front-end in front of the perceiver (utils.py:95-158 TorchMelSpectrogram = torchaudio MelSpectrogram, power 2, centre /
// reflect padding, + log(clamp(., 1e-5)) / mel_norms).  One CTA per (frame, batch element): the windowed frame and the
// n_fft twiddle factors live in shared memory, thread f accumulates bin f by direct summation over the non-zero part of
// the window (four independent partial sums: shorter dependency chain and smaller rounding growth than one running sum),
// the power spectrum stays in shared memory, thread m < n_mels applies the filterbank column, log, normalisation.
// 1.2 GFLOP for 6 s of reference audio at n_fft = 2048 -- the point is one launch and no intermediate tensors, not an FFT.
// window: [n_fft] (Hann of win_length centred, zeros outside [w_lo, w_hi)); tw: [n_fft][2] = cos, sin(2 pi k / n_fft);
// fbank: [n_fft / 2 + 1][n_mels]; norms: [n_mels] or NULL; mel: [B, n_mels, T].
// ---------------------------------------------------------------------------------------------------------------------
namespace gv {
__global__ void __launch_bounds__(256)
mel_frontend_kernel(const float* __restrict__ wav, int N, const float* __restrict__ window, const float* __restrict__ tw,
                    const float* __restrict__ fbank, const float* __restrict__ norms, float* __restrict__ mel, int n_fft, int hop,
                    int n_mels, int T, int w_lo, int w_hi, float clamp_min) {
    extern __shared__ float sm[];
    float* xs = sm;                 // [n_fft]
    float* tc = xs + n_fft;         // [n_fft]
    float* ts = tc + n_fft;         // [n_fft]
    float* pw = ts + n_fft;         // [n_fft / 2 + 1]
    const int t = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const int n_freq = n_fft / 2 + 1, mask = n_fft - 1;
    const float* x = wav + (size_t)b * N;
    for (int n = tid; n < n_fft; n += blockDim.x) {
        float v = 0.0f;
        if (n >= w_lo && n < w_hi) {
            int idx = t * hop + n - n_fft / 2;
            if (idx < 0) idx = -idx;                      // reflect padding (torch.stft centre=True, pad_mode="reflect")
            if (idx >= N) idx = 2 * (N - 1) - idx;
            v = x[idx] * window[n];
        }
        xs[n] = v;
        tc[n] = tw[2 * n];
        ts[n] = tw[2 * n + 1];
    }
    __syncthreads();
    for (int f = tid; f < n_freq; f += blockDim.x) {
        float re[4] = {0.f, 0.f, 0.f, 0.f}, im[4] = {0.f, 0.f, 0.f, 0.f};
        int k = (int)(((long long)f * w_lo) & mask);
        int n = w_lo;
        for (; n + 4 <= w_hi; n += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float xv = xs[n + u];
                re[u] = fmaf(xv, tc[k], re[u]);
                im[u] = fmaf(xv, ts[k], im[u]);
                k = (k + f) & mask;
            }
        }
        for (; n < w_hi; ++n) {
            re[0] = fmaf(xs[n], tc[k], re[0]);
            im[0] = fmaf(xs[n], ts[k], im[0]);
            k = (k + f) & mask;
        }
        const float r = (re[0] + re[1]) + (re[2] + re[3]), i = (im[0] + im[1]) + (im[2] + im[3]);
        pw[f] = r * r + i * i;
    }
    __syncthreads();
    for (int m = tid; m < n_mels; m += blockDim.x) {
        float acc = 0.0f;
        for (int f = 0; f < n_freq; ++f) acc = fmaf(pw[f], __ldg(fbank + (size_t)f * n_mels + m), acc);
        float v = logf(fmaxf(acc, clamp_min));
        if (norms) v /= __ldg(norms + m);
        mel[((size_t)b * n_mels + m) * T + t] = v;
    }
}
}  // namespace gv

extern "C" int genvc_mel_spectrogram(const float* wav, int B, int N, const float* window, const float* twiddle, const float* fbank,
                                     const float* norms, float* mel, int n_fft, int hop, int win_lo, int win_hi, int n_mels,
                                     float clamp_min, void* stream) {
    if (!wav || !window || !twiddle || !fbank || !mel || B <= 0 || N <= 0 || n_fft < 64 || n_fft > 4096 || (n_fft & (n_fft - 1)) ||
        hop <= 0 || n_mels <= 0 || win_lo < 0 || win_hi > n_fft || win_lo >= win_hi || N <= n_fft / 2)
        return GENVC_E_INVALID;  // (reflect padding needs more than n_fft / 2 samples, as torch.stft does)
    const int T = 1 + N / hop;
    const size_t smem = ((size_t)3 * n_fft + n_fft / 2 + 1) * sizeof(float);
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(gv::mel_frontend_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return GENVC_E_CUDA;
    gv::mel_frontend_kernel<<<dim3(T, B), 256, smem, (cudaStream_t)stream>>>(wav, N, window, twiddle, fbank, norms, mel, n_fft, hop, n_mels,
                                                                              T, win_lo, win_hi, clamp_min);
    return cudaGetLastError() == cudaSuccess ? GENVC_OK : GENVC_E_CUDA;
}
