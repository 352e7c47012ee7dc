// Per-op fp32 kernels (see ops.cuh).  Everything here is CUDA-core fp32 with fp32 accumulation:
// greedy token parity against the fp32 CPU reference needs logits good to ~1e-5, which
// rules TF32/bf16 out for this path.
#include "attn_decode.cuh"
#include "common.cuh"
#include <cstdlib>

#include "ops.cuh"
#include "sampling.cuh"

namespace gv {

#define GV_BUMP(n) \
    do {           \
        if (n) ++*(n); \
    } while (0)

// =============================================================================================
// GEMM: C = act(A W + bias) + residual      (A [M,K]; W [K,N] or [N,K])
// 64x64 / 32x64 / 16x64 output tiles, BK = 16, 256 threads, register-prefetch double buffering,
// optional deterministic split-K (partials to a workspace, reduced in fixed order).
// Requirements: K % 4 == 0, lda % 4 == 0, ldw % 4 == 0 and 16-byte aligned bases (vector loads).
// =============================================================================================
template <int BM>
__global__ void __launch_bounds__(256) gemm_kernel(GemmArgs a, float* __restrict__ ws, int k_chunk, const int* skip) {
    if (skip && *skip) return;
    constexpr int BN = 64, BK = 16, TM = BM / 16, TN = 4;
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kb = blockIdx.z * k_chunk;
    const int ke = min(a.K, kb + k_chunk);
    const int ty = tid / 16, tx = tid % 16;

    // loader mapping: A tile BM x 16 -> BM*4 float4 ; W tile 16 x 64 (KN) or 64 x 16 (NK) -> 256 float4
    const int a_row = tid / 4, a_kq = (tid % 4) * 4;
    const bool a_ld = tid < BM * 4;
    const int wk_k = tid / 16, wk_n = (tid % 16) * 4;   // KN
    const int wn_n = tid / 4, wn_kq = (tid % 4) * 4;    // NK

    float4 ra = make_float4(0, 0, 0, 0), rw = make_float4(0, 0, 0, 0);
    auto gload = [&](int k0) {
        ra = make_float4(0, 0, 0, 0);
        rw = make_float4(0, 0, 0, 0);
        if (a_ld) {
            int m = m0 + a_row, k = k0 + a_kq;
            if (m < a.M && k < ke) ra = *reinterpret_cast<const float4*>(a.A + (size_t)m * a.lda + k);
        }
        if (a.w_nk) {
            int n = n0 + wn_n, k = k0 + wn_kq;
            if (n < a.N && k < ke) rw = *reinterpret_cast<const float4*>(a.W + (size_t)n * a.ldw + k);
        } else {
            int k = k0 + wk_k, n = n0 + wk_n;
            if (k < ke) {
                if (n + 3 < a.N) {
                    rw = *reinterpret_cast<const float4*>(a.W + (size_t)k * a.ldw + n);
                } else {
                    const float* p = a.W + (size_t)k * a.ldw;
                    if (n < a.N) rw.x = p[n];
                    if (n + 1 < a.N) rw.y = p[n + 1];
                    if (n + 2 < a.N) rw.z = p[n + 2];
                }
            }
        }
    };
    auto sstore = [&](int buf) {
        if (a_ld) {
            As[buf][a_kq + 0][a_row] = ra.x;
            As[buf][a_kq + 1][a_row] = ra.y;
            As[buf][a_kq + 2][a_row] = ra.z;
            As[buf][a_kq + 3][a_row] = ra.w;
        }
        if (a.w_nk) {
            Bs[buf][wn_kq + 0][wn_n] = rw.x;
            Bs[buf][wn_kq + 1][wn_n] = rw.y;
            Bs[buf][wn_kq + 2][wn_n] = rw.z;
            Bs[buf][wn_kq + 3][wn_n] = rw.w;
        } else {
            *reinterpret_cast<float4*>(&Bs[buf][wk_k][wk_n]) = rw;
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

    int buf = 0;
    gload(kb);
    sstore(0);
    __syncthreads();
    for (int k0 = kb; k0 < ke; k0 += BK) {
        const bool more = (k0 + BK) < ke;
        if (more) gload(k0 + BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float av[TM], bv[TN];
            if constexpr (TM == 4) {
                float4 t = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
                av[0] = t.x; av[1] = t.y; av[2] = t.z; av[3] = t.w;
            } else if constexpr (TM == 2) {
                float2 t = *reinterpret_cast<const float2*>(&As[buf][k][ty * 2]);
                av[0] = t.x; av[1] = t.y;
            } else {
                av[0] = As[buf][k][ty];
            }
            float4 t = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            bv[0] = t.x; bv[1] = t.y; bv[2] = t.z; bv[3] = t.w;
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (more) {
            sstore(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }

    const bool split = gridDim.z > 1;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + ty * TM + i;
        if (m >= a.M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + tx * TN + j;
            if (n >= a.N) continue;
            float v = acc[i][j];
            if (split) {
                ws[((size_t)blockIdx.z * a.M + m) * a.N + n] = v;
            } else {
                if (a.bias) v += a.bias[n];
                if (a.act == ACT_GELU_NEW) v = gelu_new(v);
                if (a.residual) v += a.residual[(size_t)m * a.ldr + n];
                a.C[(size_t)m * a.ldc + n] = v;
            }
        }
    }
}

__global__ void splitk_epilogue_kernel(GemmArgs a, const float* __restrict__ ws, int splits, const int* skip) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // a following tcgen05 GEMM may begin its prologue / weight prefetch
    if (skip && *skip) return;
    const size_t total = (size_t)a.M * a.N;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int m = (int)(i / a.N), n = (int)(i % a.N);
        float v = 0.0f;
        for (int z = 0; z < splits; ++z) v += ws[(size_t)z * total + i];
        if (a.bias) v += a.bias[n];
        if (a.act == ACT_GELU_NEW) v = gelu_new(v);
        if (a.residual) v += a.residual[(size_t)m * a.ldr + n];
        a.C[(size_t)m * a.ldc + n] = v;
        if (a.kv_k != nullptr) {  // fused K/V cache append (what kv_scatter_kernel does from the finished rows)
            const int Dkv = a.N / 3;
            if (n >= Dkv) {
                const int c = (n - Dkv) % Dkv, hd = Dkv / a.kv_H;
                const int b = m / a.kv_rows, r = m % a.kv_rows;
                float* dst = (n < 2 * Dkv) ? a.kv_k : a.kv_v;
                dst[(size_t)b * a.kv_bs + ((size_t)(c / hd) * a.kv_S_max + a.kv_pos0 + r) * hd + (c % hd)] = v;
            }
        }
    }
}

// split-K reduction + bias + residual of one row per block (thread = 4 columns, all split loads in flight), then
// LayerNorm of that row (two-pass, like layernorm_kernel).  N = D <= 1024.
__global__ void __launch_bounds__(256) splitk_ln_epilogue_kernel(GemmArgs a, const float* __restrict__ ws, int splits, const int* skip) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // a following tcgen05 GEMM may begin its prologue / weight prefetch
    if (skip && *skip) return;
    __shared__ float red[2][8];
    const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = tid * 4;
    const bool valid = n < a.N;
    const size_t total = (size_t)a.M * a.N;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) {
        const float* src = ws + (size_t)row * a.N + n;
        int z = 0;
        for (; z + 4 <= splits; z += 4) {  // four independent loads in flight, summed in split order
            const float4 t0 = *reinterpret_cast<const float4*>(src + (size_t)(z + 0) * total);
            const float4 t1 = *reinterpret_cast<const float4*>(src + (size_t)(z + 1) * total);
            const float4 t2 = *reinterpret_cast<const float4*>(src + (size_t)(z + 2) * total);
            const float4 t3 = *reinterpret_cast<const float4*>(src + (size_t)(z + 3) * total);
            acc.x += t0.x; acc.y += t0.y; acc.z += t0.z; acc.w += t0.w;
            acc.x += t1.x; acc.y += t1.y; acc.z += t1.z; acc.w += t1.w;
            acc.x += t2.x; acc.y += t2.y; acc.z += t2.z; acc.w += t2.w;
            acc.x += t3.x; acc.y += t3.y; acc.z += t3.z; acc.w += t3.w;
        }
        for (; z < splits; ++z) {
            const float4 t = *reinterpret_cast<const float4*>(src + (size_t)z * total);
            acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        }
        if (a.bias) {
            const float4 b = *reinterpret_cast<const float4*>(a.bias + n);
            acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
        }
        if (a.act == ACT_GELU_NEW) {
            acc.x = gelu_new(acc.x); acc.y = gelu_new(acc.y); acc.z = gelu_new(acc.z); acc.w = gelu_new(acc.w);
        }
        if (a.residual) {
            const float4 r = *reinterpret_cast<const float4*>(a.residual + (size_t)row * a.ldr + n);
            acc.x += r.x; acc.y += r.y; acc.z += r.z; acc.w += r.w;
        }
        *reinterpret_cast<float4*>(a.C + (size_t)row * a.ldc + n) = acc;
    }
    float s = valid ? ((acc.x + acc.y) + (acc.z + acc.w)) : 0.0f;
    s = warp_sum(s);
    if (lane == 0) red[0][warp] = s;
    __syncthreads();
    float tot = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += red[0][w];
    const float mean = tot / (float)a.N;
    const float d0 = acc.x - mean, d1 = acc.y - mean, d2 = acc.z - mean, d3 = acc.w - mean;
    float q = valid ? (fmaf(d0, d0, d1 * d1) + fmaf(d2, d2, d3 * d3)) : 0.0f;
    q = warp_sum(q);
    if (lane == 0) red[1][warp] = q;
    __syncthreads();
    float qt = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) qt += red[1][w];
    const float rstd = 1.0f / sqrtf(qt / (float)a.N + 1e-5f);
    if (valid) {
        const float4 ww = *reinterpret_cast<const float4*>(a.ln_w + n);
        const float4 bb = *reinterpret_cast<const float4*>(a.ln_b + n);
        *reinterpret_cast<float4*>(a.ln_out + (size_t)row * a.ld_ln + n) =
            make_float4(d0 * rstd * ww.x + bb.x, d1 * rstd * ww.y + bb.y, d2 * rstd * ww.z + bb.z, d3 * rstd * ww.w + bb.w);
    }
}
static bool launch_splitk_ln(const GemmArgs& a, const float* ws, int splits, const int* skip, cudaStream_t st) {
    if ((a.N & 3) || a.N > 1024 || (a.ldc & 3) || (a.ldr & 3) || (a.ld_ln & 3)) return false;
    splitk_ln_epilogue_kernel<<<a.M, 256, 0, st>>>(a, ws, splits, skip);
    return true;
}

// GENVC_TC=0 in the environment keeps every GEMM on the fp32 CUDA-core kernel (debug / A-B comparison)
static bool use_tensor_cores() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("GENVC_TC");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

// forward declarations of the follow-up launchers used below
cudaError_t launch_kv_scatter(const float* qkv, int B, int M, int D, int H, float* kcache, float* vcache, long batch_stride,
                              int S_max, int pos0, const int* skip, cudaStream_t st, unsigned long long* nlaunch);

// follow-ups that could not be fused into a split-K epilogue: separate launches from the finished C
static cudaError_t gemm_followups(const GemmArgs& a, bool ln_done, bool kv_done, const int* skip, cudaStream_t st,
                                  unsigned long long* nlaunch) {
    if (a.ln_w && !ln_done) {
        cudaError_t e = launch_layernorm(a.C, a.ldc, 0, a.ln_out, a.ld_ln, 0, a.M, a.M, a.N, a.ln_w, a.ln_b, nullptr, nullptr, skip, st,
                                         nlaunch);
        if (e != cudaSuccess) return e;
    }
    if (a.kv_k && !kv_done) {
        const int Dkv = a.N / 3;
        if (a.ldc != a.N) return cudaErrorInvalidValue;
        cudaError_t e = launch_kv_scatter(a.C, a.M / a.kv_rows, a.kv_rows, Dkv, a.kv_H, a.kv_k, a.kv_v, a.kv_bs, a.kv_S_max, a.kv_pos0, skip,
                                          st, nlaunch);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// split-K reduction (+ fused LayerNorm / KV append when requested)
static cudaError_t run_splitk_epilogue(const GemmArgs& a, const float* ws, int splits, const int* skip, cudaStream_t st,
                                       unsigned long long* nlaunch) {
    bool ln_done = false, kv_done = false;
    if (a.ln_w && !a.kv_k && launch_splitk_ln(a, ws, splits, skip, st)) {
        ln_done = true;
    } else {
        size_t total = (size_t)a.M * a.N;
        int blocks = (int)min((size_t)1184, (total + 255) / 256);
        splitk_epilogue_kernel<<<blocks, 256, 0, st>>>(a, ws, splits, skip);
        kv_done = a.kv_k != nullptr;
    }
    GV_BUMP(nlaunch);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return gemm_followups(a, ln_done, kv_done, skip, st, nlaunch);
}

cudaError_t launch_gemm(const GemmArgs& a, float* ws, size_t ws_floats, const int* skip, cudaStream_t st,
                        unsigned long long* nlaunch) {
    if (a.M <= 0 || a.N <= 0 || a.K <= 0) return cudaErrorInvalidValue;
    if ((a.K & 3) || (a.lda & 3) || (a.ldw & 3)) return cudaErrorInvalidValue;
    if (use_tensor_cores() && gemm_tc_eligible(a)) {
        // dense contraction with enough rows: tcgen05 3xTF32 kernel (gemm_tc.cu)
        int splits = 1;
        cudaError_t e = launch_gemm_tc(a, ws, ws_floats, skip, st, &splits);
        if (e != cudaSuccess) return e;
        GV_BUMP(nlaunch);
        if (splits > 1) return run_splitk_epilogue(a, ws, splits, skip, st, nlaunch);
        return gemm_followups(a, false, false, skip, st, nlaunch);
    }
    const int BM = a.M <= 16 ? 16 : (a.M <= 32 ? 32 : 64);
    dim3 grid((a.N + 63) / 64, (a.M + BM - 1) / BM, 1);
    // deterministic split-K when the tile grid cannot fill the machine
    int splits = 1;
    const int tiles = grid.x * grid.y;
    if (tiles < 120 && a.K >= 512 && ws) {
        splits = min(min(8, 296 / tiles), a.K / 256);
        while (splits > 1 && (size_t)splits * a.M * a.N > ws_floats) --splits;
        if (splits < 1) splits = 1;
    }
    int k_chunk = a.K;
    if (splits > 1) {
        k_chunk = ((a.K + splits - 1) / splits + 15) / 16 * 16;
        splits = (a.K + k_chunk - 1) / k_chunk;
    }
    grid.z = splits;
    if (BM == 16) gemm_kernel<16><<<grid, 256, 0, st>>>(a, ws, k_chunk, skip);
    else if (BM == 32) gemm_kernel<32><<<grid, 256, 0, st>>>(a, ws, k_chunk, skip);
    else gemm_kernel<64><<<grid, 256, 0, st>>>(a, ws, k_chunk, skip);
    GV_BUMP(nlaunch);
    if (splits > 1) return run_splitk_epilogue(a, ws, splits, skip, st, nlaunch);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return gemm_followups(a, false, false, skip, st, nlaunch);
}

// =============================================================================================
// LayerNorm over rows (one warp per row, row held in registers, two-pass variance like ATen);
// optional second LayerNorm chained on the result (ln_f -> final_norm, layers/gpt_inference.py:18).
// =============================================================================================
template <int NCH>  // float4 chunks per lane: D = NCH * 128
__global__ void __launch_bounds__(128) layernorm_kernel(const float* __restrict__ X, long x_stride, long x_gs,
                                                        float* __restrict__ Y, long y_stride, long y_gs, int rows,
                                                        int rpg, const float* __restrict__ w1,
                                                        const float* __restrict__ b1, const float* __restrict__ w2,
                                                        const float* __restrict__ b2, const int* skip) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // a following tcgen05 GEMM may begin its prologue / weight prefetch
    if (skip && *skip) return;
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    constexpr int D = NCH * 128;
    float v[NCH * 4];
    const int grp = row / rpg, rin = row % rpg;
    const float* x = X + (size_t)grp * x_gs + (size_t)rin * x_stride;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        float4 t = *reinterpret_cast<const float4*>(x + c * 128 + lane * 4);
        v[c * 4] = t.x; v[c * 4 + 1] = t.y; v[c * 4 + 2] = t.z; v[c * 4 + 3] = t.w;
    }
    auto norm = [&](const float* w, const float* b) {
        float s = 0.0f;
#pragma unroll
        for (int i = 0; i < NCH * 4; ++i) s += v[i];
        const float mean = warp_sum(s) * (1.0f / D);
        float q = 0.0f;
#pragma unroll
        for (int i = 0; i < NCH * 4; ++i) {
            float d = v[i] - mean;
            q = fmaf(d, d, q);
        }
        const float var = warp_sum(q) * (1.0f / D);
        const float rstd = 1.0f / sqrtf(var + 1e-5f);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            float4 ww = *reinterpret_cast<const float4*>(w + c * 128 + lane * 4);
            float4 bb = *reinterpret_cast<const float4*>(b + c * 128 + lane * 4);
            v[c * 4 + 0] = (v[c * 4 + 0] - mean) * rstd * ww.x + bb.x;
            v[c * 4 + 1] = (v[c * 4 + 1] - mean) * rstd * ww.y + bb.y;
            v[c * 4 + 2] = (v[c * 4 + 2] - mean) * rstd * ww.z + bb.z;
            v[c * 4 + 3] = (v[c * 4 + 3] - mean) * rstd * ww.w + bb.w;
        }
    };
    norm(w1, b1);
    if (w2) norm(w2, b2);
    float* y = Y + (size_t)grp * y_gs + (size_t)rin * y_stride;
#pragma unroll
    for (int c = 0; c < NCH; ++c)
        *reinterpret_cast<float4*>(y + c * 128 + lane * 4) = make_float4(v[c * 4], v[c * 4 + 1], v[c * 4 + 2], v[c * 4 + 3]);
}

cudaError_t launch_layernorm(const float* X, long x_stride, long x_gs, float* Y, long y_stride, long y_gs, int rows, int rpg,
                             int D, const float* w1, const float* b1, const float* w2, const float* b2, const int* skip,
                             cudaStream_t st, unsigned long long* nlaunch) {
    if (D % 128 || D > 1024 || rows <= 0 || rpg <= 0) return cudaErrorInvalidValue;
    dim3 grid((rows + 3) / 4);
    switch (D / 128) {
#define GV_LN_CASE(n) \
    case n: layernorm_kernel<n><<<grid, 128, 0, st>>>(X, x_stride, x_gs, Y, y_stride, y_gs, rows, rpg, w1, b1, w2, b2, skip); break;
        GV_LN_CASE(1) GV_LN_CASE(2) GV_LN_CASE(3) GV_LN_CASE(4) GV_LN_CASE(5) GV_LN_CASE(6) GV_LN_CASE(7) GV_LN_CASE(8)
#undef GV_LN_CASE
    }
    GV_BUMP(nlaunch);
    return cudaGetLastError();
}

// RMSNorm of the perceiver (layers/perceiver_encoder.py:177-179): F.normalize(x, dim=-1) * sqrt(D) * gamma
__global__ void rmsnorm_kernel(const float* __restrict__ X, float* __restrict__ Y, int rows, int D,
                               const float* __restrict__ gamma) {
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const float* x = X + (size_t)row * D;
    float s = 0.0f;
    for (int i = lane; i < D; i += 32) s = fmaf(x[i], x[i], s);
    s = warp_sum(s);
    const float denom = fmaxf(sqrtf(s), 1e-12f);
    const float scale = sqrtf((float)D);
    for (int i = lane; i < D; i += 32) Y[(size_t)row * D + i] = x[i] / denom * scale * gamma[i];
}
cudaError_t launch_rmsnorm(const float* X, float* Y, int rows, int D, const float* gamma, cudaStream_t st,
                           unsigned long long* nlaunch) {
    rmsnorm_kernel<<<(rows + 3) / 4, 128, 0, st>>>(X, Y, rows, D, gamma);
    GV_BUMP(nlaunch);
    return cudaGetLastError();
}

// =============================================================================================
// Multi-query attention (prefill, latent pass, perceiver cross-attention).  One CTA = 8 query rows
// of one (batch, head); key/value tiles of 32 rows staged in shared memory and shared by the 8
// queries; online softmax per warp.  HD in {32, 64, 128, 256}.
// =============================================================================================
template <int HD>
__global__ void __launch_bounds__(256) attention_kernel(AttnArgs a, const int* skip) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // a following tcgen05 GEMM may begin its prologue / weight prefetch
    if (skip && *skip) return;
    constexpr int QT = 8, KT = 32, LD = HD + 4, DPL = HD / 32;
    extern __shared__ __align__(16) float sm[];
    float* Qs = sm;                 // [QT][HD]
    float* Ks = Qs + QT * HD;       // [KT][LD]
    float* Vs = Ks + KT * LD;       // [KT][LD]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.z, h = blockIdx.y, i0 = blockIdx.x * QT;
    const float* Qb = a.Q + (size_t)b * a.q_bs + (size_t)h * a.q_hs;
    const float* Kb = a.K + (size_t)b * a.k_bs + (size_t)h * a.k_hs;
    const float* Vb = a.V + (size_t)b * a.v_bs + (size_t)h * a.v_hs;

    for (int idx = tid; idx < QT * HD / 4; idx += 256) {
        int r = idx / (HD / 4), c = (idx % (HD / 4)) * 4;
        float4 t = make_float4(0, 0, 0, 0);
        if (i0 + r < a.M) t = *reinterpret_cast<const float4*>(Qb + (size_t)(i0 + r) * a.q_rs + c);
        *reinterpret_cast<float4*>(Qs + r * HD + c) = t;
    }
    const int i = i0 + warp;               // this warp's query row
    const bool active = i < a.M;
    const int my_keys = a.causal ? min(a.n_keys, a.pos0 + i + 1) : a.n_keys;  // keys [0, my_keys)
    const int last_row = min(a.M, i0 + QT) - 1;
    const int cta_keys = a.causal ? min(a.n_keys, a.pos0 + last_row + 1) : a.n_keys;

    float m = -INFINITY, l = 0.0f, o[DPL];
#pragma unroll
    for (int d = 0; d < DPL; ++d) o[d] = 0.0f;
    const float* qrow = Qs + warp * HD;

    for (int j0 = 0; j0 < cta_keys; j0 += KT) {
        __syncthreads();  // previous tile fully consumed (also covers the Qs fill on the first pass)
        for (int idx = tid; idx < KT * HD / 4; idx += 256) {
            int r = idx / (HD / 4), c = (idx % (HD / 4)) * 4;
            float4 kk = make_float4(0, 0, 0, 0), vv = make_float4(0, 0, 0, 0);
            if (j0 + r < cta_keys) {
                kk = *reinterpret_cast<const float4*>(Kb + (size_t)(j0 + r) * a.k_rs + c);
                vv = *reinterpret_cast<const float4*>(Vb + (size_t)(j0 + r) * a.v_rs + c);
            }
            *reinterpret_cast<float4*>(Ks + r * LD + c) = kk;
            *reinterpret_cast<float4*>(Vs + r * LD + c) = vv;
        }
        __syncthreads();
        if (!active || j0 >= my_keys) continue;
        // lane j scores key j0 + lane
        float s = 0.0f;
        const float* krow = Ks + lane * LD;
#pragma unroll 8
        for (int c = 0; c < HD; c += 4) {
            float4 qq = *reinterpret_cast<const float4*>(qrow + c);
            float4 kk = *reinterpret_cast<const float4*>(krow + c);
            s = fmaf(qq.x, kk.x, s);
            s = fmaf(qq.y, kk.y, s);
            s = fmaf(qq.z, kk.z, s);
            s = fmaf(qq.w, kk.w, s);
        }
        s = (j0 + lane < my_keys) ? s * a.scale : -INFINITY;
        const float mnew = fmaxf(m, warp_max(s));
        const float corr = expf(m - mnew);
        const float p = expf(s - mnew);
        l = l * corr + warp_sum(p);
#pragma unroll
        for (int d = 0; d < DPL; ++d) o[d] *= corr;
        const int nk = min(KT, my_keys - j0);
        for (int j = 0; j < nk; ++j) {
            const float pj = __shfl_sync(0xffffffffu, p, j);
#pragma unroll
            for (int d = 0; d < DPL; ++d) o[d] = fmaf(pj, Vs[j * LD + lane + 32 * d], o[d]);
        }
        m = mnew;
    }
    if (active) {
        float* orow = a.O + (size_t)b * a.o_bs + (size_t)i * a.o_rs + (size_t)h * a.o_hs;
        const float inv = 1.0f / l;
#pragma unroll
        for (int d = 0; d < DPL; ++d) orow[lane + 32 * d] = o[d] * inv;
    }
}

// Few query rows (the prefill of a 1 s segment is 48 rows x 4 heads): one CTA per (query row, head, batch element) instead of
// one per 8 rows -- 192 CTAs instead of 24 -- and no shared-memory staging: warp w takes keys w, w + 8, ... straight from
// L2 (a K / V row is one coalesced 128-byte-per-8-lanes read), lanes split the head dimension, one shuffle tree per key,
// online softmax per warp; the eight per-warp (max, sum, o[hd]) meet in shared memory behind one barrier and are merged in
// warp order by thread d < hd.  16 us -> ~4 us for the 48-row prefill; K / V are re-read per query row, so the tiled kernel
// above keeps the long-M cases (latent pass, perceiver).
template <int HD>
__global__ void __launch_bounds__(256) attention_row_kernel(AttnArgs a, const int* skip) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (skip && *skip) return;
    constexpr int DPL = HD / 32;  // consecutive dims per lane
    __shared__ float s_m[8], s_l[8];
    __shared__ __align__(16) float s_o[8][HD];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int i = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const float* Qb = a.Q + (size_t)b * a.q_bs + (size_t)h * a.q_hs + (size_t)i * a.q_rs + lane * DPL;
    const float* Kb = a.K + (size_t)b * a.k_bs + (size_t)h * a.k_hs + lane * DPL;
    const float* Vb = a.V + (size_t)b * a.v_bs + (size_t)h * a.v_hs + lane * DPL;
    const int my_keys = a.causal ? min(a.n_keys, a.pos0 + i + 1) : a.n_keys;
    auto load = [&](const float* p, float* r) {
        if constexpr (DPL % 4 == 0) {
#pragma unroll
            for (int c = 0; c < DPL; c += 4) {
                const float4 t = *reinterpret_cast<const float4*>(p + c);
                r[c] = t.x; r[c + 1] = t.y; r[c + 2] = t.z; r[c + 3] = t.w;
            }
        } else if constexpr (DPL == 2) {
            const float2 t = *reinterpret_cast<const float2*>(p);
            r[0] = t.x; r[1] = t.y;
        } else {
            r[0] = *p;
        }
    };
    float q[DPL], o[DPL];
    load(Qb, q);
#pragma unroll
    for (int d = 0; d < DPL; ++d) o[d] = 0.0f;
    float m = -INFINITY, l = 0.0f;
    for (int j = warp; j < my_keys; j += 16) {  // two keys of this warp in flight
        const bool two = j + 8 < my_keys;
        float k0[DPL], v0[DPL], k1[DPL], v1[DPL];
        load(Kb + (size_t)j * a.k_rs, k0);
        load(Vb + (size_t)j * a.v_rs, v0);
        if (two) {
            load(Kb + (size_t)(j + 8) * a.k_rs, k1);
            load(Vb + (size_t)(j + 8) * a.v_rs, v1);
        }
        float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
        for (int d = 0; d < DPL; ++d) {
            s0 = fmaf(q[d], k0[d], s0);
            if (two) s1 = fmaf(q[d], k1[d], s1);
        }
#pragma unroll
        for (int x = 16; x > 0; x >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, x);
            s1 += __shfl_xor_sync(0xffffffffu, s1, x);
        }
        s0 *= a.scale;
        s1 = two ? s1 * a.scale : -INFINITY;
        const float mn = fmaxf(m, fmaxf(s0, s1));
        const float c = expf(m - mn), p0 = expf(s0 - mn), p1 = expf(s1 - mn);  // exp(-inf) = 0
        l = l * c + p0 + p1;
#pragma unroll
        for (int d = 0; d < DPL; ++d) o[d] = fmaf(p1, two ? v1[d] : 0.0f, fmaf(p0, v0[d], o[d] * c));
        m = mn;
    }
    if (lane == 0) {
        s_m[warp] = m;
        s_l[warp] = l;
    }
#pragma unroll
    for (int d = 0; d < DPL; ++d) s_o[warp][lane * DPL + d] = o[d];
    __syncthreads();
    if (tid < HD) {
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < 8; ++w) M = fmaxf(M, s_m[w]);
        float L = 0.0f, acc = 0.0f;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const float cw = expf(s_m[w] - M);  // warps without keys: exp(-inf) = 0
            L = fmaf(s_l[w], cw, L);
            acc = fmaf(s_o[w][tid], cw, acc);
        }
        a.O[(size_t)b * a.o_bs + (size_t)i * a.o_rs + (size_t)h * a.o_hs + tid] = acc / L;
    }
}

cudaError_t launch_attention(const AttnArgs& a, const int* skip, cudaStream_t st, unsigned long long* nlaunch) {
    static const bool rows_ok = [] { const char* e = getenv("GENVC_ATT_ROWS"); return !(e && e[0] == '0'); }();
    if (rows_ok && a.M <= 64 && a.n_keys <= 4096) {
        const dim3 g(a.M, a.H, a.B);
        switch (a.hd) {
            case 32: attention_row_kernel<32><<<g, 256, 0, st>>>(a, skip); break;
            case 64: attention_row_kernel<64><<<g, 256, 0, st>>>(a, skip); break;
            case 128: attention_row_kernel<128><<<g, 256, 0, st>>>(a, skip); break;
            case 256: attention_row_kernel<256><<<g, 256, 0, st>>>(a, skip); break;
            default: return cudaErrorInvalidValue;
        }
        GV_BUMP(nlaunch);
        return cudaGetLastError();
    }
    dim3 grid((a.M + 7) / 8, a.H, a.B);
    auto smem = [](int hd) { return (size_t)(8 * hd + 2 * 32 * (hd + 4)) * sizeof(float); };
    cudaError_t e = cudaSuccess;
    switch (a.hd) {
#define GV_ATT_CASE(n)                                                                                         \
    case n:                                                                                                    \
        e = cudaFuncSetAttribute(attention_kernel<n>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem(n)); \
        if (e != cudaSuccess) return e;                                                                        \
        attention_kernel<n><<<grid, 256, smem(n), st>>>(a, skip);                                              \
        break;
        GV_ATT_CASE(32) GV_ATT_CASE(64) GV_ATT_CASE(128) GV_ATT_CASE(256)
#undef GV_ATT_CASE
        default: return cudaErrorInvalidValue;
    }
    GV_BUMP(nlaunch);
    return cudaGetLastError();
}

// qkv rows [B*M, 3D] -> K/V caches [b][h][pos][hd] at positions pos0 .. pos0+M-1
__global__ void kv_scatter_kernel(const float* __restrict__ qkv, int B, int M, int D, int H, float* __restrict__ kc,
                                  float* __restrict__ vc, long batch_stride, int S_max, int pos0, const int* skip) {
    if (skip && *skip) return;
    const int hd = D / H;
    const size_t total = (size_t)B * M * (D / 4);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % (D / 4)) * 4;
        const size_t row = i / (D / 4);
        const int b = (int)(row / M), r = (int)(row % M);
        const int h = c / hd, d = c % hd;
        const float* src = qkv + row * 3 * (size_t)D;
        const size_t dst = (size_t)b * batch_stride + ((size_t)h * S_max + pos0 + r) * hd + d;
        *reinterpret_cast<float4*>(kc + dst) = *reinterpret_cast<const float4*>(src + D + c);
        *reinterpret_cast<float4*>(vc + dst) = *reinterpret_cast<const float4*>(src + 2 * D + c);
    }
}
cudaError_t launch_kv_scatter(const float* qkv, int B, int M, int D, int H, float* kcache, float* vcache, long batch_stride,
                              int S_max, int pos0, const int* skip, cudaStream_t st, unsigned long long* nlaunch) {
    size_t total = (size_t)B * M * (D / 4);
    int blocks = (int)min((size_t)1184, (total + 255) / 256);
    kv_scatter_kernel<<<blocks, 256, 0, st>>>(qkv, B, M, D, H, kcache, vcache, batch_stride, S_max, pos0, skip);
    GV_BUMP(nlaunch);
    return cudaGetLastError();
}

// =============================================================================================
// embeddings (layers/gpt.py:572-592; layers/gpt_inference.py:81-96)
// =============================================================================================
__global__ void embed_prefix_kernel(const float* __restrict__ cond, const long long* __restrict__ text_ids, int B, int T,
                                    int n_lat, int D, const float* __restrict__ text_emb,
                                    const float* __restrict__ text_pos, int start_text, int stop_text,
                                    float* __restrict__ out, long out_bs, int n_vocab, int* bad) {
    const int P = n_lat + T + 2;
    const int row = blockIdx.x;  // b * P + r
    const int b = row / P, r = row % P;
    float* o = out + (size_t)b * out_bs + (size_t)r * D;
    if (r < n_lat) {
        const float* c = cond + ((size_t)b * n_lat + r) * D;
        for (int i = threadIdx.x; i < D; i += blockDim.x) o[i] = c[i];
    } else {
        const int j = r - n_lat;  // position in [start, codes..., stop]
        long long id = (j == 0) ? start_text : (j == T + 1 ? stop_text : text_ids[(size_t)b * T + (j - 1)]);
        if (id < 0 || id >= n_vocab) {  // the C ABI never reads outside the table: clamp, and flag it for the next status read
            id = id < 0 ? 0 : n_vocab - 1;
            if (bad != nullptr && threadIdx.x == 0) atomicOr(bad, 1);
        }
        const float* e = text_emb + (size_t)id * D;
        const float* p = text_pos + (size_t)j * D;
        for (int i = threadIdx.x; i < D; i += blockDim.x) o[i] = e[i] + p[i];
    }
}
cudaError_t launch_embed_prefix(const float* cond, const long long* text_ids, int B, int T, int n_lat, int D,
                                const float* text_emb, const float* text_pos, int start_text, int stop_text, float* out,
                                long out_bs, int n_vocab, int* bad, cudaStream_t st, unsigned long long* nlaunch) {
    embed_prefix_kernel<<<B * (n_lat + T + 2), 128, 0, st>>>(cond, text_ids, B, T, n_lat, D, text_emb, text_pos,
                                                             start_text, stop_text, out, out_bs, n_vocab, bad);
    GV_BUMP(nlaunch);
    return cudaGetLastError();
}

__global__ void embed_mel_rows_kernel(const long long* __restrict__ codes, int B, int R, int M, int first_tok, int pad_tok,
                                      int pos0, int D, const float* __restrict__ mel_emb,
                                      const float* __restrict__ mel_pos, float* __restrict__ out, long out_bs, int n_vocab,
                                      int* bad) {
    const int b = blockIdx.x / R, r = blockIdx.x % R;
    long long id = (r == 0) ? first_tok : ((codes && r - 1 < M) ? codes[(size_t)b * M + (r - 1)] : pad_tok);
    if (id < 0 || id >= n_vocab) {
        id = id < 0 ? 0 : n_vocab - 1;
        if (bad != nullptr && threadIdx.x == 0) atomicOr(bad, 1);
    }
    const float* e = mel_emb + (size_t)id * D;
    const float* p = mel_pos + (size_t)(pos0 + r) * D;
    float* o = out + (size_t)b * out_bs + (size_t)r * D;
    for (int i = threadIdx.x; i < D; i += blockDim.x) o[i] = e[i] + p[i];
}
cudaError_t launch_embed_mel_rows(const long long* codes, int B, int R, int M, int first_tok, int pad_tok, int pos0, int D,
                                  const float* mel_emb, const float* mel_pos, float* out, long out_bs, int n_vocab, int* bad,
                                  cudaStream_t st, unsigned long long* nlaunch) {
    embed_mel_rows_kernel<<<B * R, 128, 0, st>>>(codes, B, R, M, first_tok, pad_tok, pos0, D, mel_emb, mel_pos, out, out_bs,
                                                 n_vocab, bad);
    GV_BUMP(nlaunch);
    return cudaGetLastError();
}

__global__ void embed_last_token_kernel(const GenState* __restrict__ stt, int pos, int D,
                                        const float* __restrict__ mel_emb, const float* __restrict__ mel_pos,
                                        float* __restrict__ out) {
    if (stt->done) return;
    const int b = blockIdx.x;
    const long long id = stt->last_tok[b];
    const float* e = mel_emb + (size_t)id * D;
    const float* p = mel_pos + (size_t)pos * D;
    for (int i = threadIdx.x; i < D; i += blockDim.x) out[(size_t)b * D + i] = e[i] + p[i];
}
cudaError_t launch_embed_last_token(const GenState* stt, int B, int pos, int D, const float* mel_emb, const float* mel_pos,
                                    float* out, cudaStream_t st, unsigned long long* nlaunch) {
    embed_last_token_kernel<<<B, 128, 0, st>>>(stt, pos, D, mel_emb, mel_pos, out);
    GV_BUMP(nlaunch);
    return cudaGetLastError();
}

__global__ void copy_rows_kernel(const float* __restrict__ src, long src_bs, float* __restrict__ dst, long dst_bs, int B,
                                 long row_floats) {
    const size_t total = (size_t)B * row_floats;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / row_floats, r = i % row_floats;
        dst[b * dst_bs + r] = src[b * src_bs + r];
    }
}
cudaError_t launch_copy_rows(const float* src, long src_bs, float* dst, long dst_bs, int B, long row_floats, cudaStream_t st,
                             unsigned long long* nlaunch) {
    size_t total = (size_t)B * row_floats;
    int blocks = (int)min((size_t)1184, (total + 255) / 256);
    copy_rows_kernel<<<blocks, 256, 0, st>>>(src, src_bs, dst, dst_bs, B, row_floats);
    GV_BUMP(nlaunch);
    return cudaGetLastError();
}

// mel [B,C,S] -> [B,S,C]  (cond_input.permute(0,2,1), layers/gpt.py:369)
// mel [B, C, S] -> out [B, S, Cp] (Cp >= C: row pitch padded with zeros to the k granularity of the tensor-core GEMM)
__global__ void transpose_mel_kernel(const float* __restrict__ mel, int C, int Cp, int S, float* __restrict__ out) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int s0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const float* src = mel + (size_t)b * C * S;
    float* dst = out + (size_t)b * Cp * S;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int c = c0 + r, s = s0 + threadIdx.x;
        tile[r][threadIdx.x] = (c < C && s < S) ? src[(size_t)c * S + s] : 0.0f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int s = s0 + r, c = c0 + threadIdx.x;
        if (s < S && c < Cp) dst[(size_t)s * Cp + c] = tile[threadIdx.x][r];  // (columns C .. Cp-1 were loaded as zeros)
    }
}
cudaError_t launch_transpose_mel(const float* mel, int B, int C, int Cp, int S, float* out, cudaStream_t st,
                                 unsigned long long* nlaunch) {
    dim3 grid((S + 31) / 32, (Cp + 31) / 32, B);
    transpose_mel_kernel<<<grid, dim3(32, 8), 0, st>>>(mel, C, Cp, S, out);
    GV_BUMP(nlaunch);
    return cudaGetLastError();
}

// GEGLU (layers/perceiver_encoder.py:205-208): h = [x | gate]; out = gelu_erf(gate) * x; padded cols zeroed
__global__ void geglu_kernel(const float* __restrict__ h, int rows, int inner, int inner_pad, float* __restrict__ out) {
    const size_t total = (size_t)rows * inner_pad;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / inner_pad;
        const int c = (int)(i % inner_pad);
        float v = 0.0f;
        if (c < inner) {
            const float x = h[r * 2 * (size_t)inner + c];
            const float g = h[r * 2 * (size_t)inner + inner + c];
            v = gelu_erf(g) * x;
        }
        out[i] = v;
    }
}
cudaError_t launch_geglu(const float* h, int rows, int inner, int inner_pad, float* out, cudaStream_t st,
                         unsigned long long* nlaunch) {
    size_t total = (size_t)rows * inner_pad;
    int blocks = (int)min((size_t)1184, (total + 255) / 256);
    geglu_kernel<<<blocks, 256, 0, st>>>(h, rows, inner, inner_pad, out);
    GV_BUMP(nlaunch);
    return cudaGetLastError();
}

// =============================================================================================
// generation state + sampling (layers/stream_generator.py:809-881)
// =============================================================================================
__global__ void init_state_kernel(GenState* st, unsigned char* seen, int B, int P, int V, int Vpad, int start_audio) {
    // input_ids starts as the fake prefix [1]*P + [start_audio] (layers/gpt.py:582-592): both ids are
    // "seen" by the repetition penalty from step 0 on.
    for (int i = threadIdx.x; i < B * Vpad; i += blockDim.x) {
        int t = i % Vpad;
        seen[i] = (t == 1 || t == start_audio) ? 1 : 0;
    }
    if (threadIdx.x == 0) {
        st->n_emitted = 0;
        st->done = 0;
        st->has_pending = 1;
        st->P = P;
        st->B = B;
        for (int b = 0; b < GV_MAX_BATCH; ++b) {
            st->finished[b] = 0;
            st->last_tok[b] = start_audio;
        }
    }
}
cudaError_t launch_init_state(GenState* st, unsigned char* seen, int B, int P, int V, int Vpad, int start_audio,
                              cudaStream_t s, unsigned long long* nlaunch) {
    init_state_kernel<<<1, 256, 0, s>>>(st, seen, B, P, V, Vpad, start_audio);
    GV_BUMP(nlaunch);
    return cudaGetLastError();
}

struct BlockSync {
    __device__ __forceinline__ void operator()() const { __syncthreads(); }
};

// One CTA per row; the last CTA to finish (atomic ticket) updates the shared step bookkeeping.
__global__ void __launch_bounds__(GV_SAMPLE_THREADS) sample_kernel(SampleArgs a, int* ticket) {
    GenState* st = a.st;
    if (st->done) {
        if (blockIdx.x == 0 && threadIdx.x == 0) a.status[1] = 1;
        return;
    }
    __shared__ unsigned long long keys[GV_SORT_N];
    __shared__ float fscr[16];
    __shared__ int iscr[16];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float* lg = a.logits + (size_t)b * a.V;
    unsigned char* seen = a.seen + (size_t)b * a.Vpad;
    SampleCfg c{a.V, a.top_k, a.top_p, a.top_p_threshold, a.temperature, a.rep_penalty};
    const int n = st->n_emitted;
    int tok = sample_token([&](int e) { return ldcg(lg + e); }, seen, c, a.noise ? a.noise + (size_t)b * a.V : nullptr, a.seed, (uint32_t)n, (uint32_t)b,
                           keys, fscr, iscr, tid, BlockSync());
    if (a.forced) {
        const long long f = a.forced[b];
        tok = (f >= 0 && f < (long long)a.V) ? (int)f : a.stop_token;  // out-of-range ids never index seen[] / mel_emb
        if (tok != (int)f && tid == 0 && a.bad_ids != nullptr) atomicOr(a.bad_ids, 1);
    }
    const int was_finished = st->finished[b];
    if (!a.ignore_eos && was_finished) tok = a.stop_token;  // finished rows emit the pad (== eos) token
    // emit (token, latent[, logits]) — the yield of sample_stream happens before the EOS test
    if (tid == 0) a.ids_out[b] = tok;
    for (int i = tid; i < a.D; i += blockDim.x) a.latents_out[(size_t)b * a.D + i] = ldcg(a.latent + (size_t)b * a.D + i);
    if (a.logits_out)
        for (int i = tid; i < a.V; i += blockDim.x) a.logits_out[(size_t)b * a.V + i] = ldcg(lg + i);
    __syncthreads();
    if (tid == 0) {
        seen[tok] = 1;
        st->last_tok[b] = tok;
        if (!a.ignore_eos && tok == a.stop_token) st->finished[b] = 1;
        __threadfence();
        const int t = atomicAdd(ticket, 1);
        if (t == gridDim.x - 1) {  // last row done: step bookkeeping
            *ticket = 0;
            __threadfence();
            int all_fin = 1;
            for (int r = 0; r < (int)gridDim.x; ++r) all_fin &= (*(volatile int*)&st->finished[r]);
            st->n_emitted = n + 1;
            st->has_pending = 0;
            a.status[0] = a.step_in_call + 1;
            if (a.bad_ids != nullptr && atomicExch(a.bad_ids, 0) != 0) a.status[2] = 1;  // out-of-range ids since the last status
            if (all_fin || n + 1 >= a.max_total) {
                st->done = 1;
                a.status[1] = 1;
            }
        }
    }
}
cudaError_t launch_sample(const SampleArgs& a, int B, cudaStream_t st, unsigned long long* nlaunch) {
    // the ticket lives right after the GenState struct (workspace reserves room)
    int* ticket = reinterpret_cast<int*>(reinterpret_cast<char*>(a.st) + sizeof(GenState));
    sample_kernel<<<B, GV_SAMPLE_THREADS, 0, st>>>(a, ticket);
    GV_BUMP(nlaunch);
    return cudaGetLastError();
}

// =============================================================================================
// Single-query KV-cache attention: one CTA (8 warps) per (cache row b, head h).  Used by the
// batched per-op decode step and by the KV microbenchmark (BASELINE configs[4]).
//   q   + b*q_bs + h*HD          k,v + b*kv_bs + h*S_max*HD          out + b*o_bs + h*HD
// =============================================================================================
template <int HD>
__global__ void __launch_bounds__(256) kv_attention_kernel(const float* __restrict__ q, long q_bs,
                                                           const float* __restrict__ k, const float* __restrict__ v,
                                                           long kv_bs, int S, int S_max, float* __restrict__ out, long o_bs,
                                                           const int* skip) {
    if (skip && *skip) return;
    __shared__ float sm[GV_ATT_WARPS * (HD + 2)];
    const size_t b = blockIdx.y, h = blockIdx.x;
    const float* Kc = k + b * kv_bs + h * (size_t)S_max * HD;
    const float* Vc = v + b * kv_bs + h * (size_t)S_max * HD;
    attn_decode_item<HD>(q + b * q_bs + h * HD, Kc, Vc, 0, S, sqrtf((float)HD), sm, threadIdx.x, BlockSync(),
                         out + b * o_bs + h * HD, nullptr);
}
cudaError_t launch_kv_attention(const float* q, long q_bs, const float* k, const float* v, long kv_bs, int B, int H, int hd,
                                int S, int S_max, float* out, long o_bs, const int* skip, cudaStream_t st,
                                unsigned long long* nlaunch) {
    if (B <= 0 || H <= 0 || S <= 0 || S > S_max) return cudaErrorInvalidValue;
    dim3 grid(H, B);
    switch (hd) {
#define GV_KVA_CASE(n) \
    case n: kv_attention_kernel<n><<<grid, 256, 0, st>>>(q, q_bs, k, v, kv_bs, S, S_max, out, o_bs, skip); break;
        GV_KVA_CASE(32) GV_KVA_CASE(64) GV_KVA_CASE(128) GV_KVA_CASE(256)
#undef GV_KVA_CASE
        default: return cudaErrorInvalidValue;
    }
    GV_BUMP(nlaunch);
    return cudaGetLastError();
}

}  // namespace gv
