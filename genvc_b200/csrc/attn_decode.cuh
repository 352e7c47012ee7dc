// Single-query KV-cache attention for one (cache, head) item over a key range, executed by 8 warps.
// Arithmetic of HF GPT2Attention._attn for q_len == 1 (modeling_gpt2.py eager path, transformers 4.33):
//   s_j = (q . k_j) / sqrt(hd);  p = softmax_j(s);  o = sum_j p_j v_j
// evaluated as an online softmax: keys are dealt round-robin to the warps, each warp keeps a
// running (max, sum, o[hd]) and the 8 partial states are merged through shared memory.  The
// work is HBM/L2-bound (K and V rows are read exactly once, 128-bit coalesced per lane);
// reductions are warp shuffles.  Shared by the fused decode kernel and the KV microbenchmark.
#pragma once
#include "common.cuh"

namespace gv {

#define GV_ATT_WARPS 8

template <int HD>
struct AttLane {
    static constexpr int VEC = (HD >= 128) ? 4 : (HD / 32);  // floats per lane per chunk
    static constexpr int NCH = HD / (32 * VEC);              // chunks per lane
    static constexpr int DPL = VEC * NCH;                    // dims per lane
};

template <int VEC>
__device__ __forceinline__ void ld_vec(const float* p, float* r) {
    if constexpr (VEC == 4) {
        float4 v = ldcg4(p);
        r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
    } else if constexpr (VEC == 2) {
        float2 v = ldcg2(p);
        r[0] = v.x; r[1] = v.y;
    } else {
        r[0] = ldcg(p);
    }
}

// smem: GV_ATT_WARPS * (HD + 2) floats.
// Writes o_out[HD] (normalised when ml_out == nullptr, otherwise un-normalised with ml_out = {m, l}).
template <int HD, class Sync>
__device__ void attn_decode_item(const float* __restrict__ q, const float* __restrict__ Kc,
                                 const float* __restrict__ Vc, int j0, int j1, float sqrt_hd, float* smem, int tid,
                                 Sync sync, float* o_out, float* ml_out) {
    using L = AttLane<HD>;
    constexpr int VEC = L::VEC, NCH = L::NCH, DPL = L::DPL;
    const int warp = tid >> 5, lane = tid & 31;
    float qr[DPL];
#pragma unroll
    for (int c = 0; c < NCH; ++c) ld_vec<VEC>(q + (c * 32 + lane) * VEC, qr + c * VEC);

    float m = -INFINITY, l = 0.0f;
    float o[DPL];
#pragma unroll
    for (int i = 0; i < DPL; ++i) o[i] = 0.0f;

    constexpr int UNR = 4;
    for (int jb = j0 + warp * UNR; jb < j1; jb += GV_ATT_WARPS * UNR) {
        float kr[UNR][DPL], vr[UNR][DPL];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int j = jb + u;
            if (j < j1) {
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    ld_vec<VEC>(Kc + (size_t)j * HD + (c * 32 + lane) * VEC, kr[u] + c * VEC);
                    ld_vec<VEC>(Vc + (size_t)j * HD + (c * 32 + lane) * VEC, vr[u] + c * VEC);
                }
            } else {
#pragma unroll
                for (int i = 0; i < DPL; ++i) { kr[u][i] = 0.0f; vr[u][i] = 0.0f; }
            }
        }
        float s[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            float d = 0.0f;
#pragma unroll
            for (int i = 0; i < DPL; ++i) d = fmaf(qr[i], kr[u][i], d);
            s[u] = d;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
            for (int u = 0; u < UNR; ++u) s[u] += __shfl_xor_sync(0xffffffffu, s[u], off);
        }
        float mnew = m;
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            s[u] = (jb + u < j1) ? s[u] / sqrt_hd : -INFINITY;
            mnew = fmaxf(mnew, s[u]);
        }
        const float corr = expf(m - mnew);  // m == -inf on the first group -> 0
        l *= corr;
#pragma unroll
        for (int i = 0; i < DPL; ++i) o[i] *= corr;
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const float p = expf(s[u] - mnew);  // masked -> exp(-inf) = 0
            l += p;
#pragma unroll
            for (int i = 0; i < DPL; ++i) o[i] = fmaf(p, vr[u][i], o[i]);
        }
        m = mnew;
    }

    // merge the 8 warp states
    float* so = smem;                         // [warp][HD]
    float* sml = smem + GV_ATT_WARPS * HD;    // [warp][2]
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int v = 0; v < VEC; ++v) so[warp * HD + (c * 32 + lane) * VEC + v] = o[c * VEC + v];
    if (lane == 0) {
        sml[warp * 2] = m;
        sml[warp * 2 + 1] = l;
    }
    sync();
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < GV_ATT_WARPS; ++w) M = fmaxf(M, sml[w * 2]);
    float Lsum = 0.0f;
    float wgt[GV_ATT_WARPS];
#pragma unroll
    for (int w = 0; w < GV_ATT_WARPS; ++w) {
        wgt[w] = (sml[w * 2] == -INFINITY) ? 0.0f : expf(sml[w * 2] - M);
        Lsum += sml[w * 2 + 1] * wgt[w];
    }
    for (int d = tid; d < HD; d += GV_ATT_WARPS * 32) {
        float acc = 0.0f;
#pragma unroll
        for (int w = 0; w < GV_ATT_WARPS; ++w) acc = fmaf(so[w * HD + d], wgt[w], acc);
        o_out[d] = ml_out ? acc : acc / Lsum;
    }
    if (ml_out && tid == 0) {
        ml_out[0] = M;
        ml_out[1] = Lsum;
    }
    sync();
}

}  // namespace gv
