// Fused persistent decode kernel (B = 1): the whole sample()/sample_stream() loop of
// layers/stream_generator.py:809-881 — per step the cached GPT-2 forward of
// layers/gpt_inference.py:92-112 (30 pre-LN blocks, ln_f, final_norm, mel_head) followed by the HF
// sampling chain — runs in ONE cooperative launch for up to n_steps tokens, tokens fed back on-chip.
//
// Why this shape.  A decode step at batch 1 reads every weight exactly once (1.516 GB fp32) and does
// 2 flops per weight: an HBM-streaming problem with 30 x 5 serial all-to-all dependencies ("hops").
// The step time is max(HBM time, sum of hop latencies + per-phase compute chains), so the design
// keeps HBM busy across the hops and keeps every chain short:
//   * one persistent CTA per SM (cooperative launch), each owning a fixed slice of every matrix;
//     the slices are pre-packed so a CTA reads one contiguous byte stream (stream_layout.h);
//   * a producer thread per CTA streams that region HBM -> shared memory with 1-D bulk TMA copies
//     (cp.async.bulk, completion on mbarriers) into an 11 x 16 KB ring.  The weight stream does
//     not depend on activations, so it runs ahead across phase boundaries and across tokens: HBM
//     does not idle while the consumers exchange activations.  At most `window` tiles are in
//     flight per SM (enough to cover the bandwidth-delay product, few enough not to queue in front
//     of the latency-critical exchange traffic);
//   * 16 consumer warps do the GEMVs out of shared memory in fp32 FMA.  K = D matrices (QKV, attn
//     proj, FC, logits head): a unit is one output column, one warp per unit, the activation vector
//     in 32 registers per lane, 8 LDS.128 + 32 FFMA per lane per unit and one shuffle tree — no
//     cross-warp reduction.  mlp.c_proj (K = 4D) is split along K instead: the CTA that computed
//     u_k owns ROW k of W_proj2 and accumulates u_k * W[k, :] into a D-wide partial (thread t owns
//     outputs 2t, 2t+1: no reduction at all), so the 4D-wide activation never crosses CTAs; the G
//     partials are summed in a fixed order by D/8 reducer CTAs (deterministic, no float atomics);
//   * hops go through L2: producers store {value, tag} words (tag = global hop number) and bump an
//     arrival counter (one relaxed red per CTA, no fence); ONE thread per CTA spins on the counter,
//     a block barrier releases the others, which load the words they need and re-poll the rare word
//     whose tag is stale.  The counter is only a hint — validity comes from the tags;
//   * single-token attention is split over (head, key range) items run by the first H*nsplit CTAs:
//     the item's K/V rows are requested from the cache BEFORE the hop wait, online softmax with
//     warp-shuffle reductions, 16 warp states merged through shared memory; the item that covers
//     the newest position takes k/v from the exchange buffer and appends them to the cache;
//   * sampling is computed redundantly by every CTA (same data, same code => same token), so the
//     next token needs no broadcast.
// No tensor cores: at M = 1 there is no reuse to feed them; fp32 keeps greedy parity.
#include "common.cuh"
#include "mega.cuh"
#include "sampling.cuh"
#include "stream_layout.h"

namespace gv {

#define MEGA_WARPS 16
#define MEGA_CONSUMERS (MEGA_WARPS * 32)
#define MEGA_THREADS (MEGA_CONSUMERS + 32)  // + one producer warp (one working thread)
#define MEGA_SPIN_LIMIT (1u << 26)
#define NSLOT GV_MEGA_NSLOT
#define UPT GV_MEGA_UPT
#define MEGA_SCRATCH_BYTES (16384 + 512)

enum { TG_XQ = 0, TG_AO = 1, TG_X1 = 2, TG_PP = 3, TG_X2 = 4 };

struct Ring {
    float* slots;
    uint64_t* full;
    uint64_t* empty;
    int slot_floats;
};

// ---------------------------------------------------------------------------------------------
// stream packing (init time): gather the reference-layout matrices into the per-CTA streams
//   w_nk = 0: unit n = column n of W [K = D, N]  (HF Conv1D);  w_nk = 1: unit n = row n of W [N, D]
// ---------------------------------------------------------------------------------------------
__global__ void pack_stream_kernel(StreamDims s, int layer, int ph, const float* __restrict__ W,
                                   const float* __restrict__ bias, int w_nk, float* __restrict__ stream) {
    const int N = ph_N(s, ph), D = s.D;
    const int n = blockIdx.y;
    const int c = col_owner(N, n, s.G);
    long long dst = cta_base(s, c);
    if (ph == PH_HEAD) dst += (long long)s.L * cta_layer_floats(s, c);
    else dst += (long long)layer * cta_layer_floats(s, c) + ph_offset_in_layer(s, ph, c);
    dst += (long long)(n - col_begin(N, c, s.G)) * unit_floats(D);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < D + 4; k += gridDim.x * blockDim.x) {
        float v = 0.0f;
        if (k < D) v = w_nk ? W[(size_t)n * D + k] : W[(size_t)k * N + n];
        else if (k == D && bias != nullptr) v = bias[n];
        stream[dst + k] = v;
    }
}

cudaError_t launch_pack_stream(const StreamDims& s, int layer, int ph, const float* W, const float* bias, int w_nk,
                               float* stream, cudaStream_t st) {
    dim3 grid((s.D + 4 + 255) / 256, ph_N(s, ph));
    pack_stream_kernel<<<grid, 256, 0, st>>>(s, layer, ph, W, bias, w_nk, stream);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// tagged exchange through L2: element i of a buffer lives at floats [2i, 2i+1] = {value, tag}
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_tagged(float* buf, int idx, float v, uint32_t tag) {
    asm volatile("st.relaxed.gpu.global.v2.b32 [%0], {%1, %2};" ::"l"(buf + 2 * (size_t)idx), "r"(__float_as_uint(v)),
                 "r"(tag)
                 : "memory");
}
// elements idx, idx+1 (idx even) in one 16-byte store
__device__ __forceinline__ void st_tagged2(float* buf, int idx, float v0, float v1, uint32_t tag) {
    asm volatile("st.relaxed.gpu.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(buf + 2 * (size_t)idx),
                 "r"(__float_as_uint(v0)), "r"(tag), "r"(__float_as_uint(v1)), "r"(tag)
                 : "memory");
}
__device__ __forceinline__ uint4 ld_x16(const float* p) {
    uint4 r;
    asm volatile("ld.relaxed.gpu.global.v4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p)
                 : "memory");
    return r;
}
__device__ __forceinline__ uint2 ld_x8(const float* p) {
    uint2 r;
    asm volatile("ld.relaxed.gpu.global.v2.b32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ bool tags_ok(const uint4& a, uint32_t tag, uint32_t tmask) {
    return (((a.y ^ tag) | (a.w ^ tag)) & tmask) == 0u;
}
// elements idx, idx+1 (idx even): spin until both carry `tag`
__device__ __forceinline__ float2 ld_tagged2(const float* buf, int idx, uint32_t tag, uint32_t tmask) {
    const float* p = buf + 2 * (size_t)idx;
    uint4 a = ld_x16(p);
    uint32_t spins = 0;
    while (!tags_ok(a, tag, tmask)) {
        if (++spins > MEGA_SPIN_LIMIT) __trap();
        a = ld_x16(p);
    }
    return make_float2(__uint_as_float(a.x), __uint_as_float(a.z));
}
__device__ __forceinline__ float ld_tagged1(const float* buf, int idx, uint32_t tag, uint32_t tmask) {
    const float* p = buf + 2 * (size_t)idx;
    uint2 a = ld_x8(p);
    uint32_t spins = 0;
    while (((a.y ^ tag) & tmask) != 0u) {
        if (++spins > MEGA_SPIN_LIMIT) __trap();
        a = ld_x8(p);
    }
    return __uint_as_float(a.x);
}
// NE (1, 2 or 4) consecutive elements starting at idx (idx % NE == 0)
template <int NE>
__device__ __forceinline__ void ld_tagged_vec(const float* buf, int idx, uint32_t tag, uint32_t tmask, float* out) {
    if constexpr (NE == 1) {
        out[0] = ld_tagged1(buf, idx, tag, tmask);
    } else if constexpr (NE == 2) {
        const float2 v = ld_tagged2(buf, idx, tag, tmask);
        out[0] = v.x;
        out[1] = v.y;
    } else {
        static_assert(NE == 4, "NE must be 1, 2 or 4");
        const float* p = buf + 2 * (size_t)idx;
        uint4 a = ld_x16(p), b = ld_x16(p + 4);
        uint32_t spins = 0;
        while (!(tags_ok(a, tag, tmask) && tags_ok(b, tag, tmask))) {
            if (++spins > MEGA_SPIN_LIMIT) __trap();
            a = ld_x16(p);
            b = ld_x16(p + 4);
        }
        out[0] = __uint_as_float(a.x);
        out[1] = __uint_as_float(a.z);
        out[2] = __uint_as_float(b.x);
        out[3] = __uint_as_float(b.z);
    }
}

// ---------------------------------------------------------------------------------------------
// hops: arrival counter (a hint: one poller per CTA) + block barrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_relaxed_add(unsigned* p, unsigned v) {
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// every consumer thread's exchange stores are issued (program order before the barrier), then one arrival
__device__ __forceinline__ void hop_arrive(unsigned* cnt, int tid) {
    bar_sync(1, MEGA_CONSUMERS);
    if (tid == 0) red_relaxed_add(cnt, 1u);
}
__device__ __forceinline__ void hop_wait(const unsigned* cnt, unsigned target, int tid, uint32_t tmask) {
    if (tid == 0 && tmask != 0u) {
        uint32_t spins = 0;
        while (ld_relaxed_u32(cnt) < target) {
            if (++spins > MEGA_SPIN_LIMIT) __trap();
        }
    }
    bar_sync(1, MEGA_CONSUMERS);
}

// ---------------------------------------------------------------------------------------------
// weight ring (consumer side).  Every consumer warp walks every tile in order and arrives on its
// empty barrier (count = 16 warps) whether or not it read the tile; only readers wait for `full`.
// ---------------------------------------------------------------------------------------------
struct Cons {
    uint32_t slot, phase;  // ring position of the next tile
    uint32_t tiles;        // tiles consumed so far (same in every consumer thread)
};
__device__ __forceinline__ const float* tile_ptr(const Ring& r, const Cons& cs) {
    return r.slots + (size_t)cs.slot * r.slot_floats;
}
__device__ __forceinline__ void tile_wait(const Ring& r, const Cons& cs) { mbar_wait(&r.full[cs.slot], cs.phase); }
__device__ __forceinline__ void tile_release(const Ring& r, Cons& cs, int lane) {
    __syncwarp();
    if (lane == 0) mbar_arrive(&r.empty[cs.slot]);
    cs.tiles += 1;
    if (++cs.slot == NSLOT) {
        cs.slot = 0;
        cs.phase ^= 1u;
    }
}

// sum of the per-warp partials red[0..15] (written before a block barrier)
__device__ __forceinline__ float sum16(const float* red) {
    const float4 a = *reinterpret_cast<const float4*>(red), b = *reinterpret_cast<const float4*>(red + 4);
    const float4 c = *reinterpret_cast<const float4*>(red + 8), d = *reinterpret_cast<const float4*>(red + 12);
    return (((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w))) +
           (((c.x + c.y) + (c.z + c.w)) + ((d.x + d.y) + (d.z + d.w)));
}

// ---------------------------------------------------------------------------------------------
// LayerNorm of the vector held two elements per thread (thread t owns x[2t], x[2t+1]; valid: 2t < D).
// Two-pass statistics (mean, then centred sum of squares), two block barriers.  `red` is a 32-float
// scratch; w / b in shared memory.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void ln_pair(float& x0, float& x1, bool valid, int D, const float* w, const float* b, float* red,
                                        int tid) {
    const int lane = tid & 31, warp = tid >> 5;
    float s = valid ? (x0 + x1) : 0.0f;
    s = warp_sum(s);
    if (lane == 0) red[warp] = s;
    bar_sync(1, MEGA_CONSUMERS);
    const float mean = sum16(red) / (float)D;
    const float d0 = x0 - mean, d1 = x1 - mean;
    float q = valid ? fmaf(d0, d0, d1 * d1) : 0.0f;
    q = warp_sum(q);
    if (lane == 0) red[16 + warp] = q;
    bar_sync(1, MEGA_CONSUMERS);
    const float var = sum16(red + 16) / (float)D;
    const float rstd = 1.0f / sqrtf(var + 1e-5f);
    if (valid) {
        const float2 ww = *reinterpret_cast<const float2*>(w + 2 * tid);
        const float2 bb = *reinterpret_cast<const float2*>(b + 2 * tid);
        x0 = d0 * rstd * ww.x + bb.x;
        x1 = d1 * rstd * ww.y + bb.y;
    }
}

// ---------------------------------------------------------------------------------------------
// GEMV, K = D, one warp per unit: unit u of the phase is handled by warp u % 16 (u < nunits <= 32).
// `xs` is the activation vector in shared memory (complete: the caller synchronised).  For every
// unit, lane 0 (u < 16) or lane 16 (u >= 16) of the owning warp calls epi(u, y) with
//   y = bias + sum_k x_k W[k][unit].
// ---------------------------------------------------------------------------------------------
template <int NXV, class Epi>
__device__ __forceinline__ void gemv_dot(const Ring& ring, Cons& cs, int nunits, const float* xs, int warp, int lane, Epi epi) {
    constexpr int D = NXV * 128, UF = D + 4;
    float4 xv[NXV];
#pragma unroll
    for (int i = 0; i < NXV; ++i) xv[i] = *reinterpret_cast<const float4*>(xs + (i * 32 + lane) * 4);
    float tot0 = 0.0f, tot1 = 0.0f;
    const int ntiles = (nunits + UPT - 1) / UPT;
    for (int t = 0; t < ntiles; ++t) {
        if ((t & 3) == (warp >> 2) && t * UPT + (warp & 3) < nunits) {
            tile_wait(ring, cs);
            const float* col = tile_ptr(ring, cs) + (warp & 3) * UF;
            float a0 = (lane == 0) ? col[D] : 0.0f;  // bias folded into the first partial
            float a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll
            for (int i = 0; i < NXV; ++i) {
                const float4 wv = *reinterpret_cast<const float4*>(col + (i * 32 + lane) * 4);
                a0 = fmaf(wv.x, xv[i].x, a0);
                a1 = fmaf(wv.y, xv[i].y, a1);
                a2 = fmaf(wv.z, xv[i].z, a2);
                a3 = fmaf(wv.w, xv[i].w, a3);
            }
            const float sum = (a0 + a1) + (a2 + a3);
            if (t < 4) tot0 = sum;
            else tot1 = sum;
        }
        tile_release(ring, cs, lane);
    }
    // lanes 0..15 reduce tot0, lanes 16..31 reduce tot1
    const bool up = (lane & 16) != 0;
    const float keep = up ? tot1 : tot0, send = up ? tot0 : tot1;
    float v = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    const int u = warp + (up ? 16 : 0);
    if ((lane & 15) == 0 && u < nunits) epi(u, v);
}

// mlp.c_proj split along K: unit k = row of W_proj2 owned by this CTA; acc{0,1} += u_k * W[k][2t, 2t+1]
template <int NXV>
__device__ __forceinline__ void gemv_outer(const Ring& ring, Cons& cs, int nunits, const float* us, int tid, int lane,
                                           float& acc0, float& acc1) {
    constexpr int D = NXV * 128, UF = D + 4;
    const bool valid = 2 * tid < D;
    const int ntiles = (nunits + UPT - 1) / UPT;
    for (int t = 0; t < ntiles; ++t) {
        tile_wait(ring, cs);
        const float* base = tile_ptr(ring, cs) + 2 * tid;
#pragma unroll
        for (int r = 0; r < UPT; ++r) {
            const int k = t * UPT + r;
            if (k < nunits && valid) {
                const float uk = us[k];
                const float2 w = *reinterpret_cast<const float2*>(base + r * UF);
                acc0 = fmaf(uk, w.x, acc0);
                acc1 = fmaf(uk, w.y, acc1);
            }
        }
        tile_release(ring, cs, lane);
    }
}

// ---------------------------------------------------------------------------------------------
// single-query attention over one (head, key range) item, 16 warps.  Arithmetic of HF
// GPT2Attention._attn for q_len == 1:  s_j = (q . k_j) / sqrt(hd);  p = softmax_j(s);  o = sum_j p_j v_j
// as an online softmax: key j0 + w + 16 g belongs to warp w; the 16 warp states are merged through
// shared memory.  The first group's K/V rows are requested before the hop wait for q.
// ---------------------------------------------------------------------------------------------
template <int HD>
struct AttLane {
    static constexpr int VEC = (HD >= 128) ? 4 : (HD / 32);  // floats per lane per chunk
    static constexpr int NCH = HD / (32 * VEC);              // chunks per lane
    static constexpr int DPL = VEC * NCH;                    // dims per lane
};
template <int VEC>
__device__ __forceinline__ void ld_vec(const float* p, float* r) {
    if constexpr (VEC == 4) {
        const float4 v = ldcg4(p);
        r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
    } else if constexpr (VEC == 2) {
        const float2 v = ldcg2(p);
        r[0] = v.x; r[1] = v.y;
    } else {
        r[0] = ldcg(p);
    }
}
template <int VEC>
__device__ __forceinline__ void st_vec(float* p, const float* r) {
    if constexpr (VEC == 4) __stcg(reinterpret_cast<float4*>(p), make_float4(r[0], r[1], r[2], r[3]));
    else if constexpr (VEC == 2) __stcg(reinterpret_cast<float2*>(p), make_float2(r[0], r[1]));
    else __stcg(p, r[0]);
}

template <int HD>
__device__ void att_item(float* __restrict__ Kc, float* __restrict__ Vc, const float* xq, int D, int h, int j0, int j1, int S,
                         uint32_t tag_in, const unsigned* cnt_in, unsigned target_in, float* so, float* sml, int tid,
                         float* o_out, float* ml_out, int item, uint32_t tag_out, uint32_t tmask) {
    using L = AttLane<HD>;
    constexpr int VEC = L::VEC, NCH = L::NCH, DPL = L::DPL, UNR = 2;
    const int warp = tid >> 5, lane = tid & 31;
    const float sqrt_hd = sqrtf((float)HD);
    float m = -INFINITY, l = 0.0f;
    float o[DPL], qr[DPL];
#pragma unroll
    for (int i = 0; i < DPL; ++i) {
        o[i] = 0.0f;
        qr[i] = 0.0f;
    }
    float kr[UNR][DPL], vr[UNR][DPL];
    auto load_group = [&](int jb) {
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int j = jb + u * MEGA_WARPS;
            if (j < j1 && j != S - 1) {
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    ld_vec<VEC>(Kc + (size_t)j * HD + (c * 32 + lane) * VEC, kr[u] + c * VEC);
                    ld_vec<VEC>(Vc + (size_t)j * HD + (c * 32 + lane) * VEC, vr[u] + c * VEC);
                }
            } else {
#pragma unroll
                for (int i = 0; i < DPL; ++i) {
                    kr[u][i] = 0.0f;
                    vr[u][i] = 0.0f;
                }
            }
        }
    };
    load_group(j0 + warp);       // cache rows: independent of this step's q
    hop_wait(cnt_in, target_in, tid, tmask);
#pragma unroll
    for (int c = 0; c < NCH; ++c) ld_tagged_vec<VEC>(xq, h * HD + (c * 32 + lane) * VEC, tag_in, tmask, qr + c * VEC);
    for (int jb = j0 + warp; jb < j1; jb += UNR * MEGA_WARPS) {
        if (jb != j0 + warp) load_group(jb);
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int j = jb + u * MEGA_WARPS;
            if (j == S - 1 && j < j1) {  // the position being decoded: k/v of this very step, appended to the cache
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const int e = h * HD + (c * 32 + lane) * VEC;
                    ld_tagged_vec<VEC>(xq, D + e, tag_in, tmask, kr[u] + c * VEC);
                    ld_tagged_vec<VEC>(xq, 2 * D + e, tag_in, tmask, vr[u] + c * VEC);
                    st_vec<VEC>(Kc + (size_t)j * HD + (c * 32 + lane) * VEC, kr[u] + c * VEC);
                    st_vec<VEC>(Vc + (size_t)j * HD + (c * 32 + lane) * VEC, vr[u] + c * VEC);
                }
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
            s[u] = (jb + u * MEGA_WARPS < j1) ? s[u] / sqrt_hd : -INFINITY;
            mnew = fmaxf(mnew, s[u]);
        }
        const float corr = expf(m - mnew);  // m == -inf on the first group -> 0
        l *= corr;
#pragma unroll
        for (int i = 0; i < DPL; ++i) o[i] *= corr;
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const float pr = expf(s[u] - mnew);  // masked -> exp(-inf) = 0
            l += pr;
#pragma unroll
            for (int i = 0; i < DPL; ++i) o[i] = fmaf(pr, vr[u][i], o[i]);
        }
        m = mnew;
    }
    // merge the 16 warp states: so [warp][HD], sml [warp][2]
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int v = 0; v < VEC; ++v) so[warp * HD + (c * 32 + lane) * VEC + v] = o[c * VEC + v];
    if (lane == 0) {
        sml[warp * 2] = m;
        sml[warp * 2 + 1] = l;
    }
    bar_sync(1, MEGA_CONSUMERS);
    if (tid < HD) {
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < MEGA_WARPS; ++w) M = fmaxf(M, sml[w * 2]);
        float Lsum = 0.0f, acc = 0.0f;
#pragma unroll
        for (int w = 0; w < MEGA_WARPS; ++w) {
            const float wgt = (sml[w * 2] == -INFINITY) ? 0.0f : expf(sml[w * 2] - M);
            Lsum += sml[w * 2 + 1] * wgt;
            acc = fmaf(so[w * HD + tid], wgt, acc);
        }
        st_tagged(o_out, item * HD + tid, acc, tag_out);
        if (tid == 0) st_tagged2(ml_out, item * 2, M, Lsum, tag_out);
    }
}

// ---------------------------------------------------------------------------------------------
// producer: one thread walks the CTA's weight stream through the ring
// ---------------------------------------------------------------------------------------------
struct Producer {
    const Ring& ring;
    uint32_t t = 0, slot = 0, phase = 0;
    uint32_t window;  // at most this many tiles requested but not landed
    volatile int* stop;
    uint64_t policy;
    __device__ Producer(const Ring& r, volatile int* s, uint32_t w) : ring(r), window(w), stop(s) {
        policy = l2_policy_evict_first();
    }
    // returns false when the consumers asked to stop
    __device__ bool issue(const float* src, uint32_t floats, bool stream_once) {
        uint32_t spins = 0;
        while (!mbar_try_wait(&ring.empty[slot], phase ^ 1u)) {
            if (*stop) return false;
            if (++spins > MEGA_SPIN_LIMIT) __trap();
        }
        if (*stop) return false;
        if (t >= window) {  // tile t - window must have landed
            const uint32_t o = t - window;
            spins = 0;
            while (!mbar_try_wait(&ring.full[o % NSLOT], (o / NSLOT) & 1u)) {
                if (++spins > MEGA_SPIN_LIMIT) __trap();
            }
        }
        mbar_arrive_expect_tx(&ring.full[slot], floats * 4u);
        float* dst = ring.slots + (size_t)slot * ring.slot_floats;
        if (stream_once) bulk_g2s_hint(dst, src, floats * 4u, &ring.full[slot], policy);
        else bulk_g2s(dst, src, floats * 4u, &ring.full[slot]);
        ++t;
        if (++slot == NSLOT) {
            slot = 0;
            phase ^= 1u;
        }
        return true;
    }
    __device__ bool issue_units(const float*& src, int nunits, int uf) {
        for (int u0 = 0; u0 < nunits; u0 += UPT) {
            const int nu = min(UPT, nunits - u0);
            if (!issue(src, (uint32_t)(nu * uf), true)) return false;
            src += (long long)nu * uf;
        }
        return true;
    }
};

__device__ bool produce_forward(Producer& pr, const MegaParams& p, const StreamDims& sd, int cta) {
    const float* base = p.stream + cta_base(sd, cta);
    const long long lfl = cta_layer_floats(sd, cta);
    const int D = p.D, uf = unit_floats(D);
    int nun[5];
    for (int ph = 0; ph < 5; ++ph) nun[ph] = ph_units(sd, ph, cta);
    for (int l = 0; l < p.L; ++l) {
        const float* lw = base + (long long)l * lfl;
        if (!pr.issue(p.blob + p.ln1_off + (long long)l * p.layer_stride, 2 * D, false)) return false;
        if (!pr.issue_units(lw, nun[PH_QKV], uf)) return false;
        if (!pr.issue_units(lw, nun[PH_PROJ], uf)) return false;
        if (!pr.issue(p.blob + p.ln2_off + (long long)l * p.layer_stride, 2 * D, false)) return false;
        if (!pr.issue_units(lw, nun[PH_FC], uf)) return false;
        if (!pr.issue_units(lw, nun[PH_P2], uf)) return false;
    }
    if (!pr.issue(p.blob + p.lnf_off, 4 * D, false)) return false;
    const float* hw = base + (long long)p.L * lfl;
    return pr.issue_units(hw, nun[PH_HEAD], uf);
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <int NXV>
__global__ void __launch_bounds__(MEGA_THREADS, 1) decode_mega_kernel(MegaParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int D = NXV * 128;
    const int tid_all = threadIdx.x;
    const int cta = blockIdx.x;
    const int G = gridDim.x;
    const StreamDims sd{p.L, D, p.V, G};

    // ---- shared memory carve-up (mirrored by mega_smem_bytes) ----
    Ring ring;
    ring.slot_floats = slot_floats(D);
    size_t off = 0;
    ring.slots = reinterpret_cast<float*>(smem_raw);
    off += (size_t)NSLOT * ring.slot_floats * sizeof(float);
    off = (off + 127) & ~size_t(127);
    // scratch region: sampling sort keys | attention merge buffer | partial-sum gather (never live together)
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw + off);
    float* att_so = reinterpret_cast<float*>(smem_raw + off);            // [16][HD]  (<= 16 KB)
    float* att_sml = reinterpret_cast<float*>(smem_raw + off + 16384);   // [16][2]
    float* gat = reinterpret_cast<float*>(smem_raw + off);               // [G][8]
    off += MEGA_SCRATCH_BYTES;
    float* slog = reinterpret_cast<float*>(smem_raw + off);  // [Vpad] logits of the step being sampled
    off += (size_t)p.Vpad * sizeof(float);
    float* xres0 = reinterpret_cast<float*>(smem_raw + off);  // [D] residual stream entering the block
    off += (size_t)D * sizeof(float);
    float* xres1 = reinterpret_cast<float*>(smem_raw + off);  // [D] residual stream after attention
    off += (size_t)D * sizeof(float);
    float* xn = reinterpret_cast<float*>(smem_raw + off);  // [D] GEMV input vector
    off += (size_t)D * sizeof(float);
    ring.full = reinterpret_cast<uint64_t*>(smem_raw + off);
    off += 16 * sizeof(uint64_t);
    ring.empty = reinterpret_cast<uint64_t*>(smem_raw + off);
    off += 16 * sizeof(uint64_t);
    float* us = reinterpret_cast<float*>(smem_raw + off);  // [32] this CTA's gelu(fc) values
    off += 32 * sizeof(float);
    float* red = reinterpret_cast<float*>(smem_raw + off);  // [32] LayerNorm statistics
    off += 32 * sizeof(float);
    float* fscr = reinterpret_cast<float*>(smem_raw + off);
    off += 16 * sizeof(float);
    int* iscr = reinterpret_cast<int*>(smem_raw + off);
    off += 16 * sizeof(int);
    volatile int* ctl = reinterpret_cast<volatile int*>(smem_raw + off);  // [0] stop flag, [1] tiles consumed, [2] token
    off += 4 * sizeof(int);
    unsigned char* seen = smem_raw + off;  // [Vpad]

    if (tid_all == 0) {
        for (int i = 0; i < NSLOT; ++i) {
            mbar_init(&ring.full[i], 1);
            mbar_init(&ring.empty[i], MEGA_WARPS);
        }
        ctl[0] = 0;
        ctl[1] = 0;
        ctl[2] = 0;
        mbar_fence_init();
    }
    for (int i = tid_all; i < p.Vpad; i += MEGA_THREADS) seen[i] = p.seen[i];
    __syncthreads();

    GenState* st = p.st;
    const int had_pending = st->has_pending;
    const int n_start = st->n_emitted;
    if (st->done) {  // uniform: nothing to do
        if (cta == 0 && tid_all == 0) p.status[1] = 1;
        return;
    }

    if (tid_all >= MEGA_CONSUMERS) {
        // ================= producer warp =================
        if (tid_all == MEGA_CONSUMERS) {
            Producer pr(ring, ctl, (uint32_t)max(1, min(p.window, NSLOT)));
            bool ok = true;
            for (int i = 0; i < p.n_steps && ok; ++i) {
                if (i == 0 && had_pending) continue;
                ok = produce_forward(pr, p, sd, cta);
            }
            // drain: every bulk copy issued must have landed before the CTA may exit.  Wait for the
            // consumers to finish (they may stop early on EOS with copies still in flight), then for
            // the full-barrier of every tile that was issued but never consumed.
            uint32_t spins = 0;
            while (!ctl[0]) {
                if (++spins > (1u << 30)) __trap();
                __nanosleep(64);
            }
            const uint32_t consumed = (uint32_t)ctl[1];
            for (uint32_t t = consumed; t < pr.t; ++t) mbar_wait(&ring.full[t % NSLOT], (t / NSLOT) & 1u);
        }
        return;
    }

    // ================= consumer warps =================
    const int tid = tid_all;
    const int lane = tid & 31, warp = tid >> 5;
    Cons cs{0u, 0u, 0u};
    const uint32_t tmask = p.dbg_nosync ? 0u : 0xffffffffu;  // debug: 0 = do not wait for exchange data
    const bool xvalid = 2 * tid < D;
    const int H = p.H, HD = D / H;
    int nun[5], ubeg[5];
    for (int ph = 0; ph < 5; ++ph) {
        nun[ph] = ph_units(sd, ph, cta);
        ubeg[ph] = (int)col_begin(ph_N(sd, ph), cta, G);
    }
    const int n_red = D / 8;  // reducer CTAs of the mlp.c_proj partial sums (8 outputs each)
    const SampleCfg scfg{p.V, p.top_k, p.top_p, p.top_p_threshold, p.temperature, p.rep_penalty};
    int n = n_start;  // tokens emitted so far
    long long last_tok = st->last_tok[0];
    int finished = st->finished[0];
    int emitted = 0, done = 0;
    uint32_t fwd = 0;  // forwards executed by this launch
    const uint32_t tags_per_fwd = (uint32_t)GV_TAGS_PER_LAYER * (uint32_t)p.L + 1u;
    const float* mel_emb = p.blob + p.mel_emb_off;
    const float* mel_pos = p.blob + p.mel_pos_off;
    unsigned* const hc = p.hops;
    unsigned t_xq = 0, t_ao = 0, t_x1 = 0, t_pp = 0, t_x2 = 0, t_lg = 0;  // hop counter targets (counters are zero at launch)

    for (int i = 0; i < p.n_steps; ++i) {
        const bool tr = p.trace != nullptr && i == p.trace_step && tid == 0;
        unsigned long long* trow = p.trace + (size_t)cta * p.trace_slots;
        auto stamp = [&](int slot) {
            if (tr && slot < p.trace_slots) trow[slot] = globaltimer_ns();
        };
        stamp(p.L * GV_TRACE_PER_LAYER + 4);
        float lat0 = 0.0f, lat1 = 0.0f;  // final_norm(ln_f(x)): the latent of this step (elements 2 tid, 2 tid + 1)
        if (!(i == 0 && had_pending)) {
            // ------------- forward of token `last_tok` at mel position n, cache row P + n -------------
            const uint32_t tbase = p.tag0 + fwd * tags_per_fwd;
            fwd += 1;
            const int pos = p.P + n;
            const int S = pos + 1;
            const int chunk = att_chunk(S), nsplit = att_nsplit(S);
            const int n_items = H * nsplit;
            for (int l = 0; l < p.L; ++l) {
                float* kc = p.kv + ((size_t)l * 2 + 0) * p.kv_layer_stride;
                float* vc = p.kv + ((size_t)l * 2 + 1) * p.kv_layer_stride;
                const uint32_t tg = tbase + (uint32_t)GV_TAGS_PER_LAYER * (uint32_t)l;
                const int ts = l * GV_TRACE_PER_LAYER;
                // ---- QKV: LN1 -> [q|k|v] columns ----
                {
                    float x0 = 0.0f, x1 = 0.0f;
                    if (l == 0) {
                        if (xvalid) {
                            const float2 a = *reinterpret_cast<const float2*>(mel_emb + (size_t)last_tok * D + 2 * tid);
                            const float2 b = *reinterpret_cast<const float2*>(mel_pos + (size_t)n * D + 2 * tid);
                            x0 = a.x + b.x;
                            x1 = a.y + b.y;
                        }
                    } else {
                        hop_wait(hc + HC_X2 * GV_HOP_STRIDE, t_x2, tid, tmask);
                        if (xvalid) {
                            const float2 v = ld_tagged2(p.x2, 2 * tid, tg - (uint32_t)GV_TAGS_PER_LAYER + TG_X2, tmask);
                            x0 = v.x;
                            x1 = v.y;
                        }
                    }
                    if (xvalid) *reinterpret_cast<float2*>(xres0 + 2 * tid) = make_float2(x0, x1);
                    stamp(ts + 0);
                    tile_wait(ring, cs);
                    const float* lnp = tile_ptr(ring, cs);
                    ln_pair(x0, x1, xvalid, D, lnp, lnp + D, red, tid);
                    tile_release(ring, cs, lane);
                    if (xvalid) *reinterpret_cast<float2*>(xn + 2 * tid) = make_float2(x0, x1);
                    bar_sync(1, MEGA_CONSUMERS);
                    gemv_dot<NXV>(ring, cs, nun[PH_QKV], xn, warp, lane,
                                  [&](int u, float y) { st_tagged(p.xq, ubeg[PH_QKV] + u, y, tg + TG_XQ); });
                    stamp(ts + 1);
                    hop_arrive(hc + HC_XQ * GV_HOP_STRIDE, tid);
                    t_xq += (unsigned)G;
                }
                // ---- ATT: (head, key-range) items on the first n_items CTAs ----
                if (cta < n_items) {
                    const int h = cta / nsplit, sp = cta % nsplit;
                    const int j0 = sp * chunk, j1 = min(S, j0 + chunk);
                    float* kh = kc + (size_t)h * p.S_max * HD;
                    float* vh = vc + (size_t)h * p.S_max * HD;
#define GV_ATT_CASE(hd)                                                                                                     \
    case hd:                                                                                                                \
        att_item<hd>(kh, vh, p.xq, D, h, j0, j1, S, tg + TG_XQ, hc + HC_XQ * GV_HOP_STRIDE, t_xq, att_so, att_sml, tid,     \
                     p.att_o, p.att_ml, cta, tg + TG_AO, tmask);                                                            \
        break;
                    switch (HD) {
                        GV_ATT_CASE(32) GV_ATT_CASE(64) GV_ATT_CASE(128) GV_ATT_CASE(256)
                        default: break;
                    }
#undef GV_ATT_CASE
                    hop_arrive(hc + HC_AO * GV_HOP_STRIDE, tid);
                }
                t_ao += (unsigned)n_items;
                stamp(ts + 2);
                // ---- PROJ: merge attention partials -> o ; x1 = x + o . W_proj + b ----
                {
                    hop_wait(hc + HC_AO * GV_HOP_STRIDE, t_ao, tid, tmask);
                    float o0 = 0.0f, o1 = 0.0f;
                    if (xvalid) {
                        const int h = (2 * tid) / HD, d = (2 * tid) % HD;
                        float M = -INFINITY, den = 0.0f;
                        for (int s2 = 0; s2 < nsplit; ++s2) {  // running merge in split order
                            const int it = h * nsplit + s2;
                            const float2 ml = ld_tagged2(p.att_ml, it * 2, tg + TG_AO, tmask);
                            const float2 ov = ld_tagged2(p.att_o, it * HD + d, tg + TG_AO, tmask);
                            const float Mn = fmaxf(M, ml.x);
                            const float c_old = expf(M - Mn);  // first split: exp(-inf) = 0
                            const float c_new = expf(ml.x - Mn);
                            den = den * c_old + ml.y * c_new;
                            o0 = o0 * c_old + ov.x * c_new;
                            o1 = o1 * c_old + ov.y * c_new;
                            M = Mn;
                        }
                        o0 = o0 / den;
                        o1 = o1 / den;
                        *reinterpret_cast<float2*>(xn + 2 * tid) = make_float2(o0, o1);
                    }
                    stamp(ts + 3);
                    bar_sync(1, MEGA_CONSUMERS);
                    gemv_dot<NXV>(ring, cs, nun[PH_PROJ], xn, warp, lane, [&](int u, float y) {
                        const int col = ubeg[PH_PROJ] + u;
                        st_tagged(p.x1, col, xres0[col] + y, tg + TG_X1);
                    });
                    stamp(ts + 4);
                    hop_arrive(hc + HC_X1 * GV_HOP_STRIDE, tid);
                    t_x1 += (unsigned)G;
                }
                // ---- FC + P2: LN2 -> u = gelu_new(. W_fc + b) (kept in this CTA) -> partial of u . W_proj2 ----
                {
                    hop_wait(hc + HC_X1 * GV_HOP_STRIDE, t_x1, tid, tmask);
                    float x0 = 0.0f, x1 = 0.0f;
                    if (xvalid) {
                        const float2 v = ld_tagged2(p.x1, 2 * tid, tg + TG_X1, tmask);
                        x0 = v.x;
                        x1 = v.y;
                        *reinterpret_cast<float2*>(xres1 + 2 * tid) = v;
                    }
                    stamp(ts + 5);
                    tile_wait(ring, cs);
                    const float* lnp = tile_ptr(ring, cs);
                    ln_pair(x0, x1, xvalid, D, lnp, lnp + D, red, tid);
                    tile_release(ring, cs, lane);
                    if (xvalid) *reinterpret_cast<float2*>(xn + 2 * tid) = make_float2(x0, x1);
                    bar_sync(1, MEGA_CONSUMERS);
                    gemv_dot<NXV>(ring, cs, nun[PH_FC], xn, warp, lane, [&](int u, float y) { us[u] = gelu_new(y); });
                    bar_sync(1, MEGA_CONSUMERS);
                    float acc0 = 0.0f, acc1 = 0.0f;
                    gemv_outer<NXV>(ring, cs, nun[PH_P2], us, tid, lane, acc0, acc1);
                    if (xvalid) st_tagged2(p.pp, cta * D + 2 * tid, acc0, acc1, tg + TG_PP);
                    stamp(ts + 6);
                    hop_arrive(hc + HC_PP * GV_HOP_STRIDE, tid);
                    t_pp += (unsigned)G;
                }
                // ---- RED: x2 = x1 + b + sum over CTAs of the partials (8 outputs per reducer CTA) ----
                if (cta < n_red) {
                    float b2 = 0.0f;
                    if (warp < 8 && lane == 0) b2 = __ldg(p.blob + p.proj2_b_off + (long long)l * p.layer_stride + cta * 8 + warp);
                    hop_wait(hc + HC_PP * GV_HOP_STRIDE, t_pp, tid, tmask);
                    stamp(ts + 7);
                    for (int q = tid; q < G * 4; q += MEGA_CONSUMERS) {
                        const int c = q >> 2, part = q & 3;
                        const float2 v = ld_tagged2(p.pp, c * D + cta * 8 + 2 * part, tg + TG_PP, tmask);
                        *reinterpret_cast<float2*>(gat + c * 8 + 2 * part) = v;
                    }
                    bar_sync(1, MEGA_CONSUMERS);
                    if (warp < 8) {
                        float s = 0.0f;
                        for (int c = lane; c < G; c += 32) s += gat[c * 8 + warp];
                        s = warp_sum(s);
                        if (lane == 0) {
                            const int col = cta * 8 + warp;
                            st_tagged(p.x2, col, (xres1[col] + b2) + s, tg + TG_X2);
                        }
                    }
                    stamp(ts + 8);
                    hop_arrive(hc + HC_X2 * GV_HOP_STRIDE, tid);
                }
                t_x2 += (unsigned)n_red;
            }
            // ---- HEAD: ln_f -> final_norm -> latent z ; logits = z . mel_head^T + b ----
            {
                const uint32_t tg = tbase + (uint32_t)GV_TAGS_PER_LAYER * (uint32_t)p.L;
                const int ts = p.L * GV_TRACE_PER_LAYER;
                hop_wait(hc + HC_X2 * GV_HOP_STRIDE, t_x2, tid, tmask);
                if (xvalid) {
                    const float2 v = ld_tagged2(p.x2, 2 * tid, tg - (uint32_t)GV_TAGS_PER_LAYER + TG_X2, tmask);
                    lat0 = v.x;
                    lat1 = v.y;
                }
                stamp(ts + 0);
                tile_wait(ring, cs);
                const float* lnp = tile_ptr(ring, cs);
                ln_pair(lat0, lat1, xvalid, D, lnp, lnp + D, red, tid);
                bar_sync(1, MEGA_CONSUMERS);  // `red` is reused by the second LayerNorm
                ln_pair(lat0, lat1, xvalid, D, lnp + 2 * D, lnp + 3 * D, red, tid);
                tile_release(ring, cs, lane);
                if (xvalid) *reinterpret_cast<float2*>(xn + 2 * tid) = make_float2(lat0, lat1);
                bar_sync(1, MEGA_CONSUMERS);
                gemv_dot<NXV>(ring, cs, nun[PH_HEAD], xn, warp, lane,
                              [&](int u, float y) { st_tagged(p.lg, ubeg[PH_HEAD] + u, y, tg); });
                stamp(ts + 1);
                hop_arrive(hc + HC_LG * GV_HOP_STRIDE, tid);
                t_lg += (unsigned)G;
                hop_wait(hc + HC_LG * GV_HOP_STRIDE, t_lg, tid, tmask);
                for (int e = 2 * tid; e < p.V; e += 2 * MEGA_CONSUMERS) {
                    if (e + 1 < p.V) {
                        const float2 v = ld_tagged2(p.lg, e, tg, tmask);
                        slog[e] = v.x;
                        slog[e + 1] = v.y;
                    } else {
                        slog[e] = ld_tagged1(p.lg, e, tg, tmask);
                    }
                }
                stamp(ts + 2);
            }
        } else {
            // logits / latent left pending by the prefill (per-op kernels; plain arrays)
            for (int e = tid; e < p.V; e += MEGA_CONSUMERS) slog[e] = ldcg(p.pend_logits + e);
            if (xvalid) {
                const float2 a = ldcg2(p.pend_latent + 2 * tid);
                lat0 = a.x;
                lat1 = a.y;
            }
        }
        bar_sync(1, MEGA_CONSUMERS);  // slog complete; attention / gather scratch (aliasing `keys`) is dead
        // ------------- sample + emit (every CTA computes the same token; warps 0-7 run the chain) -------------
        if (tid < GV_SAMPLE_THREADS) {
            const int t = sample_token([&](int e) { return slog[e]; }, seen, scfg, p.noise ? p.noise + (size_t)i * p.V : nullptr,
                                       p.seed, (uint32_t)n, 0u, keys, fscr, iscr, tid, [] { bar_sync(2, GV_SAMPLE_THREADS); });
            if (tid == 0) ctl[2] = t;
        }
        bar_sync(1, MEGA_CONSUMERS);
        int tok = ctl[2];
        stamp(p.L * GV_TRACE_PER_LAYER + 3);
        if (p.forced) tok = (int)p.forced[i];
        if (!p.ignore_eos && finished) tok = p.stop_token;
        if (cta == 0) {
            if (tid == 0) p.ids_out[i] = tok;
            if (xvalid) *reinterpret_cast<float2*>(p.latents_out + (size_t)i * D + 2 * tid) = make_float2(lat0, lat1);
            if (p.logits_out)
                for (int q = tid; q < p.V; q += MEGA_CONSUMERS) p.logits_out[(size_t)i * p.V + q] = slog[q];
        }
        if (tid == 0) seen[tok] = 1;
        last_tok = tok;
        if (!p.ignore_eos && tok == p.stop_token) finished = 1;
        n += 1;
        emitted += 1;
        bar_sync(1, MEGA_CONSUMERS);  // seen[] update visible to the next step's sampler; slog / keys free
        if (finished || n >= p.max_total) {
            done = 1;
            break;
        }
    }
    // tell the producer to stop (it may be blocked on a full ring or still have copies in flight)
    if (tid == 0) {
        ctl[1] = (int)cs.tiles;
        __threadfence_block();
        ctl[0] = 1;
    }
    if (cta == 0) {
        bar_sync(1, MEGA_CONSUMERS);
        for (int q = tid; q < p.Vpad; q += MEGA_CONSUMERS) p.seen[q] = seen[q];
        if (tid == 0) {
            st->n_emitted = n;
            st->done = done;
            st->has_pending = 0;
            st->finished[0] = finished;
            st->last_tok[0] = last_tok;
            p.status[0] = emitted;
            p.status[1] = done;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host launcher
// ---------------------------------------------------------------------------------------------
size_t mega_smem_bytes(int D, int Vpad) {
    size_t off = (size_t)NSLOT * slot_floats(D) * sizeof(float);
    off = (off + 127) & ~size_t(127);
    off += MEGA_SCRATCH_BYTES;
    off += (size_t)Vpad * sizeof(float) + 3 * (size_t)D * sizeof(float);
    off += 32 * sizeof(uint64_t);
    off += (32 + 32 + 16) * sizeof(float) + 16 * sizeof(int) + 4 * sizeof(int);
    off += Vpad;
    return (off + 15) & ~size_t(15);
}

template <int NXV>
static cudaError_t launch_nxv(const MegaParams& p, int grid, size_t smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(decode_mega_kernel<NXV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    MegaParams pp = p;
    void* args[] = {&pp};
    return cudaLaunchCooperativeKernel((void*)decode_mega_kernel<NXV>, dim3(grid), dim3(MEGA_THREADS), args, smem, st);
}

cudaError_t launch_decode_mega(const MegaParams& p, int grid, cudaStream_t st) {
    const int hd = p.D / p.H;
    if (p.D % 128 || p.D > 1024 || !(hd == 32 || hd == 64 || hd == 128 || hd == 256)) return cudaErrorInvalidValue;
    if ((size_t)grid * 8 * sizeof(float) > 16384 || p.H * 8 > grid) return cudaErrorInvalidValue;
    const size_t smem = mega_smem_bytes(p.D, p.Vpad);
    switch (p.D / 128) {
        case 1: return launch_nxv<1>(p, grid, smem, st);
        case 2: return launch_nxv<2>(p, grid, smem, st);
        case 4: return launch_nxv<4>(p, grid, smem, st);
        case 8: return launch_nxv<8>(p, grid, smem, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace gv
