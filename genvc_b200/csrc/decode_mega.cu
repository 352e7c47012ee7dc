// Fused persistent decode kernel (B = 1): the whole sample()/sample_stream() loop of
// layers/stream_generator.py:809-881 — per step the cached GPT-2 forward of
// layers/gpt_inference.py:92-112 (30 pre-LN blocks, ln_f, final_norm, mel_head) followed by the HF
// sampling chain — runs in ONE cooperative launch for up to n_steps tokens, tokens fed back on-chip.
//
// Why this shape.  A decode step at batch 1 reads every weight exactly once (1.516 GB fp32) and does
// 2 flops per weight: it is an HBM-streaming problem with 30 x 5 serial all-to-all dependencies.  So:
//   * one persistent CTA per SM (cooperative launch), each owning a fixed column slice of every
//     matrix; the slices are pre-packed so each CTA reads one contiguous byte stream (stream_layout.h);
//   * a dedicated producer thread per CTA streams that region HBM -> shared memory with 1-D bulk TMA
//     copies (cp.async.bulk, completion on mbarriers) into a 6 x 32 KB ring.  The weight stream does
//     not depend on activations, so the producer keeps running ahead across phase boundaries and even
//     across tokens: HBM never idles while the consumers exchange activations;
//   * 8 consumer warps do the GEMV out of shared memory in fp32 FMA.  For the K = D matrices a tile
//     is 8 columns and every warp owns ONE column of it: the activation vector sits in 32 registers
//     per lane, a column costs 8 LDS.128 + 32 FFMA per lane and one short shuffle tree — no cross-warp
//     reduction, no block barrier.  The K = 4D matrix (mlp.c_proj) is split over all 256 threads
//     along K with a transposing butterfly + one pass through shared memory.  LayerNorm (one-pass
//     statistics), gelu_new, residual and bias are fused around the GEMVs;
//   * the activation vector between two phases (<= 4096 floats) is exchanged through L2 with a
//     flag-in-data protocol: every float travels as an 8-byte {value, tag} word, the tag being the
//     global phase number.  Writers store their columns and move on; readers spin on the very words
//     they need.  One L2 round trip per phase instead of fence + atomic + poll + load, and no
//     grid-wide barrier anywhere;
//   * single-token attention is split over (head, 32-key range) items; the item's K/V rows are
//     requested from the cache BEFORE the query is polled, so their latency hides behind the QKV
//     exchange; online softmax with warp-shuffle reductions; the newest K/V row comes from the
//     exchange buffer, and the column owners also append it to the cache for later steps;
//   * sampling is computed redundantly by every CTA (same data, same code => same token), so the
//     next token needs no broadcast and the next step's embedding starts without an exchange.
// No tensor cores: at M = 1 there is no reuse to feed them; fp32 keeps greedy parity.
#include "attn_decode.cuh"
#include "common.cuh"
#include "mega.cuh"
#include "sampling.cuh"
#include "stream_layout.h"

namespace gv {

#define MEGA_CONSUMERS 256
// 2 consumer warpgroups + 1 producer warpgroup (one working thread): a full warpgroup so that
// setmaxnreg can hand the producer's registers to the consumers (168 -> 232 per thread)
#define MEGA_THREADS (MEGA_CONSUMERS + 128)
#define MEGA_SPIN_LIMIT (1u << 24)
#define NSLOT GV_MEGA_NSLOT

// exchange tags inside a layer: tag(layer l, buffer b) = tbase + 5 l + b
enum { TG_QKV = 0, TG_ATT = 1, TG_X1 = 2, TG_U = 3, TG_X2 = 4 };

struct ConsumerSync {
    __device__ __forceinline__ void operator()() const { bar_sync(1, MEGA_CONSUMERS); }
};

struct Ring {
    float* slots;
    uint64_t* full;
    uint64_t* empty;
    int slot_floats;
    // debug tile timeline (this thread's row of MegaParams::trace2, or null): tiles [tr_lo, tr_hi)
    unsigned long long* tr2;
    uint32_t tr_lo, tr_hi;
};

// ---------------------------------------------------------------------------------------------
// stream packing (init time): gather the reference-layout matrices into the per-CTA streams
// ---------------------------------------------------------------------------------------------
__global__ void pack_stream_kernel(StreamDims s, int layer, int ph, const float* __restrict__ W,
                                   const float* __restrict__ bias, int w_nk, float* __restrict__ stream) {
    const int N = ph_N(s, ph), K = ph_K(s, ph);
    const int n = blockIdx.y;
    const int c = col_owner(N, n, s.G);
    long long dst = cta_base(s, c);
    if (ph == PH_HEAD) dst += (long long)s.L * cta_layer_floats(s, c);
    else dst += (long long)layer * cta_layer_floats(s, c) + ph_offset_in_layer(s, ph, c);
    dst += (long long)(n - col_begin(N, c, s.G)) * (K + 4);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < K + 4; k += gridDim.x * blockDim.x) {
        float v = 0.0f;
        if (k < K) v = w_nk ? W[(size_t)n * K + k] : W[(size_t)k * N + n];
        else if (k == K) v = bias[n];
        stream[dst + k] = v;
    }
}

cudaError_t launch_pack_stream(const StreamDims& s, int layer, int ph, const float* W, const float* bias, int w_nk,
                               float* stream, cudaStream_t st) {
    const int N = ph_N(s, ph), K = ph_K(s, ph);
    dim3 grid((K + 4 + 255) / 256, N);
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
__device__ __forceinline__ uint4 ld_poll16(const float* p) {
    uint4 r;
    asm volatile("ld.relaxed.gpu.global.v4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p)
                 : "memory");
    return r;
}
__device__ __forceinline__ uint2 ld_poll8(const float* p) {
    uint2 r;
    asm volatile("ld.relaxed.gpu.global.v2.b32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p) : "memory");
    return r;
}
// Spin until the NE consecutive elements starting at `idx` (NE in {1,2,4}; idx % NE == 0) carry `tag`.
template <int NE>
__device__ __forceinline__ void poll_vals(const float* buf, int idx, uint32_t tag, float* out, uint32_t tmask = 0xffffffffu) {
    const float* p = buf + 2 * (size_t)idx;
    uint32_t spins = 0;
    if constexpr (NE == 1) {
        uint2 a;
        while (true) {
            a = ld_poll8(p);
            if (((a.y ^ tag) & tmask) == 0u) break;
            if (++spins > MEGA_SPIN_LIMIT) __trap();
        }
        out[0] = __uint_as_float(a.x);
    } else if constexpr (NE == 2) {
        uint4 a;
        while (true) {
            a = ld_poll16(p);
            if ((((a.y ^ tag) | (a.w ^ tag)) & tmask) == 0u) break;
            if (++spins > MEGA_SPIN_LIMIT) __trap();
        }
        out[0] = __uint_as_float(a.x);
        out[1] = __uint_as_float(a.z);
    } else {
        static_assert(NE == 4, "NE must be 1, 2 or 4");
        uint4 a, b;
        while (true) {
            a = ld_poll16(p);
            b = ld_poll16(p + 4);
            if ((((a.y ^ tag) | (a.w ^ tag) | (b.y ^ tag) | (b.w ^ tag)) & tmask) == 0u) break;
            if (++spins > MEGA_SPIN_LIMIT) __trap();
        }
        out[0] = __uint_as_float(a.x);
        out[1] = __uint_as_float(a.z);
        out[2] = __uint_as_float(b.x);
        out[3] = __uint_as_float(b.z);
    }
}

// ---------------------------------------------------------------------------------------------
// weight ring (consumer side)
// ---------------------------------------------------------------------------------------------
struct ConsumerState {
    uint32_t slot, phase;  // ring position of the next tile
    uint32_t tiles;        // tiles consumed so far (same in every consumer thread)
    int flip;              // double-buffer index of the statistics scratch
    int xflip;             // double-buffer index of the GEMV input vector in shared memory
};

__device__ __forceinline__ const float* tile_acquire(const Ring& r, const ConsumerState& cs) {
    const bool tr = r.tr2 != nullptr && cs.tiles >= r.tr_lo && cs.tiles < r.tr_hi;
    if (tr) r.tr2[(cs.tiles - r.tr_lo) * 3 + 1] = globaltimer_ns();
    mbar_wait(&r.full[cs.slot], cs.phase);
    if (tr) r.tr2[(cs.tiles - r.tr_lo) * 3 + 2] = globaltimer_ns();
    return r.slots + (size_t)cs.slot * r.slot_floats;
}
__device__ __forceinline__ void tile_release(const Ring& r, ConsumerState& cs, int lane) {
    __syncwarp();
    if (lane == 0) mbar_arrive(&r.empty[cs.slot]);
    cs.tiles += 1;
    if (++cs.slot == NSLOT) {
        cs.slot = 0;
        cs.phase ^= 1u;
    }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm of the register-resident vector (thread owns x[4*tid .. 4*tid+3]); weight/bias in smem.
// One-pass statistics (sum, sum of squares; the final E[x^2] - mean^2 in double), one block barrier.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void ln_regs(float (&x)[4], bool valid, int D, const float* w, const float* b, float* scratch,
                                        int& flip, int tid) {
    float s1 = 0.0f, s2 = 0.0f;
    if (valid) {
        s1 = (x[0] + x[1]) + (x[2] + x[3]);
        s2 = fmaf(x[0], x[0], fmaf(x[1], x[1], fmaf(x[2], x[2], x[3] * x[3])));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    float* s = scratch + flip * 16;
    flip ^= 1;
    if ((tid & 31) == 0) {
        s[tid >> 5] = s1;
        s[8 + (tid >> 5)] = s2;
    }
    bar_sync(1, MEGA_CONSUMERS);
    const float4 a0 = *reinterpret_cast<const float4*>(s), a1 = *reinterpret_cast<const float4*>(s + 4);
    const float4 q0 = *reinterpret_cast<const float4*>(s + 8), q1 = *reinterpret_cast<const float4*>(s + 12);
    const float t1 = ((a0.x + a0.y) + (a0.z + a0.w)) + ((a1.x + a1.y) + (a1.z + a1.w));
    const float t2 = ((q0.x + q0.y) + (q0.z + q0.w)) + ((q1.x + q1.y) + (q1.z + q1.w));
    const double inv = 1.0 / (double)D;
    const double meand = (double)t1 * inv;
    double vard = (double)t2 * inv - meand * meand;
    if (vard < 0.0) vard = 0.0;
    const float mean = (float)meand;
    const float rstd = 1.0f / sqrtf((float)vard + 1e-5f);
    if (valid) {
        const float4 ww = *reinterpret_cast<const float4*>(w + 4 * tid);
        const float4 bb = *reinterpret_cast<const float4*>(b + 4 * tid);
        x[0] = (x[0] - mean) * rstd * ww.x + bb.x;
        x[1] = (x[1] - mean) * rstd * ww.y + bb.y;
        x[2] = (x[2] - mean) * rstd * ww.z + bb.z;
        x[3] = (x[3] - mean) * rstd * ww.w + bb.w;
    }
}

// ---------------------------------------------------------------------------------------------
// GEMV, K = D: tiles of 8 columns, warp w owns column w of every tile.  `xs` is the activation
// vector in shared memory (already complete: the caller synchronised).  For every column this CTA
// owns, lane 0/8/16/24 of the owning warp calls epi(local column index, y) with
//   y = bias + sum_k x_k W[k][col].
// FULL: D == 1024 (8 float4 of x per lane, no predicates).
// ---------------------------------------------------------------------------------------------
template <bool FULL, class Epi>
__device__ __forceinline__ void gemv_cols(const Ring& ring, ConsumerState& cs, int ncols, int D, const float* xs, int warp,
                                          int lane, Epi epi) {
    const int nxv = FULL ? 8 : D / 128;  // float4 chunks of x per lane
    const int ntiles = (ncols + 7) >> 3;
    float4 xv[8];
    if (warp < ncols) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (FULL || i < nxv) xv[i] = *reinterpret_cast<const float4*>(xs + (i * 32 + lane) * 4);
    }
    float tot[4] = {0.f, 0.f, 0.f, 0.f};
    const int cstride = D + 4;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        if (t < ntiles) {
            const float* w = tile_acquire(ring, cs);
            if (t * 8 + warp < ncols) {
                const float* col = w + warp * cstride;
                float a0 = (lane == 0) ? col[D] : 0.0f;  // bias folded into the first partial
                float a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (FULL || i < nxv) {
                        const float4 wv = *reinterpret_cast<const float4*>(col + (i * 32 + lane) * 4);
                        a0 = fmaf(wv.x, xv[i].x, a0);
                        a1 = fmaf(wv.y, xv[i].y, a1);
                        a2 = fmaf(wv.z, xv[i].z, a2);
                        a3 = fmaf(wv.w, xv[i].w, a3);
                    }
                }
                tot[t] = (a0 + a1) + (a2 + a3);
            }
            tile_release(ring, cs, lane);
        }
    }
    if (warp >= ncols) return;  // this warp owns no column of this phase
    // 4 per-lane partials -> lane L holds the warp total of column (L >> 3)
    {
        const bool up = (lane & 16) != 0;
        const float k0 = up ? tot[2] : tot[0], s0 = up ? tot[0] : tot[2];
        const float k1 = up ? tot[3] : tot[1], s1 = up ? tot[1] : tot[3];
        tot[0] = k0 + __shfl_xor_sync(0xffffffffu, s0, 16);
        tot[1] = k1 + __shfl_xor_sync(0xffffffffu, s1, 16);
    }
    float v;
    {
        const bool up = (lane & 8) != 0;
        const float k0 = up ? tot[1] : tot[0], s0 = up ? tot[0] : tot[1];
        v = k0 + __shfl_xor_sync(0xffffffffu, s0, 8);
    }
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    const int c = (lane >> 3) * 8 + warp;
    if ((lane & 7) == 0 && c < ncols) epi(c, v);
}

// Transposing warp reduction of 8 per-lane partial sums: lane L gets the warp total of element L >> 2.
__device__ __forceinline__ float warp_reduce8(float (&r)[8], int lane) {
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        const int n = 8 >> s, off = 16 >> s;
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float keep = upper ? r[i + n / 2] : r[i];
            const float send = upper ? r[i] : r[i + n / 2];
            r[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    float v = r[0];
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// GEMV, K = 4D (mlp.c_proj): tiles of 2 columns; K is split over the 256 threads (thread t owns
// k = 4 (t + 256 v) .. +3, v < 4, held in ur[]).  After the call thread j (< ncols <= 8) holds y_j.
template <bool FULL>
__device__ __forceinline__ float gemv_k4(const Ring& ring, ConsumerState& cs, int ncols, int K, const float (&ur)[16],
                                         float* red, int tid) {
    const int lane = tid & 31, warp = tid >> 5;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
    const int ntiles = (ncols + 1) >> 1;
    const int cstride = K + 4;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        if (t < ntiles) {
            const float* w = tile_acquire(ring, cs);
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                if (t * 2 + cc < ncols) {
                    const float* col = w + cc * cstride;
                    float a[4];
                    a[0] = (tid == 0) ? col[K] : 0.0f;  // bias
                    a[1] = a[2] = a[3] = 0.0f;
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        const int k = (tid + MEGA_CONSUMERS * v) * 4;
                        if (FULL || k < K) {
                            const float4 wv = *reinterpret_cast<const float4*>(col + k);
                            a[v] = fmaf(wv.x, ur[v * 4 + 0], a[v]);
                            a[v] = fmaf(wv.y, ur[v * 4 + 1], a[v]);
                            a[v] = fmaf(wv.z, ur[v * 4 + 2], a[v]);
                            a[v] = fmaf(wv.w, ur[v * 4 + 3], a[v]);
                        }
                    }
                    acc[t * 2 + cc] = (a[0] + a[1]) + (a[2] + a[3]);
                }
            }
            tile_release(ring, cs, lane);
        }
    }
    const float v = warp_reduce8(acc, lane);
    if ((lane & 3) == 0) red[warp * 8 + (lane >> 2)] = v;
    bar_sync(1, MEGA_CONSUMERS);
    float y = 0.0f;
    if (tid < ncols) {
        const float* r = red + tid;
        y = ((r[0] + r[8]) + (r[16] + r[24])) + ((r[32] + r[40]) + (r[48] + r[56]));
    }
    return y;
}

// ---------------------------------------------------------------------------------------------
// single-query attention over one (head, key range) item; see attn_decode.cuh for the arithmetic.
// K/V rows of the cache are requested first, then the query (and, for the newest position, the
// new k/v row) is polled from the exchange buffer `xq` = tagged [q | k | v] of this step.
// ---------------------------------------------------------------------------------------------
template <int HD>
__device__ void att_item(const float* __restrict__ Kc, const float* __restrict__ Vc, const float* xq, int D, int h, int j0,
                         int j1, int S, uint32_t tag_in, float sqrt_hd, float* so, float* sml, int tid, float* o_out,
                         float* ml_out, int item, uint32_t tag_out, uint32_t tmask) {
    using L = AttLane<HD>;
    constexpr int VEC = L::VEC, NCH = L::NCH, DPL = L::DPL, UNR = 4;
    const int warp = tid >> 5, lane = tid & 31;
    float m = -INFINITY, l = 0.0f;
    float o[DPL], qr[DPL];
#pragma unroll
    for (int i = 0; i < DPL; ++i) {
        o[i] = 0.0f;
        qr[i] = 0.0f;
    }
    bool have_q = false;
    for (int jb = j0 + warp * UNR; jb < j1; jb += GV_ATT_WARPS * UNR) {
        float kr[UNR][DPL], vr[UNR][DPL];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int j = jb + u;
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
        if (!have_q) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) poll_vals<VEC>(xq, h * HD + (c * 32 + lane) * VEC, tag_in, qr + c * VEC, tmask);
            have_q = true;
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            if (jb + u == S - 1 && jb + u < j1) {  // the position being decoded: k/v of this very step
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    poll_vals<VEC>(xq, D + h * HD + (c * 32 + lane) * VEC, tag_in, kr[u] + c * VEC, tmask);
                    poll_vals<VEC>(xq, 2 * D + h * HD + (c * 32 + lane) * VEC, tag_in, vr[u] + c * VEC, tmask);
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
    // merge the 8 warp states: so [warp][HD], sml [warp][2]
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int v = 0; v < VEC; ++v) so[warp * HD + (c * 32 + lane) * VEC + v] = o[c * VEC + v];
    if (lane == 0) {
        sml[warp * 2] = m;
        sml[warp * 2 + 1] = l;
    }
    bar_sync(1, MEGA_CONSUMERS);
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
    for (int d = tid; d < HD; d += MEGA_CONSUMERS) {
        float acc = 0.0f;
#pragma unroll
        for (int w = 0; w < GV_ATT_WARPS; ++w) acc = fmaf(so[w * HD + d], wgt[w], acc);
        st_tagged(o_out, item * HD + d, acc, tag_out);
    }
    if (tid == 0) {
        st_tagged(ml_out, item * 2, M, tag_out);
        st_tagged(ml_out, item * 2 + 1, Lsum, tag_out);
    }
    bar_sync(1, MEGA_CONSUMERS);  // so / sml are reused by the next item
}

// ---------------------------------------------------------------------------------------------
// producer: one thread walks the CTA's weight stream through the ring
// ---------------------------------------------------------------------------------------------
struct Producer {
    const Ring& ring;
    uint32_t t = 0, slot = 0, phase = 0;
    uint32_t window;  // at most this many tiles requested but not landed (bounds the queueing delay the
                      // bulk requests impose on this SM's latency-critical exchange loads/stores)
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
        if (ring.tr2 != nullptr && t >= ring.tr_lo && t < ring.tr_hi) ring.tr2[(t - ring.tr_lo) * 3] = globaltimer_ns();
        ++t;
        if (++slot == NSLOT) {
            slot = 0;
            phase ^= 1u;
        }
        return true;
    }
};

__device__ bool produce_forward(Producer& pr, const MegaParams& p, const StreamDims& sd, int cta) {
    const float* base = p.stream + cta_base(sd, cta);
    const long long lfl = cta_layer_floats(sd, cta);
    const int D = p.D;
    int ncol[5];
    for (int ph = 0; ph < 5; ++ph) ncol[ph] = ph_cols(sd, ph, cta);
    for (int l = 0; l < p.L; ++l) {
        const float* lw = base + (long long)l * lfl;
        for (int ph = PH_QKV; ph <= PH_PROJ2; ++ph) {
            if (ph == PH_QKV) {
                if (!pr.issue(p.blob + p.ln1_off + (long long)l * p.layer_stride, 2 * D, false)) return false;
            } else if (ph == PH_FC) {
                if (!pr.issue(p.blob + p.ln2_off + (long long)l * p.layer_stride, 2 * D, false)) return false;
            }
            const int ct = tile_cols(ph), cstride = ph_K(sd, ph) + 4;
            for (int c0 = 0; c0 < ncol[ph]; c0 += ct) {
                const int nc = min(ct, ncol[ph] - c0);
                if (!pr.issue(lw, (uint32_t)(nc * cstride), true)) return false;
                lw += (long long)nc * cstride;
            }
        }
    }
    if (!pr.issue(p.blob + p.lnf_off, 4 * D, false)) return false;
    const float* hw = base + (long long)p.L * lfl;
    const int ct = tile_cols(PH_HEAD);
    for (int c0 = 0; c0 < ncol[PH_HEAD]; c0 += ct) {
        const int nc = min(ct, ncol[PH_HEAD] - c0);
        if (!pr.issue(hw, (uint32_t)(nc * (D + 4)), true)) return false;
        hw += (long long)nc * (D + 4);
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <int HD, bool FULL>
__global__ void __launch_bounds__(MEGA_THREADS, 1) decode_mega_kernel(MegaParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid_all = threadIdx.x;
    const int cta = blockIdx.x;
    const int G = gridDim.x;
    const StreamDims sd{p.L, p.D, p.V, G};
    const int D = FULL ? 1024 : p.D;

    // ---- shared memory carve-up (mirrored by mega_smem_bytes) ----
    Ring ring;
    ring.slot_floats = slot_floats(D);
    size_t off = 0;
    ring.slots = reinterpret_cast<float*>(smem_raw);
    off += (size_t)NSLOT * ring.slot_floats * sizeof(float);
    off = (off + 127) & ~size_t(127);
    // 16 KB region: sampling sort keys, aliased (outside sampling) by the attention merge buffer and
    // the two residual-stream stashes
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw + off);
    float* att_so = reinterpret_cast<float*>(smem_raw + off);           // [8][HD]      (<= 8 KB)
    float* xs_a = reinterpret_cast<float*>(smem_raw + off + 8192);      // [D] residual stream entering the block
    float* xs_b = reinterpret_cast<float*>(smem_raw + off + 12288);     // [D] residual stream after attention
    off += GV_SORT_N * sizeof(unsigned long long);
    float* slog = reinterpret_cast<float*>(smem_raw + off);  // [Vpad] logits of the step being sampled
    off += (size_t)p.Vpad * sizeof(float);
    float* xn = reinterpret_cast<float*>(smem_raw + off);  // [2][D] GEMV input vector (double-buffered)
    off += 2 * (size_t)D * sizeof(float);
    ring.full = reinterpret_cast<uint64_t*>(smem_raw + off);
    off += 8 * sizeof(uint64_t);
    ring.empty = reinterpret_cast<uint64_t*>(smem_raw + off);
    off += 8 * sizeof(uint64_t);
    float* red = reinterpret_cast<float*>(smem_raw + off);  // [8][8]
    off += 64 * sizeof(float);
    float* scratch = reinterpret_cast<float*>(smem_raw + off);  // [2][16] LayerNorm statistics
    off += 32 * sizeof(float);
    float* att_sml = reinterpret_cast<float*>(smem_raw + off);  // [8][2]
    off += 16 * sizeof(float);
    float* fscr = reinterpret_cast<float*>(smem_raw + off);
    off += 16 * sizeof(float);
    int* iscr = reinterpret_cast<int*>(smem_raw + off);
    off += 16 * sizeof(int);
    volatile int* ctl = reinterpret_cast<volatile int*>(smem_raw + off);  // [0] stop flag, [1] tiles consumed
    off += 4 * sizeof(int);
    unsigned char* seen = smem_raw + off;  // [Vpad]

    ring.tr2 = nullptr;
    ring.tr_lo = ring.tr_hi = 0;
    if (p.trace2 != nullptr && (tid_all == 0 || tid_all == MEGA_CONSUMERS)) {
        int per_layer = 2, head = 1;
        for (int ph = PH_QKV; ph <= PH_PROJ2; ++ph) per_layer += (ph_cols(sd, ph, cta) + tile_cols(ph) - 1) / tile_cols(ph);
        head += (ph_cols(sd, PH_HEAD, cta) + tile_cols(PH_HEAD) - 1) / tile_cols(PH_HEAD);
        const int fwd_idx = p.trace_step - (p.st->has_pending ? 1 : 0);
        if (fwd_idx >= 0) {
            ring.tr2 = p.trace2 + (size_t)cta * GV_TRACE2_TILES * 3;
            ring.tr_lo = (uint32_t)(fwd_idx * (p.L * per_layer + head) + min(10, p.L - 1) * per_layer);
            ring.tr_hi = ring.tr_lo + (uint32_t)min(GV_TRACE2_TILES, 2 * per_layer);
        }
    }
    if (tid_all == 0) {
        for (int i = 0; i < NSLOT; ++i) {
            mbar_init(&ring.full[i], 1);
            mbar_init(&ring.empty[i], MEGA_CONSUMERS / 32);
        }
        ctl[0] = 0;
        ctl[1] = 0;
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
        // ================= producer warpgroup =================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
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
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int tid = tid_all;
    const int lane = tid & 31, warp = tid >> 5;
    ConsumerState cs{0u, 0u, 0u, 0, 0};
    const uint32_t tmask = p.dbg_nosync ? 0u : 0xffffffffu;  // debug: 0 = do not wait for exchange data
    const bool xvalid = 4 * tid < D;
    const int H = p.H;
    const float sqrt_hd = sqrtf((float)HD);
    int ncol[5];
    int cbeg[5];
    for (int ph = 0; ph < 5; ++ph) {
        ncol[ph] = ph_cols(sd, ph, cta);
        cbeg[ph] = (int)col_begin(ph_N(sd, ph), cta, G);
    }
    const SampleCfg scfg{p.V, p.top_k, p.top_p, p.top_p_threshold, p.temperature, p.rep_penalty};
    int n = n_start;  // tokens emitted so far
    long long last_tok = st->last_tok[0];
    int finished = st->finished[0];
    int emitted = 0, done = 0;
    uint32_t fwd = 0;  // forwards executed by this launch
    const uint32_t phases_per_fwd = 5u * (uint32_t)p.L + 1u;
    const float* mel_emb = p.blob + p.mel_emb_off;
    const float* mel_pos = p.blob + p.mel_pos_off;

    // publish the GEMV input vector (this thread's 4 elements) and return the buffer all warps read
    auto publish = [&](const float (&v)[4]) -> const float* {
        float* buf = xn + (size_t)cs.xflip * D;
        cs.xflip ^= 1;
        if (xvalid) *reinterpret_cast<float4*>(buf + 4 * tid) = make_float4(v[0], v[1], v[2], v[3]);
        bar_sync(1, MEGA_CONSUMERS);
        return buf;
    };

    for (int i = 0; i < p.n_steps; ++i) {
        const bool tr = p.trace != nullptr && i == p.trace_step && tid == 0;
        unsigned long long* trow = p.trace + (size_t)cta * p.trace_slots;
        auto stamp = [&](int slot) {
            if (tr && slot < p.trace_slots) trow[slot] = globaltimer_ns();
        };
        stamp(p.L * 10 + 3);
        float lat[4] = {0.f, 0.f, 0.f, 0.f};  // final_norm(ln_f(x)): the latent of this step
        if (!(i == 0 && had_pending)) {
            // ------------- forward of token `last_tok` at mel position n, cache row P + n -------------
            const uint32_t tbase = p.tag0 + fwd * phases_per_fwd;  // tag of (layer l, phase ph) = tbase + 5 l + ph
            fwd += 1;
            const int pos = p.P + n;
            const int S = pos + 1;
            const int nsplit = min((S + 31) / 32, max(1, G / H));
            const int chunk = (S + nsplit - 1) / nsplit;
            for (int l = 0; l < p.L; ++l) {
                float* kc = p.kv + ((size_t)l * 2 + 0) * p.kv_layer_stride;
                float* vc = p.kv + ((size_t)l * 2 + 1) * p.kv_layer_stride;
                const uint32_t tg = tbase + 5u * (uint32_t)l;
                // ---- QKV: LN1 -> [q|k|v] columns ----
                {
                    float xr[4] = {0.f, 0.f, 0.f, 0.f};
                    if (xvalid) {
                        if (l == 0) {
                            const float4 a = *reinterpret_cast<const float4*>(mel_emb + (size_t)last_tok * D + 4 * tid);
                            const float4 b = *reinterpret_cast<const float4*>(mel_pos + (size_t)n * D + 4 * tid);
                            xr[0] = a.x + b.x; xr[1] = a.y + b.y; xr[2] = a.z + b.z; xr[3] = a.w + b.w;
                        } else {
                            poll_vals<4>(p.x2, 4 * tid, tg - 1u, xr, tmask);
                        }
                        *reinterpret_cast<float4*>(xs_a + 4 * tid) = make_float4(xr[0], xr[1], xr[2], xr[3]);
                    }
                    if (l > 0) stamp((l - 1) * 10 + 9);
                    const float* lnp = tile_acquire(ring, cs);
                    ln_regs(xr, xvalid, D, lnp, lnp + D, scratch, cs.flip, tid);
                    tile_release(ring, cs, lane);
                    const float* xin = publish(xr);
                    gemv_cols<FULL>(ring, cs, ncol[PH_QKV], D, xin, warp, lane, [&](int c, float y) {
                        const int ncolg = cbeg[PH_QKV] + c;
                        st_tagged(p.xq, ncolg, y, tg + TG_QKV);
                        if (ncolg >= D) {  // append to the cache for the following steps
                            const int c2 = (ncolg - D) % D;
                            float* dstc = (ncolg < 2 * D) ? kc : vc;
                            dstc[((size_t)(c2 / HD) * p.S_max + pos) * HD + (c2 % HD)] = y;
                            __threadfence();  // ordered before this CTA's later exchange stores
                        }
                    });
                    stamp(l * 10 + 0);
                }
                // ---- ATT: (head, key-range) items ----
                stamp(l * 10 + 1);
                for (int item = cta; item < H * nsplit; item += G) {
                    const int h = item / nsplit, sp = item % nsplit;
                    const int j0 = sp * chunk, j1 = min(S, j0 + chunk);
                    att_item<HD>(kc + (size_t)h * p.S_max * HD, vc + (size_t)h * p.S_max * HD, p.xq, D, h, j0, j1, S,
                                 tg + TG_QKV, sqrt_hd, att_so, att_sml, tid, p.att_o, p.att_ml, item, tg + TG_ATT, tmask);
                }
                stamp(l * 10 + 2);
                // ---- PROJ: merge attention partials -> o ; x1 = x + o . W_proj + b ----
                {
                    float xr[4] = {0.f, 0.f, 0.f, 0.f};
                    if (xvalid) {
                        const int h = (4 * tid) / HD, d = (4 * tid) % HD;
                        float M = -INFINITY, den = 0.0f;
                        for (int s2 = 0; s2 < nsplit; ++s2) {  // running merge in split order
                            const int it = h * nsplit + s2;
                            float ml[2], ov[4];
                            poll_vals<2>(p.att_ml, it * 2, tg + TG_ATT, ml, tmask);
                            poll_vals<4>(p.att_o, it * HD + d, tg + TG_ATT, ov, tmask);
                            const float Mn = fmaxf(M, ml[0]);
                            const float c_old = expf(M - Mn);  // first split: exp(-inf) = 0
                            const float c_new = expf(ml[0] - Mn);
                            den = den * c_old + ml[1] * c_new;
                            xr[0] = xr[0] * c_old + ov[0] * c_new;
                            xr[1] = xr[1] * c_old + ov[1] * c_new;
                            xr[2] = xr[2] * c_old + ov[2] * c_new;
                            xr[3] = xr[3] * c_old + ov[3] * c_new;
                            M = Mn;
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q) xr[q] = xr[q] / den;
                    }
                    stamp(l * 10 + 3);
                    const float* xin = publish(xr);
                    gemv_cols<FULL>(ring, cs, ncol[PH_PROJ], D, xin, warp, lane, [&](int c, float y) {
                        const int ncolg = cbeg[PH_PROJ] + c;
                        st_tagged(p.x1, ncolg, xs_a[ncolg] + y, tg + TG_X1);
                    });
                    stamp(l * 10 + 4);
                }
                // ---- FC: LN2 -> u = gelu_new(. W_fc + b) ----
                {
                    float xr[4] = {0.f, 0.f, 0.f, 0.f};
                    if (xvalid) {
                        poll_vals<4>(p.x1, 4 * tid, tg + TG_X1, xr, tmask);
                        *reinterpret_cast<float4*>(xs_b + 4 * tid) = make_float4(xr[0], xr[1], xr[2], xr[3]);
                    }
                    stamp(l * 10 + 5);
                    const float* lnp = tile_acquire(ring, cs);
                    ln_regs(xr, xvalid, D, lnp, lnp + D, scratch, cs.flip, tid);
                    tile_release(ring, cs, lane);
                    const float* xin = publish(xr);
                    gemv_cols<FULL>(ring, cs, ncol[PH_FC], D, xin, warp, lane, [&](int c, float y) {
                        st_tagged(p.u, cbeg[PH_FC] + c, gelu_new(y), tg + TG_U);
                    });
                    stamp(l * 10 + 6);
                }
                // ---- PROJ2: x2 = x1 + u . W_proj2 + b ----
                {
                    float ur[16];
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        const int k = (tid + MEGA_CONSUMERS * v) * 4;
                        if (FULL || k < 4 * D) {
                            poll_vals<4>(p.u, k, tg + TG_U, ur + 4 * v, tmask);
                        } else {
                            ur[v * 4 + 0] = 0.f; ur[v * 4 + 1] = 0.f; ur[v * 4 + 2] = 0.f; ur[v * 4 + 3] = 0.f;
                        }
                    }
                    stamp(l * 10 + 7);
                    const float y = gemv_k4<FULL>(ring, cs, ncol[PH_PROJ2], 4 * D, ur, red, tid);
                    if (tid < ncol[PH_PROJ2]) {
                        const int ncolg = cbeg[PH_PROJ2] + tid;
                        st_tagged(p.x2, ncolg, xs_b[ncolg] + y, tg + TG_X2);
                    }
                    stamp(l * 10 + 8);
                }
            }
            // ---- HEAD: ln_f -> final_norm -> latent z ; logits = z . mel_head^T + b ----
            {
                const uint32_t tg = tbase + 5u * (uint32_t)p.L;
                if (xvalid) poll_vals<4>(p.x2, 4 * tid, tg - 1u, lat, tmask);
                stamp((p.L - 1) * 10 + 9);
                const float* lnp = tile_acquire(ring, cs);
                ln_regs(lat, xvalid, D, lnp, lnp + D, scratch, cs.flip, tid);
                ln_regs(lat, xvalid, D, lnp + 2 * D, lnp + 3 * D, scratch, cs.flip, tid);
                tile_release(ring, cs, lane);
                const float* xin = publish(lat);
                gemv_cols<FULL>(ring, cs, ncol[PH_HEAD], D, xin, warp, lane,
                                [&](int c, float y) { st_tagged(p.lg, cbeg[PH_HEAD] + c, y, tg); });
                stamp(p.L * 10 + 0);
                for (int e = tid; e < p.V; e += MEGA_CONSUMERS) poll_vals<1>(p.lg, e, tg, slog + e, tmask);
            }
        } else {
            // logits / latent left pending by the prefill (per-op kernels; plain arrays)
            for (int e = tid; e < p.V; e += MEGA_CONSUMERS) slog[e] = ldcg(p.pend_logits + e);
            if (xvalid) {
                const float4 a = ldcg4(p.pend_latent + 4 * tid);
                lat[0] = a.x; lat[1] = a.y; lat[2] = a.z; lat[3] = a.w;
            }
        }
        bar_sync(1, MEGA_CONSUMERS);  // slog complete; stashes / attention scratch (aliasing `keys`) are dead
        stamp(p.L * 10 + 1);
        // ------------- sample + emit (every CTA computes the same token) -------------
        int tok = sample_token([&](int e) { return slog[e]; }, seen, scfg, p.noise ? p.noise + (size_t)i * p.V : nullptr,
                               p.seed, (uint32_t)n, 0u, keys, fscr, iscr, tid, ConsumerSync());
        stamp(p.L * 10 + 2);
        if (p.forced) tok = (int)p.forced[i];
        if (!p.ignore_eos && finished) tok = p.stop_token;
        if (cta == 0) {
            if (tid == 0) p.ids_out[i] = tok;
            if (xvalid)
                *reinterpret_cast<float4*>(p.latents_out + (size_t)i * D + 4 * tid) = make_float4(lat[0], lat[1], lat[2], lat[3]);
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
    off += GV_SORT_N * sizeof(unsigned long long);
    off += (size_t)Vpad * sizeof(float) + 2 * (size_t)D * sizeof(float);
    off += 16 * sizeof(uint64_t);
    off += (64 + 32 + 16 + 16) * sizeof(float) + 16 * sizeof(int) + 4 * sizeof(int);
    off += Vpad;
    return (off + 15) & ~size_t(15);
}

template <int HD, bool FULL>
static cudaError_t launch_hd(const MegaParams& p, int grid, size_t smem, cudaStream_t st) {
    cudaError_t e =
        cudaFuncSetAttribute(decode_mega_kernel<HD, FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    MegaParams pp = p;
    void* args[] = {&pp};
    return cudaLaunchCooperativeKernel((void*)decode_mega_kernel<HD, FULL>, dim3(grid), dim3(MEGA_THREADS), args, smem, st);
}

cudaError_t launch_decode_mega(const MegaParams& p, int grid, cudaStream_t st) {
    if (p.D % 128 || p.D > 1024 || p.D / p.H > 1024) return cudaErrorInvalidValue;
    const size_t smem = mega_smem_bytes(p.D, p.Vpad);
    const bool full = p.D == 1024;
    switch (p.D / p.H) {
#define GV_MEGA_CASE(hd) \
    case hd: return full ? launch_hd<hd, true>(p, grid, smem, st) : launch_hd<hd, false>(p, grid, smem, st);
        GV_MEGA_CASE(32) GV_MEGA_CASE(64) GV_MEGA_CASE(128) GV_MEGA_CASE(256)
#undef GV_MEGA_CASE
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace gv
