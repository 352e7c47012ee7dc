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
//     (cp.async.bulk, completion on mbarriers) into a 12 x 16 KB ring.  The weight stream does
//     not depend on activations, so it runs ahead across phase boundaries and across tokens: HBM
//     does not idle while the consumers exchange activations.  At most `window` tiles are in
//     flight per SM (enough to cover the bandwidth-delay product, few enough not to queue in front
//     of the latency-critical exchange traffic);
//   * 8 consumer warps (two warpgroups at 232 registers: the producer warpgroup hands its registers over with
//     setmaxnreg) do the GEMVs out of shared memory in fp32 FMA.  K = D matrices (QKV, attn
//     proj, FC, logits head): a unit is one output column, one warp per unit, the activation vector
//     in 32 registers per lane, 8 LDS.128 + 32 FFMA per lane per unit and one shuffle tree — no
//     cross-warp reduction.  mlp.c_proj (K = 4D) is split along K instead: the CTA that computed
//     u_k owns ROW k of W_proj2 and accumulates u_k * W[k, :] into a D-wide partial (a thread owns
//     eight outputs per warp group), so the 4D-wide activation never crosses CTAs; the G
//     partials are summed in a fixed order by D/8 reducer CTAs (deterministic, no float atomics);
//   * hops go through L2: producers store {value, tag} words (tag = global hop number) and bump an
//     arrival counter (one relaxed red per CTA, no fence, no barrier); ONE thread per CTA spins on the
//     counter and releases the CTA when all but a few arrivals are in; the threads load the words they
//     need and spin on those whose tag is stale.  The counter is only a hint — validity comes from the tags;
//   * single-token attention is split over (head, key range) items run by the first H*nsplit CTAs:
//     the item's first K/V rows are requested from the cache BEFORE q is polled, online softmax per
//     warp with warp-shuffle reductions, 8 warp states merged through shared memory; the item that
//     covers the newest position takes k/v from the exchange buffer and appends them to the cache;
//   * sampling is computed redundantly by every CTA (same data, same code => same token), so the
//     next token needs no broadcast.
// No tensor cores: at M = 1 there is no reuse to feed them; fp32 keeps greedy parity.
#include "common.cuh"
#include "mega.cuh"
#include "sampling.cuh"
#include "stream_layout.h"

namespace gv {

#define MEGA_WARPS 8
#define MEGA_CONSUMERS (MEGA_WARPS * 32)
#define MEGA_THREADS (MEGA_CONSUMERS + 128)  // + one producer warpgroup (one working thread): register reallocation is per warpgroup
#define MEGA_SPIN_LIMIT (1u << 26)
#define NSLOT GV_MEGA_NSLOT
#define UPT GV_MEGA_UPT
// scratch region (never live together): attention per-warp (max, sum) + PV partials (1 KB + 8 * hd floats <= 9 KB) |
// mlp.c_proj group partials [2][D] floats | partial-sum gather [G][8] floats | sampling sort keys [GV_SORT_N] u64 = 16 KB.
// The sort keys run over into the two residual vectors that follow the scratch region (dead while a token is sampled),
// which is what makes room for a twelfth ring slot.
__host__ __device__ inline size_t mega_scratch_bytes(int D) {
    const size_t spill = 2 * (size_t)D * sizeof(float);                    // xres0 + xres1
    const size_t need = (size_t)GV_SORT_N * 8 > spill ? (size_t)GV_SORT_N * 8 - spill : 0;
    return need > 9728 ? need : 9728;
}

enum { TG_XQ = 0, TG_AO = 1, TG_X1 = 2, TG_PP = 3, TG_X2 = 4 };

struct Ring {
    float* slots;
    uint64_t* full;
    uint64_t* empty;
    uint32_t* landed;  // number of leading tiles the producer has SEEN complete (monotonic hint for the consumers)
    int slot_floats;
};
__device__ __forceinline__ uint32_t ld_acquire_cta_shared(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_cta_shared(uint32_t* p, uint32_t v) {
    asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Tile `idx` is ready: either the producer already saw its full barrier complete (one shared-memory load,
// the common case: tiles are requested microseconds ahead) or this thread waits on the barrier itself
// (mbarrier.try_wait costs several hundred cycles even when the phase is long complete).
__device__ __forceinline__ void tile_ready_wait(const Ring& r, uint32_t idx) {
    if (ld_acquire_cta_shared(r.landed) > idx) return;
    mbar_wait(&r.full[idx % NSLOT], (idx / NSLOT) & 1u);
}

// ---------------------------------------------------------------------------------------------
// stream packing (init time): gather the reference-layout matrices into the per-CTA streams.
//   w_nk = 0: unit n = column n of W [K = D, N]  (HF Conv1D);  w_nk = 1: unit n = row n of W [N, D]
// LayerNorm folding (lnw != null): the block computes  LN(x) . W + b  with
//   LN(x)_k = (x_k - mean) * rstd * lnw_k + lnb_k,  so
//   y_n = rstd * ( sum_k x_k (lnw_k W_kn)  -  mean * c1_n ) + c2_n,
//   c1_n = sum_k lnw_k W_kn,   c2_n = sum_k lnb_k W_kn + b_n.
// The unit stores W'_kn = lnw_k W_kn followed by {c2_n, c1_n, 0, 0}: the GEMV runs on the RAW
// activation vector while the statistics are still being reduced; mean / rstd enter in the epilogue.
// Without folding the pad is {b_n, 0, 0, 0}.
// ---------------------------------------------------------------------------------------------
__global__ void pack_stream_kernel(StreamDims s, int layer, int ph, const float* __restrict__ W,
                                   const float* __restrict__ bias, int w_nk, const float* __restrict__ lnw,
                                   const float* __restrict__ lnb, float* __restrict__ stream) {
    const int N = ph_N(s, ph), D = s.D;
    const int n = blockIdx.x;
    const int c = col_owner(N, n, s.G);
    long long dst = cta_base(s, c);
    if (ph == PH_HEAD) dst += (long long)s.L * cta_layer_floats(s, c);
    else dst += (long long)layer * cta_layer_floats(s, c) + ph_offset_in_layer(s, ph, c);
    dst += (long long)(n - col_begin(N, c, s.G)) * unit_floats(D);
    double c1 = 0.0, c2 = 0.0;
    for (int k = threadIdx.x; k < D; k += blockDim.x) {
        const float w = w_nk ? W[(size_t)n * D + k] : W[(size_t)k * N + n];
        float v = w;
        if (lnw != nullptr) {
            v = lnw[k] * w;
            c1 += (double)v;
            c2 += (double)lnb[k] * (double)w;
        }
        stream[dst + k] = v;
    }
    __shared__ double r1[32], r2[32];
    for (int o = 16; o > 0; o >>= 1) {
        c1 += __shfl_xor_sync(0xffffffffu, c1, o);
        c2 += __shfl_xor_sync(0xffffffffu, c2, o);
    }
    if ((threadIdx.x & 31) == 0) {
        r1[threadIdx.x >> 5] = c1;
        r2[threadIdx.x >> 5] = c2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t1 = 0.0, t2 = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
            t1 += r1[w];
            t2 += r2[w];
        }
        if (bias != nullptr) t2 += (double)bias[n];
        stream[dst + D + 0] = (float)t2;
        stream[dst + D + 1] = (float)t1;
        stream[dst + D + 2] = 0.0f;
        stream[dst + D + 3] = 0.0f;
    }
}

cudaError_t launch_pack_stream(const StreamDims& s, int layer, int ph, const float* W, const float* bias, int w_nk,
                               const float* lnw, const float* lnb, float* stream, cudaStream_t st) {
    pack_stream_kernel<<<ph_N(s, ph), 256, 0, st>>>(s, layer, ph, W, bias, w_nk, lnw, lnb, stream);
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
// self-counting fixed-point all-reduce (mlp.c_proj partial sums).  Every CTA adds, per output, the 64-bit word
//   (round(partial * 2^34) << 8) + 1
// into one accumulator with a relaxed L2 reduction.  Integer addition is associative, so the sum does not depend
// on arrival order (bit-reproducible, unlike float atomics) and is exact to 2^-35 per term; the low byte counts
// contributors, so a reader knows from the word itself when all G (< 256) CTAs have arrived — no fence, no
// separate reducer CTAs, one hop instead of three.
// ---------------------------------------------------------------------------------------------
#ifndef GV_PP_COUNTER
#define GV_PP_COUNTER 1  /* 0 = reducer CTAs poll the partial-sum tags directly: measured slower (0.643 vs 0.631 ms/token) */
#endif
#ifndef GV_ATOMIC_RED
#define GV_ATOMIC_RED 0  /* measured on B200: 148 x 1024 u64 reductions onto 8 KB cost ~6 us per layer (L2 serialises per line); the reducer-CTA path below is faster */
#endif
#define GV_FIX_SCALE 17179869184.0f              /* 2^34 */
#define GV_FIX_INV (1.0 / 17179869184.0)
__device__ __forceinline__ void red_fix_add(unsigned long long* acc, float v) {
    const long long q = __float2ll_rn(v * GV_FIX_SCALE);
    const unsigned long long w = ((unsigned long long)q << 8) + 1ull;
    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(acc), "l"(w) : "memory");
}
__device__ __forceinline__ ulonglong2 ld_x2u64(const unsigned long long* p) {
    ulonglong2 r;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ float fix_value(unsigned long long w) {
    return (float)((double)((long long)w >> 8) * GV_FIX_INV);
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
// One arrival per CTA, by thread 0 as soon as ITS warp's exchange stores are issued: the counter is only a hint (the
// tags validate the data), so it may run ahead of the other warps' stores by their skew -- no block barrier here.
// Every hop_arrive is followed by a hop_wait (with its barrier) before shared memory is reused.
__device__ __forceinline__ void hop_arrive(unsigned* cnt, int tid) {
    if (tid == 0) red_relaxed_add(cnt, 1u);
}
// `near`: the CTA is released when all but `near` arrivals are in; its threads then spin on the tagged words themselves
// for the tail, so the data loads overlap the last arrivals instead of following the counter by an L2 round trip
// (measured: 4 of 148 is the sweet spot, 0.442 -> 0.427 ms/token; 8 and more lose again -- many pollers on lines that
// are still being written slow the writers).  The tags, not the counter, validate the data: any `near` is correct.
// `settle_ns`: the arrival counter is bumped without a fence, so it can overtake the last data stores on their way to
// L2; a first data load that finds a stale tag costs a whole extra round trip, a short pause after the counter
// reaches its target is cheaper.
__device__ __forceinline__ void hop_wait(const unsigned* cnt, unsigned target, int tid, uint32_t tmask, unsigned settle_ns,
                                         volatile int* hold = nullptr, unsigned near = 0u) {
    if (tid == 0 && tmask != 0u) {
        if (hold) *hold = 1;  // the weight producer stops issuing bulk copies: they slow this SM's ordinary loads down
        uint32_t spins = 0;
        while (ld_relaxed_u32(cnt) + near < target) {
            if (++spins > MEGA_SPIN_LIMIT) __trap();
        }
        if (settle_ns) __nanosleep(settle_ns);
    }
    bar_sync(1, MEGA_CONSUMERS);
}

// ---------------------------------------------------------------------------------------------
// weight ring (consumer side).  Tile t of a phase holds units 4t .. 4t+3 and is read by exactly the
// four warps of group t % 2 (warp = 4 * group + unit % 4); each of them arrives once on the tile's
// empty barrier (count = 4).  Tiles are addressed by their global index (ring slot = index % NSLOT).
// ---------------------------------------------------------------------------------------------
struct Cons {
    uint32_t gt;               // global index of the first tile of the current phase (same in every thread)
    unsigned long long* wacc;  // debug: accumulates ns spent waiting for weight tiles (null = off)
};
__device__ __forceinline__ const float* slot_ptr(const Ring& r, uint32_t idx) {
    return r.slots + (size_t)(idx % NSLOT) * r.slot_floats;
}
// Wait for all tiles t = g, g + 2, g + 4, g + 6 (< ntiles) of this warp's group at once: lane k waits
// for tile g + 2k (32 lanes polling one mbarrier would serialise), __syncwarp orders the rest.
__device__ __forceinline__ void group_wait(const Ring& r, const Cons& cs, int g, int ntiles, int lane) {
    long long t0 = 0;
    if (cs.wacc != nullptr && lane == 0) t0 = clock64();  // debug timeline
    if (lane < 4) {
        const int t = g + 2 * lane;
        if (t < ntiles) tile_ready_wait(r, cs.gt + (uint32_t)t);
    }
    __syncwarp();
    if (cs.wacc != nullptr && lane == 0) *cs.wacc += (unsigned long long)(clock64() - t0);
}
// Release them (this warp's arrival on each tile's empty barrier) once the warp has read its units.
__device__ __forceinline__ void group_release(const Ring& r, const Cons& cs, int g, int ntiles, int lane) {
    __syncwarp();
    if (lane < 4) {
        const int t = g + 2 * lane;
        if (t < ntiles) mbar_arrive(&r.empty[(cs.gt + (uint32_t)t) % NSLOT]);
    }
}
// single tile read by all warps (logits-head LayerNorm parameters)
__device__ __forceinline__ const float* tile_wait(const Ring& r, const Cons& cs, uint32_t idx, int lane) {
    if (lane == 0) {
        if (cs.wacc != nullptr) {  // debug timeline
            const long long t0 = clock64();
            tile_ready_wait(r, idx);
            *cs.wacc += (unsigned long long)(clock64() - t0);
        } else {
            tile_ready_wait(r, idx);
        }
    }
    __syncwarp();
    return slot_ptr(r, idx);
}
__device__ __forceinline__ void tile_release(const Ring& r, uint32_t idx, int lane, uint32_t count = 1u) {
    __syncwarp();
    if (lane == 0)
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&r.empty[idx % NSLOT])), "r"(count) : "memory");
}

// sum of the per-warp partials red[0..7] (written before a block barrier)
__device__ __forceinline__ float sum8(const float* red) {
    const float4 a = *reinterpret_cast<const float4*>(red), b = *reinterpret_cast<const float4*>(red + 4);
    return ((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w));
}

// ---------------------------------------------------------------------------------------------
// LayerNorm statistics of the vector held four elements per thread (thread t owns x[4t .. 4t+3]).
// One pass over data shifted by `shift` (any value near the mean keeps E[d^2] - E[d]^2 free of
// cancellation; the callers pass the mean this LayerNorm saw one layer earlier):
//   stats_partial : per-warp sums of d and d^2 -> red[0..7], red[8..15]   (then ONE block barrier)
//   stats_finish  : mean, rstd from the 16 partials (every thread, after the barrier)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void stats_partial(const float4& x, bool valid, float shift, float* red, int lane, int warp) {
    float s1 = 0.0f, s2 = 0.0f;
    if (valid) {
        const float d0 = x.x - shift, d1 = x.y - shift, d2 = x.z - shift, d3 = x.w - shift;
        s1 = (d0 + d1) + (d2 + d3);
        s2 = fmaf(d0, d0, d1 * d1) + fmaf(d2, d2, d3 * d3);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (lane == 0) {
        red[warp] = s1;
        red[8 + warp] = s2;
    }
}
__device__ __forceinline__ void stats_finish(const float* red, float inv_d, float shift, float& mean, float& rstd) {
    const float m = sum8(red) * inv_d;
    float var = fmaf(-m, m, sum8(red + 8) * inv_d);
    var = fmaxf(var, 0.0f);
    mean = shift + m;
    rstd = rsqrtf(var + 1e-5f);
    rstd = rstd * fmaf(-0.5f * (var + 1e-5f) * rstd, rstd, 1.5f);  // one Newton step: full fp32 accuracy
}
// explicit LayerNorm (logits head only: its output is the latent handed to the vocoder); two block barriers
__device__ __forceinline__ void ln_quad(float4& x, bool valid, int D, const float* w, const float* b, float* red, int tid) {
    const int lane = tid & 31, warp = tid >> 5;
    float s = valid ? ((x.x + x.y) + (x.z + x.w)) : 0.0f;
    s = warp_sum(s);
    if (lane == 0) red[warp] = s;
    bar_sync(1, MEGA_CONSUMERS);
    const float mean = sum8(red) / (float)D;
    const float d0 = x.x - mean, d1 = x.y - mean, d2 = x.z - mean, d3 = x.w - mean;
    float q = valid ? (fmaf(d0, d0, d1 * d1) + fmaf(d2, d2, d3 * d3)) : 0.0f;
    q = warp_sum(q);
    if (lane == 0) red[8 + warp] = q;
    bar_sync(1, MEGA_CONSUMERS);
    const float var = sum8(red + 8) / (float)D;
    const float rstd = 1.0f / sqrtf(var + 1e-5f);
    if (valid) {
        const float4 ww = *reinterpret_cast<const float4*>(w + 4 * tid);
        const float4 bb = *reinterpret_cast<const float4*>(b + 4 * tid);
        x.x = d0 * rstd * ww.x + bb.x;
        x.y = d1 * rstd * ww.y + bb.y;
        x.z = d2 * rstd * ww.z + bb.z;
        x.w = d3 * rstd * ww.w + bb.w;
    }
}

// ---------------------------------------------------------------------------------------------
// GEMV, K = D, one warp per unit: unit u of the phase is handled by warp u % 8 (u < nunits <= 32,
// up to four units per warp).  `xs` is the activation vector in shared memory (complete: the
// caller synchronised).  For every unit, lane 8q (q = u / 8) of the owning warp calls
// epi(u, dot, c2, c1) with  dot = sum_k x_k W'[k][unit]  and the unit's two epilogue constants
// (pack_stream_kernel).
// ---------------------------------------------------------------------------------------------
// EARLY: release every tile right after its unit (the producer refills while this phase goes on: needed when another
// GEMV follows without a hop, i.e. FC -> P2); otherwise one release for all of the warp's tiles at the end (each
// __syncwarp + elected mbarrier.arrive costs ~250 cycles of this warp's chain).
// Registers: at 168 per thread (the launch allocation of a 12-warp CTA) ptxas sinks every LDS.128 next to its FFMAs (one
// or two shared-memory loads in flight per warp).  The CTA therefore launches with a 4-warp producer warpgroup that
// gives its registers away (setmaxnreg.dec 40) and the two consumer warpgroups take 232 each (setmaxnreg.inc): ptxas
// then batches 8 LDS.128 per warp and every phase of the layer gets ~40 % shorter (0.60 -> 0.45 ms/token).
template <int NXV, bool EARLY, class Epi>
__device__ __forceinline__ void gemv_dot(const Ring& ring, const Cons& cs, int nunits, const float* xs, int warp, int lane,
                                         Epi epi) {
    constexpr int D = NXV * 128, UF = D + 4;
    float4 xv[NXV];
#pragma unroll
    for (int i = 0; i < NXV; ++i) xv[i] = *reinterpret_cast<const float4*>(xs + (i * 32 + lane) * 4);
    float tot[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    float2 cc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) cc[k] = make_float2(0.f, 0.f);
    const int ntiles = (nunits + UPT - 1) / UPT;
    const int g = warp >> 2, r = warp & 3;
    group_wait(ring, cs, g, ntiles, lane);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int t = g + 2 * k;
        if (t * UPT + r < nunits) {
            const float* col = slot_ptr(ring, cs.gt + (uint32_t)t) + r * UF;
            cc[k] = *reinterpret_cast<const float2*>(col + D);
            float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll
            for (int i = 0; i < NXV; ++i) {
                const float4 wv = *reinterpret_cast<const float4*>(col + (i * 32 + lane) * 4);
                a0 = fmaf(wv.x, xv[i].x, a0);
                a1 = fmaf(wv.y, xv[i].y, a1);
                a2 = fmaf(wv.z, xv[i].z, a2);
                a3 = fmaf(wv.w, xv[i].w, a3);
            }
            tot[k] = (a0 + a1) + (a2 + a3);
        }
        if constexpr (EARLY) {
            if (t < ntiles) tile_release(ring, cs.gt + (uint32_t)t, lane);
        }
    }
    if constexpr (!EARLY) group_release(ring, cs, g, ntiles, lane);
    // transposing butterfly: lanes [8q, 8q+8) end up reducing tot[q]
    const bool up16 = (lane & 16) != 0, up8 = (lane & 8) != 0;
    const float k0 = up16 ? tot[2] : tot[0], s0 = up16 ? tot[0] : tot[2];
    const float k1 = up16 ? tot[3] : tot[1], s1 = up16 ? tot[1] : tot[3];
    const float h0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 16);  // tot[0] (lanes < 16) / tot[2]
    const float h1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 16);  // tot[1] (lanes < 16) / tot[3]
    const float keep = up8 ? h1 : h0, send = up8 ? h0 : h1;
    float v = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    const int q = lane >> 3;
    const int u = warp + 8 * q;
    const float2 c = up16 ? (up8 ? cc[3] : cc[2]) : (up8 ? cc[1] : cc[0]);
    if ((lane & 7) == 0 && u < nunits) epi(u, v, c.x, c.y);
}

// mlp.c_proj split along K: unit k = row of W_proj2 owned by this CTA.  Warp group g takes tiles
// t = g, g + 2, ...; thread j of the group (0..127) owns outputs [4j, 4j+4) and [D/2 + 4j, D/2 + 4j + 4) and
// accumulates u_k * W[k][.] over the group's rows; the two group partials land in part[g][D].
template <int NXV>
__device__ __forceinline__ void gemv_outer(const Ring& ring, const Cons& cs, int nunits, const float* us, int tid, int lane,
                                           int warp, float* part) {
    constexpr int D = NXV * 128, UF = D + 4, HALF = D / 2;
    const int g = warp >> 2, j = tid & 127;
    const bool valid = 4 * j < HALF;
    const int ntiles = (nunits + UPT - 1) / UPT;
    float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int t = g + 2 * k;
        if (t < ntiles) {
            const float* base = tile_wait(ring, cs, cs.gt + (uint32_t)t, lane) + 4 * j;
#pragma unroll
            for (int r = 0; r < UPT; ++r) {
                const int kk = t * UPT + r;
                if (kk < nunits && valid) {
                    const float uk = us[kk];
                    const float4 w0 = *reinterpret_cast<const float4*>(base + r * UF);
                    const float4 w1 = *reinterpret_cast<const float4*>(base + r * UF + HALF);
                    acc0.x = fmaf(uk, w0.x, acc0.x); acc0.y = fmaf(uk, w0.y, acc0.y);
                    acc0.z = fmaf(uk, w0.z, acc0.z); acc0.w = fmaf(uk, w0.w, acc0.w);
                    acc1.x = fmaf(uk, w1.x, acc1.x); acc1.y = fmaf(uk, w1.y, acc1.y);
                    acc1.z = fmaf(uk, w1.z, acc1.z); acc1.w = fmaf(uk, w1.w, acc1.w);
                }
            }
        }
        if (t < ntiles) tile_release(ring, cs.gt + (uint32_t)t, lane);
    }
    if (valid) {
        *reinterpret_cast<float4*>(part + g * D + 4 * j) = acc0;
        *reinterpret_cast<float4*>(part + g * D + HALF + 4 * j) = acc1;
    }
}

// ---------------------------------------------------------------------------------------------
// single-query attention over one (head, key range) item, 8 warps, built for a short dependency chain after q
// arrives.  Arithmetic of HF GPT2Attention._attn for q_len == 1:
//   s_j = (q . k_j) / sqrt(hd);  p = softmax_j(s);  o = sum_j p_j v_j       (un-normalised here)
//   * keys are cut into batches of 32: warp w owns rows 4w .. 4w+3 of every batch and keeps its own running
//     (max, sum, o[hd]) -- flash-decoding inside the CTA: no score buffer, no barrier before the softmax;
//   * before q exists: K and V rows of batch 0 are requested from the cache into registers;
//   * then the CTA polls the tagged q words (only the <= H * 8 attention CTAs read xq, so they poll the data directly:
//     one L2 round trip less than counter-then-data; warp w polls an eighth of q and the CTA assembles it in shared
//     memory: -0.6 % ms/token against every warp polling all of q); the warp that owns the position being decoded also
//     polls k / v of this very step from the exchange buffer and appends them to the cache; that position rides along
//     as a fifth row of its batch (its cache row is loaded as zeros and masked);
//   * per batch: 4 (+1) dots and one interleaved shuffle tree, online-softmax update, o += p v; K / V rows of the next
//     batch are requested as soon as the registers of the current one are free;
//   * the 8 per-warp partials meet in shared memory behind ONE barrier and are merged by thread d < hd.
// ---------------------------------------------------------------------------------------------
#define ATT_ROWS 4        // K rows per warp per batch (32 keys / 8 warps)
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
// DPL tagged elements of one lane (element e at base[2e] = {value, tag}); false while any tag is stale
template <int VEC, int NCH>
__device__ __forceinline__ bool ld_tagged_lane(const float* base, int lane, uint32_t tag, uint32_t tmask, float* r) {
    bool ok = true;
    if constexpr (VEC == 1) {
        const uint2 a = ld_x8(base + 2 * lane);
        ok = ((a.y ^ tag) & tmask) == 0u;
        r[0] = __uint_as_float(a.x);
    } else {
        uint4 a[NCH * (VEC / 2)];
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
            for (int w2 = 0; w2 < VEC / 2; ++w2) a[c * (VEC / 2) + w2] = ld_x16(base + 2 * ((c * 32 + lane) * VEC + 2 * w2));
#pragma unroll
        for (int w2 = 0; w2 < NCH * (VEC / 2); ++w2) {
            ok = ok && tags_ok(a[w2], tag, tmask);
            r[2 * w2] = __uint_as_float(a[w2].x);
            r[2 * w2 + 1] = __uint_as_float(a[w2].z);
        }
    }
    return ok;
}

template <int HD, bool DBG>
__device__ __noinline__ void att_item(float* __restrict__ Kc, float* __restrict__ Vc, const float* xq, int D, int h, int j0, int j1, int S,
                         uint32_t tag_in, float* wml, float* wpart, int tid,
                         float* o_out, float* ml_out, int item, uint32_t tag_out, uint32_t tmask, unsigned long long* dbg) {
    using L = AttLane<HD>;
    long long ck[8];
    if constexpr (DBG) ck[0] = clock64();
    constexpr int VEC = L::VEC, NCH = L::NCH, DPL = L::DPL;
    const int warp = tid >> 5, lane = tid & 31;
    const int nk = j1 - j0;
    const int nb = (nk + 31) >> 5;
    const int jn = S - 1 - j0;  // relative index of the position being decoded (inside this item iff 0 <= jn < nk)
    const bool mine = jn >= 0 && jn < nk && ((jn & 31) >> 2) == warp;  // this warp owns the new position
    const int bn = jn >> 5;
    const float sqrt_hd = sqrtf((float)HD);

    auto load_row = [&](const float* base, int jr, float* dst) {  // cache row of relative key jr (zeros outside / new key)
        if (jr < nk && jr != jn) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) ld_vec<VEC>(base + (size_t)(j0 + jr) * HD + (c * 32 + lane) * VEC, dst + c * VEC);
        } else {
#pragma unroll
            for (int i = 0; i < DPL; ++i) dst[i] = 0.0f;
        }
    };
    // ---- batch 0 from the cache (independent of this step's q) ----
    float kr[ATT_ROWS][DPL], vr[ATT_ROWS][DPL];
#pragma unroll
    for (int u = 0; u < ATT_ROWS; ++u) load_row(Kc, warp * ATT_ROWS + u, kr[u]);
#pragma unroll
    for (int u = 0; u < ATT_ROWS; ++u) load_row(Vc, warp * ATT_ROWS + u, vr[u]);
    if constexpr (DBG) ck[1] = ck[2] = ck[7] = clock64();
    // ---- this step's q (and k / v of the position being decoded) ----
    float qr[DPL], knew[DPL], vnew[DPL];
#pragma unroll
    for (int i = 0; i < DPL; ++i) knew[i] = vnew[i] = 0.0f;
    {
        const float* qp = xq + 2 * (size_t)(h * HD);
        const float* kp = xq + 2 * (size_t)(D + h * HD);
        const float* vp = xq + 2 * (size_t)(2 * D + h * HD);
        // q is the same for every warp: warp w polls only elements [w * hd/8, (w+1) * hd/8) and the CTA assembles q in
        // shared memory (an eighth of the polling traffic on lines the QKV epilogues are still writing)
        constexpr int EPW = HD / MEGA_WARPS;
        float qmine = 0.0f;
        uint32_t spins = 0;
        bool ok = false;
        while (!ok) {
            ok = true;
            if (lane < EPW) {
                const uint2 a = ld_x8(qp + 2 * (warp * EPW + lane));
                ok = ((a.y ^ tag_in) & tmask) == 0u;
                qmine = __uint_as_float(a.x);
            }
            if (mine) {
                const bool ok_k = ld_tagged_lane<VEC, NCH>(kp, lane, tag_in, tmask, knew);
                const bool ok_v = ld_tagged_lane<VEC, NCH>(vp, lane, tag_in, tmask, vnew);
                ok = ok && ok_k && ok_v;
            }
            ok = __all_sync(0xffffffffu, ok);
            if (++spins > MEGA_SPIN_LIMIT) __trap();
        }
        float* qs = wml + 16;
        if (lane < EPW) qs[warp * EPW + lane] = qmine;
        bar_sync(1, MEGA_CONSUMERS);
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
            for (int i = 0; i < VEC; ++i) qr[c * VEC + i] = qs[(c * 32 + lane) * VEC + i];
    }
    if constexpr (DBG) ck[3] = clock64() + (long long)(qr[0] == 123.f);
    if (mine) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            st_vec<VEC>(Kc + (size_t)(S - 1) * HD + (c * 32 + lane) * VEC, knew + c * VEC);
            st_vec<VEC>(Vc + (size_t)(S - 1) * HD + (c * 32 + lane) * VEC, vnew + c * VEC);
        }
    }
    // ---- batches: scores, online softmax, PV (per warp) ----
    float m = -INFINITY, lsum = 0.0f, o[DPL];
#pragma unroll
    for (int i = 0; i < DPL; ++i) o[i] = 0.0f;
    for (int b = 0; b < nb; ++b) {
        const int r0 = b * 32 + warp * ATT_ROWS;
        const bool with_new = mine && b == bn;  // warp-uniform
        float sc[ATT_ROWS + 1];
#pragma unroll
        for (int u = 0; u < ATT_ROWS; ++u) {
            float a = 0.0f;
#pragma unroll
            for (int i = 0; i < DPL; ++i) a = fmaf(qr[i], kr[u][i], a);
            sc[u] = a;
        }
        {
            float a = 0.0f;
#pragma unroll
            for (int i = 0; i < DPL; ++i) a = fmaf(qr[i], knew[i], a);
            sc[ATT_ROWS] = a;
        }
        if (b + 1 < nb) {  // K registers are free: request the next batch
#pragma unroll
            for (int u = 0; u < ATT_ROWS; ++u) load_row(Kc, r0 + 32 + u, kr[u]);
        }
#pragma unroll
        for (int x = 16; x > 0; x >>= 1) {
#pragma unroll
            for (int u = 0; u <= ATT_ROWS; ++u) sc[u] += __shfl_xor_sync(0xffffffffu, sc[u], x);
        }
        if constexpr (DBG) {
            if (b == 0) ck[7] = clock64() + (long long)(sc[0] == 123.f) + (long long)(sc[3] == 123.f);
        }
        float mb = -INFINITY;
#pragma unroll
        for (int u = 0; u <= ATT_ROWS; ++u) {
            const bool valid = u < ATT_ROWS ? (r0 + u < nk && r0 + u != jn) : with_new;
            // s / sqrt(hd): for hd = 64, 256 the divisor is a power of two and the product with its reciprocal is the same
            const float sv = (HD == 64 || HD == 256) ? sc[u] * (1.0f / sqrt_hd) : sc[u] / sqrt_hd;
            sc[u] = valid ? sv : -INFINITY;
            mb = fmaxf(mb, sc[u]);
        }
        if (mb > -INFINITY) {  // warp-uniform
            const float mn = fmaxf(m, mb);
            const float c = expf(m - mn);  // first batch: exp(-inf) = 0
            float pj[ATT_ROWS + 1];
            lsum *= c;
#pragma unroll
            for (int u = 0; u <= ATT_ROWS; ++u) {
                pj[u] = expf(sc[u] - mn);  // rows outside the range: exp(-inf) = 0
                lsum += pj[u];
            }
#pragma unroll
            for (int i = 0; i < DPL; ++i) {
                float a = o[i] * c;
#pragma unroll
                for (int u = 0; u < ATT_ROWS; ++u) a = fmaf(pj[u], vr[u][i], a);
                o[i] = fmaf(pj[ATT_ROWS], vnew[i], a);  // p = 0 unless this warp holds the new position in this batch
            }
            m = mn;
        }
        if constexpr (DBG) {
            if (b == 0) ck[2] = clock64() + (long long)(lsum == 123.f);
        }
        if (b + 1 < nb) {
#pragma unroll
            for (int u = 0; u < ATT_ROWS; ++u) load_row(Vc, r0 + 32 + u, vr[u]);
        }
    }
    if constexpr (DBG) ck[4] = clock64() + (long long)(lsum == 123.f);
    // ---- merge the 8 per-warp partials ----
    if (lane == 0) {
        wml[2 * warp] = m;
        wml[2 * warp + 1] = lsum;
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        float* dst = wpart + warp * HD + (c * 32 + lane) * VEC;
        if constexpr (VEC == 4) *reinterpret_cast<float4*>(dst) = make_float4(o[c * 4], o[c * 4 + 1], o[c * 4 + 2], o[c * 4 + 3]);
        else if constexpr (VEC == 2) *reinterpret_cast<float2*>(dst) = make_float2(o[0], o[1]);
        else *dst = o[0];
    }
    bar_sync(1, MEGA_CONSUMERS);
    if constexpr (DBG) ck[5] = clock64();
    if (tid < HD) {
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < MEGA_WARPS; ++w) M = fmaxf(M, wml[2 * w]);
        float Lsum = 0.0f, os = 0.0f;
#pragma unroll
        for (int w = 0; w < MEGA_WARPS; ++w) {
            const float cw = expf(wml[2 * w] - M);  // warps without keys: exp(-inf) = 0
            Lsum = fmaf(wml[2 * w + 1], cw, Lsum);
            os = fmaf(wpart[w * HD + tid], cw, os);
        }
        st_tagged(o_out, item * HD + tid, os, tag_out);
        if (tid == 0) st_tagged2(ml_out, item * 2, M, Lsum, tag_out);
    }
    if constexpr (DBG) {
      ck[6] = clock64();
      if (dbg != nullptr && tid == 0) {
        // [prefetch issue, poll until q seen, dots + shuffles of batch 0, softmax weights + PV, later batches, barrier + merge + store]
        dbg[0] = (unsigned long long)(ck[1] - ck[0]);
        dbg[1] = (unsigned long long)(ck[3] - ck[1]);
        dbg[2] = (unsigned long long)(ck[7] - ck[3]);
        dbg[3] = (unsigned long long)(ck[2] - ck[7]);
        dbg[4] = (unsigned long long)(ck[4] - ck[2]);
        dbg[5] = (unsigned long long)(ck[6] - ck[4]);
      }
    }
}

// ---------------------------------------------------------------------------------------------
// producer: one thread walks the CTA's weight stream through the ring
// ---------------------------------------------------------------------------------------------
struct Producer {
    const Ring& ring;
    const float* region = nullptr;   // this CTA's contiguous weight stream [region, region + region_floats)
    long long region_floats = 0;
    long long ahead_floats = 0;      // L2 prefetch distance (0 = off)
    uint32_t t = 0, slot = 0, phase = 0;
    uint32_t land = 0;  // tiles [0, land) have been seen complete (published to the consumers through ring.landed)
    uint32_t window;    // at most this many tiles requested but not landed
    volatile int* stop;
    volatile int* hold = nullptr;  // non-null: do not issue new copies while *hold != 0 (consumers are polling / loading a hop)
    uint64_t policy;
    __device__ Producer(const Ring& r, volatile int* s, uint32_t w) : ring(r), window(w), stop(s) {
        policy = l2_policy_evict_first();
    }
    // non-blocking: move `land` over every tile whose full barrier has completed and publish it
    __device__ void advance() {
        const uint32_t l0 = land;
        while (land < t && mbar_test_wait(&ring.full[land % NSLOT], (land / NSLOT) & 1u)) ++land;
        if (land != l0) st_release_cta_shared(ring.landed, land);
    }
    // returns false when the consumers asked to stop
    __device__ bool issue(const float* src, uint32_t floats, bool stream_once) {
        uint32_t spins = 0;
        while (!mbar_test_wait(&ring.empty[slot], phase ^ 1u)) {
            advance();
            if (*stop) return false;
            if (++spins > MEGA_SPIN_LIMIT) __trap();
        }
        if (*stop) return false;
        if (hold != nullptr) {
            spins = 0;
            while (*hold) {
                advance();
                if (*stop) return false;
                if (++spins > MEGA_SPIN_LIMIT) __trap();
            }
        }
        spins = 0;
        while (t - land >= window) {  // at most `window` tiles requested but not landed
            advance();
            if (++spins > MEGA_SPIN_LIMIT) __trap();
        }
        mbar_arrive_expect_tx(&ring.full[slot], floats * 4u);
        float* dst = ring.slots + (size_t)slot * ring.slot_floats;
        if (stream_once) {
            bulk_g2s_hint(dst, src, floats * 4u, &ring.full[slot], policy);
            if (ahead_floats > 0) {
                // HBM -> L2 for the bytes this CTA will want `ahead_floats` later (wraps to the next token's pass):
                // the smem ring then refills at L2 latency, and HBM streaming no longer depends on ring depth
                long long o = (src - region) + ahead_floats;
                if (o >= region_floats) o -= region_floats;
                const long long n = min((long long)floats, region_floats - o);
                bulk_prefetch_l2(region + o, (uint32_t)n * 4u);
                if (n < (long long)floats) bulk_prefetch_l2(region, (uint32_t)((long long)floats - n) * 4u);
            }
        } else {
            bulk_g2s(dst, src, floats * 4u, &ring.full[slot]);
        }
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
        for (int ph = PH_QKV; ph <= PH_P2; ++ph)
            if (!pr.issue_units(lw, nun[ph], uf)) return false;
    }
    if (!pr.issue(p.blob + p.lnf_off, 4 * D, false)) return false;  // ln_f and final_norm parameters
    const float* hw = base + (long long)p.L * lfl;
    return pr.issue_units(hw, nun[PH_HEAD], uf);
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <int NXV, bool TRACE>
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
    // scratch region: sampling sort keys | attention scores + PV partials | mlp.c_proj group partials |
    // partial-sum gather (never live together)
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw + off);
    float* att_sc = reinterpret_cast<float*>(smem_raw + off);            // [8][2] per-warp (max, sum)
    float* att_op = reinterpret_cast<float*>(smem_raw + off + 1280);     // [8][hd] per-warp PV partials (q staging before it)
    float* part = reinterpret_cast<float*>(smem_raw + off);              // [2][D]
    float* gat = reinterpret_cast<float*>(smem_raw + off);               // [G][8]
    off += mega_scratch_bytes(D);
    float* xres0 = reinterpret_cast<float*>(smem_raw + off);  // [D] residual stream entering the block (= QKV GEMV input)
    off += (size_t)D * sizeof(float);
    float* xres1 = reinterpret_cast<float*>(smem_raw + off);  // [D] residual stream after attention (= FC GEMV input)
    off += (size_t)D * sizeof(float);
    float* xo = reinterpret_cast<float*>(smem_raw + off);  // [D] attention output (PROJ input) / latent (head input)
    off += (size_t)D * sizeof(float);
    float* slog = reinterpret_cast<float*>(smem_raw + off);  // [Vpad] logits of the step being sampled
    off += (size_t)p.Vpad * sizeof(float);
    ring.full = reinterpret_cast<uint64_t*>(smem_raw + off);
    off += 16 * sizeof(uint64_t);
    ring.empty = reinterpret_cast<uint64_t*>(smem_raw + off);
    off += 16 * sizeof(uint64_t);
    float* us = reinterpret_cast<float*>(smem_raw + off);  // [32] this CTA's gelu(fc) values
    off += 32 * sizeof(float);
    float* red = reinterpret_cast<float*>(smem_raw + off);  // [2][16] LayerNorm statistics (ln_1 | ln_2)
    off += 64 * sizeof(float);
    float* fscr = reinterpret_cast<float*>(smem_raw + off);
    off += 16 * sizeof(float);
    int* iscr = reinterpret_cast<int*>(smem_raw + off);
    off += 16 * sizeof(int);
    volatile int* ctl = reinterpret_cast<volatile int*>(smem_raw + off);  // [0] stop flag, [1] tiles consumed, [2] token
    ring.landed = reinterpret_cast<uint32_t*>(smem_raw + off) + 3;        // [3] tiles the producer has seen landed
    volatile int* hold = reinterpret_cast<volatile int*>(smem_raw + off) + 4;  // [4] consumers are inside a hop: producer pauses
    off += 8 * sizeof(int);
    unsigned char* seen = smem_raw + off;  // [Vpad]

    if (tid_all == 0) {
        for (int i = 0; i < NSLOT; ++i) {
            mbar_init(&ring.full[i], 1);
            mbar_init(&ring.empty[i], 4);  // the four reader warps of a tile
        }
        ctl[0] = 0;
        ctl[1] = 0;
        ctl[2] = 0;
        ctl[3] = 0;
        ctl[4] = 0;
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
        // ================= producer warpgroup: hands its registers to the consumers =================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (tid_all == MEGA_CONSUMERS) {
            Producer pr(ring, ctl, (uint32_t)max(1, min(p.window, NSLOT)));
            if (p.hop_hold) pr.hold = hold;
            pr.region = p.stream + cta_base(sd, cta);
            pr.region_floats = cta_base(sd, cta + 1) - cta_base(sd, cta);
            pr.ahead_floats = min((long long)p.l2_ahead_tiles * slot_floats(D), pr.region_floats - slot_floats(D));
            if (pr.ahead_floats < 0) pr.ahead_floats = 0;
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
                pr.advance();
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
    Cons cs{0u, nullptr};
    unsigned long long wait_ns = 0;
    const uint32_t tmask = p.dbg_nosync ? 0u : 0xffffffffu;  // debug: 0 = do not wait for exchange data
    const bool xvalid = 4 * tid < D;                         // thread t owns elements 4t .. 4t+3 of every D-vector
    const float inv_d = 1.0f / (float)D;
    const int H = p.H, HD = D / H;
    int nun[5], ubeg[5], ntl[5];
    for (int ph = 0; ph < 5; ++ph) {
        nun[ph] = ph_units(sd, ph, cta);
        ubeg[ph] = (int)col_begin(ph_N(sd, ph), cta, G);
        ntl[ph] = (nun[ph] + UPT - 1) / UPT;
    }
    const int n_red = D / 8;  // reducer CTAs of the mlp.c_proj partial sums (8 outputs each, one per warp)
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
    const unsigned settle = (unsigned)p.hop_settle_ns;
    volatile int* const hold_c = p.hop_hold ? hold : nullptr;
    // early-release margins of the hops (hop_wait): grid-wide hops, and the attention-output hop (n_items arrivals)
    const unsigned near = p.hop_near >= 0 ? (unsigned)p.hop_near : (unsigned)max(G / 37, 1);
    // hop counter targets (counters are zero at launch): x1 / pp / x2 advance by a fixed amount per layer, so they are
    // derived from one layer counter; only the attention target (items vary with S) and the logits target are running sums
    unsigned lc = 0, t_ao = 0, t_lg = 0;
    float shift1 = 0.0f, shift2 = 0.0f;  // statistics shifts of ln_1 / ln_2: the means seen one layer earlier

    for (int i = 0; i < p.n_steps; ++i) {
        // debug timeline (compiled out of the production instantiation)
        const bool tr = TRACE && p.trace != nullptr && i == p.trace_step && tid == 0;
        unsigned long long* trow = TRACE ? p.trace + (size_t)cta * p.trace_slots : nullptr;
        auto stamp = [&](int slot) {
            if constexpr (TRACE) {
                if (tr && slot < p.trace_slots) trow[slot] = globaltimer_ns();
            }
        };
        stamp(p.L * GV_TRACE_PER_LAYER + 4);
        if constexpr (TRACE) cs.wacc = tr ? &wait_ns : nullptr;
        auto stamp_wait = [&](int slot) {  // stores the weight-wait time accumulated since the previous call
            if constexpr (TRACE) {
                if (tr && slot < p.trace_slots) trow[slot] = wait_ns;
                wait_ns = 0;
            }
        };
        float4 lat = make_float4(0.f, 0.f, 0.f, 0.f);  // final_norm(ln_f(x)): the latent of this step (elements 4 tid ..)
        float4 xnext = make_float4(0.f, 0.f, 0.f, 0.f);  // residual stream leaving a block (GV_ATOMIC_RED: every CTA has it)
        if (!(i == 0 && had_pending)) {
            // ------------- forward of token `last_tok` at mel position n, cache row P + n -------------
            const uint32_t tbase = p.tag0 + fwd * tags_per_fwd;
            fwd += 1;
            const int pos = p.P + n;
            const int S = pos + 1;
            const int chunk = att_chunk(S), nsplit = att_nsplit(S);
            const int n_items = H * nsplit;
            const unsigned near_ao = p.hop_near_ao >= 0 ? (unsigned)p.hop_near_ao : (unsigned)min(max(n_items / 3, 1), 4);
            for (int l = 0; l < p.L; ++l) {
                float* kc = p.kv + ((size_t)l * 2 + 0) * p.kv_layer_stride;
                float* vc = p.kv + ((size_t)l * 2 + 1) * p.kv_layer_stride;
                const uint32_t tg = tbase + (uint32_t)GV_TAGS_PER_LAYER * (uint32_t)l;
                const int ts = l * GV_TRACE_PER_LAYER;
                // ---- QKV: [q|k|v] = LN1(x) . W_attn + b  (LN folded: GEMV on raw x, statistics in the epilogue) ----
                {
                    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (l == 0) {
                        if (xvalid) {
                            const float4 a = *reinterpret_cast<const float4*>(mel_emb + (size_t)last_tok * D + 4 * tid);
                            const float4 b = *reinterpret_cast<const float4*>(mel_pos + (size_t)n * D + 4 * tid);
                            x = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
                        }
#if GV_ATOMIC_RED
                        {   // clear the accumulator set the NEXT forward will use (idle since the previous forward ended)
                            unsigned long long* z = p.acc + (size_t)(fwd & 1u) * p.L * D;  // fwd was incremented: (fwd & 1) = next parity
                            const int total = p.L * D, per = (total + G - 1) / G;
                            for (int q = cta * per + tid; q < min(total, (cta + 1) * per); q += MEGA_CONSUMERS) __stcg(z + q, 0ull);
                        }
                    } else {
                        x = xnext;
                    }
#else
                    } else {
                        hop_wait(hc + HC_X2 * GV_HOP_STRIDE, lc * (unsigned)n_red, tid, tmask, settle, hold_c, near);
                        if (xvalid) ld_tagged_vec<4>(p.x2, 4 * tid, tg - (uint32_t)GV_TAGS_PER_LAYER + TG_X2, tmask, &x.x);
                    }
#endif
                    if (xvalid) *reinterpret_cast<float4*>(xres0 + 4 * tid) = x;
                    if (hold_c && tid == 0) *hold_c = 0;  // hop data is in registers: the producer may stream again
                    stamp(ts + 0);
                    stats_partial(x, xvalid, shift1, red, lane, warp);
                    bar_sync(1, MEGA_CONSUMERS);
                    float mean, rstd;
                    stats_finish(red, inv_d, shift1, mean, rstd);
                    shift1 = mean;
                    stamp(ts + 1);
                    gemv_dot<NXV, false>(ring, cs, nun[PH_QKV], xres0, warp, lane, [&](int u, float dot, float c2, float c1) {
                        st_tagged(p.xq, ubeg[PH_QKV] + u, fmaf(rstd, fmaf(-mean, c1, dot), c2), tg + TG_XQ);
                    });
                    cs.gt += (uint32_t)ntl[PH_QKV];
                    stamp(ts + 2);
                }
                // ---- ATT: (head, key-range) items on the first n_items CTAs ----
                if (cta < n_items) {
                    const int h = cta / nsplit, sp = cta % nsplit;
                    const int j0 = sp * chunk, j1 = min(S, j0 + chunk);
                    float* kh = kc + (size_t)h * p.S_max * HD;
                    float* vh = vc + (size_t)h * p.S_max * HD;
#define GV_ATT_CASE(hd)                                                                                                     \
    case hd:                                                                                                                \
        att_item<hd, TRACE>(kh, vh, p.xq, D, h, j0, j1, S, tg + TG_XQ, att_sc, att_op, tid,                                       \
                     p.att_o, p.att_ml, cta, tg + TG_AO, tmask, (TRACE && tr && ts + 20 <= p.trace_slots) ? trow + ts + 14 : nullptr); \
        break;
                    switch (HD) {
                        GV_ATT_CASE(32) GV_ATT_CASE(64) GV_ATT_CASE(128) GV_ATT_CASE(256)
                        default: break;
                    }
#undef GV_ATT_CASE
                    hop_arrive(hc + HC_AO * GV_HOP_STRIDE, tid);
                }
                t_ao += (unsigned)n_items;
                stamp(ts + 3);
                // ---- PROJ: merge attention partials -> o ; x1 = x + o . W_proj + b ----
                {
                    hop_wait(hc + HC_AO * GV_HOP_STRIDE, t_ao, tid, tmask, settle, hold_c, near_ao);
                    if (xvalid) {
                        const int h = (4 * tid) / HD, d = (4 * tid) % HD;
                        const uint32_t tga = tg + TG_AO;
                        float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f, M = -INFINITY, den = 0.0f;
                        for (int s0 = 0; s0 < nsplit; s0 += 4) {  // loads of four splits in flight, merged in split order
                            uint4 a[4], b[4], c[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                if (s0 + q < nsplit) {
                                    const int it = h * nsplit + s0 + q;
                                    a[q] = ld_x16(p.att_ml + 2 * (size_t)(it * 2));
                                    b[q] = ld_x16(p.att_o + 2 * (size_t)(it * HD + d));
                                    c[q] = ld_x16(p.att_o + 2 * (size_t)(it * HD + d + 2));
                                }
                            }
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                if (s0 + q < nsplit) {
                                    const int it = h * nsplit + s0 + q;
                                    uint32_t spins = 0;
                                    while (!(tags_ok(a[q], tga, tmask) && tags_ok(b[q], tga, tmask) && tags_ok(c[q], tga, tmask))) {
                                        if (++spins > MEGA_SPIN_LIMIT) __trap();
                                        a[q] = ld_x16(p.att_ml + 2 * (size_t)(it * 2));
                                        b[q] = ld_x16(p.att_o + 2 * (size_t)(it * HD + d));
                                        c[q] = ld_x16(p.att_o + 2 * (size_t)(it * HD + d + 2));
                                    }
                                    const float mq = __uint_as_float(a[q].x), lq = __uint_as_float(a[q].z);
                                    const float Mn = fmaxf(M, mq);
                                    const float c_old = expf(M - Mn);  // first split: exp(-inf) = 0
                                    const float c_new = expf(mq - Mn);
                                    den = den * c_old + lq * c_new;
                                    o0 = o0 * c_old + __uint_as_float(b[q].x) * c_new;
                                    o1 = o1 * c_old + __uint_as_float(b[q].z) * c_new;
                                    o2 = o2 * c_old + __uint_as_float(c[q].x) * c_new;
                                    o3 = o3 * c_old + __uint_as_float(c[q].z) * c_new;
                                    M = Mn;
                                }
                            }
                        }
                        *reinterpret_cast<float4*>(xo + 4 * tid) = make_float4(o0 / den, o1 / den, o2 / den, o3 / den);
                    }
                    if (hold_c && tid == 0) *hold_c = 0;  // hop data is in registers: the producer may stream again
                    stamp(ts + 4);
                    bar_sync(1, MEGA_CONSUMERS);
                    gemv_dot<NXV, false>(ring, cs, nun[PH_PROJ], xo, warp, lane, [&](int u, float dot, float c2, float) {
                        const int col = ubeg[PH_PROJ] + u;
                        st_tagged(p.x1, col, xres0[col] + (dot + c2), tg + TG_X1);
                    });
                    cs.gt += (uint32_t)ntl[PH_PROJ];
                    stamp(ts + 5);
                    stamp_wait(ts + 12);
                    hop_arrive(hc + HC_X1 * GV_HOP_STRIDE, tid);
                }
                // ---- FC + P2: u = gelu_new(LN2(x1) . W_fc + b) (kept in this CTA) -> partial of u . W_proj2 ----
                {
                    hop_wait(hc + HC_X1 * GV_HOP_STRIDE, (lc + 1u) * (unsigned)G, tid, tmask, settle, hold_c, near);
                    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (xvalid) {
                        ld_tagged_vec<4>(p.x1, 4 * tid, tg + TG_X1, tmask, &x.x);
                        *reinterpret_cast<float4*>(xres1 + 4 * tid) = x;
                    }
                    if (hold_c && tid == 0) *hold_c = 0;  // hop data is in registers: the producer may stream again
                    stamp(ts + 6);
                    stats_partial(x, xvalid, shift2, red + 16, lane, warp);
                    bar_sync(1, MEGA_CONSUMERS);
                    float mean, rstd;
                    stats_finish(red + 16, inv_d, shift2, mean, rstd);
                    shift2 = mean;
                    stamp(ts + 7);
                    gemv_dot<NXV, true>(ring, cs, nun[PH_FC], xres1, warp, lane, [&](int u, float dot, float c2, float c1) {
                        us[u] = gelu_new(fmaf(rstd, fmaf(-mean, c1, dot), c2));
                    });
                    cs.gt += (uint32_t)ntl[PH_FC];
                    bar_sync(1, MEGA_CONSUMERS);
                    stamp(ts + 8);
                    gemv_outer<NXV>(ring, cs, nun[PH_P2], us, tid, lane, warp, part);
                    cs.gt += (uint32_t)ntl[PH_P2];
                    bar_sync(1, MEGA_CONSUMERS);
#if GV_ATOMIC_RED
                    unsigned long long* accl = p.acc + ((size_t)((fwd - 1u) & 1u) * p.L + l) * D + 4 * tid;
                    float4 b2 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (xvalid) {
                        const float4 p0 = *reinterpret_cast<const float4*>(part + 4 * tid);
                        const float4 p1 = *reinterpret_cast<const float4*>(part + D + 4 * tid);
                        red_fix_add(accl + 0, p0.x + p1.x);
                        red_fix_add(accl + 1, p0.y + p1.y);
                        red_fix_add(accl + 2, p0.z + p1.z);
                        red_fix_add(accl + 3, p0.w + p1.w);
                        b2 = __ldg(reinterpret_cast<const float4*>(p.blob + p.proj2_b_off + (long long)l * p.layer_stride + 4 * tid));
                    }
                    stamp(ts + 9);
                    stamp_wait(ts + 13);
                    hop_arrive(hc + HC_PP * GV_HOP_STRIDE, tid);
                    // x2 = x1 + b + sum of the G partials: every CTA reads the finished accumulators itself
                    hop_wait(hc + HC_PP * GV_HOP_STRIDE, (lc + 1u) * (unsigned)G, tid, tmask, settle, hold_c, near);
                    if (xvalid) {
                        const unsigned long long full_count = (unsigned long long)(G & 0xff);
                        ulonglong2 w0 = ld_x2u64(accl), w1 = ld_x2u64(accl + 2);
                        uint32_t spins = 0;
                        while (tmask != 0u && (((w0.x & 0xffull) != full_count) || ((w0.y & 0xffull) != full_count) ||
                                               ((w1.x & 0xffull) != full_count) || ((w1.y & 0xffull) != full_count))) {
                            if (++spins > MEGA_SPIN_LIMIT) __trap();
                            w0 = ld_x2u64(accl);
                            w1 = ld_x2u64(accl + 2);
                        }
                        const float4 x1v = *reinterpret_cast<const float4*>(xres1 + 4 * tid);
                        xnext = make_float4((x1v.x + b2.x) + fix_value(w0.x), (x1v.y + b2.y) + fix_value(w0.y),
                                            (x1v.z + b2.z) + fix_value(w1.x), (x1v.w + b2.w) + fix_value(w1.y));
                    }
                    stamp(ts + 11);
                }
                lc += 1u;
#else
                    if (xvalid) {
                        const float4 p0 = *reinterpret_cast<const float4*>(part + 4 * tid);
                        const float4 p1 = *reinterpret_cast<const float4*>(part + D + 4 * tid);
                        st_tagged2(p.pp, cta * D + 4 * tid, p0.x + p1.x, p0.y + p1.y, tg + TG_PP);
                        st_tagged2(p.pp, cta * D + 4 * tid + 2, p0.z + p1.z, p0.w + p1.w, tg + TG_PP);
                    }
                    stamp(ts + 9);
                    stamp_wait(ts + 13);
#if GV_PP_COUNTER
                    hop_arrive(hc + HC_PP * GV_HOP_STRIDE, tid);
#else
                    bar_sync(1, MEGA_CONSUMERS);  // `part` / `gat` alias: all reads of `part` precede the gather below
#endif
                }
                // ---- RED: x2 = x1 + b + sum over CTAs of the partials (8 outputs per reducer CTA) ----
                if (cta < n_red) {
                    float b2 = 0.0f;
                    if (lane == 0) b2 = __ldg(p.blob + p.proj2_b_off + (long long)l * p.layer_stride + cta * 8 + warp);
#if GV_PP_COUNTER
                    hop_wait(hc + HC_PP * GV_HOP_STRIDE, (lc + 1u) * (unsigned)G, tid, tmask, settle, hold_c, near);
#endif
                    stamp(ts + 10);
                    {   // load q: 16 bytes {v, tag, v, tag} of source CTA q / 4, outputs 2 (q % 4), +1; three rounds in flight
                        const uint32_t tgp = tg + TG_PP;
                        const int nq = G * 4;
                        uint4 a[3];
                        const float* ap[3];
#pragma unroll
                        for (int rr = 0; rr < 3; ++rr) {
                            const int q = tid + rr * MEGA_CONSUMERS;
                            ap[rr] = p.pp + 2 * (size_t)((q < nq ? (q >> 2) : 0) * D + cta * 8 + 2 * (q & 3));
                            a[rr] = make_uint4(0u, tgp, 0u, tgp);
                            if (q < nq) a[rr] = ld_x16(ap[rr]);
                        }
                        uint32_t spins = 0;
                        while (!(tags_ok(a[0], tgp, tmask) && tags_ok(a[1], tgp, tmask) && tags_ok(a[2], tgp, tmask))) {
                            if (++spins > MEGA_SPIN_LIMIT) __trap();
#pragma unroll
                            for (int rr = 0; rr < 3; ++rr)
                                if (tid + rr * MEGA_CONSUMERS < nq) a[rr] = ld_x16(ap[rr]);
                        }
#pragma unroll
                        for (int rr = 0; rr < 3; ++rr) {
                            const int q = tid + rr * MEGA_CONSUMERS;
                            if (q < nq)
                                *reinterpret_cast<float2*>(gat + (q >> 2) * 8 + 2 * (q & 3)) =
                                    make_float2(__uint_as_float(a[rr].x), __uint_as_float(a[rr].z));
                        }
                    }
                    if (hold_c && tid == 0) *hold_c = 0;  // hop data is in registers: the producer may stream again
                    bar_sync(1, MEGA_CONSUMERS);
                    {
                        float s = 0.0f;
                        for (int c = lane; c < G; c += 32) s += gat[c * 8 + warp];
                        s = warp_sum(s);
                        if (lane == 0) {
                            const int col = cta * 8 + warp;
                            st_tagged(p.x2, col, (xres1[col] + b2) + s, tg + TG_X2);
                        }
                    }
                    stamp(ts + 11);
                    hop_arrive(hc + HC_X2 * GV_HOP_STRIDE, tid);
                }
                lc += 1u;
#endif
            }
            // ---- HEAD: ln_f -> final_norm -> latent z ; logits = z . mel_head^T + b ----
            {
                const uint32_t tg = tbase + (uint32_t)GV_TAGS_PER_LAYER * (uint32_t)p.L;
                const int ts = p.L * GV_TRACE_PER_LAYER;
#if GV_ATOMIC_RED
                lat = xnext;
#else
                hop_wait(hc + HC_X2 * GV_HOP_STRIDE, lc * (unsigned)n_red, tid, tmask, settle, hold_c, near);
                if (xvalid) ld_tagged_vec<4>(p.x2, 4 * tid, tg - (uint32_t)GV_TAGS_PER_LAYER + TG_X2, tmask, &lat.x);
#endif
                if (hold_c && tid == 0) *hold_c = 0;
                stamp(ts + 0);
                const float* lnp = tile_wait(ring, cs, cs.gt, lane);  // all warps read the parameter tile
                ln_quad(lat, xvalid, D, lnp, lnp + D, red, tid);
                bar_sync(1, MEGA_CONSUMERS);  // `red` is reused by the second LayerNorm
                ln_quad(lat, xvalid, D, lnp + 2 * D, lnp + 3 * D, red, tid);
                if (xvalid) *reinterpret_cast<float4*>(xo + 4 * tid) = lat;
                bar_sync(1, MEGA_CONSUMERS);
                if (warp == 0) tile_release(ring, cs.gt, lane, 4u);
                cs.gt += 1u;
                gemv_dot<NXV, false>(ring, cs, nun[PH_HEAD], xo, warp, lane,
                              [&](int u, float dot, float c2, float) { st_tagged(p.lg, ubeg[PH_HEAD] + u, dot + c2, tg); });
                cs.gt += (uint32_t)ntl[PH_HEAD];
                stamp(ts + 1);
                hop_arrive(hc + HC_LG * GV_HOP_STRIDE, tid);
                t_lg += (unsigned)G;
                hop_wait(hc + HC_LG * GV_HOP_STRIDE, t_lg, tid, tmask, settle, hold_c, near);
                for (int e = 2 * tid; e < p.V; e += 2 * MEGA_CONSUMERS) {
                    if (e + 1 < p.V) {
                        const float2 v = ld_tagged2(p.lg, e, tg, tmask);
                        slog[e] = v.x;
                        slog[e + 1] = v.y;
                    } else {
                        slog[e] = ld_tagged1(p.lg, e, tg, tmask);
                    }
                }
                if (hold_c && tid == 0) *hold_c = 0;
                stamp(ts + 2);
            }
        } else {
            // logits / latent left pending by the prefill (per-op kernels; plain arrays)
            for (int e = tid; e < p.V; e += MEGA_CONSUMERS) slog[e] = ldcg(p.pend_logits + e);
            if (xvalid) lat = ldcg4(p.pend_latent + 4 * tid);
        }
        bar_sync(1, MEGA_CONSUMERS);  // slog complete; attention / gather scratch (aliasing `keys`) is dead
        // ------------- sample + emit (every CTA computes the same token) -------------
        int tok = sample_token([&](int e) { return slog[e]; }, seen, scfg, p.noise ? p.noise + (size_t)i * p.V : nullptr, p.seed,
                               (uint32_t)n, 0u, keys, fscr, iscr, tid, [] { bar_sync(1, MEGA_CONSUMERS); });
        stamp(p.L * GV_TRACE_PER_LAYER + 3);
        if (p.forced) tok = (int)p.forced[i];
        if (!p.ignore_eos && finished) tok = p.stop_token;
        if (cta == 0) {
            if (tid == 0) p.ids_out[i] = tok;
            if (xvalid) *reinterpret_cast<float4*>(p.latents_out + (size_t)i * D + 4 * tid) = lat;
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
        ctl[1] = (int)cs.gt;
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
    off += mega_scratch_bytes(D);
    off += (size_t)Vpad * sizeof(float) + 3 * (size_t)D * sizeof(float);
    off += 32 * sizeof(uint64_t);
    off += (32 + 64 + 16) * sizeof(float) + 16 * sizeof(int) + 8 * sizeof(int);
    off += Vpad;
    return (off + 15) & ~size_t(15);
}

template <int NXV, bool TRACE>
static cudaError_t launch_nxv2(const MegaParams& p, int grid, size_t smem, cudaStream_t st) {
    cudaError_t e =
        cudaFuncSetAttribute(decode_mega_kernel<NXV, TRACE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    MegaParams pp = p;
    void* args[] = {&pp};
    return cudaLaunchCooperativeKernel((void*)decode_mega_kernel<NXV, TRACE>, dim3(grid), dim3(MEGA_THREADS), args, smem, st);
}
template <int NXV>
static cudaError_t launch_nxv(const MegaParams& p, int grid, size_t smem, cudaStream_t st) {
    return p.trace != nullptr ? launch_nxv2<NXV, true>(p, grid, smem, st) : launch_nxv2<NXV, false>(p, grid, smem, st);
}

cudaError_t launch_decode_mega(const MegaParams& p, int grid, cudaStream_t st) {
    const int hd = p.D / p.H;
    if (p.D % 128 || p.D > 1024 || !(hd == 32 || hd == 64 || hd == 128 || hd == 256)) return cudaErrorInvalidValue;
    if ((size_t)grid * 8 * sizeof(float) > 9728 || p.H * 8 > grid) return cudaErrorInvalidValue;
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
