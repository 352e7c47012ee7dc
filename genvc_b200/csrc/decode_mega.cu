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
#define GV_RING_NSLOT GV_MEGA_NSLOT
#define GV_MEGA_NS mega1
#define GV_UNIFORM_TILE_WAIT 1
// weight loads of a phase hoisted above the hop that delivers its activations (gemv_preload / gemv_finish); switchable for A/B
#ifndef GV_PRE_QKV
#define GV_PRE_QKV 1
#endif
#ifndef GV_PRE_PROJ
#define GV_PRE_PROJ 1
#endif
#ifndef GV_PRE_HEAD
#define GV_PRE_HEAD 1
#endif
#include "mega_dev.cuh"

namespace gv {
using namespace mega1;

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
// Projected-value variant (PVW): attn c_proj applied to the VALUE of the position being decoded, per head:
//   vw[h][u] = sum_{d < hd} v[h hd + d] W_proj[h hd + d][unit u]
// (the attention output of the step is then sum_j p_{h,j} vw_j[h][u] over the cached projected values: the
// projection is linear, so it commutes with the softmax-weighted sum, and it no longer waits for the softmax).
// Unit u of the phase is handled by warp u (<= 8 units); a tile is read by the four warps 4t .. 4t+3, each arrives
// once.  Lane layout of a head's K range as in the attention item (AttLane): conflict-free vector loads.
// epi(u, h, value) on lane 8 (h % 4) of warp u; bias(u, c2) once per unit on lane 0.
// ---------------------------------------------------------------------------------------------
// L2 load that stays where it is written (the compiler sinks a plain __ldcg to its first use; these are issued early on
// purpose, microseconds before the values are needed)
__device__ __forceinline__ float ld_cg_early(const float* p) {
    float v;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

// `fill_n` > 0 (first forward after a prefill): the projected values of the fill_n cached positions are computed here
// too, while this CTA's attn c_proj columns are in shared memory anyway -- warp w takes positions w, w + 8, ..., all
// units, values straight from the V cache (vfill: [H][S_max][hd]), results to out_fill[(pos H + h) 8 + unit].
#ifdef GV_PROG
// debug: per-warp progress markers (tools/hang_dump.py reads them from a side stream while a launch is stuck)
__device__ unsigned g_prog[160 * 8 * 4];
#define GV_MARK(code) do { if ((threadIdx.x & 31) == 0) { unsigned* g_ = g_prog + (blockIdx.x * 8 + (threadIdx.x >> 5)) * 4; g_[0] = (code); g_[1] = (unsigned)lc; } } while (0)
#else
#define GV_MARK(code) do { } while (0)
#endif
template <int NXV, int HD, class Epi, class Bias>
__device__ __forceinline__ void gemv_heads(const Ring& ring, const Cons& cs, int nunits, const float* xs, int warp, int lane, Epi epi,
                                           Bias bias, int fill_n, const float* vfill, int S_max, float* out_fill) {
    constexpr int D = NXV * 128, UF = D + 4, H = D / HD;
    using L = AttLane<HD>;
    constexpr int VEC = L::VEC, NCH = L::NCH, DPL = L::DPL;
    const int ntiles = (nunits + UPT - 1) / UPT;
    if (fill_n > 0) {  // uniform over the CTA
        if (ntiles > 0) tile_ready_wait_u(ring, cs.gt + (uint32_t)(ntiles - 1));
        __syncwarp();
        // (Once per generated sequence; the host enables this variant only for generations long enough to amortise it.
        // A software-pipelined version with two positions in flight per warp needed 64 more live registers and slowed
        // the whole kernel down.)
        for (int pos = warp; pos < fill_n; pos += MEGA_WARPS) {
            float vr[H][DPL];
#pragma unroll
            for (int h = 0; h < H; ++h)
#pragma unroll
                for (int c = 0; c < NCH; ++c)
                    ld_vec<VEC>(vfill + ((size_t)h * S_max + pos) * HD + (c * 32 + lane) * VEC, vr[h] + c * VEC);
            for (int u = 0; u < nunits; ++u) {
                const float* col = slot_ptr(ring, cs.gt + (uint32_t)(u >> 2)) + (u & 3) * UF;
#pragma unroll
                for (int hg = 0; hg < H; hg += 4) {
                    float tot[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float a = 0.0f;
                        if (hg + q < H) {
#pragma unroll
                            for (int c = 0; c < NCH; ++c) {
                                const int off = (hg + q) * HD + (c * 32 + lane) * VEC;
#pragma unroll
                                for (int e = 0; e < VEC; ++e) a = fmaf(col[off + e], vr[(hg + q) < H ? (hg + q) : 0][c * VEC + e], a);
                            }
                        }
                        tot[q] = a;
                    }
                    const bool up16 = (lane & 16) != 0, up8 = (lane & 8) != 0;
                    const float k0 = up16 ? tot[2] : tot[0], s0 = up16 ? tot[0] : tot[2];
                    const float k1 = up16 ? tot[3] : tot[1], s1 = up16 ? tot[1] : tot[3];
                    const float h0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 16);
                    const float h1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 16);
                    const float keep = up8 ? h1 : h0, send = up8 ? h0 : h1;
                    float v = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                    v += __shfl_xor_sync(0xffffffffu, v, 4);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    const int q = lane >> 3;
                    if ((lane & 7) == 0 && hg + q < H) __stcg(out_fill + ((size_t)pos * H + hg + q) * 8 + u, v);
                }
            }
        }
        BAR1();  // every warp read both tiles: nobody releases one before all are done
    }
    const int t = warp >> 2;
    if (t >= ntiles) return;
    tile_ready_wait_u(ring, cs.gt + (uint32_t)t);
    __syncwarp();
    if (warp < nunits) {
        const float* col = slot_ptr(ring, cs.gt + (uint32_t)t) + (warp & 3) * UF;
        if (lane == 0) bias(warp, col[D]);
#pragma unroll
        for (int hg = 0; hg < H; hg += 4) {
            float tot[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float a = 0.0f;
                if (hg + q < H) {
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        const int off = (hg + q) * HD + (c * 32 + lane) * VEC;
                        if constexpr (VEC == 4) {
                            const float4 wv = *reinterpret_cast<const float4*>(col + off);
                            const float4 xv = *reinterpret_cast<const float4*>(xs + off);
                            a = fmaf(wv.x, xv.x, a);
                            a = fmaf(wv.y, xv.y, a);
                            a = fmaf(wv.z, xv.z, a);
                            a = fmaf(wv.w, xv.w, a);
                        } else if constexpr (VEC == 2) {
                            const float2 wv = *reinterpret_cast<const float2*>(col + off);
                            const float2 xv = *reinterpret_cast<const float2*>(xs + off);
                            a = fmaf(wv.x, xv.x, a);
                            a = fmaf(wv.y, xv.y, a);
                        } else {
                            a = fmaf(col[off], xs[off], a);
                        }
                    }
                }
                tot[q] = a;
            }
            // transposing butterfly: lanes [8q, 8q+8) end up reducing tot[q]
            const bool up16 = (lane & 16) != 0, up8 = (lane & 8) != 0;
            const float k0 = up16 ? tot[2] : tot[0], s0 = up16 ? tot[0] : tot[2];
            const float k1 = up16 ? tot[3] : tot[1], s1 = up16 ? tot[1] : tot[3];
            const float h0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 16);
            const float h1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 16);
            const float keep = up8 ? h1 : h0, send = up8 ? h0 : h1;
            float v = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            const int q = lane >> 3;
            if ((lane & 7) == 0 && hg + q < H) epi(warp, hg + q, v);
        }
    }
    tile_release(ring, cs.gt + (uint32_t)t, lane);
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <int NXV, bool TRACE, bool PVW>
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
    float* vwn = reinterpret_cast<float*>(smem_raw + off);  // PVW: [H][8] projected value of the position being decoded
    float* pbias = vwn + 32 * 8;                             // PVW: [8] attn c_proj bias of this CTA's columns
    off += (32 * 8 + 8) * sizeof(float);
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
#ifdef GV_PROG
        for (int q = 0; q < 8; ++q) bar_cnt()[q] = 0u;
#endif
        ctl[0] = 0;
        ctl[1] = 0;
        ctl[2] = 0;
        ctl[3] = 0;
        ctl[4] = 0;
        mbar_fence_init();
    }
    for (int i = tid_all; i < p.Vpad; i += MEGA_THREADS) seen[i] = p.seen[i];
    __syncthreads();

    // The state is read from p.st / p.seen and written to p.st_out / p.seen_out (double-buffered by the host): CTA 0
    // may finish a launch that runs no forward before a late CTA has read the state.
    const GenState* st = p.st;
    const int had_pending = st->has_pending;
    const int n_start = st->n_emitted;
    if (st->done) {  // uniform: nothing to do
        if (cta == 0) {
            for (int q = tid_all; q < p.Vpad; q += MEGA_THREADS) p.seen_out[q] = seen[q];
            if (tid_all == 0) {
                GenState* so = p.st_out;
                so->n_emitted = n_start;
                so->done = 1;
                so->has_pending = st->has_pending;
                so->P = st->P;
                so->B = st->B;
                so->finished[0] = st->finished[0];
                so->last_tok[0] = st->last_tok[0];
                p.status[1] = 1;
            }
        }
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
                ok = produce_forward(pr, p.stream, p.blob, p.lnf_off, p.L, D, sd, cta);
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
                    // this warp's c_attn units go to registers while the x2 hop of the previous block is in flight
                    GemvRegs<NXV, GemvKU<NXV>::QKV> wq;
#if GV_PRE_QKV
                    gemv_preload(ring, cs, nun[PH_QKV], warp, lane, wq);
                    cs.gt += (uint32_t)ntl[PH_QKV];
#endif
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
                        ld_tagged_vec_u<4>(p.x2, 4 * tid, xvalid, tg - (uint32_t)GV_TAGS_PER_LAYER + TG_X2, tmask, &x.x);
                        if (!xvalid) x = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#endif
                    if (xvalid) *reinterpret_cast<float4*>(xres0 + 4 * tid) = x;
                    if (hold_c && tid == 0) *hold_c = 0;  // hop data is in registers: the producer may stream again
                    stamp(ts + 0);
                    GV_MARK(1u);
                    stats_partial(x, xvalid, shift1, red, lane, warp);
                    BAR1();
                    GV_MARK(2u);
                    float mean, rstd;
                    stats_finish(red, inv_d, shift1, mean, rstd);
                    shift1 = mean;
                    stamp(ts + 1);
#if !GV_PRE_QKV
                    gemv_preload(ring, cs, nun[PH_QKV], warp, lane, wq);
                    cs.gt += (uint32_t)ntl[PH_QKV];
#endif
                    gemv_finish(nun[PH_QKV], xres0, warp, lane, wq, [&](int u, float dot, float c2, float c1) {
                        st_tagged(p.xq, ubeg[PH_QKV] + u, fmaf(rstd, fmaf(-mean, c1, dot), c2), tg + TG_XQ);
                    });
                    GV_MARK(3u);
                    stamp(ts + 2);
                }
                if constexpr (PVW) {
                // ---- projected-value variant: scores-only items, attn c_proj applied to the new VALUE (off the softmax's
                // critical path), softmax + weighted sum of the cached projected values in every CTA for its own columns ----
                hop_arrive(hc + HC_XQ * GV_HOP_STRIDE, tid);  // q | k | v of this layer are on their way (hint; tags validate)
                if (cta < n_items) {
                    const int h = cta / nsplit, sp = cta % nsplit;
                    const int j0 = sp * chunk, j1 = min(S, j0 + chunk);
                    float* kh = kc + (size_t)h * p.S_max * HD;
                    float* vh = vc + (size_t)h * p.S_max * HD;
#define GV_ATT_CASE(hd)                                                                                                       \
    case hd:                                                                                                                  \
        score_item<hd>(kh, vh, p.xq, D, h, j0, j1, S, tg + TG_XQ, att_sc + 16, tid, p.sbuf + 2 * (size_t)h * p.S_max, tg + TG_AO, \
                       tmask);                                                                                                \
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
                {
                    float* vw_slice = p.vw + ((size_t)l * G + cta) * (size_t)p.S_max * H * 8;  // [pos][h][8]
                    float* psm = part;                         // [heads of the pass][S] probabilities (scratch region)
                    const int psm_cap = 9728 / 4 - 64;
                    float* redsm = part + psm_cap;             // [8 warps][8 columns]
                    const int HG = max(1, min(H, psm_cap / S));
                    const int colj = tid & 7, g = tid >> 3;
                    // request this thread's first eight cached projected values now: their addresses depend only on the
                    // position, and the rows come from HBM (the weight stream sweeps L2 once per token)
                    // first forward after a prefill: the cached positions have no projected values yet (computed below)
                    const int fill_n = (p.vw_fill && fwd == 1u) ? S - 1 : 0;
                    const bool prefetched = HG == H && fill_n == 0;
                    float vwv[8];
                    if (prefetched) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const int qq = g + 32 * k;
                            vwv[k] = (qq < (S - 1) * H) ? ld_cg_early(vw_slice + (size_t)qq * 8 + colj) : 0.0f;
                        }
                    }
                    // v of the position being decoded -> xo (the input of the per-head attn c_proj GEMV)
                    hop_wait(hc + HC_XQ * GV_HOP_STRIDE, (lc + 1u) * (unsigned)G, tid, tmask, settle, hold_c, near);
                    stamp(ts + 14);
                    {
                        float4 v4;
                        ld_tagged_vec_u<4>(p.xq, 2 * D + 4 * tid, xvalid, tg + TG_XQ, tmask, &v4.x);
                        if (xvalid) *reinterpret_cast<float4*>(xo + 4 * tid) = v4;
                    }
                    if (hold_c && tid == 0) *hold_c = 0;
                    BAR1();
                    stamp(ts + 15);
                    const int np = nun[PH_PROJ];
                    float* vw_new = vw_slice + (size_t)(S - 1) * H * 8;
                    // The scores are usually complete by now (the items need q only): see them arrive, request this warp's
                    // first 128 scores, and let the loads fly while the per-head GEMV runs.
                    // (no counter wait: a score that is not there yet is spun on below -- the items only need q, so the
                    // scores land well before the value hop above has completed)
                    stamp(ts + 16);
                    uint2 sa[4];
                    if (warp < min(HG, H)) {
                        const float* sb = p.sbuf + 2 * (size_t)warp * p.S_max;
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (lane + 32 * k < S) sa[k] = ld_x8(sb + 2 * (size_t)(lane + 32 * k));
                    }
                    auto epi = [&](int u, int h, float v) {
                        vwn[h * 8 + u] = v;
                        __stcg(vw_new + h * 8 + u, v);
                    };
                    auto bia = [&](int u, float c2) { pbias[u] = c2; };
#define GV_HEADS_CASE(hd)                                                                    \
    case hd:                                                                                 \
        if constexpr (D % hd == 0 && D / hd <= 32)                                           \
            gemv_heads<NXV, hd>(ring, cs, np, xo, warp, lane, epi, bia, fill_n, vc, p.S_max, vw_slice); \
        break;
                    switch (HD) {
                        GV_HEADS_CASE(32) GV_HEADS_CASE(64) GV_HEADS_CASE(128) GV_HEADS_CASE(256)
                        default: break;
                    }
#undef GV_HEADS_CASE
                    cs.gt += (uint32_t)ntl[PH_PROJ];
                    stamp(ts + 4);
                    // scores of all heads -> softmax -> weighted sum of the projected values (this CTA's columns)
                    float acc = 0.0f;
                    const int hshift = (H & (H - 1)) == 0 ? __ffs(H) - 1 : -1;  // pairs -> (position, head) without a division
                    for (int h0 = 0; h0 < H; h0 += HG) {
                        const int nh = min(HG, H - h0);
                        // one head per warp: scores straight from L2 (four tagged loads in flight per lane), fp32 softmax as
                        // HF computes it, probabilities -> shared memory
                        for (int hh = warp; hh < nh; hh += MEGA_WARPS) {
                            float* ph = psm + hh * S;
                            const float* sb = p.sbuf + 2 * (size_t)(h0 + hh) * p.S_max;
                            const uint32_t tgs = tg + TG_AO;
                            float m = -INFINITY;
                            for (int j0 = 0; j0 < S; j0 += 128) {  // uniform trip count: the warp polls as one
                                const int j4 = j0 + lane;
                                uint2 a[4];
                                const bool pre = h0 == 0 && hh == warp && j0 == 0;  // requested before the GEMV
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    a[k] = make_uint2(0u, tgs);
                                    if (j4 + 32 * k < S) a[k] = pre ? sa[k] : ld_x8(sb + 2 * (size_t)(j4 + 32 * k));
                                }
                                uint32_t spins = 0;
                                for (;;) {
                                    bool ok = true;
#pragma unroll
                                    for (int k = 0; k < 4; ++k) ok = ok && ((a[k].y ^ tgs) & tmask) == 0u;
                                    if (__all_sync(0xffffffffu, ok)) break;
                                    GV_SPIN(spins, __LINE__, 0u, 0u);
#pragma unroll
                                    for (int k = 0; k < 4; ++k)
                                        if (j4 + 32 * k < S && ((a[k].y ^ tgs) & tmask) != 0u) a[k] = ld_x8(sb + 2 * (size_t)(j4 + 32 * k));
                                }
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    if (j4 + 32 * k < S) {
                                        const float sv = __uint_as_float(a[k].x);
                                        ph[j4 + 32 * k] = sv;
                                        m = fmaxf(m, sv);
                                    }
                                }
                            }
                            m = warp_max(m);
                            float sum = 0.0f;
                            for (int jj = lane; jj < S; jj += 32) {
                                const float e = expf(ph[jj] - m);
                                ph[jj] = e;
                                sum += e;
                            }
                            sum = warp_sum(sum);
                            const float inv = 1.0f / sum;
                            for (int jj = lane; jj < S; jj += 32) ph[jj] *= inv;
                        }
                        if (hold_c && tid == 0) *hold_c = 0;
                        BAR1();
                        stamp(ts + 18);
                        // cached positions: (j, head) pairs dealt to the 32 thread groups; 8 consecutive threads read one
                        // 32-byte row segment of the slice.  The first eight pairs of every thread were requested from
                        // the cache at the start of the section (single pass only).
                        const int npairs = (S - 1) * nh;
                        int q = g;
                        if (prefetched) {
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                const int qq = g + 32 * k;
                                if (qq < npairs) {  // (single pass: nh == H)
                                    const int jj = hshift >= 0 ? (qq >> hshift) : qq / nh, hh = qq - jj * nh;
                                    acc = fmaf(psm[hh * S + jj], vwv[k], acc);
                                }
                            }
                            q = g + 256;
                        }
                        for (; q < npairs; q += 32) {
                            const int jj = q / nh, hh = q - jj * nh;
                            acc = fmaf(psm[hh * S + jj], ldcg(vw_slice + ((size_t)jj * H + h0 + hh) * 8 + colj), acc);
                        }
                        if (g == 0)  // the position being decoded: its projected value is still in shared memory
                            for (int hh = 0; hh < nh; ++hh) acc = fmaf(psm[hh * S + S - 1], vwn[(h0 + hh) * 8 + colj], acc);
                        BAR1();  // psm is rewritten by the next pass / redsm follows
                        stamp(ts + 19);
                    }
                    acc += __shfl_xor_sync(0xffffffffu, acc, 8);
                    acc += __shfl_xor_sync(0xffffffffu, acc, 16);
                    if (lane < 8) redsm[warp * 8 + lane] = acc;
                    BAR1();
                    if (tid < np) {
                        float o = 0.0f;
#pragma unroll
                        for (int w = 0; w < MEGA_WARPS; ++w) o += redsm[w * 8 + tid];
                        const int col = ubeg[PH_PROJ] + tid;
                        st_tagged(p.x1, col, xres0[col] + (o + pbias[tid]), tg + TG_X1);
                    }
                    stamp(ts + 5);
                    stamp_wait(ts + 12);
                    hop_arrive(hc + HC_X1 * GV_HOP_STRIDE, tid);
                }
                } else {
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
                    GV_MARK(4u);
                    hop_arrive(hc + HC_AO * GV_HOP_STRIDE, tid);
                }
                t_ao += (unsigned)n_items;
                stamp(ts + 3);
                // ---- PROJ: merge attention partials -> o ; x1 = x + o . W_proj + b ----
                {
                    // the attn c_proj unit of this warp goes to registers while the CTA waits for the attention items
                    GemvRegs<NXV, GemvKU<NXV>::PROJ> wpj;
#if GV_PRE_PROJ
                    gemv_preload(ring, cs, nun[PH_PROJ], warp, lane, wpj);
                    cs.gt += (uint32_t)ntl[PH_PROJ];
#endif
                    GV_MARK(5u);
                    hop_wait(hc + HC_AO * GV_HOP_STRIDE, t_ao, tid, tmask, settle, hold_c, near_ao);
                    GV_MARK(6u);
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
                                    // the warp leaves the poll as one (xvalid is warp-uniform: D is a multiple of 128)
                                    while (!__all_sync(0xffffffffu, tags_ok(a[q], tga, tmask) && tags_ok(b[q], tga, tmask) &&
                                                                        tags_ok(c[q], tga, tmask))) {
                                        GV_SPIN(spins, __LINE__, 0u, 0u);
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
                    BAR1();
#if !GV_PRE_PROJ
                    gemv_preload(ring, cs, nun[PH_PROJ], warp, lane, wpj);
                    cs.gt += (uint32_t)ntl[PH_PROJ];
#endif
                    gemv_finish(nun[PH_PROJ], xo, warp, lane, wpj, [&](int u, float dot, float c2, float) {
                        const int col = ubeg[PH_PROJ] + u;
                        st_tagged(p.x1, col, xres0[col] + (dot + c2), tg + TG_X1);
                    });
                    stamp(ts + 5);
                    stamp_wait(ts + 12);
                    hop_arrive(hc + HC_X1 * GV_HOP_STRIDE, tid);
                }
                }
                // ---- FC + P2: u = gelu_new(LN2(x1) . W_fc + b) (kept in this CTA) -> partial of u . W_proj2 ----
                {
                    // the c_fc units of this warp go to registers while the x1 hop is in flight (the weights do not depend on it)
                    GemvRegs<NXV, GemvKU<NXV>::FC> wfc;
                    gemv_preload(ring, cs, nun[PH_FC], warp, lane, wfc);
                    cs.gt += (uint32_t)ntl[PH_FC];
                    hop_wait(hc + HC_X1 * GV_HOP_STRIDE, (lc + 1u) * (unsigned)G, tid, tmask, settle, hold_c, near);
                    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                    ld_tagged_vec_u<4>(p.x1, 4 * tid, xvalid, tg + TG_X1, tmask, &x.x);
                    if (xvalid) *reinterpret_cast<float4*>(xres1 + 4 * tid) = x;
                    if (hold_c && tid == 0) *hold_c = 0;  // hop data is in registers: the producer may stream again
                    stamp(ts + 6);
                    stats_partial(x, xvalid, shift2, red + 16, lane, warp);
                    BAR1();
                    float mean, rstd;
                    stats_finish(red + 16, inv_d, shift2, mean, rstd);
                    shift2 = mean;
                    stamp(ts + 7);
                    gemv_finish(nun[PH_FC], xres1, warp, lane, wfc, [&](int u, float dot, float c2, float c1) {
                        us[u] = gelu_new(fmaf(rstd, fmaf(-mean, c1, dot), c2));
                    });
                    BAR1();
                    stamp(ts + 8);
                    gemv_outer<NXV>(ring, cs, nun[PH_P2], us, tid, lane, warp, part);
                    cs.gt += (uint32_t)ntl[PH_P2];
                    BAR1();
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
                            GV_SPIN(spins, __LINE__, 0u, 0u);
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
                    BAR1();  // `part` / `gat` alias: all reads of `part` precede the gather below
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
                        while (!__all_sync(0xffffffffu, tags_ok(a[0], tgp, tmask) && tags_ok(a[1], tgp, tmask) && tags_ok(a[2], tgp, tmask))) {
                            GV_SPIN(spins, __LINE__, 0u, 0u);
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
                    BAR1();
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
                // the logits-head units (they follow the LayerNorm-parameter tile in the stream) go to registers before the hop
                GemvRegs<NXV, GemvKU<NXV>::HEAD> wh;
#if GV_PRE_HEAD
                gemv_preload(ring, Cons{cs.gt + 1u, cs.wacc}, nun[PH_HEAD], warp, lane, wh);
#endif
#if GV_ATOMIC_RED
                lat = xnext;
#else
                hop_wait(hc + HC_X2 * GV_HOP_STRIDE, lc * (unsigned)n_red, tid, tmask, settle, hold_c, near);
                ld_tagged_vec_u<4>(p.x2, 4 * tid, xvalid, tg - (uint32_t)GV_TAGS_PER_LAYER + TG_X2, tmask, &lat.x);
                if (!xvalid) lat = make_float4(0.f, 0.f, 0.f, 0.f);
#endif
                if (hold_c && tid == 0) *hold_c = 0;
                stamp(ts + 0);
                const float* lnp = tile_wait(ring, cs, cs.gt, lane);  // all warps read the parameter tile
                ln_quad(lat, xvalid, D, lnp, lnp + D, red, tid);
                BAR1();  // `red` is reused by the second LayerNorm
                ln_quad(lat, xvalid, D, lnp + 2 * D, lnp + 3 * D, red, tid);
                if (xvalid) *reinterpret_cast<float4*>(xo + 4 * tid) = lat;
                BAR1();
                if (warp == 0) tile_release(ring, cs.gt, lane, 4u);
                cs.gt += 1u;
#if !GV_PRE_HEAD
                gemv_preload(ring, Cons{cs.gt, cs.wacc}, nun[PH_HEAD], warp, lane, wh);
#endif
                gemv_finish(nun[PH_HEAD], xo, warp, lane, wh,
                                 [&](int u, float dot, float c2, float) { st_tagged(p.lg, ubeg[PH_HEAD] + u, dot + c2, tg); });
                cs.gt += (uint32_t)ntl[PH_HEAD];
                stamp(ts + 1);
                hop_arrive(hc + HC_LG * GV_HOP_STRIDE, tid);
                t_lg += (unsigned)G;
                hop_wait(hc + HC_LG * GV_HOP_STRIDE, t_lg, tid, tmask, settle, hold_c, near);
                for (int e0 = 0; e0 < p.Vpad; e0 += 2 * MEGA_CONSUMERS) {  // uniform trip count; lg holds Vpad (even) tagged words
                    const int e = e0 + 2 * tid;
                    const bool v0 = e < p.V, v1 = e + 1 < p.V;
                    uint4 a = make_uint4(0u, tg, 0u, tg);
                    uint32_t spins = 0;
                    for (;;) {  // the warp leaves the poll as one (see ld_tagged_vec_u)
                        bool ok = true;
                        if (v0) {
                            a = ld_x16(p.lg + 2 * (size_t)e);
                            ok = ((a.y ^ tg) & tmask) == 0u && (!v1 || ((a.w ^ tg) & tmask) == 0u);
                        }
                        if (__all_sync(0xffffffffu, ok)) break;
                        GV_SPIN(spins, __LINE__, (unsigned)e, tg - a.y);
                    }
                    if (v0) slog[e] = __uint_as_float(a.x);
                    if (v1) slog[e + 1] = __uint_as_float(a.z);
                }
                if (hold_c && tid == 0) *hold_c = 0;
                stamp(ts + 2);
            }
        } else {
            // logits / latent left pending by the prefill (per-op kernels; plain arrays)
            for (int e = tid; e < p.V; e += MEGA_CONSUMERS) slog[e] = ldcg(p.pend_logits + e);
            if (xvalid) lat = ldcg4(p.pend_latent + 4 * tid);
        }
        BAR1();  // slog complete; attention / gather scratch (aliasing `keys`) is dead
        // ------------- sample + emit (every CTA computes the same token) -------------
        int tok = sample_token([&](int e) { return slog[e]; }, seen, scfg, p.noise ? p.noise + (size_t)i * p.V : nullptr, p.seed,
                               (uint32_t)n, 0u, keys, fscr, iscr, tid, [] { BAR1(); });
        stamp(p.L * GV_TRACE_PER_LAYER + 3);
        if (p.forced) {
            const long long f = p.forced[i];
            tok = (f >= 0 && f < (long long)p.V) ? (int)f : p.stop_token;  // out-of-range ids never index seen[] / mel_emb
            if (tok != (int)f && cta == 0 && tid == 0) atomicOr(p.bad_ids, 1);
        }
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
        BAR1();  // seen[] update visible to the next step's sampler; slog / keys free
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
        BAR1();
        for (int q = tid; q < p.Vpad; q += MEGA_CONSUMERS) p.seen_out[q] = seen[q];
        if (tid == 0) {
            GenState* so = p.st_out;
            so->n_emitted = n;
            so->done = done;
            so->has_pending = 0;
            so->P = st->P;
            so->B = st->B;
            so->finished[0] = finished;
            so->last_tok[0] = last_tok;
            p.status[0] = emitted;
            p.status[1] = done;
            if (atomicExch(p.bad_ids, 0) != 0) p.status[2] = 1;  // an id was clamped since the last status (embedding kernels / forced ids)
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
    off += (32 * 8 + 8) * sizeof(float);
    off += Vpad;
    return (off + 15) & ~size_t(15);
}

template <int NXV, bool TRACE, bool PVW>
static cudaError_t launch_nxv2(const MegaParams& p, int grid, size_t smem, cudaStream_t st) {
    cudaError_t e =
        cudaFuncSetAttribute(decode_mega_kernel<NXV, TRACE, PVW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    MegaParams pp = p;
    void* args[] = {&pp};
    return cudaLaunchCooperativeKernel((void*)decode_mega_kernel<NXV, TRACE, PVW>, dim3(grid), dim3(MEGA_THREADS), args, smem, st);
}
template <int NXV>
static cudaError_t launch_nxv(const MegaParams& p, int grid, size_t smem, cudaStream_t st) {
    if (p.vw != nullptr)
        return p.trace != nullptr ? launch_nxv2<NXV, true, true>(p, grid, smem, st) : launch_nxv2<NXV, false, true>(p, grid, smem, st);
    return p.trace != nullptr ? launch_nxv2<NXV, true, false>(p, grid, smem, st) : launch_nxv2<NXV, false, false>(p, grid, smem, st);
}

cudaError_t launch_decode_mega(const MegaParams& p, int grid, cudaStream_t st) {
    const int hd = p.D / p.H;
    if (p.D % 128 || p.D > 1024 || !(hd == 32 || hd == 64 || hd == 128 || hd == 256)) return cudaErrorInvalidValue;
    if ((size_t)grid * 8 * sizeof(float) > 9728 || p.H * 8 > grid) return cudaErrorInvalidValue;
    if (grid < 128 || (p.V + grid - 1) / grid > 16) return cudaErrorInvalidValue;  // GemvKU: register sets per warp and phase
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

#ifdef GV_PROG
extern "C" int genvc_debug_prog_copy(unsigned* pinned_host, void* stream) {
    cudaMemcpyFromSymbolAsync(pinned_host + 160 * 8 * 4, gv::mega1::g_slip, sizeof(unsigned) * (8 + 8 * 64), 0, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    return (int)cudaMemcpyFromSymbolAsync(pinned_host, gv::g_prog, sizeof(unsigned) * 160 * 8 * 4, 0, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
}
#endif
#if GV_WAIT_DIAG
// debug builds only (tools/wait_diag.py): binds the pinned host buffer the wait records of the single-row kernel go to
extern "C" int genvc_debug_wait_bind(unsigned* pinned_host) {
    return (int)cudaMemcpyToSymbol(gv::mega1::g_wait_buf, &pinned_host, sizeof(pinned_host));
}
#endif
