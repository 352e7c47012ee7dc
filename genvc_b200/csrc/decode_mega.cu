// Fused persistent decode kernel (B = 1): the whole sample()/sample_stream() loop of
// layers/stream_generator.py:809-881 — per step the cached GPT-2 forward of
// layers/gpt_inference.py:92-112 (30 pre-LN blocks, ln_f, final_norm, mel_head) followed by the HF
// sampling chain — runs in ONE cooperative launch for up to n_steps tokens, tokens fed back on-chip.
//
// Why this shape.  A decode step at batch 1 reads every weight exactly once (1.516 GB fp32) and does
// 2 flops per weight: it is an HBM-streaming problem with 30 x 5 serial data dependencies.  So:
//   * one persistent CTA per SM (cooperative launch), each owning a fixed column slice of every
//     matrix; the slices are pre-packed so each CTA reads one contiguous byte stream (stream_layout.h);
//   * a dedicated producer thread per CTA streams that region HBM -> shared memory with 1-D bulk TMA
//     copies (cp.async.bulk, completion on mbarriers) into a 12-deep ring.  The weight stream does
//     not depend on activations, so the producer keeps running ahead across phase boundaries, grid
//     barriers and even across tokens: HBM never idles while the consumers synchronise;
//   * 8 consumer warps do the GEMV out of shared memory in fp32 FMA (x held in registers, one
//     128-bit LDS per 4 weights), LayerNorm / gelu_new / residual / bias fused around it;
//   * single-token attention is split over (head, key-range) items, online softmax with warp-shuffle
//     reductions, K/V read 128-bit coalesced straight from the cache (attn_decode.cuh); the new K/V
//     rows are written by the CTA that produced those QKV columns;
//   * phases are separated by a grid barrier (one atomic per CTA on an L2-resident counter);
//   * sampling is computed redundantly by every CTA (same data, same code => same token), so the
//     next token needs no broadcast and the next step's embedding starts without a barrier.
// No tensor cores: at M = 1 there is no reuse to feed them; fp32 keeps greedy parity.
#include "attn_decode.cuh"
#include "common.cuh"
#include "mega.cuh"
#include "sampling.cuh"
#include "stream_layout.h"

namespace gv {

#define MEGA_CONSUMERS 256
#define MEGA_THREADS (MEGA_CONSUMERS + 32)

struct ConsumerSync {
    __device__ __forceinline__ void operator()() const { bar_sync(1, MEGA_CONSUMERS); }
};

struct Ring {
    float* slots;
    uint64_t* full;
    uint64_t* empty;
    int nslot, slot_floats;
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
// device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& epoch, unsigned G, int tid) {
    bar_sync(1, MEGA_CONSUMERS);  // this CTA's global writes of the phase are issued
    if (tid == 0) {
        epoch += 1;
        const unsigned target = epoch * G;
        __threadfence();
        red_release_gpu_add(counter, 1u);
        unsigned spins = 0;
        while (ld_acquire_gpu(counter) < target) {
            if (++spins > (1u << 24)) __trap();
        }
    }
    bar_sync(1, MEGA_CONSUMERS);
}

// sum over the 256 consumer threads; scratch is double-buffered so one barrier per call suffices
__device__ __forceinline__ float block_sum(float v, float* scratch, int& flip, int tid) {
    v = warp_sum(v);
    float* s = scratch + flip * 8;
    flip ^= 1;
    if ((tid & 31) == 0) s[tid >> 5] = v;
    bar_sync(1, MEGA_CONSUMERS);
    float t = s[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) t += s[w];
    return t;
}

struct ConsumerState {
    uint32_t tile;  // tiles consumed so far (same in every consumer thread)
    int flip;
};

__device__ __forceinline__ const float* tile_acquire(const Ring& r, uint32_t t) {
    const uint32_t slot = t % (uint32_t)r.nslot;
    mbar_wait(&r.full[slot], (t / (uint32_t)r.nslot) & 1u);
    return r.slots + (size_t)slot * r.slot_floats;
}
__device__ __forceinline__ void tile_release(const Ring& r, uint32_t t, int lane) {
    __syncwarp();
    if (lane == 0) mbar_arrive(&r.empty[t % (uint32_t)r.nslot]);
}

// LayerNorm of the register-resident vector (thread owns x[4*tid .. 4*tid+3]) with weight/bias in smem
__device__ __forceinline__ void ln_regs(float (&x)[4], bool valid, int D, const float* w, const float* b, float* scratch,
                                        int& flip, int tid) {
    float s = valid ? (x[0] + x[1] + x[2] + x[3]) : 0.0f;
    const float mean = block_sum(s, scratch, flip, tid) / (float)D;
    float q = 0.0f;
    if (valid) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float d = x[i] - mean;
            q = fmaf(d, d, q);
        }
    }
    const float var = block_sum(q, scratch, flip, tid) / (float)D;
    const float rstd = 1.0f / sqrtf(var + 1e-5f);
    if (valid) {
        const float4 ww = *reinterpret_cast<const float4*>(w + 4 * tid);
        const float4 bb = *reinterpret_cast<const float4*>(b + 4 * tid);
        x[0] = (x[0] - mean) * rstd * ww.x + bb.x;
        x[1] = (x[1] - mean) * rstd * ww.y + bb.y;
        x[2] = (x[2] - mean) * rstd * ww.z + bb.z;
        x[3] = (x[3] - mean) * rstd * ww.w + bb.w;
    }
}

// GEMV over this CTA's column slice of one phase.  CT = columns per tile, NV = 128-bit chunks of x
// per thread.  After the call thread j (< ncols) holds y_j = bias_j + sum_k x_k W[k][col_j] in `y`.
template <int CT, int NV>
__device__ __forceinline__ float gemv_phase(const Ring& ring, ConsumerState& cs, int ncols, int K,
                                            const float (&xr)[NV * 4], float* red, int tid) {
    constexpr int MAXT = 8;  // <= 32 columns per CTA for CT = 4, <= 8 for CT = 1
    const int lane = tid & 31, warp = tid >> 5;
    float acc[MAXT * CT];
#pragma unroll
    for (int i = 0; i < MAXT * CT; ++i) acc[i] = 0.0f;
    const int ntiles = (ncols + CT - 1) / CT;
    const int cstride = K + 4;
#pragma unroll
    for (int ti = 0; ti < MAXT; ++ti) {
        if (ti < ntiles) {
            const float* w = tile_acquire(ring, cs.tile);
            const int nc = min(CT, ncols - ti * CT);
#pragma unroll
            for (int cc = 0; cc < CT; ++cc) {
                if (cc < nc) {
                    const float* col = w + cc * cstride;
                    float a = (tid == 0) ? col[K] : 0.0f;  // bias folded into the first partial
#pragma unroll
                    for (int v = 0; v < NV; ++v) {
                        const int k = (tid + MEGA_CONSUMERS * v) * 4;
                        if (k < K) {
                            const float4 wv = *reinterpret_cast<const float4*>(col + k);
                            a = fmaf(wv.x, xr[v * 4 + 0], a);
                            a = fmaf(wv.y, xr[v * 4 + 1], a);
                            a = fmaf(wv.z, xr[v * 4 + 2], a);
                            a = fmaf(wv.w, xr[v * 4 + 3], a);
                        }
                    }
                    acc[ti * CT + cc] = a;
                }
            }
            tile_release(ring, cs.tile, lane);
            cs.tile += 1;
        }
    }
    // cross-thread reduction: warp butterflies, then 8 warp partials per column through smem
#pragma unroll
    for (int j = 0; j < MAXT * CT; ++j) {
        if (j < ncols) {
            const float v = warp_sum(acc[j]);
            if (lane == 0) red[warp * 32 + j] = v;
        }
    }
    bar_sync(1, MEGA_CONSUMERS);
    float y = 0.0f;
    if (tid < ncols) {
#pragma unroll
        for (int w = 0; w < 8; ++w) y += red[w * 32 + tid];
    }
    return y;
}

// ---------------------------------------------------------------------------------------------
// producer: one thread walks the CTA's weight stream through the ring
// ---------------------------------------------------------------------------------------------
struct Producer {
    const Ring& ring;
    uint32_t t = 0;
    volatile int* stop;
    uint64_t policy;
    __device__ Producer(const Ring& r, volatile int* s) : ring(r), stop(s) { policy = l2_policy_evict_first(); }
    // returns false when the consumers asked to stop
    __device__ bool issue(const float* src, uint32_t floats, bool stream_once) {
        const uint32_t slot = t % (uint32_t)ring.nslot;
        const uint32_t par = ((t / (uint32_t)ring.nslot) & 1u) ^ 1u;
        uint32_t spins = 0;
        while (!mbar_try_wait(&ring.empty[slot], par)) {
            if (*stop) return false;
            if (++spins > (1u << 24)) __trap();
        }
        if (*stop) return false;
        mbar_arrive_expect_tx(&ring.full[slot], floats * 4u);
        float* dst = ring.slots + (size_t)slot * ring.slot_floats;
        if (stream_once) bulk_g2s_hint(dst, src, floats * 4u, &ring.full[slot], policy);
        else bulk_g2s(dst, src, floats * 4u, &ring.full[slot]);
        ++t;
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
    for (int c0 = 0; c0 < ncol[PH_HEAD]; c0 += 4) {
        const int nc = min(4, ncol[PH_HEAD] - c0);
        if (!pr.issue(hw, (uint32_t)(nc * (D + 4)), true)) return false;
        hw += (long long)nc * (D + 4);
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(MEGA_THREADS, 1) decode_mega_kernel(MegaParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid_all = threadIdx.x;
    const int cta = blockIdx.x;
    const int G = gridDim.x;
    const StreamDims sd{p.L, p.D, p.V, G};
    const int D = p.D;

    // ---- shared memory carve-up ----
    Ring ring;
    ring.nslot = p.nslot;
    ring.slot_floats = slot_floats(D);
    size_t off = 0;
    ring.slots = reinterpret_cast<float*>(smem_raw);
    off += (size_t)ring.nslot * ring.slot_floats * sizeof(float);
    off = (off + 127) & ~size_t(127);
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw + off);  // sampling sort keys;
    float* att_smem = reinterpret_cast<float*>(smem_raw + off);                          // aliased by attention merge
    off += GV_SORT_N * sizeof(unsigned long long);
    ring.full = reinterpret_cast<uint64_t*>(smem_raw + off);
    off += 16 * sizeof(uint64_t);
    ring.empty = reinterpret_cast<uint64_t*>(smem_raw + off);
    off += 16 * sizeof(uint64_t);
    float* red = reinterpret_cast<float*>(smem_raw + off);  // [8][32]
    off += 8 * 32 * sizeof(float);
    float* scratch = reinterpret_cast<float*>(smem_raw + off);  // [2][8] block_sum
    off += 16 * sizeof(float);
    float* fscr = reinterpret_cast<float*>(smem_raw + off);
    off += 16 * sizeof(float);
    int* iscr = reinterpret_cast<int*>(smem_raw + off);
    off += 16 * sizeof(int);
    volatile int* ctl = reinterpret_cast<volatile int*>(smem_raw + off);  // [0] stop flag, [1] tiles consumed
    off += 4 * sizeof(int);
    unsigned char* seen = smem_raw + off;  // [Vpad]

    if (tid_all == 0) {
        for (int i = 0; i < ring.nslot; ++i) {
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
        // ================= producer warp =================
        if (tid_all == MEGA_CONSUMERS) {
            Producer pr(ring, ctl);
            bool ok = true;
            for (int i = 0; i < p.n_steps && ok; ++i) {
                if (i == 0 && had_pending) continue;
                ok = produce_forward(pr, p, sd, cta);
            }
            // drain: every bulk copy issued must have landed before the CTA may exit.  Wait for the
            // consumers to finish (they may stop early on EOS with copies still in flight), then for
            // the full-barrier of every tile that was issued but never consumed.
            {
                uint32_t spins = 0;
                while (!ctl[0]) {
                    if (++spins > (1u << 30)) __trap();
                    __nanosleep(64);
                }
                const uint32_t consumed = (uint32_t)ctl[1];
                for (uint32_t t = consumed; t < pr.t; ++t)
                    mbar_wait(&ring.full[t % (uint32_t)ring.nslot], (t / (uint32_t)ring.nslot) & 1u);
            }
        }
    } else {
        // ================= consumer warps =================
        const int tid = tid_all;
        const int lane = tid & 31;
        ConsumerState cs{0u, 0};
        unsigned epoch = 0;
        const bool xvalid = 4 * tid < D;
        const int H = p.H;
        const float sqrt_hd = sqrtf((float)HD);
        int ncol[5];
        long long cbeg[5];
        for (int ph = 0; ph < 5; ++ph) {
            ncol[ph] = ph_cols(sd, ph, cta);
            cbeg[ph] = col_begin(ph_N(sd, ph), cta, G);
        }
        const SampleCfg scfg{p.V, p.top_k, p.top_p, p.top_p_threshold, p.temperature, p.rep_penalty};
        int n = n_start;                 // tokens emitted so far
        long long last_tok = st->last_tok[0];
        int finished = st->finished[0];
        int emitted = 0, done = 0;
        const float* mel_emb = p.blob + p.mel_emb_off;
        const float* mel_pos = p.blob + p.mel_pos_off;

        for (int i = 0; i < p.n_steps; ++i) {
            const bool tr = p.trace != nullptr && i == p.trace_step && tid == 0;
            unsigned long long* trow = p.trace + (size_t)cta * p.trace_slots;
            auto stamp = [&](int slot) {
                if (tr && slot < p.trace_slots) trow[slot] = globaltimer_ns();
            };
            stamp(p.L * 10 + 3);
            if (!(i == 0 && had_pending)) {
                // ------------- forward of token `last_tok` at mel position n, cache row P + n -------------
                const int pos = p.P + n;
                const int S = pos + 1;
                const float* e_tok = mel_emb + (size_t)last_tok * D;
                const float* e_pos = mel_pos + (size_t)n * D;
                for (int l = 0; l < p.L; ++l) {
                    float* kc = p.kv + ((size_t)l * 2 + 0) * p.kv_layer_stride;
                    float* vc = p.kv + ((size_t)l * 2 + 1) * p.kv_layer_stride;
                    // ---- QKV: LN1 -> [q|k|v] columns ----
                    {
                        float xr[4] = {0.f, 0.f, 0.f, 0.f};
                        if (xvalid) {
                            if (l == 0) {
                                const float4 a = *reinterpret_cast<const float4*>(e_tok + 4 * tid);
                                const float4 b = *reinterpret_cast<const float4*>(e_pos + 4 * tid);
                                xr[0] = a.x + b.x; xr[1] = a.y + b.y; xr[2] = a.z + b.z; xr[3] = a.w + b.w;
                            } else {
                                const float4 a = ldcg4(p.x + 4 * tid);
                                xr[0] = a.x; xr[1] = a.y; xr[2] = a.z; xr[3] = a.w;
                            }
                        }
                        const float* lnp = tile_acquire(ring, cs.tile);
                        ln_regs(xr, xvalid, D, lnp, lnp + D, scratch, cs.flip, tid);
                        tile_release(ring, cs.tile, lane);
                        cs.tile += 1;
                        const float y = gemv_phase<4, 1>(ring, cs, ncol[PH_QKV], D, xr, red, tid);
                        if (tid < ncol[PH_QKV]) {
                            const int ncolg = (int)cbeg[PH_QKV] + tid;
                            if (ncolg < D) {
                                p.qbuf[ncolg] = y;
                            } else {
                                const int c2 = (ncolg - D) % D;
                                float* dstc = (ncolg < 2 * D) ? kc : vc;
                                dstc[((size_t)(c2 / HD) * p.S_max + pos) * HD + (c2 % HD)] = y;
                            }
                        }
                    }
                    stamp(l * 10 + 0);
                    grid_barrier(p.barrier, epoch, G, tid);
                    stamp(l * 10 + 1);
                    // ---- ATT: (head, key-range) items ----
                    const int nsplit = min((S + 31) / 32, max(1, G / H));
                    const int chunk = (S + nsplit - 1) / nsplit;
                    for (int item = cta; item < H * nsplit; item += G) {
                        const int h = item / nsplit, sp = item % nsplit;
                        const int j0 = sp * chunk, j1 = min(S, j0 + chunk);
                        attn_decode_item<HD>(p.qbuf + h * HD, kc + (size_t)h * p.S_max * HD, vc + (size_t)h * p.S_max * HD,
                                             j0, j1, sqrt_hd, att_smem, tid, ConsumerSync(), p.att_o + (size_t)item * HD,
                                             p.att_ml + (size_t)item * 2);
                    }
                    stamp(l * 10 + 2);
                    grid_barrier(p.barrier, epoch, G, tid);
                    stamp(l * 10 + 3);
                    // ---- PROJ: merge attention partials -> o ; x += o . W_proj + b ----
                    {
                        float xr[4] = {0.f, 0.f, 0.f, 0.f};
                        if (xvalid) {
                            const int h = (4 * tid) / HD, d = (4 * tid) % HD;
                            float M = -INFINITY;
                            for (int s2 = 0; s2 < nsplit; ++s2) M = fmaxf(M, ldcg(p.att_ml + (size_t)(h * nsplit + s2) * 2));
                            float den = 0.0f;
                            for (int s2 = 0; s2 < nsplit; ++s2) {
                                const size_t it = (size_t)h * nsplit + s2;
                                const float2 ml = ldcg2(p.att_ml + it * 2);
                                const float wgt = (ml.x == -INFINITY) ? 0.0f : expf(ml.x - M);
                                const float4 ov = ldcg4(p.att_o + it * HD + d);
                                den = fmaf(ml.y, wgt, den);
                                xr[0] = fmaf(ov.x, wgt, xr[0]);
                                xr[1] = fmaf(ov.y, wgt, xr[1]);
                                xr[2] = fmaf(ov.z, wgt, xr[2]);
                                xr[3] = fmaf(ov.w, wgt, xr[3]);
                            }
#pragma unroll
                            for (int q = 0; q < 4; ++q) xr[q] = xr[q] / den;
                        }
                        const float y = gemv_phase<4, 1>(ring, cs, ncol[PH_PROJ], D, xr, red, tid);
                        if (tid < ncol[PH_PROJ]) {
                            const int ncolg = (int)cbeg[PH_PROJ] + tid;
                            const float xres = (l == 0) ? (e_tok[ncolg] + e_pos[ncolg]) : ldcg(p.x + ncolg);
                            p.x[ncolg] = xres + y;
                        }
                    }
                    stamp(l * 10 + 4);
                    grid_barrier(p.barrier, epoch, G, tid);
                    stamp(l * 10 + 5);
                    // ---- FC: LN2 -> u = gelu_new(. W_fc + b) ----
                    {
                        float xr[4] = {0.f, 0.f, 0.f, 0.f};
                        if (xvalid) {
                            const float4 a = ldcg4(p.x + 4 * tid);
                            xr[0] = a.x; xr[1] = a.y; xr[2] = a.z; xr[3] = a.w;
                        }
                        const float* lnp = tile_acquire(ring, cs.tile);
                        ln_regs(xr, xvalid, D, lnp, lnp + D, scratch, cs.flip, tid);
                        tile_release(ring, cs.tile, lane);
                        cs.tile += 1;
                        const float y = gemv_phase<4, 1>(ring, cs, ncol[PH_FC], D, xr, red, tid);
                        if (tid < ncol[PH_FC]) p.ubuf[cbeg[PH_FC] + tid] = gelu_new(y);
                    }
                    stamp(l * 10 + 6);
                    grid_barrier(p.barrier, epoch, G, tid);
                    stamp(l * 10 + 7);
                    // ---- PROJ2: x += u . W_proj2 + b ----
                    {
                        float ur[16];
#pragma unroll
                        for (int v = 0; v < 4; ++v) {
                            const int k = (tid + MEGA_CONSUMERS * v) * 4;
                            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (k < 4 * D) a = ldcg4(p.ubuf + k);
                            ur[v * 4 + 0] = a.x; ur[v * 4 + 1] = a.y; ur[v * 4 + 2] = a.z; ur[v * 4 + 3] = a.w;
                        }
                        const float y = gemv_phase<1, 4>(ring, cs, ncol[PH_PROJ2], 4 * D, ur, red, tid);
                        if (tid < ncol[PH_PROJ2]) {
                            const int ncolg = (int)cbeg[PH_PROJ2] + tid;
                            p.x[ncolg] = ldcg(p.x + ncolg) + y;
                        }
                    }
                    stamp(l * 10 + 8);
                    grid_barrier(p.barrier, epoch, G, tid);
                    stamp(l * 10 + 9);
                }
                // ---- HEAD: ln_f -> final_norm -> latent z ; logits = z . mel_head^T + b ----
                {
                    float xr[4] = {0.f, 0.f, 0.f, 0.f};
                    if (xvalid) {
                        const float4 a = ldcg4(p.x + 4 * tid);
                        xr[0] = a.x; xr[1] = a.y; xr[2] = a.z; xr[3] = a.w;
                    }
                    const float* lnp = tile_acquire(ring, cs.tile);
                    ln_regs(xr, xvalid, D, lnp, lnp + D, scratch, cs.flip, tid);
                    ln_regs(xr, xvalid, D, lnp + 2 * D, lnp + 3 * D, scratch, cs.flip, tid);
                    tile_release(ring, cs.tile, lane);
                    cs.tile += 1;
                    if (cta == 0 && xvalid)
                        *reinterpret_cast<float4*>(p.pend_latent + 4 * tid) = make_float4(xr[0], xr[1], xr[2], xr[3]);
                    const float y = gemv_phase<4, 1>(ring, cs, ncol[PH_HEAD], D, xr, red, tid);
                    if (tid < ncol[PH_HEAD]) p.pend_logits[cbeg[PH_HEAD] + tid] = y;
                }
                stamp(p.L * 10 + 0);
                grid_barrier(p.barrier, epoch, G, tid);
                stamp(p.L * 10 + 1);
            }
            // ------------- sample + emit (every CTA computes the same token) -------------
            int tok = sample_token(p.pend_logits, seen, scfg, p.noise ? p.noise + (size_t)i * p.V : nullptr, p.seed,
                                   (uint32_t)n, 0u, keys, fscr, iscr, tid, ConsumerSync());
            stamp(p.L * 10 + 2);
            if (p.forced) tok = (int)p.forced[i];
            if (!p.ignore_eos && finished) tok = p.stop_token;
            if (cta == 0) {
                if (tid == 0) p.ids_out[i] = tok;
                for (int q = tid; q < D; q += MEGA_CONSUMERS) p.latents_out[(size_t)i * D + q] = ldcg(p.pend_latent + q);
                if (p.logits_out)
                    for (int q = tid; q < p.V; q += MEGA_CONSUMERS) p.logits_out[(size_t)i * p.V + q] = ldcg(p.pend_logits + q);
            }
            if (tid == 0) seen[tok] = 1;
            last_tok = tok;
            if (!p.ignore_eos && tok == p.stop_token) finished = 1;
            n += 1;
            emitted += 1;
            bar_sync(1, MEGA_CONSUMERS);  // seen[] update visible to the next step's sampler
            if (finished || n >= p.max_total) {
                done = 1;
                break;
            }
        }
        // tell the producer to stop (it may be blocked on a full ring or still have copies in flight)
        if (tid == 0) {
            ctl[1] = (int)cs.tile;
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
}

// ---------------------------------------------------------------------------------------------
// host launcher
// ---------------------------------------------------------------------------------------------
size_t mega_smem_bytes(int D, int nslot, int Vpad) {
    size_t off = (size_t)nslot * slot_floats(D) * sizeof(float);
    off = (off + 127) & ~size_t(127);
    off += GV_SORT_N * sizeof(unsigned long long);
    off += 32 * sizeof(uint64_t);
    off += 8 * 32 * sizeof(float) + 16 * sizeof(float) + 16 * sizeof(float) + 16 * sizeof(int) + 4 * sizeof(int);
    off += Vpad;
    return (off + 15) & ~size_t(15);
}

template <int HD>
static cudaError_t launch_hd(const MegaParams& p, int grid, size_t smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(decode_mega_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    MegaParams pp = p;
    void* args[] = {&pp};
    return cudaLaunchCooperativeKernel((void*)decode_mega_kernel<HD>, dim3(grid), dim3(MEGA_THREADS), args, smem, st);
}

cudaError_t launch_decode_mega(const MegaParams& p, int grid, cudaStream_t st) {
    const size_t smem = mega_smem_bytes(p.D, p.nslot, p.Vpad);
    cudaError_t e = cudaMemsetAsync(p.barrier, 0, sizeof(unsigned), st);
    if (e != cudaSuccess) return e;
    switch (p.D / p.H) {
        case 32: return launch_hd<32>(p, grid, smem, st);
        case 64: return launch_hd<64>(p, grid, smem, st);
        case 128: return launch_hd<128>(p, grid, smem, st);
        case 256: return launch_hd<256>(p, grid, smem, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace gv
