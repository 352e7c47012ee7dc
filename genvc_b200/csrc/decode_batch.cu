// Batched fused persistent decode kernel: up to GV_BATCH_ROWS (8) equal-length rows share ONE pass of the weight
// stream per generated token.  Same loop as decode_mega.cu (layers/stream_generator.py:809-881 around the cached
// forward of layers/gpt_inference.py:92-112), with the reference's batch semantics: every row uses the same
// position index (gpt_inference.py:92-96), rows finish independently and finished rows emit the stop token
// until the last one ends (stream_generator.py:860-881).
//
// What changes against the single-row kernel:
//   * GEMV -> skinny GEMM out of the same ring, still fp32 FFMA: a warp takes a whole tile (4 units), holds its weights in
//     registers (128 per lane at D = 1024) and streams the rows of the activation buffer against them
//     (8 LDS.128 + 128 FFMA per row); one transposing shuffle tree per row, no cross-warp reduction.
//   * mlp.c_proj stays split along K (the CTA that computed u[:, k] owns row k of W_proj2): thread j accumulates outputs
//     4j .. 4j+3 of every row in registers; the per-CTA partials [rows][D] go to the reducer CTAs.
//   * one activation buffer xs[8][D + 4] in shared memory is the B operand of every phase (x -> attention output
//     -> x1 -> latent); LayerNorm statistics per row by one warp each.
//   * attention items = (row, head, key range), H * B * nsplit <= grid; the item code is the single-row one.
//   * row r is sampled by CTA r (its own repetition-penalty set in shared memory) and the token is published
//     through a tagged word; every CTA reads the B tokens for the next step's embeddings.
#define GV_RING_NSLOT GV_BATCH_NSLOT
#define GV_MEGA_NS megab
#define GV_UNIFORM_TILE_WAIT 1  // consumer warps never split on a wait (mega_dev.cuh: tile_ready_wait_u, ld_tagged_vec_u)
#include "mega_dev.cuh"

namespace gv {
using namespace megab;

#define NBR GV_BATCH_ROWS
#define B_SCR_BYTES 9728  // attention scratch (as the single-row kernel) | reducer partials

// ---------------------------------------------------------------------------------------------
// dot phase (QKV, attn c_proj, FC, logits head): for every unit u < nunits of this CTA and every row r < nb,
//   dot[u][r] = sum_k W'[k][u] xs[r][k];   epi(u, r, dot, c2, c1)  is called once per (u, r).
// A warp takes whole ring tiles (4 units): the tile's weights go to registers once (4 units x D / 32 values per lane),
// the tile is released, then every row of xs is read from shared memory against them: 8 LDS.128 + 128 FFMA per row and
// one transposing shuffle tree for the four units -- no cross-warp reduction, fp32 FMA like the single-row kernel.
// (A first version ran these phases on the warp-level tensor path -- mma.sync m16n8k8, 3xTF32 -- measured on B200 at
// 2 cycles per mma per SM = 512 MAC/clk/SM; with three products per fp32 MAC and the hi/lo splits that is no faster than
// the 128 FFMA/clk/SM of the CUDA cores, and it was 4-5x slower in practice: tools/mma_probe.cu)
// Only the owning warp reads a tile: it arrives on the tile's empty barrier with the full count.
// ---------------------------------------------------------------------------------------------
template <int NXV, class Epi>
__device__ __forceinline__ void dot_phase(const Ring& ring, const Cons& cs, int nunits, const float* xs, int nb, int tid, Epi epi) {
    constexpr int D = NXV * 128, UF = D + 4;
    const int warp = tid >> 5, lane = tid & 31;
    const int ntiles = (nunits + UPT - 1) / UPT;
    for (int t = warp; t < ntiles; t += MEGA_WARPS) {
        long long tw0 = 0;
        if (cs.wacc != nullptr && tid == 0) tw0 = clock64();  // debug profile: time spent waiting for weight tiles
        const float* base = tile_wait(ring, Cons{cs.gt, nullptr}, cs.gt + (uint32_t)t, lane);
        if (cs.wacc != nullptr && tid == 0) *cs.wacc += (unsigned long long)(clock64() - tw0);
        const int nu = min(UPT, nunits - t * UPT);
        float4 w[UPT][NXV];
        float2 cc[UPT];
#pragma unroll
        for (int u = 0; u < UPT; ++u) {
            const float* col = base + (u < nu ? u : 0) * UF;  // units past the end of the phase: unit 0 again (results dropped)
            cc[u] = *reinterpret_cast<const float2*>(col + D);
#pragma unroll
            for (int i = 0; i < NXV; ++i) w[u][i] = *reinterpret_cast<const float4*>(col + (i * 32 + lane) * 4);
        }
        tile_release(ring, cs.gt + (uint32_t)t, lane, (uint32_t)MEGA_WARPS);
        // two rows per iteration (independent accumulator sets, eight LDS.128 in flight): the tile is handled by ONE warp, so
        // the instruction-level parallelism has to come from inside the warp
        const bool up16 = (lane & 16) != 0, up8 = (lane & 8) != 0;
        const int q = lane >> 3;
        const float2 c = up16 ? (up8 ? cc[3] : cc[2]) : (up8 ? cc[1] : cc[0]);
        for (int r0 = 0; r0 < nb; r0 += 2) {
            const float* xa = xs + r0 * UF + lane * 4;
            const float* xb = xa + UF;  // row r0 + 1 (always inside xs; dropped when r0 + 1 == nb)
            float a0[UPT], a1[UPT], b0[UPT], b1[UPT];
#pragma unroll
            for (int u = 0; u < UPT; ++u) a0[u] = a1[u] = b0[u] = b1[u] = 0.0f;
            constexpr int HB = NXV >= 2 ? NXV / 2 : 1;
#pragma unroll
            for (int h = 0; h < NXV / HB; ++h) {
                float4 va[HB], vb[HB];
#pragma unroll
                for (int i = 0; i < HB; ++i) {
                    va[i] = *reinterpret_cast<const float4*>(xa + (h * HB + i) * 128);
                    vb[i] = *reinterpret_cast<const float4*>(xb + (h * HB + i) * 128);
                }
#pragma unroll
                for (int i = 0; i < HB; ++i) {
#pragma unroll
                    for (int u = 0; u < UPT; ++u) {
                        const float4 wv = w[u][h * HB + i];
                        a0[u] = fmaf(wv.x, va[i].x, a0[u]);
                        a1[u] = fmaf(wv.y, va[i].y, a1[u]);
                        b0[u] = fmaf(wv.x, vb[i].x, b0[u]);
                        b1[u] = fmaf(wv.y, vb[i].y, b1[u]);
                        a0[u] = fmaf(wv.z, va[i].z, a0[u]);
                        a1[u] = fmaf(wv.w, va[i].w, a1[u]);
                        b0[u] = fmaf(wv.z, vb[i].z, b0[u]);
                        b1[u] = fmaf(wv.w, vb[i].w, b1[u]);
                    }
                }
            }
            // transposing butterflies (both rows interleaved): lanes [8q, 8q+8) end up reducing unit q
            float ta[UPT], tb[UPT];
#pragma unroll
            for (int u = 0; u < UPT; ++u) {
                ta[u] = a0[u] + a1[u];
                tb[u] = b0[u] + b1[u];
            }
            const float ka0 = up16 ? ta[2] : ta[0], sa0 = up16 ? ta[0] : ta[2];
            const float ka1 = up16 ? ta[3] : ta[1], sa1 = up16 ? ta[1] : ta[3];
            const float kb0 = up16 ? tb[2] : tb[0], sb0 = up16 ? tb[0] : tb[2];
            const float kb1 = up16 ? tb[3] : tb[1], sb1 = up16 ? tb[1] : tb[3];
            const float ha0 = ka0 + __shfl_xor_sync(0xffffffffu, sa0, 16);
            const float ha1 = ka1 + __shfl_xor_sync(0xffffffffu, sa1, 16);
            const float hb0 = kb0 + __shfl_xor_sync(0xffffffffu, sb0, 16);
            const float hb1 = kb1 + __shfl_xor_sync(0xffffffffu, sb1, 16);
            float va_ = (up8 ? ha1 : ha0) + __shfl_xor_sync(0xffffffffu, up8 ? ha0 : ha1, 8);
            float vb_ = (up8 ? hb1 : hb0) + __shfl_xor_sync(0xffffffffu, up8 ? hb0 : hb1, 8);
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
                va_ += __shfl_xor_sync(0xffffffffu, va_, o);
                vb_ += __shfl_xor_sync(0xffffffffu, vb_, o);
            }
            if ((lane & 7) == 0 && q < nu) {
                epi(t * UPT + q, r0, va_, c.x, c.y);
                if (r0 + 1 < nb) epi(t * UPT + q, r0 + 1, vb_, c.x, c.y);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// mlp.c_proj split along K: unit k = row k of W_proj2 (owned by the CTA that computed u[:, k]).
//   part[r][n] = sum_k usT[k][r] W[k][n]    for all D outputs n, rows r < nb
// Thread j owns outputs 4j .. 4j+3 for every row (accumulators in registers); per unit one LDS.128 of the row of
// W_proj2 and two broadcast LDS.128 of the unit's 8 row values.  Every warp reads every tile (one arrival each).
// Results go straight to the exchange buffer pp[cta][n / 8][row][n % 8] (tagged).
// ---------------------------------------------------------------------------------------------
template <int NXV>
__device__ __forceinline__ void outer_phase(const Ring& ring, const Cons& cs, int nunits, const float* usT, int nb, int tid,
                                            float* pp_cta, uint32_t tag) {
    constexpr int D = NXV * 128, UF = D + 4;
    const int lane = tid & 31;
    const bool valid = 4 * tid < D;
    const int ntiles = (nunits + UPT - 1) / UPT;
    float4 acc[NBR];
#pragma unroll
    for (int r = 0; r < NBR; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = 0; t < ntiles; ++t) {
        long long tw0 = 0;
        if (cs.wacc != nullptr && tid == 0) tw0 = clock64();
        const float* base = tile_wait(ring, Cons{cs.gt, nullptr}, cs.gt + (uint32_t)t, lane);
        if (cs.wacc != nullptr && tid == 0) *cs.wacc += (unsigned long long)(clock64() - tw0);
        if (valid) {
            float4 wv[UPT], ua[UPT], ub[UPT];
#pragma unroll
            for (int u = 0; u < UPT; ++u) {  // the tile's four weight rows and their row values: twelve LDS.128 in flight
                const int k = min(t * UPT + u, nunits - 1);  // (units past the end: values unused)
                wv[u] = *reinterpret_cast<const float4*>(base + (t * UPT + u < nunits ? u : 0) * UF + 4 * tid);
                ua[u] = *reinterpret_cast<const float4*>(usT + k * NBR);
                ub[u] = *reinterpret_cast<const float4*>(usT + k * NBR + 4);
            }
#pragma unroll
            for (int u = 0; u < UPT; ++u) {
                if (t * UPT + u < nunits) {
                    const float ur[NBR] = {ua[u].x, ua[u].y, ua[u].z, ua[u].w, ub[u].x, ub[u].y, ub[u].z, ub[u].w};
#pragma unroll
                    for (int r = 0; r < NBR; ++r) {
                        if (r < nb) {
                            acc[r].x = fmaf(ur[r], wv[u].x, acc[r].x);
                            acc[r].y = fmaf(ur[r], wv[u].y, acc[r].y);
                            acc[r].z = fmaf(ur[r], wv[u].z, acc[r].z);
                            acc[r].w = fmaf(ur[r], wv[u].w, acc[r].w);
                        }
                    }
                }
            }
        }
        tile_release(ring, cs.gt + (uint32_t)t, lane);
    }
    if (cs.wacc != nullptr && tid == 0) cs.wacc[1] -= (unsigned long long)clock64();  // debug profile [14]: stores
    if (valid) {
        const int n = 4 * tid;
#pragma unroll
        for (int r = 0; r < NBR; ++r) {
            if (r < nb) {
                const int idx = ((n >> 3) * NBR + r) * 8 + (n & 7);
                st_tagged2(pp_cta, idx, acc[r].x, acc[r].y, tag);
                st_tagged2(pp_cta, idx + 2, acc[r].z, acc[r].w, tag);
            }
        }
    }
    if (cs.wacc != nullptr && tid == 0) cs.wacc[1] += (unsigned long long)clock64();
}

// activation rows [8][D + 4]; the sampling sort keys (16 KB) alias them (and the scratch region that follows, dead then)
__host__ __device__ inline size_t batch_xs_bytes(int D) {
    const size_t rows = (size_t)NBR * (D + 4) * sizeof(float);
    const size_t keys = (size_t)GV_SORT_N * 8 > B_SCR_BYTES ? (size_t)GV_SORT_N * 8 - B_SCR_BYTES : 0;
    return rows > keys ? rows : (keys + 15) / 16 * 16;
}

// rows 0 .. nb-1 of a tagged [rows][D] exchange buffer -> xs (thread t owns elements 4t .. 4t+3 of every row); all loads
// are issued before the first tag is tested
template <int D>
__device__ __forceinline__ void load_rows(const float* buf, uint32_t tag, float* xs, int nb, int tid) {
    constexpr int UF = D + 4;
    if (4 * tid >= D) return;
#pragma unroll 1
    for (int r0 = 0; r0 < nb; r0 += 4) {  // four rows (eight 16-byte loads) in flight per thread
        uint4 a[4], b[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (r0 + q < nb) {
                const float* p = buf + 2 * ((size_t)(r0 + q) * D + 4 * tid);
                a[q] = ld_x16(p);
                b[q] = ld_x16(p + 4);
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (r0 + q < nb) {
                const float* p = buf + 2 * ((size_t)(r0 + q) * D + 4 * tid);
                uint32_t spins = 0;
                // the warp leaves the poll as one (4 * tid < D is warp-uniform: D is a multiple of 128)
                while (!__all_sync(0xffffffffu, tags_ok(a[q], tag, 0xffffffffu) && tags_ok(b[q], tag, 0xffffffffu))) {
                    GV_SPIN(spins, __LINE__, 0u, 0u);
                    a[q] = ld_x16(p);
                    b[q] = ld_x16(p + 4);
                }
                *reinterpret_cast<float4*>(xs + (r0 + q) * UF + 4 * tid) =
                    make_float4(__uint_as_float(a[q].x), __uint_as_float(a[q].z), __uint_as_float(b[q].x), __uint_as_float(b[q].z));
            }
        }
    }
}

// LayerNorm statistics of row `warp` of xs (one warp per row, two passes over registers): stats[2 warp] = mean, rstd
template <int NXV>
__device__ __forceinline__ void row_stats(const float* xs, float* stats, int nb, int warp, int lane) {
    constexpr int D = NXV * 128, UF = D + 4;
    if (warp >= nb) return;
    float4 v[NXV];
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < NXV; ++i) {
        v[i] = *reinterpret_cast<const float4*>(xs + warp * UF + (i * 32 + lane) * 4);
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    s = warp_sum(s);
    const float mean = s / (float)D;
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < NXV; ++i) {
        const float d0 = v[i].x - mean, d1 = v[i].y - mean, d2 = v[i].z - mean, d3 = v[i].w - mean;
        q += fmaf(d0, d0, d1 * d1) + fmaf(d2, d2, d3 * d3);
    }
    q = warp_sum(q);
    if (lane == 0) {
        stats[2 * warp] = mean;
        stats[2 * warp + 1] = 1.0f / sqrtf(q / (float)D + 1e-5f);
    }
}

// explicit LayerNorm of row `warp` of xs in place (logits head: ln_f, then final_norm)
template <int NXV>
__device__ __forceinline__ void row_layernorm2(float* xs, const float* lnp, int nb, int warp, int lane) {
    constexpr int D = NXV * 128, UF = D + 4;
    if (warp >= nb) return;
    float4 v[NXV];
#pragma unroll
    for (int i = 0; i < NXV; ++i) v[i] = *reinterpret_cast<const float4*>(xs + warp * UF + (i * 32 + lane) * 4);
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        const float* w = lnp + 2 * pass * D;
        const float* b = w + D;
        float s = 0.0f;
#pragma unroll
        for (int i = 0; i < NXV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        s = warp_sum(s);
        const float mean = s / (float)D;
        float q = 0.0f;
#pragma unroll
        for (int i = 0; i < NXV; ++i) {
            v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
            q += fmaf(v[i].x, v[i].x, v[i].y * v[i].y) + fmaf(v[i].z, v[i].z, v[i].w * v[i].w);
        }
        q = warp_sum(q);
        const float rstd = 1.0f / sqrtf(q / (float)D + 1e-5f);
#pragma unroll
        for (int i = 0; i < NXV; ++i) {
            const float4 ww = *reinterpret_cast<const float4*>(w + (i * 32 + lane) * 4);
            const float4 bb = *reinterpret_cast<const float4*>(b + (i * 32 + lane) * 4);
            v[i].x = v[i].x * rstd * ww.x + bb.x;
            v[i].y = v[i].y * rstd * ww.y + bb.y;
            v[i].z = v[i].z * rstd * ww.z + bb.z;
            v[i].w = v[i].w * rstd * ww.w + bb.w;
        }
    }
#pragma unroll
    for (int i = 0; i < NXV; ++i) *reinterpret_cast<float4*>(xs + warp * UF + (i * 32 + lane) * 4) = v[i];
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <int NXV>
__global__ void __launch_bounds__(MEGA_THREADS, 1) decode_batch_kernel(MegaParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int D = NXV * 128, UF = D + 4;
    const int tid_all = threadIdx.x;
    const int cta = blockIdx.x;
    const int G = gridDim.x;
    const int NB = p.B;
    const StreamDims sd{p.L, D, p.V, G};

    // ---- shared memory carve-up (mirrored by batch_smem_bytes) ----
    Ring ring;
    ring.slot_floats = slot_floats(D);
    size_t off = 0;
    ring.slots = reinterpret_cast<float*>(smem_raw);
    off += (size_t)NSLOT * ring.slot_floats * sizeof(float);
    off = (off + 127) & ~size_t(127);
    float* xs = reinterpret_cast<float*>(smem_raw + off);  // [8][D + 4] activation rows: the B operand of every phase
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw + off);  // sampling sort keys (xs is dead then)
    off += batch_xs_bytes(D);
    float* att_sc = reinterpret_cast<float*>(smem_raw + off);         // attention: per-warp (max, sum) + q staging
    float* att_op = reinterpret_cast<float*>(smem_raw + off + 1280);  // attention: [8][hd] per-warp PV partials
    float* red4 = reinterpret_cast<float*>(smem_raw + off);           // reducer CTAs: [8 warps][64] partial sums
    off += B_SCR_BYTES;
    float* us = reinterpret_cast<float*>(smem_raw + off);  // [32 units][8 rows] gelu(fc) of this CTA's units (zero past nb)
    off += (size_t)32 * NBR * sizeof(float);
    float* stats = reinterpret_cast<float*>(smem_raw + off);  // [8][2] mean, rstd of the LayerNorm folded into the phase
    off += 2 * NBR * sizeof(float);
    float* resid = reinterpret_cast<float*>(smem_raw + off);  // [8 rows][8] block input at this CTA's attn c_proj columns
    off += 8 * NBR * sizeof(float);
    float* slog = reinterpret_cast<float*>(smem_raw + off);  // [Vpad] logits of the row this CTA samples
    off += (size_t)p.Vpad * sizeof(float);
    ring.full = reinterpret_cast<uint64_t*>(smem_raw + off);
    off += 16 * sizeof(uint64_t);
    ring.empty = reinterpret_cast<uint64_t*>(smem_raw + off);
    off += 16 * sizeof(uint64_t);
    float* fscr = reinterpret_cast<float*>(smem_raw + off);
    off += 16 * sizeof(float);
    int* iscr = reinterpret_cast<int*>(smem_raw + off);
    off += 16 * sizeof(int);
    volatile int* ctl = reinterpret_cast<volatile int*>(smem_raw + off);  // [0] stop flag, [1] tiles consumed
    ring.landed = reinterpret_cast<uint32_t*>(smem_raw + off) + 3;        // [3] tiles the producer has seen landed
    off += 8 * sizeof(int);
    int* ltok = reinterpret_cast<int*>(smem_raw + off);  // [8] last token of every row
    off += NBR * sizeof(int);
    int* fin = reinterpret_cast<int*>(smem_raw + off);  // [8] row finished
    off += NBR * sizeof(int);
    unsigned long long* prof = reinterpret_cast<unsigned long long*>(smem_raw + off);  // [16] debug: cycles per phase (thread 0)
    off += 16 * sizeof(unsigned long long);
    unsigned char* seen = smem_raw + off;  // [Vpad] ids present in the row this CTA samples (repetition penalty)

    if (tid_all == 0) {
        for (int i = 0; i < NSLOT; ++i) {
            mbar_init(&ring.full[i], 1);
            mbar_init(&ring.empty[i], MEGA_WARPS);  // every consumer warp reads every tile
        }
        ctl[0] = 0;
        ctl[1] = 0;
        ctl[2] = 0;
        ctl[3] = 0;
        ctl[4] = 0;
        mbar_fence_init();
    }
    const GenState* st = p.st;
    if (cta < NB)
        for (int i = tid_all; i < p.Vpad; i += MEGA_THREADS) seen[i] = p.seen[(size_t)cta * p.Vpad + i];
    for (int i = tid_all; i < NBR * UF; i += MEGA_THREADS) xs[i] = 0.0f;
    for (int i = tid_all; i < 32 * NBR; i += MEGA_THREADS) us[i] = 0.0f;
    if (tid_all < NBR) {
        ltok[tid_all] = tid_all < NB ? (int)st->last_tok[tid_all] : 0;
        fin[tid_all] = tid_all < NB ? st->finished[tid_all] : 1;
    }
    __syncthreads();

    const int had_pending = st->has_pending;
    const int n_start = st->n_emitted;
    if (st->done) {  // uniform: nothing to do; the state moves to the output copy unchanged
        if (cta < NB)
            for (int q = tid_all; q < p.Vpad; q += MEGA_THREADS) p.seen_out[(size_t)cta * p.Vpad + q] = seen[q];
        if (cta == 0 && tid_all == 0) {
            GenState* so = p.st_out;
            so->n_emitted = n_start;
            so->done = 1;
            so->has_pending = had_pending;
            so->P = st->P;
            so->B = st->B;
            for (int r = 0; r < NB; ++r) {
                so->finished[r] = st->finished[r];
                so->last_tok[r] = st->last_tok[r];
            }
            p.status[1] = 1;
        }
        return;
    }

    if (tid_all >= MEGA_CONSUMERS) {
        // ================= producer warpgroup: hands its registers to the consumers =================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (tid_all == MEGA_CONSUMERS) {
            Producer pr(ring, ctl, (uint32_t)max(1, min(p.window, NSLOT)));
            pr.region = p.stream + cta_base(sd, cta);
            pr.region_floats = cta_base(sd, cta + 1) - cta_base(sd, cta);
            bool ok = true;
            for (int i = 0; i < p.n_steps && ok; ++i) {
                if (i == 0 && had_pending) continue;
                ok = produce_forward(pr, p.stream, p.blob, p.lnf_off, p.L, D, sd, cta);
            }
            uint32_t spins = 0;
            while (!ctl[0]) {  // drain: every bulk copy issued must have landed before the CTA exits
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
    const uint32_t tmask = 0xffffffffu;
    const bool xvalid = 4 * tid < D;
    const int H = p.H, HD = D / H;
    int nun[5], ubeg[5], ntl[5];
    for (int ph = 0; ph < 5; ++ph) {
        nun[ph] = ph_units(sd, ph, cta);
        ubeg[ph] = (int)col_begin(ph_N(sd, ph), cta, G);
        ntl[ph] = (nun[ph] + UPT - 1) / UPT;
    }
    const int n_red = D / 8;  // reducer CTAs of the mlp.c_proj partial sums (8 outputs x 8 rows each)
    const SampleCfg scfg{p.V, p.top_k, p.top_p, p.top_p_threshold, p.temperature, p.rep_penalty};
    int n = n_start;
    int emitted = 0, done = 0;
    const uint32_t tags_per_step = (uint32_t)GV_TAGS_PER_LAYER * (uint32_t)p.L + 1u + (uint32_t)GV_BATCH_TAGS_EXTRA;
    const float* mel_emb = p.blob + p.mel_emb_off;
    const float* mel_pos = p.blob + p.mel_pos_off;
    unsigned* const hc = p.hops;
    const unsigned near = p.hop_near >= 0 ? (unsigned)p.hop_near : (unsigned)max(G / 37, 1);
    const unsigned settle = (unsigned)p.hop_settle_ns;
    unsigned lc = 0, t_ao = 0, t_lg = 0, t_cnt = 0;
    // debug phase profile (genvc_debug_trace): thread 0 accumulates the cycles between consecutive marks
    const bool profiling = p.trace != nullptr && tid == 0;
    long long pclk = 0;
    if (profiling) {
        for (int k = 0; k < 16; ++k) prof[k] = 0ull;
        pclk = clock64();
        cs.wacc = prof + 12;  // [12] tile waits of the dot phases, [13] of mlp.c_proj
    }
    auto mark = [&](int k) {
        if (profiling) {
            const long long c = clock64();
            prof[k] += (unsigned long long)(c - pclk);
            pclk = c;
        }
    };
    const size_t xq_row = 2 * (size_t)3 * D, x_row = 2 * (size_t)D, lg_row = 2 * (size_t)p.Vpad;

    for (int i = 0; i < p.n_steps; ++i) {
        const uint32_t tbase = p.tag0 + (uint32_t)i * tags_per_step;
        const uint32_t tg_tok = tbase + (uint32_t)GV_TAGS_PER_LAYER * (uint32_t)p.L + 1u;
        if (!(i == 0 && had_pending)) {
            // ------------- forward of the rows' last tokens at mel position n, cache row P + n -------------
            const int pos = p.P + n;
            const int S = pos + 1;
            const int nsplit0 = att_nsplit_b(S, NB * H, G);
            const int chunk = att_chunk_b(S, nsplit0);
            const int nsplit = (S + chunk - 1) / chunk;
            const int n_items = NB * H * nsplit;
            const unsigned near_ao = p.hop_near_ao >= 0 ? (unsigned)p.hop_near_ao : (unsigned)min(max(NB * H / 3, 1), 4);
            for (int l = 0; l < p.L; ++l) {
                float* kc = p.kv + ((size_t)l * 2 + 0) * p.kv_layer_stride;
                float* vc = p.kv + ((size_t)l * 2 + 1) * p.kv_layer_stride;
                const uint32_t tg = tbase + (uint32_t)GV_TAGS_PER_LAYER * (uint32_t)l;
                // ---- block input rows -> xs ----
                if (l == 0) {
                    if (xvalid) {
                        const float4 b = *reinterpret_cast<const float4*>(mel_pos + (size_t)n * D + 4 * tid);
                        for (int r = 0; r < NB; ++r) {
                            const float4 a = *reinterpret_cast<const float4*>(mel_emb + (size_t)ltok[r] * D + 4 * tid);
                            *reinterpret_cast<float4*>(xs + r * UF + 4 * tid) = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
                        }
                    }
                } else {
                    hop_wait(hc + HC_X2 * GV_HOP_STRIDE, lc * (unsigned)n_red, tid, tmask, settle, nullptr, near);
                    load_rows<D>(p.x2, tg - (uint32_t)GV_TAGS_PER_LAYER + TG_X2, xs, NB, tid);
                }
                bar_sync(1, MEGA_CONSUMERS);
                mark(0);  // x2 hop + load
                // ---- QKV: [q|k|v] = LN1(x) . W_attn + b  (LN folded into the packed weights: statistics enter in the epilogue) ----
                row_stats<NXV>(xs, stats, NB, warp, lane);
                if (tid < NBR * 8) {  // block input at this CTA's attn c_proj columns (residual of the PROJ epilogue)
                    const int r = tid >> 3, u = tid & 7;
                    if (r < NB && u < nun[PH_PROJ]) resid[tid] = xs[r * UF + ubeg[PH_PROJ] + u];
                }
                bar_sync(1, MEGA_CONSUMERS);  // statistics visible to the epilogues
                dot_phase<NXV>(ring, cs, nun[PH_QKV], xs, NB, tid, [&](int u, int r, float dot, float c2, float c1) {
                    const float mean = stats[2 * r], rstd = stats[2 * r + 1];
                    st_tagged(p.xq + r * xq_row, ubeg[PH_QKV] + u, fmaf(rstd, fmaf(-mean, c1, dot), c2), tg + TG_XQ);
                });
                cs.gt += (uint32_t)ntl[PH_QKV];
                mark(1);  // stats + QKV
                // ---- ATT: (row, head, key-range) items on the first n_items CTAs.  The CTA of range 0 merges the nsplit
                // partials of its (row, head) in split order (the other ranges bump a per-(row, head) counter, one thread
                // of the merger polls it, the tags validate the data) and publishes the normalised output segment: every
                // CTA then loads B x D finished values instead of B x nsplit x D partials (measured: the per-CTA merge
                // of all rows was 10 us per layer at B = 8) ----
                if (cta < n_items) {
                    const int rh = cta / nsplit, sp = cta % nsplit;
                    const int r = rh / H, h = rh % H;
                    const int j0 = sp * chunk, j1 = min(S, j0 + chunk);
                    float* kh = kc + ((size_t)r * H + h) * p.S_max * HD;
                    float* vh = vc + ((size_t)r * H + h) * p.S_max * HD;
                    if (nsplit == 1) {
#define GV_ATT_CASE(hd)                                                                                                    \
    case hd:                                                                                                               \
        att_item<hd, false, true>(kh, vh, p.xq + r * xq_row, D, h, j0, j1, S, tg + TG_XQ, att_sc, att_op, tid, p.ao + r * x_row, \
                                  nullptr, h, tg + TG_AO, tmask, nullptr);                                                 \
        break;
                        switch (HD) {
                            GV_ATT_CASE(32) GV_ATT_CASE(64) GV_ATT_CASE(128) GV_ATT_CASE(256)
                            default: break;
                        }
#undef GV_ATT_CASE
                        hop_arrive(hc + HC_AO * GV_HOP_STRIDE, tid);
                    } else {
#define GV_ATT_CASE(hd)                                                                                                  \
    case hd:                                                                                                             \
        att_item<hd, false>(kh, vh, p.xq + r * xq_row, D, h, j0, j1, S, tg + TG_XQ, att_sc, att_op, tid, p.att_o, p.att_ml, \
                            cta, tg + TG_AO, tmask, nullptr);                                                            \
        break;
                        switch (HD) {
                            GV_ATT_CASE(32) GV_ATT_CASE(64) GV_ATT_CASE(128) GV_ATT_CASE(256)
                            default: break;
                        }
#undef GV_ATT_CASE
                        if (sp != 0) {
                            // arrival on the (row, head) counter: a hint for the merging CTA (the tags validate the data)
                            if (tid == 0) red_relaxed_add(p.att_cnt + (size_t)rh * GV_ATTCNT_STRIDE, 1u);
                        } else {  // the CTA of split 0 merges the (row, head)
                            if (tid == 0) {  // one poller on the counter line; nobody polls lines that are still being written
                                const unsigned* cnt = p.att_cnt + (size_t)rh * GV_ATTCNT_STRIDE;
                                uint32_t spins = 0;
                                while (ld_relaxed_u32(cnt) < t_cnt + (unsigned)(nsplit - 1)) {
                                    GV_SPIN(spins, __LINE__, 0u, 0u);
                                }
                            }
                            bar_sync(1, MEGA_CONSUMERS);
                            if (tid < HD) {
                                const uint32_t tga = tg + TG_AO;
                                float M = -INFINITY, den = 0.0f, o = 0.0f;
                                for (int s0 = 0; s0 < nsplit; s0 += 4) {
                                    uint4 a[4];
                                    uint2 b[4];
#pragma unroll
                                    for (int q = 0; q < 4; ++q) {
                                        if (s0 + q < nsplit) {
                                            const int it = rh * nsplit + s0 + q;
                                            a[q] = ld_x16(p.att_ml + 2 * (size_t)(it * 2));
                                            b[q] = ld_x8(p.att_o + 2 * (size_t)(it * HD + tid));
                                        }
                                    }
#pragma unroll
                                    for (int q = 0; q < 4; ++q) {
                                        if (s0 + q < nsplit) {
                                            const int it = rh * nsplit + s0 + q;
                                            uint32_t spins = 0;
                                            while (!__all_sync(0xffffffffu, tags_ok(a[q], tga, tmask) && b[q].y == tga)) {  // tid < HD: whole warps
                                                GV_SPIN(spins, __LINE__, 0u, 0u);
                                                a[q] = ld_x16(p.att_ml + 2 * (size_t)(it * 2));
                                                b[q] = ld_x8(p.att_o + 2 * (size_t)(it * HD + tid));
                                            }
                                            const float mq = __uint_as_float(a[q].x), lq = __uint_as_float(a[q].z);
                                            const float Mn = fmaxf(M, mq);
                                            const float c_old = expf(M - Mn);  // first split: exp(-inf) = 0
                                            const float c_new = expf(mq - Mn);
                                            den = den * c_old + lq * c_new;
                                            o = o * c_old + __uint_as_float(b[q].x) * c_new;
                                            M = Mn;
                                        }
                                    }
                                }
                                st_tagged(p.ao + r * x_row, h * HD + tid, o / den, tga);
                            }
                            hop_arrive(hc + HC_AO * GV_HOP_STRIDE, tid);
                        }
                    }
                }
                if (nsplit > 1) t_cnt += (unsigned)(nsplit - 1);  // arrivals per (row, head) counter and layer
                t_ao += (unsigned)(NB * H);
                mark(2);  // attention item
                // ---- PROJ: attention output rows -> xs; x1 = x + o . W_proj + b ----
                hop_wait(hc + HC_AO * GV_HOP_STRIDE, t_ao, tid, tmask, settle, nullptr, near_ao);
                mark(3);  // AO hop
                load_rows<D>(p.ao, tg + TG_AO, xs, NB, tid);
                bar_sync(1, MEGA_CONSUMERS);
                mark(4);  // merge
                dot_phase<NXV>(ring, cs, nun[PH_PROJ], xs, NB, tid, [&](int u, int r, float dot, float c2, float) {
                    st_tagged(p.x1 + r * x_row, ubeg[PH_PROJ] + u, resid[r * 8 + u] + (dot + c2), tg + TG_X1);
                });
                cs.gt += (uint32_t)ntl[PH_PROJ];
                hop_arrive(hc + HC_X1 * GV_HOP_STRIDE, tid);
                mark(5);  // PROJ
                // ---- FC + P2: u = gelu_new(LN2(x1) . W_fc + b) (kept in this CTA) -> partial of u . W_proj2 ----
                hop_wait(hc + HC_X1 * GV_HOP_STRIDE, (lc + 1u) * (unsigned)G, tid, tmask, settle, nullptr, near);
                load_rows<D>(p.x1, tg + TG_X1, xs, NB, tid);
                bar_sync(1, MEGA_CONSUMERS);
                mark(6);  // x1 hop + load
                row_stats<NXV>(xs, stats, NB, warp, lane);
                bar_sync(1, MEGA_CONSUMERS);  // statistics visible to the epilogues
                dot_phase<NXV>(ring, cs, nun[PH_FC], xs, NB, tid, [&](int u, int r, float dot, float c2, float c1) {
                    const float mean = stats[2 * r], rstd = stats[2 * r + 1];
                    us[u * NBR + r] = gelu_new(fmaf(rstd, fmaf(-mean, c1, dot), c2));
                });
                cs.gt += (uint32_t)ntl[PH_FC];
                long long tb0 = 0;
                if (profiling) tb0 = clock64();
                bar_sync(1, MEGA_CONSUMERS);
                if (profiling) prof[15] += (unsigned long long)(clock64() - tb0);  // [15]: barrier after FC (thread 0's wait)
                mark(7);  // stats + FC
                if (profiling) cs.wacc = prof + 13;
                outer_phase<NXV>(ring, cs, nun[PH_P2], us, NB, tid, p.pp + 2 * (size_t)cta * NBR * D, tg + TG_PP);
                if (profiling) cs.wacc = prof + 12;
                cs.gt += (uint32_t)ntl[PH_P2];
                hop_arrive(hc + HC_PP * GV_HOP_STRIDE, tid);
                mark(8);  // P2
                // ---- RED: x2 = x1 + b + sum over CTAs of the partials (8 outputs x 8 rows per reducer CTA) ----
                if (cta < n_red) {
                    hop_wait(hc + HC_PP * GV_HOP_STRIDE, (lc + 1u) * (unsigned)G, tid, tmask, settle, nullptr, near);
                    const uint32_t tgp = tg + TG_PP;
                    // thread: output pair o2 = lane (row o2 / 4, columns 8 cta + 2 (o2 % 4), +1), sources warp, warp + 8, ...
                    const int o2 = lane, q = warp;
                    float acc0 = 0.0f, acc1 = 0.0f;
                    if ((o2 >> 2) < NB) {
                        const float* src = p.pp + 2 * ((size_t)cta * 64 + 2 * o2);
                        const size_t sstride = 2 * (size_t)NBR * D;
                        for (int s0 = q; s0 < G; s0 += 80) {  // ten 16-byte loads in flight
                            uint4 v[10];
#pragma unroll
                            for (int k = 0; k < 10; ++k)
                                if (s0 + 8 * k < G) v[k] = ld_x16(src + (size_t)(s0 + 8 * k) * sstride);
#pragma unroll
                            for (int k = 0; k < 10; ++k) {
                                if (s0 + 8 * k < G) {
                                    uint32_t spins = 0;
                                    while (!tags_ok(v[k], tgp, tmask)) {
                                        GV_SPIN(spins, __LINE__, 0u, 0u);
                                        v[k] = ld_x16(src + (size_t)(s0 + 8 * k) * sstride);
                                    }
                                    acc0 += __uint_as_float(v[k].x);
                                    acc1 += __uint_as_float(v[k].z);
                                }
                            }
                        }
                    }
                    *reinterpret_cast<float2*>(red4 + q * 64 + 2 * o2) = make_float2(acc0, acc1);
                    bar_sync(1, MEGA_CONSUMERS);
                    if (tid < 64 && (tid >> 3) < NB) {
                        const int r = tid >> 3, col = cta * 8 + (tid & 7);
                        const float b2 = __ldg(p.blob + p.proj2_b_off + (long long)l * p.layer_stride + col);
                        float s = 0.0f;
#pragma unroll
                        for (int w = 0; w < MEGA_WARPS; ++w) s += red4[w * 64 + tid];
                        st_tagged(p.x2 + r * x_row, col, (xs[r * UF + col] + b2) + s, tg + TG_X2);
                    }
                    hop_arrive(hc + HC_X2 * GV_HOP_STRIDE, tid);
                    mark(9);  // reduce (reducer CTAs)
                }
                lc += 1u;
            }
            // ---- HEAD: ln_f -> final_norm -> latent rows z ; logits = z . mel_head^T + b ----
            {
                const uint32_t tg = tbase + (uint32_t)GV_TAGS_PER_LAYER * (uint32_t)p.L;
                hop_wait(hc + HC_X2 * GV_HOP_STRIDE, lc * (unsigned)n_red, tid, tmask, settle, nullptr, near);
                load_rows<D>(p.x2, tg - (uint32_t)GV_TAGS_PER_LAYER + TG_X2, xs, NB, tid);
                bar_sync(1, MEGA_CONSUMERS);
                const float* lnp = tile_wait(ring, Cons{cs.gt, nullptr}, cs.gt, lane);  // all warps read the parameter tile
                row_layernorm2<NXV>(xs, lnp, NB, warp, lane);
                bar_sync(1, MEGA_CONSUMERS);
                if (warp == 0) tile_release(ring, cs.gt, lane, (uint32_t)MEGA_WARPS);
                cs.gt += 1u;
                if (cta < NB && xvalid)  // the latent of the row this CTA samples
                    *reinterpret_cast<float4*>(p.latents_out + ((size_t)i * NB + cta) * D + 4 * tid) =
                        *reinterpret_cast<const float4*>(xs + cta * UF + 4 * tid);
                dot_phase<NXV>(ring, cs, nun[PH_HEAD], xs, NB, tid, [&](int u, int r, float dot, float c2, float) {
                    st_tagged(p.lg + r * lg_row, ubeg[PH_HEAD] + u, dot + c2, tg);
                });
                cs.gt += (uint32_t)ntl[PH_HEAD];
                hop_arrive(hc + HC_LG * GV_HOP_STRIDE, tid);
                t_lg += (unsigned)G;
                if (cta < NB) {
                    hop_wait(hc + HC_LG * GV_HOP_STRIDE, t_lg, tid, tmask, settle, nullptr, near);
                    const float* lgr = p.lg + cta * lg_row;
                    for (int e = 2 * tid; e < p.V; e += 2 * MEGA_CONSUMERS) {
                        if (e + 1 < p.V) {
                            const float2 v = ld_tagged2(lgr, e, tg, tmask);
                            slog[e] = v.x;
                            slog[e + 1] = v.y;
                        } else {
                            slog[e] = ld_tagged1(lgr, e, tg, tmask);
                        }
                    }
                }
            }
        } else if (cta < NB) {
            // logits / latent left pending by the prefill (per-op kernels; plain arrays)
            for (int e = tid; e < p.V; e += MEGA_CONSUMERS) slog[e] = ldcg(p.pend_logits + (size_t)cta * p.V + e);
            if (xvalid)
                *reinterpret_cast<float4*>(p.latents_out + ((size_t)i * NB + cta) * D + 4 * tid) =
                    ldcg4(p.pend_latent + (size_t)cta * D + 4 * tid);
        }
        bar_sync(1, MEGA_CONSUMERS);  // slog complete; xs (aliased by the sort keys) is dead
        mark(10);  // head (+ logits hop on the sampling CTAs)
        // ------------- sample + emit: row r by CTA r -------------
        if (cta < NB) {
            int tok = sample_token([&](int e) { return slog[e]; }, seen, scfg,
                                   p.noise ? p.noise + ((size_t)i * NB + cta) * p.V : nullptr, p.seed, (uint32_t)n, (uint32_t)cta, keys,
                                   fscr, iscr, tid, [] { bar_sync(1, MEGA_CONSUMERS); });
            if (p.forced) {
                const long long f = p.forced[(size_t)i * NB + cta];
                tok = (f >= 0 && f < (long long)p.V) ? (int)f : p.stop_token;
                if (tok != (int)f && tid == 0) atomicOr(p.bad_ids, 1);
            }
            if (!p.ignore_eos && fin[cta]) tok = p.stop_token;  // finished rows emit the pad (== eos) token
            if (tid == 0) {
                p.ids_out[(size_t)i * NB + cta] = tok;
                seen[tok] = 1;
                st_tagged(p.tokx, cta, __int_as_float(tok), tg_tok);
            }
            if (p.logits_out)
                for (int q = tid; q < p.V; q += MEGA_CONSUMERS) p.logits_out[((size_t)i * NB + cta) * p.V + q] = slog[q];
        }
        // every CTA learns the B tokens of this step
        if (tid < NB) {
            const int tok = __float_as_int(ld_tagged1(p.tokx, tid, tg_tok, tmask));
            ltok[tid] = tok;
            if (!p.ignore_eos && tok == p.stop_token) fin[tid] = 1;
        }
        n += 1;
        emitted += 1;
        bar_sync(1, MEGA_CONSUMERS);  // ltok / fin / seen visible; slog / keys free
        mark(11);  // sampling + token exchange
        int all_fin = 1;
        for (int r = 0; r < NB; ++r) all_fin &= fin[r];
        if (all_fin || n >= p.max_total) {
            done = 1;
            break;
        }
    }
    if (profiling)
        for (int k = 0; k < 16; ++k) p.trace[(size_t)cta * 16 + k] = prof[k];
    // tell the producer to stop (it may be blocked on a full ring or still have copies in flight)
    if (tid == 0) {
        ctl[1] = (int)cs.gt;
        __threadfence_block();
        ctl[0] = 1;
    }
    if (cta < NB)
        for (int q = tid; q < p.Vpad; q += MEGA_CONSUMERS) p.seen_out[(size_t)cta * p.Vpad + q] = seen[q];
    if (cta == 0 && tid == 0) {
        GenState* so = p.st_out;
        so->n_emitted = n;
        so->done = done;
        so->has_pending = 0;
        so->P = st->P;
        so->B = st->B;
        for (int r = 0; r < NB; ++r) {
            so->finished[r] = fin[r];
            so->last_tok[r] = ltok[r];
        }
        p.status[0] = emitted;
        p.status[1] = done;
        if (atomicExch(p.bad_ids, 0) != 0) p.status[2] = 1;  // an id was clamped since the last status (embedding kernels / forced ids)
    }
}

// ---------------------------------------------------------------------------------------------
// host launcher
// ---------------------------------------------------------------------------------------------
size_t batch_smem_bytes(int D, int Vpad) {
    size_t off = (size_t)NSLOT * slot_floats(D) * sizeof(float);
    off = (off + 127) & ~size_t(127);
    off += batch_xs_bytes(D);
    off += B_SCR_BYTES;
    off += (size_t)32 * NBR * sizeof(float);
    off += 2 * NBR * sizeof(float) + 8 * NBR * sizeof(float);
    off += (size_t)Vpad * sizeof(float);
    off += 32 * sizeof(uint64_t);
    off += 16 * sizeof(float) + 16 * sizeof(int) + 8 * sizeof(int);
    off += 2 * NBR * sizeof(int);
    off += 16 * sizeof(unsigned long long);
    off += Vpad;
    return (off + 15) & ~size_t(15);
}

template <int NXV>
static cudaError_t launch_b(const MegaParams& p, int grid, size_t smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(decode_batch_kernel<NXV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    MegaParams pp = p;
    void* args[] = {&pp};
    return cudaLaunchCooperativeKernel((void*)decode_batch_kernel<NXV>, dim3(grid), dim3(MEGA_THREADS), args, smem, st);
}

cudaError_t launch_decode_batch(const MegaParams& p, int grid, cudaStream_t st) {
    const int hd = p.D / p.H;
    if (p.D % 128 || p.D > 1024 || !(hd == 32 || hd == 64 || hd == 128 || hd == 256)) return cudaErrorInvalidValue;
    if (p.B < 1 || p.B > NBR || p.B * p.H > grid || p.D / 8 > grid) return cudaErrorInvalidValue;
    if ((int)(((long long)p.D + grid - 1) / grid) > 8) return cudaErrorInvalidValue;  // resid[] holds 8 columns per row
    const size_t smem = batch_smem_bytes(p.D, p.Vpad);
    switch (p.D / 128) {
        case 1: return launch_b<1>(p, grid, smem, st);
        case 2: return launch_b<2>(p, grid, smem, st);
        case 4: return launch_b<4>(p, grid, smem, st);
        case 8: return launch_b<8>(p, grid, smem, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace gv
