// Batched fused persistent decode kernel: up to GV_BATCH_ROWS (8) equal-length rows share ONE pass of the weight
// stream per generated token.  Same loop as decode_mega.cu (layers/stream_generator.py:809-881 around the cached
// forward of layers/gpt_inference.py:92-112), with the reference's batch semantics: every row uses the same
// position index (gpt_inference.py:92-96), rows finish independently and finished rows emit the stop token
// until the last one ends (stream_generator.py:860-881).
//
// What changes against the single-row kernel:
//   * GEMV -> skinny GEMM out of the same ring.  The CTA's units (output columns, K = D) are taken 16 at a time:
//     D[16 units x 8 rows] = W[16 x K] . X^T[K x 8] with warp-level mma (m16n8k8), the 8 warps splitting K.
//     fp32 parity on tensor cores by 3xTF32 (x = hi + lo, hi = tf32(x); lo.hi + hi.lo + hi.hi in an fp32
//     accumulator).  Operands come straight from shared memory with ldmatrix: the unit pitch of the stream
//     (D + 4 floats) puts the 8 row addresses of every 8 x 4-float matrix in 8 different 16-byte bank groups.
//     (tcgen05 has no shape for this: its smallest tile is 64 rows, and the activations would have to be a
//     [64 x K] operand per CTA; with 8 rows the step stays HBM-bound, the tensor work is a few % of it.)
//   * mlp.c_proj stays split along K (the CTA that computed u[:, k] owns row k of W_proj2):
//     D[16 outputs x 8 rows] += W^T[16 x 8 units] . u^T[8 units x 8 rows], outputs permuted inside 32-column
//     blocks so the operand loads are conflict-free; the per-CTA partials [8 rows][D] go to the reducer CTAs.
//   * one activation buffer xs[8][D + 4] in shared memory is the B operand of every phase (x -> attention output
//     -> x1 -> latent); LayerNorm statistics per row by one warp each.
//   * attention items = (row, head, key range), H * B * nsplit <= grid; the item code is the single-row one.
//   * row r is sampled by CTA r (its own repetition-penalty set in shared memory) and the token is published
//     through a tagged word; every CTA reads the B tokens for the next step's embeddings.
#define GV_RING_NSLOT GV_BATCH_NSLOT
#define GV_MEGA_NS megab
#include "mega_dev.cuh"

namespace gv {
using namespace megab;

#define NBR GV_BATCH_ROWS
#define B_PART_FLOATS (2 * MEGA_WARPS * 128)  // [parity][warp][16 units x 8 rows] cross-warp partials of the dot phases
#define B_US_PITCH 36                         // gelu(fc) values of this CTA: [8 rows][32 units + 4]
#define B_SCR_BYTES 9728                      // attention scratch (as the single-row kernel) | reducer partials

// ---------------------------------------------------------------------------------------------
// warp-level tensor-core primitives (legacy mma path; SASS: LDSM, HMMA.1688.F32.TF32)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr)
                 : "memory");
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr) : "memory");
}
// x = hi + lo with hi = x rounded to tf32, lo = (x - hi) rounded to tf32
__device__ __forceinline__ void split_tf32(uint32_t x, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(__uint_as_float(x)));
    const float l = __uint_as_float(x) - __uint_as_float(hi);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(l));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// c += A . B in 3xTF32 (small terms first)
__device__ __forceinline__ void mma_3x(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], const uint32_t (&bh)[2],
                                       const uint32_t (&bl)[2]) {
    mma_tf32(c, al[0], al[1], al[2], al[3], bh[0], bh[1]);
    mma_tf32(c, ah[0], ah[1], ah[2], ah[3], bl[0], bl[1]);
    mma_tf32(c, ah[0], ah[1], ah[2], ah[3], bh[0], bh[1]);
}

// ---------------------------------------------------------------------------------------------
// dot phase (QKV, attn c_proj, FC, logits head): for every unit u < nunits of this CTA and every row r,
//   dot[u][r] = sum_k W'[k][u] xs[r][k];   epi(u, r, dot, c2, c1)  is called once per (u, r < nb).
// Units are taken 16 at a time (four ring tiles); warp w covers k in [w D/8, (w+1) D/8); the eight per-warp
// partial tiles meet in `part` (double-buffered by group parity: one block barrier per group).
// Every warp reads every tile: the ring's empty barriers count 8 arrivals.
// ---------------------------------------------------------------------------------------------
template <int NXV, class Epi>
__device__ __forceinline__ void dot_phase(const Ring& ring, const Cons& cs, int nunits, const float* xs, float* part, int nb,
                                          int tid, Epi epi) {
    constexpr int D = NXV * 128, UF = D + 4, KW = D / MEGA_WARPS, KS = KW / 8;
    const int warp = tid >> 5, lane = tid & 31;
    const int ngroups = (nunits + 15) >> 4;
    const int ntiles = (nunits + UPT - 1) / UPT;
    // B operand (rows of xs): lanes 0-7 address rows 0-7 at k + 0, lanes 8-15 at k + 4 (lanes 16-31: ignored, kept valid)
    const uint32_t b_addr = smem_u32(xs + (lane & 7) * UF + warp * KW + ((lane >> 3) & 1) * 4);
    for (int g = 0; g < ngroups; ++g) {
        const int u0 = g << 4;
        const int nu = min(16, nunits - u0);
        const int t0 = g << 2, nt = min(4, ntiles - t0);
        if (lane < nt) tile_ready_wait(ring, cs.gt + (uint32_t)(t0 + lane));
        __syncwarp();
        // A operand (units): matrix m = lane / 8: units u0 + (m & 1) * 8 + lane % 8 at k + (m >> 1) * 4; rows past the
        // end of the phase re-read unit u0 (finite values; their outputs are dropped)
        int ua = u0 + ((lane >> 3) & 1) * 8 + (lane & 7);
        if (ua >= nunits) ua = u0;
        const uint32_t a_addr =
            smem_u32(slot_ptr(ring, cs.gt + (uint32_t)(ua >> 2)) + (ua & 3) * UF + warp * KW + (lane >> 4) * 4);
        // epilogue constants of the unit this thread will finish (threads 0..127: unit u0 + tid / 8)
        float2 cc = make_float2(0.f, 0.f);
        const int ue = u0 + (tid >> 3);
        if (tid < 128 && ue < nunits) cc = *reinterpret_cast<const float2*>(slot_ptr(ring, cs.gt + (uint32_t)(ue >> 2)) + (ue & 3) * UF + D);
        float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int s = 0; s < KS; ++s) {
            uint32_t a[4], b[2], ah[4], al[4], bh[2], bl[2];
            ldsm_x4(a_addr + s * 32, a[0], a[1], a[2], a[3]);
            ldsm_x2(b_addr + s * 32, b[0], b[1]);
#pragma unroll
            for (int i = 0; i < 4; ++i) split_tf32(a[i], ah[i], al[i]);
#pragma unroll
            for (int i = 0; i < 2; ++i) split_tf32(b[i], bh[i], bl[i]);
            mma_3x(c, ah, al, bh, bl);
        }
        // c[0], c[1]: unit lane / 4, rows 2 (lane % 4), +1;  c[2], c[3]: unit lane / 4 + 8
        float* pw = part + (g & 1) * (MEGA_WARPS * 128) + warp * 128;
        *reinterpret_cast<float2*>(pw + (lane >> 2) * 8 + 2 * (lane & 3)) = make_float2(c[0], c[1]);
        *reinterpret_cast<float2*>(pw + ((lane >> 2) + 8) * 8 + 2 * (lane & 3)) = make_float2(c[2], c[3]);
        __syncwarp();
        if (lane < nt) mbar_arrive(&ring.empty[(cs.gt + (uint32_t)(t0 + lane)) % NSLOT]);
        bar_sync(1, MEGA_CONSUMERS);
        if (tid < 128) {
            const int ul = tid >> 3, r = tid & 7;
            if (ul < nu && r < nb) {
                const float* pr = part + (g & 1) * (MEGA_WARPS * 128) + ul * 8 + r;
                float s = 0.0f;
#pragma unroll
                for (int w = 0; w < MEGA_WARPS; ++w) s += pr[w * 128];
                epi(u0 + ul, r, s, cc.x, cc.y);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// mlp.c_proj split along K: unit k = row k of W_proj2 (owned by the CTA that computed u[:, k]).
//   part[r][n] = sum_k us[r][k] W[k][n]    for all D outputs n, rows r
// as D[16 outputs x 8 rows] += A[16 outputs x 8 units] . B[8 units x 8 rows]: A element (output i, unit k) is
// slot(k)[col(i)] with col(i) = c0 + i % 4 + 16 (i / 4 % 2) + 4 (i / 8) inside a 32-column block, so the 32 lanes of
// one operand load (unit = lane % 4, output = lane / 4) hit 32 different banks (unit pitch = D + 4 floats).
// Warp w owns m-tiles w NXV .. w NXV + NXV - 1 (16 outputs each).  Two ring tiles (8 units) per k-step.
// Results go straight to the exchange buffer pp[cta][n / 8][row][n % 8] (tagged).
// ---------------------------------------------------------------------------------------------
template <int NXV>
__device__ __forceinline__ void outer_phase(const Ring& ring, const Cons& cs, int nunits, const float* us, int nb, int tid,
                                            float* pp_cta, uint32_t tag) {
    constexpr int D = NXV * 128, UF = D + 4;
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, j = lane & 3;
    const int ntiles = (nunits + UPT - 1) / UPT;
    const int nks = (nunits + 7) >> 3;
    float acc[NXV][4];
#pragma unroll
    for (int m = 0; m < NXV; ++m) acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0.0f;
    // output columns of this lane in m-tile T: rows g and g + 8 of the tile
    const int cg0 = (g & 3) + 16 * (g >> 2);  // col(g) - c0;  col(g + 8) = col(g) + 4
    for (int ks = 0; ks < nks; ++ks) {
        const int t0 = 2 * ks, nt = min(2, ntiles - t0);
        if (lane < nt) tile_ready_wait(ring, cs.gt + (uint32_t)(t0 + lane));
        __syncwarp();
        // B operand: b0 = us[row g][8 ks + j], b1 = us[row g][8 ks + j + 4]   (zero beyond nunits / nb)
        uint32_t bh[2], bl[2];
        split_tf32(__float_as_uint(us[g * B_US_PITCH + 8 * ks + j]), bh[0], bl[0]);
        split_tf32(__float_as_uint(us[g * B_US_PITCH + 8 * ks + j + 4]), bh[1], bl[1]);
        // A operand rows: units 8 ks + j (tile t0, row j) and 8 ks + j + 4 (tile t0 + 1, row j); units past the end
        // re-read unit 8 ks (finite weights; multiplied by us = 0)
        const int k0 = 8 * ks + j, k1 = k0 + 4;
        const float* wz = slot_ptr(ring, cs.gt + (uint32_t)t0);
        const float* w0 = (k0 < nunits) ? wz + j * UF : wz;
        const float* w1 = (k1 < nunits) ? slot_ptr(ring, cs.gt + (uint32_t)(t0 + 1)) + j * UF : wz;
#pragma unroll
        for (int m = 0; m < NXV; ++m) {
            const int T = warp * NXV + m;
            const int c0 = 32 * (T >> 1) + 8 * (T & 1) + cg0;
            uint32_t a[4], ah[4], al[4];
            a[0] = __float_as_uint(w0[c0]);
            a[1] = __float_as_uint(w0[c0 + 4]);
            a[2] = __float_as_uint(w1[c0]);
            a[3] = __float_as_uint(w1[c0 + 4]);
#pragma unroll
            for (int i = 0; i < 4; ++i) split_tf32(a[i], ah[i], al[i]);
            mma_3x(acc[m], ah, al, bh, bl);
        }
        __syncwarp();
        if (lane < nt) mbar_arrive(&ring.empty[(cs.gt + (uint32_t)(t0 + lane)) % NSLOT]);
    }
    // acc[m][0], [1]: output col(g), rows 2j, 2j+1;  acc[m][2], [3]: output col(g) + 4
#pragma unroll
    for (int m = 0; m < NXV; ++m) {
        const int T = warp * NXV + m;
        const int n0 = 32 * (T >> 1) + 8 * (T & 1) + cg0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int n = n0 + 4 * h;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int r = 2 * j + q;
                if (r < nb) st_tagged(pp_cta, ((n >> 3) * NBR + r) * 8 + (n & 7), acc[m][2 * h + q], tag);
            }
        }
    }
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
                while (!(tags_ok(a[q], tag, 0xffffffffu) && tags_ok(b[q], tag, 0xffffffffu))) {
                    if (++spins > MEGA_SPIN_LIMIT) __trap();
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
    off += (size_t)NBR * UF * sizeof(float);
    float* part = reinterpret_cast<float*>(smem_raw + off);
    off += (size_t)B_PART_FLOATS * sizeof(float);
    float* att_sc = reinterpret_cast<float*>(smem_raw + off);         // attention: per-warp (max, sum) + q staging
    float* att_op = reinterpret_cast<float*>(smem_raw + off + 1280);  // attention: [8][hd] per-warp PV partials
    float* red4 = reinterpret_cast<float*>(smem_raw + off);           // reducer CTAs: [4][64] partial sums
    off += B_SCR_BYTES;
    float* us = reinterpret_cast<float*>(smem_raw + off);  // [8][B_US_PITCH] gelu(fc) of this CTA's units
    off += (size_t)NBR * B_US_PITCH * sizeof(float);
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
    for (int i = tid_all; i < NBR * B_US_PITCH; i += MEGA_THREADS) us[i] = 0.0f;
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
    unsigned lc = 0, t_ao = 0, t_lg = 0;
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
            const unsigned near_ao = p.hop_near_ao >= 0 ? (unsigned)p.hop_near_ao : (unsigned)min(max(n_items / 3, 1), 4);
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
                    hop_wait(hc + HC_X2 * GV_HOP_STRIDE, lc * (unsigned)n_red, tid, tmask, 0u, nullptr, near);
                    load_rows<D>(p.x2, tg - (uint32_t)GV_TAGS_PER_LAYER + TG_X2, xs, NB, tid);
                }
                bar_sync(1, MEGA_CONSUMERS);
                // ---- QKV: [q|k|v] = LN1(x) . W_attn + b  (LN folded into the packed weights: statistics enter in the epilogue) ----
                row_stats<NXV>(xs, stats, NB, warp, lane);
                if (tid < NBR * 8) {  // block input at this CTA's attn c_proj columns (residual of the PROJ epilogue)
                    const int r = tid >> 3, u = tid & 7;
                    if (r < NB && u < nun[PH_PROJ]) resid[tid] = xs[r * UF + ubeg[PH_PROJ] + u];
                }
                dot_phase<NXV>(ring, cs, nun[PH_QKV], xs, part, NB, tid, [&](int u, int r, float dot, float c2, float c1) {
                    const float mean = stats[2 * r], rstd = stats[2 * r + 1];
                    st_tagged(p.xq + r * xq_row, ubeg[PH_QKV] + u, fmaf(rstd, fmaf(-mean, c1, dot), c2), tg + TG_XQ);
                });
                cs.gt += (uint32_t)ntl[PH_QKV];
                // ---- ATT: (row, head, key-range) items on the first n_items CTAs ----
                if (cta < n_items) {
                    const int rh = cta / nsplit, sp = cta % nsplit;
                    const int r = rh / H, h = rh % H;
                    const int j0 = sp * chunk, j1 = min(S, j0 + chunk);
                    float* kh = kc + ((size_t)r * H + h) * p.S_max * HD;
                    float* vh = vc + ((size_t)r * H + h) * p.S_max * HD;
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
                    hop_arrive(hc + HC_AO * GV_HOP_STRIDE, tid);
                }
                t_ao += (unsigned)n_items;
                // ---- PROJ: merge the items' partials -> attention output rows (xs); x1 = x + o . W_proj + b ----
                hop_wait(hc + HC_AO * GV_HOP_STRIDE, t_ao, tid, tmask, 0u, nullptr, near_ao);
                if (xvalid) {
                    const int h = (4 * tid) / HD, d = (4 * tid) % HD;
                    const uint32_t tga = tg + TG_AO;
                    for (int r = 0; r < NB; ++r) {
                        float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f, M = -INFINITY, den = 0.0f;
                        for (int s0 = 0; s0 < nsplit; s0 += 4) {  // loads of four splits in flight, merged in split order
                            uint4 a[4], b[4], c[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                if (s0 + q < nsplit) {
                                    const int it = (r * H + h) * nsplit + s0 + q;
                                    a[q] = ld_x16(p.att_ml + 2 * (size_t)(it * 2));
                                    b[q] = ld_x16(p.att_o + 2 * (size_t)(it * HD + d));
                                    c[q] = ld_x16(p.att_o + 2 * (size_t)(it * HD + d + 2));
                                }
                            }
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                if (s0 + q < nsplit) {
                                    const int it = (r * H + h) * nsplit + s0 + q;
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
                        *reinterpret_cast<float4*>(xs + r * UF + 4 * tid) = make_float4(o0 / den, o1 / den, o2 / den, o3 / den);
                    }
                }
                bar_sync(1, MEGA_CONSUMERS);
                dot_phase<NXV>(ring, cs, nun[PH_PROJ], xs, part, NB, tid, [&](int u, int r, float dot, float c2, float) {
                    st_tagged(p.x1 + r * x_row, ubeg[PH_PROJ] + u, resid[r * 8 + u] + (dot + c2), tg + TG_X1);
                });
                cs.gt += (uint32_t)ntl[PH_PROJ];
                hop_arrive(hc + HC_X1 * GV_HOP_STRIDE, tid);
                // ---- FC + P2: u = gelu_new(LN2(x1) . W_fc + b) (kept in this CTA) -> partial of u . W_proj2 ----
                hop_wait(hc + HC_X1 * GV_HOP_STRIDE, (lc + 1u) * (unsigned)G, tid, tmask, 0u, nullptr, near);
                load_rows<D>(p.x1, tg + TG_X1, xs, NB, tid);
                bar_sync(1, MEGA_CONSUMERS);
                row_stats<NXV>(xs, stats, NB, warp, lane);
                dot_phase<NXV>(ring, cs, nun[PH_FC], xs, part, NB, tid, [&](int u, int r, float dot, float c2, float c1) {
                    const float mean = stats[2 * r], rstd = stats[2 * r + 1];
                    us[r * B_US_PITCH + u] = gelu_new(fmaf(rstd, fmaf(-mean, c1, dot), c2));
                });
                cs.gt += (uint32_t)ntl[PH_FC];
                bar_sync(1, MEGA_CONSUMERS);
                outer_phase<NXV>(ring, cs, nun[PH_P2], us, NB, tid, p.pp + 2 * (size_t)cta * NBR * D, tg + TG_PP);
                cs.gt += (uint32_t)ntl[PH_P2];
                hop_arrive(hc + HC_PP * GV_HOP_STRIDE, tid);
                // ---- RED: x2 = x1 + b + sum over CTAs of the partials (8 outputs x 8 rows per reducer CTA) ----
                if (cta < n_red) {
                    hop_wait(hc + HC_PP * GV_HOP_STRIDE, (lc + 1u) * (unsigned)G, tid, tmask, 0u, nullptr, near);
                    const uint32_t tgp = tg + TG_PP;
                    const int o = tid & 63, q = tid >> 6;  // output (row o / 8, column 8 cta + o % 8); sources q, q + 4, ...
                    float acc = 0.0f;
                    if ((o >> 3) < NB) {
                        const float* src = p.pp + 2 * ((size_t)cta * 64 + o);
                        const size_t sstride = 2 * (size_t)NBR * D;
                        for (int s0 = q; s0 < G; s0 += 32) {
                            uint2 v[8];
#pragma unroll
                            for (int k = 0; k < 8; ++k)
                                if (s0 + 4 * k < G) v[k] = ld_x8(src + (size_t)(s0 + 4 * k) * sstride);
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                if (s0 + 4 * k < G) {
                                    uint32_t spins = 0;
                                    while (v[k].y != tgp) {
                                        if (++spins > MEGA_SPIN_LIMIT) __trap();
                                        v[k] = ld_x8(src + (size_t)(s0 + 4 * k) * sstride);
                                    }
                                    acc += __uint_as_float(v[k].x);
                                }
                            }
                        }
                    }
                    red4[q * 64 + o] = acc;
                    bar_sync(1, MEGA_CONSUMERS);
                    if (tid < 64 && (tid >> 3) < NB) {
                        const int r = tid >> 3, col = cta * 8 + (tid & 7);
                        const float b2 = __ldg(p.blob + p.proj2_b_off + (long long)l * p.layer_stride + col);
                        const float s = (red4[tid] + red4[64 + tid]) + (red4[128 + tid] + red4[192 + tid]);
                        st_tagged(p.x2 + r * x_row, col, (xs[r * UF + col] + b2) + s, tg + TG_X2);
                    }
                    hop_arrive(hc + HC_X2 * GV_HOP_STRIDE, tid);
                }
                lc += 1u;
            }
            // ---- HEAD: ln_f -> final_norm -> latent rows z ; logits = z . mel_head^T + b ----
            {
                const uint32_t tg = tbase + (uint32_t)GV_TAGS_PER_LAYER * (uint32_t)p.L;
                hop_wait(hc + HC_X2 * GV_HOP_STRIDE, lc * (unsigned)n_red, tid, tmask, 0u, nullptr, near);
                load_rows<D>(p.x2, tg - (uint32_t)GV_TAGS_PER_LAYER + TG_X2, xs, NB, tid);
                bar_sync(1, MEGA_CONSUMERS);
                const float* lnp = tile_wait(ring, cs, cs.gt, lane);  // all warps read the parameter tile
                row_layernorm2<NXV>(xs, lnp, NB, warp, lane);
                bar_sync(1, MEGA_CONSUMERS);
                if (warp == 0) tile_release(ring, cs.gt, lane, (uint32_t)MEGA_WARPS);
                cs.gt += 1u;
                if (cta < NB && xvalid)  // the latent of the row this CTA samples
                    *reinterpret_cast<float4*>(p.latents_out + ((size_t)i * NB + cta) * D + 4 * tid) =
                        *reinterpret_cast<const float4*>(xs + cta * UF + 4 * tid);
                dot_phase<NXV>(ring, cs, nun[PH_HEAD], xs, part, NB, tid, [&](int u, int r, float dot, float c2, float) {
                    st_tagged(p.lg + r * lg_row, ubeg[PH_HEAD] + u, dot + c2, tg);
                });
                cs.gt += (uint32_t)ntl[PH_HEAD];
                hop_arrive(hc + HC_LG * GV_HOP_STRIDE, tid);
                t_lg += (unsigned)G;
                if (cta < NB) {
                    hop_wait(hc + HC_LG * GV_HOP_STRIDE, t_lg, tid, tmask, 0u, nullptr, near);
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
        int all_fin = 1;
        for (int r = 0; r < NB; ++r) all_fin &= fin[r];
        if (all_fin || n >= p.max_total) {
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
    off += (size_t)NBR * (D + 4) * sizeof(float);
    off += (size_t)B_PART_FLOATS * sizeof(float);
    off += B_SCR_BYTES;
    off += (size_t)NBR * B_US_PITCH * sizeof(float);
    off += 2 * NBR * sizeof(float) + 8 * NBR * sizeof(float);
    off += (size_t)Vpad * sizeof(float);
    off += 32 * sizeof(uint64_t);
    off += 16 * sizeof(float) + 16 * sizeof(int) + 8 * sizeof(int);
    off += 2 * NBR * sizeof(int);
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
