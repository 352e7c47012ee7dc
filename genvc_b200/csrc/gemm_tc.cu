// Dense fp32-accurate GEMM on the 5th-generation tensor cores (tcgen05, TMEM accumulators) for the
// batched parts of the path: prefill (layers/gpt_inference.py:81-112, 48..110 rows), the teacher-
// forced latent pass (layers/gpt.py:375-508, <= 255 rows) and the perceiver projections
// (layers/perceiver_encoder.py:108-151, 305-319).  C[M,N] = act(A[M,K] . W + bias) + residual.
//
// fp32 parity on tensor cores: 3xTF32.  Every operand is split x = hi + lo with hi = x rounded to
// TF32 (10 mantissa bits) and lo = x - hi (exact in fp32); the kernel accumulates
//   A_lo.B_hi + A_hi.B_lo + A_hi.B_hi        (the dropped A_lo.B_lo term is ~2^-22 relative)
// with tcgen05.mma.kind::tf32 into one fp32 TMEM accumulator, which keeps the greedy token ids of
// the prefill bit-identical to the fp32 reference (a plain TF32 GEMM, ~1e-3 relative, does not).
//
// These GEMMs are weight-streaming bound (M <= 600 rows against 4..17 MB of weights), so the kernel
// is built around the weight stream, not tensor-pipe occupancy:
//   * the weights are the M = 128 side of the MMA (D^T[n, m] = W^T . A^T): one CTA owns 128 output
//     columns n over one K range and ALL rows m of its row chunk (<= 128), so every weight byte is
//     read by exactly one CTA;
//   * at load time every matrix is pre-split (hi | lo) and pre-tiled into 32 KB blocks = one
//     pipeline stage = 128 columns x 32 k in the UMMA canonical K-major no-swizzle layout (8 x 16-byte
//     core matrices), see pack_tc_kernel.  A stage is ONE bulk TMA copy (cp.async.bulk, mbarrier
//     complete_tx) issued by a producer thread into a 3-4 stage ring: no in-kernel transform of
//     weights, HBM reads fully coalesced, 96-128 KB in flight per SM;
//   * warps 0-3 (loaders, then epilogue): read the fp32 activation rows (L2-resident) one stage
//     ahead in registers, split hi/lo and store them in the same canonical layout;
//     fence.proxy.async + mbarrier arrive hands the stage to the tensor core;
//   * warp 5, one elected thread: waits for the stage (weights landed + activations stored), issues
//     4 k-steps x 3 tcgen05.mma (M = 128, N = 64 | 128, K = 8) from shared-memory descriptors;
//     tcgen05.commit releases the stage;
//   * epilogue: tcgen05.ld (lane = output column n, register = row m: global stores coalesced along
//     n) -> bias / gelu_new / residual -> C, or raw partials for the deterministic split-K reduction
//     (splitk_epilogue_kernel in ops.cu) when the tile grid alone cannot fill 148 SMs.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include <cstdlib>
#include "ops.cuh"

namespace gv {

namespace tc {

constexpr int BN = 128;   // output columns per CTA (the M = 128 side of the MMA)
constexpr int KT = 32;    // k per stage
constexpr int W_STAGE_FLOATS = 2 * BN * KT;  // hi | lo
constexpr int W_STAGE_BYTES = W_STAGE_FLOATS * 4;  // 32 KB
constexpr int LOADERS = 128, THREADS = 192;

template <int MP>
struct Cfg {
    static constexpr int STAGES = MP == 64 ? 4 : 3;
    static constexpr int X_HALF = MP * KT * 4;  // bytes of one activation half (hi or lo) per stage
    static constexpr int STAGE_BYTES = W_STAGE_BYTES + 2 * X_HALF;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */;
    // instruction descriptor: D = F32, A = B = TF32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
    static constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(MP >> 3) << 17) | ((uint32_t)(BN >> 4) << 24);
};

// hi = x rounded to nearest TF32 (ties away), lo = x - hi
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    const uint32_t u = __float_as_uint(x);
    hi = __uint_as_float((u + 0x1000u) & 0xffffe000u);
    lo = x - hi;
}

// byte offset of (row, 16-byte k-chunk) inside a K-major no-swizzle [rows x 32 k] tile: a core matrix is 8 rows x
// 16 bytes = 128 contiguous bytes; the 8 k-chunks of a row group are adjacent (LBO = 128), row groups 1 KB apart (SBO)
__host__ __device__ __forceinline__ uint32_t core_off(int row, int kc) {
    return (uint32_t)(((row >> 3) * (KT / 4) + kc) * 128 + (row & 7) * 16);
}

// shared-memory matrix descriptor, K-major, no swizzle
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
    d |= (uint64_t)(128u >> 4) << 16;                   // LBO: between the two 16-byte K chunks of a k-step
    d |= (uint64_t)(((KT / 4) * 128u) >> 4) << 32;      // SBO: between 8-row groups
    d |= (uint64_t)1 << 46;                             // descriptor version (sm_100)
    return d;                                           // base offset 0, layout type 0 = SWIZZLE_NONE
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---------------------------------------------------------------------------------------------
// load-time packing: W (reference layout) -> [n tile][k stage][hi | lo][128 rows x 32 k canonical]
// ---------------------------------------------------------------------------------------------
__global__ void pack_tc_kernel(const float* __restrict__ W, int N, int K, int ldw, int w_nk, float* __restrict__ out) {
    const int nt = blockIdx.x, ks = blockIdx.y, nks = gridDim.y;
    float* blk = out + ((size_t)nt * nks + ks) * W_STAGE_FLOATS;
    for (int e = threadIdx.x; e < BN * KT; e += blockDim.x) {
        // coalesced along the contiguous dimension of the source
        int row, k;
        if (w_nk) { row = e / KT; k = e % KT; } else { k = e / BN; row = e % BN; }
        const int n = nt * BN + row, kg = ks * KT + k;
        float x = 0.0f;
        if (n < N) x = w_nk ? W[(size_t)n * ldw + kg] : W[(size_t)kg * ldw + n];
        float hi, lo;
        split_tf32(x, hi, lo);
        const uint32_t o = core_off(row, k >> 2) / 4 + (k & 3);
        blk[o] = hi;
        blk[BN * KT + o] = lo;
    }
}

template <int MP>
__global__ void __launch_bounds__(THREADS, 1) gemm_tc_kernel(GemmArgs a, const float* __restrict__ Wt, float* __restrict__ ws,
                                                             int k_chunk, const int* skip) {
    using C = Cfg<MP>;
    if (skip) {  // (the flag is written by an earlier kernel of the stream)
        asm volatile("griddepcontrol.wait;" ::: "memory");
        if (*skip) return;
    }
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
    uint64_t* empty = full + C::STAGES;
    uint64_t* done = empty + C::STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.x * BN, m0 = blockIdx.y * MP;
    const int kb = blockIdx.z * k_chunk, ke = min(a.K, kb + k_chunk);
    const int nst = (ke - kb) / KT;  // stages of this CTA (host guarantees multiples of KT)
    const int nks = a.K / KT;

    if (tid == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(&full[s], 4 + 1);  // one arrival per loader warp + the producer's expect_tx arrival
            mbar_init(&empty[s], 1);
        }
        mbar_init(done, 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(MP)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        // ================= activation loaders =================
        // rows [m0, m0 + MP) x 32 k per stage; 8 rows x 4 chunks per warp instruction; register double buffer:
        // the loads of stage it + 1 are in flight while stage it is split and stored
        constexpr int NQ = MP / 16;  // float4 per thread per stage
        const int r0 = lane & 7, c = lane >> 3;
        float4 va[2][NQ];
        auto load_stage = [&](int it, float4* pa) {
            const int k0 = kb + it * KT;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const int row = (MP / 4) * warp + 8 * (q >> 1) + r0, kc = c + 4 * (q & 1);
                pa[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (m0 + row < a.M) pa[q] = ld_stream4(a.A + (size_t)(m0 + row) * a.lda + k0 + 4 * kc);
            }
        };
        auto store_stage = [&](int s, const float4* pa) {
            unsigned char* sx_hi = smem + s * C::STAGE_BYTES + W_STAGE_BYTES;
            unsigned char* sx_lo = sx_hi + C::X_HALF;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const int row = (MP / 4) * warp + 8 * (q >> 1) + r0, kc = c + 4 * (q & 1);
                float4 hi, lo;
                split_tf32(pa[q].x, hi.x, lo.x);
                split_tf32(pa[q].y, hi.y, lo.y);
                split_tf32(pa[q].z, hi.z, lo.z);
                split_tf32(pa[q].w, hi.w, lo.w);
                const uint32_t o = core_off(row, kc);
                *reinterpret_cast<float4*>(sx_hi + o) = hi;
                *reinterpret_cast<float4*>(sx_lo + o) = lo;
            }
        };
        // Programmatic dependent launch: this grid may start while the preceding (small) kernel still runs — barrier setup,
        // TMEM allocation and the weight producer's first stages overlap it; everything that reads or overwrites data of
        // earlier kernels (activations, residual, split-K workspace) comes after this wait.
        asm volatile("griddepcontrol.wait;" ::: "memory");
        if (nst > 0) load_stage(0, va[0]);
        for (int it = 0; it < nst; it += 2) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {  // explicit two-step unroll keeps the register buffers statically indexed
                const int cur = it + h;
                if (cur < nst) {
                    if (cur + 1 < nst) load_stage(cur + 1, va[h ^ 1]);
                    const int s = cur % C::STAGES;
                    if (lane == 0) mbar_wait(&empty[s], (((uint32_t)(cur / C::STAGES)) & 1u) ^ 1u);  // one poller per warp
                    __syncwarp();
                    store_stage(s, va[h]);
                    fence_async_smem();  // generic-proxy stores -> visible to the tensor core (async proxy)
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full[s]);
                }
            }
        }
        // ================= epilogue: lane = output column, register = row =================
        if (lane == 0) mbar_wait(done, 0u);
        __syncwarp();
        tc_fence_after();
        const int n = n0 + 32 * warp + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(32 * warp) << 16);
        const float bias = (a.bias && n < a.N) ? a.bias[n] : 0.0f;
        const int mrem = a.M - m0;  // valid rows of this chunk
        // compact address arithmetic (one pointer bump per row): the unrolled epilogue is the largest piece of code in
        // this short kernel and its instruction fetch showed up as the top stall
        float* out = (gridDim.z > 1) ? ws + (size_t)blockIdx.z * ((size_t)a.M * a.N) + (size_t)m0 * a.N + n : a.C + (size_t)m0 * a.ldc + n;
        const float* res = a.residual ? a.residual + (size_t)m0 * a.ldr + n : nullptr;
        const int ostride = (gridDim.z > 1) ? a.N : a.ldc;
        const bool direct = gridDim.z == 1;
#pragma unroll
        for (int h = 0; h < MP / 32; ++h) {
            float v[32];
            tmem_ld32(taddr + (uint32_t)(32 * h), v);
            if (n < a.N) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (32 * h + i < mrem) {
                        float y = v[i];
                        if (direct) {
                            y += bias;
                            if (a.act == ACT_GELU_NEW) y = gelu_new(y);
                            if (res) y += res[0];
                        }
                        out[0] = y;
                    }
                    out += ostride;
                    if (res) res += a.ldr;
                }
            }
        }
        tc_fence_before();
    } else if (warp == 4) {
        // ================= weight producer: one bulk TMA copy per stage =================
        if (lane == 0) {
            const float* src = Wt + ((size_t)blockIdx.x * nks + kb / KT) * W_STAGE_FLOATS;
            const uint64_t policy = l2_policy_evict_first();
            for (int it = 0; it < nst; ++it) {
                const int s = it % C::STAGES;
                mbar_wait(&empty[s], (((uint32_t)(it / C::STAGES)) & 1u) ^ 1u);
                mbar_arrive_expect_tx(&full[s], (uint32_t)W_STAGE_BYTES);
                bulk_g2s_hint(smem + s * C::STAGE_BYTES, src + (size_t)it * W_STAGE_FLOATS, (uint32_t)W_STAGE_BYTES, &full[s], policy);
            }
        }
    } else if (lane == 0) {
        // ================= MMA issuer (one thread) =================
        for (int it = 0; it < nst; ++it) {
            const int s = it % C::STAGES;
            mbar_wait(&full[s], ((uint32_t)(it / C::STAGES)) & 1u);
            tc_fence_after();
            const uint32_t sw_hi = smem_u32(smem + s * C::STAGE_BYTES), sw_lo = sw_hi + BN * KT * 4;
            const uint32_t sx_hi = sw_hi + W_STAGE_BYTES, sx_lo = sx_hi + C::X_HALF;
#pragma unroll
            for (int j = 0; j < KT / 8; ++j) {  // k-step = 8 TF32 = two 16-byte chunks = 256 bytes along the row group
                const uint32_t ko = (uint32_t)j * 256u;
                const uint64_t wh = make_desc(sw_hi + ko), wl = make_desc(sw_lo + ko);
                const uint64_t xh = make_desc(sx_hi + ko), xl = make_desc(sx_lo + ko);
                mma_tf32(tmem_base, wl, xh, C::IDESC, (it > 0 || j > 0) ? 1u : 0u);
                mma_tf32(tmem_base, wh, xl, C::IDESC, 1u);
                mma_tf32(tmem_base, wh, xh, C::IDESC, 1u);
            }
            mma_commit(&empty[s]);  // the stage is free once these MMAs have read it
        }
        mma_commit(done);
    }
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(MP) : "memory");
    }
}

// =============================================================================================
// Perceiver cross-attention on tensor cores (layers/perceiver_encoder.py:108-151, Attend: softmax(q k^T / sqrt(64)) v, no
// mask): one CTA per (batch element, head); 32 latent queries against the 32 + S context keys, head dim 64.
//   pass 1  scores: per tile of 128 keys,  S^T[128 keys x 32 queries] = K_tile[128 x 64] . Q^T   (tcgen05, M = 128, N = 32,
//           3xTF32, accumulator in TMEM) -> tcgen05.ld (lane = key) -> scaled scores to shared memory, S[query][key]
//   softmax rows of S in shared memory (fp32, one warp per 8 queries)
//   pass 2  per tile of 64 keys,  O^T[64 dims (+64 zero rows) x 32 queries] += V^T_tile[128 x 64 keys] . P_tile^T
//           accumulated in TMEM over the tiles -> tcgen05.ld (lane = dim) -> out[query][dim]
// Operands are staged by the threads (fp32 -> TF32 hi | lo, UMMA canonical K-major no-swizzle core matrices), handed to the
// async proxy with fence.proxy.async; one thread issues the MMAs, tcgen05.commit -> mbarrier tells the CTA when a tile is done.
// The work is tiny (80 MFLOP per call) and latency-bound; what the kernel buys is that scores and probabilities never
// leave the SM, and the tensor pipe does the 3xTF32 contractions.
// =============================================================================================
constexpr int PA_Q = 32, PA_HD = 64, PA_THREADS = 128;

// byte offset of (row, 16-byte k chunk) in a canonical K-major tile whose rows hold `kchunks` chunks
__device__ __forceinline__ uint32_t canon_off(int row, int kc, int kchunks) {
    return (uint32_t)(((row >> 3) * kchunks + kc) * 128 + (row & 7) * 16);
}
__device__ __forceinline__ uint64_t make_desc_k(uint32_t smem_addr, int kchunks) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
    d |= (uint64_t)(128u >> 4) << 16;                        // LBO: between the two 16-byte K chunks of a k-step
    d |= (uint64_t)(((uint32_t)kchunks * 128u) >> 4) << 32;  // SBO: between 8-row groups
    d |= (uint64_t)1 << 46;
    return d;
}
constexpr uint32_t PA_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(PA_Q >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__host__ __device__ inline int pa_sp(int RC) { return (RC + 127) / 128 * 128 + 4; }  // row pitch of S (floats)
__host__ __device__ inline size_t pa_smem_bytes(int RC) {
    return 1024 /* align */ + 2 * PA_Q * PA_HD * 4 /* Q hi|lo */ + 2 * 128 * 64 * 4 /* tile hi|lo */ + 2 * PA_Q * 64 * 4 /* P tile hi|lo */ +
           (size_t)PA_Q * pa_sp(RC) * 4 + 64;
}

__global__ void __launch_bounds__(PA_THREADS, 1) pc_attention_tc_kernel(const float* __restrict__ Q, const float* __restrict__ KV,
                                                                        float* __restrict__ O, int RC, int inner, float scale) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    unsigned char* q_hi = smem;                              // [32 x 64] canonical, 16 chunks per row
    unsigned char* q_lo = q_hi + PA_Q * PA_HD * 4;
    unsigned char* t_hi = q_lo + PA_Q * PA_HD * 4;           // [128 x 64] canonical: K tile (pass 1) | V^T tile (pass 2)
    unsigned char* t_lo = t_hi + 128 * 64 * 4;
    unsigned char* p_hi = t_lo + 128 * 64 * 4;               // [32 x 64 keys] canonical: probabilities of the tile
    unsigned char* p_lo = p_hi + PA_Q * 64 * 4;
    float* S = reinterpret_cast<float*>(p_lo + PA_Q * 64 * 4);
    const int SP = pa_sp(RC);
    uint64_t* bar = reinterpret_cast<uint64_t*>(S + (size_t)PA_Q * SP);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int h = blockIdx.x, b = blockIdx.y;
    const float* Qb = Q + (size_t)b * PA_Q * inner + h * PA_HD;           // row stride inner
    const float* Kb = KV + (size_t)b * RC * 2 * inner + h * PA_HD;        // row stride 2 inner
    const float* Vb = Kb + inner;

    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(32) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // Q -> canonical hi | lo: 32 rows x 16 chunks = 512 chunks, 4 per thread
    for (int c = tid; c < PA_Q * 16; c += PA_THREADS) {
        const int row = c >> 4, kc = c & 15;
        const float4 x = *reinterpret_cast<const float4*>(Qb + (size_t)row * inner + 4 * kc);
        float4 hi, lo;
        split_tf32(x.x, hi.x, lo.x); split_tf32(x.y, hi.y, lo.y); split_tf32(x.z, hi.z, lo.z); split_tf32(x.w, hi.w, lo.w);
        const uint32_t o = canon_off(row, kc, 16);
        *reinterpret_cast<float4*>(q_hi + o) = hi;
        *reinterpret_cast<float4*>(q_lo + o) = lo;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    uint32_t phase = 0;

    // ---------------- pass 1: scores ----------------
    const int n1 = (RC + 127) / 128;
    for (int t = 0; t < n1; ++t) {
        const int key = t * 128 + tid;  // this thread stages row `tid` of the tile
        {
            const float* src = Kb + (size_t)key * 2 * inner;
#pragma unroll
            for (int kc = 0; kc < 16; ++kc) {
                float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                if (key < RC) x = *reinterpret_cast<const float4*>(src + 4 * kc);
                float4 hi, lo;
                split_tf32(x.x, hi.x, lo.x); split_tf32(x.y, hi.y, lo.y); split_tf32(x.z, hi.z, lo.z); split_tf32(x.w, hi.w, lo.w);
                const uint32_t o = canon_off(tid, kc, 16);
                *reinterpret_cast<float4*>(t_hi + o) = hi;
                *reinterpret_cast<float4*>(t_lo + o) = lo;
            }
        }
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int j = 0; j < PA_HD / 8; ++j) {
                const uint32_t ko = (uint32_t)j * 256u;
                const uint64_t ah = make_desc_k(smem_u32(t_hi) + ko, 16), al = make_desc_k(smem_u32(t_lo) + ko, 16);
                const uint64_t bh = make_desc_k(smem_u32(q_hi) + ko, 16), bl = make_desc_k(smem_u32(q_lo) + ko, 16);
                mma_tf32(tmem, al, bh, PA_IDESC, j > 0 ? 1u : 0u);
                mma_tf32(tmem, ah, bl, PA_IDESC, 1u);
                mma_tf32(tmem, ah, bh, PA_IDESC, 1u);
            }
            mma_commit(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
        tc_fence_after();
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16), v);  // lane = key of the tile, v[q] = its score with query q
        if (key < RC) {
#pragma unroll
            for (int q = 0; q < PA_Q; ++q) S[(size_t)q * SP + key] = v[q] * scale;
        }
        tc_fence_before();
        __syncthreads();  // the tile buffer and the accumulator are reused
    }
    // ---------------- softmax over the keys (rows of S), probabilities in place; zero the padding ----------------
    const int RCP = (RC + 63) / 64 * 64;
    for (int q = warp * 8; q < warp * 8 + 8; ++q) {
        float* row = S + (size_t)q * SP;
        float m = -INFINITY;
        for (int j = lane; j < RC; j += 32) m = fmaxf(m, row[j]);
        m = warp_max(m);
        float sum = 0.0f;
        for (int j = lane; j < RC; j += 32) {
            const float e = expf(row[j] - m);
            row[j] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        for (int j = lane; j < RCP; j += 32) row[j] = j < RC ? row[j] * inv : 0.0f;
    }
    // rows 64 .. 127 of the A tile are zero in pass 2 (64 real dims)
    for (int c = tid; c < 64 * 16; c += PA_THREADS) {
        const uint32_t o = canon_off(64 + (c >> 4), c & 15, 16);
        *reinterpret_cast<float4*>(t_hi + o) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(t_lo + o) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    // ---------------- pass 2: O^T = V^T . P^T over tiles of 64 keys ----------------
    const int n2 = RCP / 64;
    for (int t = 0; t < n2; ++t) {
        {   // V^T tile: thread = (key tid % 64, dims [32 (tid / 64), +32)): element (dim d, key k) at chunk k / 4, lane k % 4
            const int k = tid & 63, d0 = (tid >> 6) * 32;
            const int key = t * 64 + k;
            const float* src = Vb + (size_t)key * 2 * inner + d0;
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
                float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                if (key < RC) x = *reinterpret_cast<const float4*>(src + 4 * c4);
                const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float hi, lo;
                    split_tf32(xs[e], hi, lo);
                    const uint32_t o = canon_off(d0 + 4 * c4 + e, k >> 2, 16) + (uint32_t)(k & 3) * 4u;
                    *reinterpret_cast<float*>(t_hi + o) = hi;
                    *reinterpret_cast<float*>(t_lo + o) = lo;
                }
            }
        }
        // P tile: 32 queries x 16 chunks of 4 keys
        for (int c = tid; c < PA_Q * 16; c += PA_THREADS) {
            const int q = c >> 4, kc = c & 15;
            const float4 x = *reinterpret_cast<const float4*>(S + (size_t)q * SP + t * 64 + 4 * kc);
            float4 hi, lo;
            split_tf32(x.x, hi.x, lo.x); split_tf32(x.y, hi.y, lo.y); split_tf32(x.z, hi.z, lo.z); split_tf32(x.w, hi.w, lo.w);
            const uint32_t o = canon_off(q, kc, 16);
            *reinterpret_cast<float4*>(p_hi + o) = hi;
            *reinterpret_cast<float4*>(p_lo + o) = lo;
        }
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int j = 0; j < 64 / 8; ++j) {
                const uint32_t ko = (uint32_t)j * 256u;
                const uint64_t ah = make_desc_k(smem_u32(t_hi) + ko, 16), al = make_desc_k(smem_u32(t_lo) + ko, 16);
                const uint64_t bh = make_desc_k(smem_u32(p_hi) + ko, 16), bl = make_desc_k(smem_u32(p_lo) + ko, 16);
                mma_tf32(tmem, al, bh, PA_IDESC, (t > 0 || j > 0) ? 1u : 0u);
                mma_tf32(tmem, ah, bl, PA_IDESC, 1u);
                mma_tf32(tmem, ah, bh, PA_IDESC, 1u);
            }
            mma_commit(bar);
        }
        mbar_wait(bar, phase);  // the operands may be overwritten (and, after the last tile, the accumulator read)
        phase ^= 1u;
        tc_fence_after();
        __syncthreads();
    }
    if (warp < 2) {  // lanes 0 .. 63 of the accumulator = the 64 output dims
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16), v);
        float* dst = O + (size_t)b * PA_Q * inner + h * PA_HD + 32 * warp + lane;
#pragma unroll
        for (int q = 0; q < PA_Q; ++q) dst[(size_t)q * inner] = v[q];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32) : "memory");
    }
}

// =============================================================================================
// Persistent fused prefill (batch 1, <= 128 rows): the 30 pre-LN blocks of layers/gpt_inference.py:81-112 in ONE
// cooperative launch instead of nine launches per block.  Same arithmetic as the per-op path, phase by phase:
//   QKV GEMM (split-K) | reduce + bias + K/V append | causal attention | attn c_proj GEMM | reduce + residual + ln_2 |
//   c_fc GEMM | reduce + bias + gelu_new | mlp c_proj GEMM | reduce + residual + ln_1 of the next block
// separated by grid barriers (one arrival counter; workers only).  What persistence buys: the TMA producer warp walks
// the weight stream of ALL phases and runs ahead of the barriers / reductions (the ring is always full of the next
// GEMM's first stages), TMEM / mbarriers are set up once, no launch gaps.
//   warps 0-7  workers: activation loaders + TMEM epilogue (0-3) during a GEMM phase, all eight in the reduce / attention phases
//   warp  8    TMA producer (one thread), free-running over the whole kernel
//   warp  9    MMA issuer (one thread): follows the full barriers; waits for the epilogue of job k before job k+1 reuses TMEM
// A GEMM phase is one wave: job = (128-column tile, K range) for CTA < tiles x splits, partition as launch_gemm_tc.
// Everything another CTA wrote is read through L2 (ld.global.cg): L1 is not coherent across a grid barrier.
// =============================================================================================
struct PrefillParams {
    int L, D, H, M, S_max;
    const float* blob;
    long long layer_stride;  // floats between the same tensor of consecutive layers
    long long ln1_w, ln1_b, attn_b, proj_b, ln2_w, ln2_b, fc_b, proj2_b;  // offsets of layer 0 in the blob
    const float* const* tcw;  // [L][4] packed tensor-core copies: c_attn, attn c_proj, c_fc, mlp c_proj
    float *X, *A, *QKV, *U, *ws;
    float* kv;                // K plane of layer l at kv + 2 l stride, V plane at kv + (2 l + 1) stride (batch row 0)
    long long kv_layer_stride;
    unsigned* gbar;           // grid-barrier arrival counter, zero at launch
    unsigned long long* prof; // debug: [grid][16] cycles per phase kind accumulated by thread 0 (null = off)
    int dbg;                  // debug (results invalid): 1 = no MMAs, 2 = no activation loads, 4 = no weight copies
};

constexpr int PF_WORKERS = 256, PF_THREADS = 320;

struct PfJob {
    int has, n0, kb, nst;
};
// phase 0..3 = c_attn, attn c_proj, c_fc, mlp c_proj;  partition as launch_gemm_tc (one row chunk)
__device__ __forceinline__ void pf_shape(int ph, int D, int& N, int& K) {
    N = ph == 0 ? 3 * D : (ph == 2 ? 4 * D : D);
    K = ph == 3 ? 4 * D : D;
}
__device__ __forceinline__ void pf_split(int N, int K, int& tiles, int& splits, int& k_chunk) {
    tiles = (N + BN - 1) / BN;
    splits = 1;
    if (tiles < 120) {
        splits = 148 / tiles;
        splits = min(splits, K / 64);
        if (splits < 1) splits = 1;
    }
    k_chunk = K;
    if (splits > 1) {
        k_chunk = ((K + splits - 1) / splits + KT - 1) / KT * KT;
        splits = (K + k_chunk - 1) / k_chunk;
    }
}
__device__ __forceinline__ PfJob pf_job(int ph, int D, int cta) {
    int N, K, tiles, splits, k_chunk;
    pf_shape(ph, D, N, K);
    pf_split(N, K, tiles, splits, k_chunk);
    PfJob j;
    j.has = cta < tiles * splits;
    const int tile = cta % tiles, z = cta / tiles;
    j.n0 = tile * BN;
    j.kb = z * k_chunk;
    j.nst = j.has ? (min(K, j.kb + k_chunk) - j.kb) / KT : 0;
    return j;
}

__device__ __forceinline__ void pf_grid_barrier(unsigned* cnt, unsigned& epoch, int G, int tid) {
    bar_sync(1, PF_WORKERS);
    epoch += 1u;
    if (tid == 0) {
        __threadfence();
        atomicAdd(cnt, 1u);
        uint32_t spins = 0;
        while (ld_acquire_gpu(cnt) < epoch * (unsigned)G) {
            if (++spins > (1u << 26)) __trap();
        }
        __threadfence();
    }
    bar_sync(1, PF_WORKERS);
}

// one (query row, head) per group of four warps: warp w of the group takes keys w, w + 4, ... in trips of four keys (K and V
// rows of a trip are requested together), online softmax per warp, the four partial states merged through shared memory.
// q, K, V rows straight from the QKV rows in L2 (coalesced: a row is hd contiguous floats); arithmetic of HF
// GPT2Attention._attn (scale 1 / sqrt(hd), causal).  sm: [4 warps][hd + 2] floats of the group.
template <int HD>
__device__ __forceinline__ void pf_attention_item(const float* __restrict__ QKV, float* __restrict__ O, int D, int i, int h, int lane,
                                                  int gw, float* sm, int bar_id) {
    constexpr int DPL = HD / 32;
    const float scale = 1.0f / sqrtf((float)HD);
    const size_t rs = 3 * (size_t)D;
    float q[DPL], o[DPL];
    auto ldrow = [&](const float* base, float* dst) {
        if constexpr (DPL >= 4) {
#pragma unroll
            for (int c = 0; c < DPL / 4; ++c) {
                const float4 t = ldcg4(base + lane * DPL + 4 * c);
                dst[4 * c] = t.x; dst[4 * c + 1] = t.y; dst[4 * c + 2] = t.z; dst[4 * c + 3] = t.w;
            }
        } else if constexpr (DPL == 2) {
            const float2 t = ldcg2(base + lane * 2);
            dst[0] = t.x; dst[1] = t.y;
        } else {
            dst[0] = ldcg(base + lane);
        }
    };
    ldrow(QKV + (size_t)i * rs + h * HD, q);
#pragma unroll
    for (int d = 0; d < DPL; ++d) o[d] = 0.0f;
    float m = -INFINITY, l = 0.0f;
    constexpr int KPT = DPL >= 8 ? 2 : 4;   // keys per trip (register budget: K and V rows of a trip are both in flight)
    for (int j0 = gw; j0 <= i; j0 += 4 * KPT) {  // this warp's keys: j0, j0 + 4, ...
        float k[KPT][DPL], v[KPT][DPL], sc[KPT];
#pragma unroll
        for (int u = 0; u < KPT; ++u) {
            if (j0 + 4 * u <= i) {
                ldrow(QKV + (size_t)(j0 + 4 * u) * rs + D + h * HD, k[u]);
                ldrow(QKV + (size_t)(j0 + 4 * u) * rs + 2 * D + h * HD, v[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < KPT; ++u) {
            sc[u] = 0.0f;
            if (j0 + 4 * u <= i) {
#pragma unroll
                for (int d = 0; d < DPL; ++d) sc[u] = fmaf(q[d], k[u][d], sc[u]);
            }
        }
#pragma unroll
        for (int x = 16; x > 0; x >>= 1)
#pragma unroll
            for (int u = 0; u < KPT; ++u) sc[u] += __shfl_xor_sync(0xffffffffu, sc[u], x);
        float mb = -INFINITY;
#pragma unroll
        for (int u = 0; u < KPT; ++u) {
            sc[u] = (j0 + 4 * u <= i) ? sc[u] * scale : -INFINITY;
            mb = fmaxf(mb, sc[u]);
        }
        const float mn = fmaxf(m, mb);
        const float c = expf(m - mn);  // first trip: exp(-inf) = 0
        l *= c;
#pragma unroll
        for (int d = 0; d < DPL; ++d) o[d] *= c;
#pragma unroll
        for (int u = 0; u < KPT; ++u) {
            if (j0 + 4 * u <= i) {
                const float pj = expf(sc[u] - mn);
                l += pj;
#pragma unroll
                for (int d = 0; d < DPL; ++d) o[d] = fmaf(pj, v[u][d], o[d]);
            }
        }
        m = mn;
    }
    // merge the four warp states (a warp without keys has m = -inf, l = 0)
    float* mine = sm + gw * (HD + 2);
#pragma unroll
    for (int d = 0; d < DPL; ++d) mine[lane * DPL + d] = o[d];
    if (lane == 0) {
        mine[HD] = m;
        mine[HD + 1] = l;
    }
    if (bar_id == 2) bar_sync(2, 128); else bar_sync(3, 128);
    if (gw == 0) {
        float M4 = -INFINITY;
#pragma unroll
        for (int w = 0; w < 4; ++w) M4 = fmaxf(M4, sm[w * (HD + 2) + HD]);
        float L4 = 0.0f, acc[DPL];
#pragma unroll
        for (int d = 0; d < DPL; ++d) acc[d] = 0.0f;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const float mw = sm[w * (HD + 2) + HD];
            const float cw = (mw == -INFINITY) ? 0.0f : expf(mw - M4);
            L4 = fmaf(sm[w * (HD + 2) + HD + 1], cw, L4);
#pragma unroll
            for (int d = 0; d < DPL; ++d) acc[d] = fmaf(sm[w * (HD + 2) + lane * DPL + d], cw, acc[d]);
        }
        const float inv = 1.0f / L4;
        float* dst = O + (size_t)i * D + h * HD + lane * DPL;
#pragma unroll
        for (int d = 0; d < DPL; ++d) dst[d] = acc[d] * inv;
    }
    if (bar_id == 2) bar_sync(2, 128); else bar_sync(3, 128);  // sm is reused by the group's next item
}

template <int MP>
__global__ void __launch_bounds__(PF_THREADS, 1) prefill_fused_kernel(PrefillParams p) {
    using C = Cfg<MP>;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
    uint64_t* empty = full + C::STAGES;
    uint64_t* done = empty + C::STAGES;   // MMAs of a job complete -> epilogue
    uint64_t* tfree = done + 1;           // epilogue has read TMEM -> the next job may overwrite it
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfree + 1);
    float* red = reinterpret_cast<float*>(tmem_slot + 2);  // [2][8] LayerNorm partials of the row-wise reductions
    float* att_sm = red + 16;                                // [2 groups][4 warps][256 + 2] attention merge scratch

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cta = blockIdx.x, G = gridDim.x;
    const int D = p.D, M = p.M, H = p.H, HD = D / H;

    if (tid == 0) {
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(&full[s], 4 + 1);  // one arrival per loader warp + the producer's expect_tx arrival
            mbar_init(&empty[s], 1);
        }
        mbar_init(done, 1);
        mbar_init(tfree, 4);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(MP)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 8) {
        // ================= weight producer: walks the stream of every GEMM phase of every layer =================
        if (lane == 0) {
            const uint64_t policy = l2_policy_evict_first();
            uint32_t gi = 0;
            for (int l = 0; l < p.L; ++l) {
                for (int ph = 0; ph < 4; ++ph) {
                    const PfJob j = pf_job(ph, D, cta);
                    if (!j.has) continue;
                    int N, K;
                    pf_shape(ph, D, N, K);
                    const float* src = p.tcw[l * 4 + ph] + ((size_t)(j.n0 / BN) * (K / KT) + j.kb / KT) * W_STAGE_FLOATS;
                    for (int it = 0; it < j.nst; ++it, ++gi) {
                        const int s = gi % C::STAGES;
                        mbar_wait(&empty[s], ((gi / C::STAGES) & 1u) ^ 1u);
                        if (p.dbg & 4) {
                            mbar_arrive(&full[s]);
                        } else {
                            mbar_arrive_expect_tx(&full[s], (uint32_t)W_STAGE_BYTES);
                            bulk_g2s_hint(smem + s * C::STAGE_BYTES, src + (size_t)it * W_STAGE_FLOATS, (uint32_t)W_STAGE_BYTES, &full[s], policy);
                        }
                    }
                }
            }
        }
    } else if (warp == 9) {
        // ================= MMA issuer (one thread) =================
        if (lane == 0) {
            uint32_t gi = 0, nj = 0;
            for (int l = 0; l < p.L; ++l) {
                for (int ph = 0; ph < 4; ++ph) {
                    const PfJob j = pf_job(ph, D, cta);
                    if (!j.has) continue;
                    if (nj > 0) mbar_wait(tfree, (nj - 1u) & 1u);  // the previous job's accumulator has been read
                    tc_fence_after();
                    for (int it = 0; it < j.nst; ++it, ++gi) {
                        const int s = gi % C::STAGES;
                        mbar_wait(&full[s], (gi / C::STAGES) & 1u);
                        tc_fence_after();
                        const uint32_t sw_hi = smem_u32(smem + s * C::STAGE_BYTES), sw_lo = sw_hi + BN * KT * 4;
                        const uint32_t sx_hi = sw_hi + W_STAGE_BYTES, sx_lo = sx_hi + C::X_HALF;
#pragma unroll
                        for (int q = 0; q < KT / 8; ++q) {
                            if (p.dbg & 1) break;
                            const uint32_t ko = (uint32_t)q * 256u;
                            const uint64_t wh = make_desc(sw_hi + ko), wl = make_desc(sw_lo + ko);
                            const uint64_t xh = make_desc(sx_hi + ko), xl = make_desc(sx_lo + ko);
                            mma_tf32(tmem_base, wl, xh, C::IDESC, (it > 0 || q > 0) ? 1u : 0u);
                            mma_tf32(tmem_base, wh, xl, C::IDESC, 1u);
                            mma_tf32(tmem_base, wh, xh, C::IDESC, 1u);
                        }
                        mma_commit(&empty[s]);
                    }
                    mma_commit(done);
                    ++nj;
                }
            }
        }
    } else {
        // ================= workers =================
        unsigned epoch = 0;
        uint32_t gi = 0, nj = 0;
        const float* blob = p.blob;
        // debug phase profile: [0..3] GEMM phase of c_attn / proj / fc / proj2 (loaders + epilogue), [4] barrier after a GEMM,
        // [5] flat reduce, [6] row-wise reduce, [7] barrier after a reduce, [8] attention, [9] barrier after attention
        unsigned long long prof[16];
        for (int k = 0; k < 16; ++k) prof[k] = 0ull;
        long long pc = clock64();
        auto mark = [&](int k) {
            if (p.prof != nullptr && tid == 0) {
                const long long c = clock64();
                prof[k] += (unsigned long long)(c - pc);
                pc = c;
            }
        };
        for (int l = 0; l < p.L; ++l) {
            const long long lo = (long long)l * p.layer_stride;
            float* kc = p.kv + ((size_t)l * 2 + 0) * p.kv_layer_stride;
            float* vc = p.kv + ((size_t)l * 2 + 1) * p.kv_layer_stride;
            for (int ph = 0; ph < 4; ++ph) {
                int N, K, tiles, splits, k_chunk;
                pf_shape(ph, D, N, K);
                pf_split(N, K, tiles, splits, k_chunk);
                const PfJob j = pf_job(ph, D, cta);
                const float* Ain = ph == 3 ? p.U : p.A;
                const int lda = ph == 3 ? 4 * D : D;
                // ---------------- GEMM phase: activation loaders + TMEM epilogue (warps 0-3) ----------------
                if (j.has && warp < 4) {
                    constexpr int NQ = MP / 16;
                    constexpr int PD = MP == 64 ? 3 : 2, NBUF = PD + 1;  // prefetch distance in stages
                    const int r0 = lane & 7, c = lane >> 3;
                    float4 va[NBUF][NQ];
                    auto load_stage = [&](int it, float4* pa) {
                        const int k0 = j.kb + it * KT;
#pragma unroll
                        for (int q = 0; q < NQ; ++q) {
                            const int row = (MP / 4) * warp + 8 * (q >> 1) + r0, kc4 = c + 4 * (q & 1);
                            pa[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (row < M && !(p.dbg & 2)) pa[q] = ldcg4(Ain + (size_t)row * lda + k0 + 4 * kc4);
                        }
                    };
                    auto store_stage = [&](int s, const float4* pa) {
                        unsigned char* sx_hi = smem + s * C::STAGE_BYTES + W_STAGE_BYTES;
                        unsigned char* sx_lo = sx_hi + C::X_HALF;
#pragma unroll
                        for (int q = 0; q < NQ; ++q) {
                            const int row = (MP / 4) * warp + 8 * (q >> 1) + r0, kc4 = c + 4 * (q & 1);
                            float4 hi, lo4;
                            split_tf32(pa[q].x, hi.x, lo4.x);
                            split_tf32(pa[q].y, hi.y, lo4.y);
                            split_tf32(pa[q].z, hi.z, lo4.z);
                            split_tf32(pa[q].w, hi.w, lo4.w);
                            const uint32_t o = core_off(row, kc4);
                            *reinterpret_cast<float4*>(sx_hi + o) = hi;
                            *reinterpret_cast<float4*>(sx_lo + o) = lo4;
                        }
                    };
                    // register ring of NBUF stages: the loads of stages cur + 1 .. cur + PD are in flight while stage cur is split
                    // and stored (with one stage of lookahead every stage cost a whole L2 round trip: 1.7 us per 32 KB stage)
                    long long sc0 = clock64();
#pragma unroll
                    for (int b = 0; b < PD; ++b)
                        if (b < j.nst) load_stage(b, va[b]);
                    for (int it = 0; it < j.nst; it += NBUF) {
#pragma unroll
                        for (int hh = 0; hh < NBUF; ++hh) {
                            const int cur = it + hh;
                            if (cur < j.nst) {
                                if (cur + PD < j.nst) load_stage(cur + PD, va[(hh + PD) % NBUF]);
                                const uint32_t g2 = gi + (uint32_t)cur;
                                const int s = g2 % C::STAGES;
                                if (lane == 0) mbar_wait(&empty[s], ((g2 / C::STAGES) & 1u) ^ 1u);
                                __syncwarp();
                                if (!(p.dbg & 16)) store_stage(s, va[hh]);
                                if (!(p.dbg & 8)) fence_async_smem();
                                __syncwarp();
                                if (lane == 0) mbar_arrive(&full[s]);
                            }
                        }
                    }
                    long long sc1 = clock64();
                    // epilogue: raw split-K partials, lane = output column, register = row
                    if (lane == 0) mbar_wait(done, nj & 1u);
                    __syncwarp();
                    long long sc2 = clock64();
                    tc_fence_after();
                    const int n = j.n0 + 32 * warp + lane;
                    const uint32_t taddr = tmem_base + ((uint32_t)(32 * warp) << 16);
                    float* out = p.ws + (size_t)(j.kb / k_chunk) * ((size_t)M * N) + n;
#pragma unroll
                    for (int hh = 0; hh < MP / 32; ++hh) {
                        float v[32];
                        tmem_ld32(taddr + (uint32_t)(32 * hh), v);
                        if (n < N) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) {
                                if (32 * hh + i < M) out[0] = v[i];
                                out += N;
                            }
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tfree);
                    if (p.prof != nullptr && tid == 0) {
                        const long long sc3 = clock64();
                        prof[11] += (unsigned long long)(sc1 - sc0);
                        prof[12] += (unsigned long long)(sc2 - sc1);
                        prof[13] += (unsigned long long)(sc3 - sc2);
                    }
                }
                if (j.has) {
                    gi += (uint32_t)j.nst;
                    nj += 1u;
                }
                mark(ph);
                pf_grid_barrier(p.gbar, epoch, G, tid);
                mark(4);
                // ---------------- reduce phase ----------------
                const size_t total = (size_t)M * N;
                if (ph == 0 || ph == 2) {
                    // flat: sum of the partials in split order + bias (+ gelu_new) -> QKV (+ K/V cache append) | U
                    const float* bias = blob + (ph == 0 ? p.attn_b : p.fc_b) + lo;
                    float* Cout = ph == 0 ? p.QKV : p.U;
                    // (four consecutive outputs per thread; the loads of all splits are issued before the first add: a loop
                    // of load-add pairs serialises on L2 latency, ~0.7 us per split)
                    for (size_t i4 = (size_t)cta * PF_WORKERS + tid; i4 < total / 4; i4 += (size_t)G * PF_WORKERS) {
                        const size_t i = 4 * i4;
                        const int mrow = (int)(i / N), n = (int)(i % N);
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        for (int z0 = 0; z0 < splits; z0 += 8) {  // eight loads in flight, summed in split order
                            float4 t[8];
#pragma unroll
                            for (int z = 0; z < 8; ++z)
                                if (z0 + z < splits) t[z] = ldcg4(p.ws + (size_t)(z0 + z) * total + i);
#pragma unroll
                            for (int z = 0; z < 8; ++z) {
                                if (z0 + z < splits) {
                                    v.x += t[z].x; v.y += t[z].y; v.z += t[z].z; v.w += t[z].w;
                                }
                            }
                        }
                        const float4 b = __ldg(reinterpret_cast<const float4*>(bias + n));
                        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
                        if (ph == 2) {
                            v.x = gelu_new(v.x); v.y = gelu_new(v.y); v.z = gelu_new(v.z); v.w = gelu_new(v.w);
                        }
                        *reinterpret_cast<float4*>(Cout + i) = v;
                        if (ph == 0 && n >= D) {  // K/V cache append (4 consecutive columns stay inside one head: hd % 4 == 0)
                            const int cc = (n - D) % D;
                            float* dst = (n < 2 * D) ? kc : vc;
                            *reinterpret_cast<float4*>(dst + ((size_t)(cc / HD) * p.S_max + mrow) * HD + (cc % HD)) = v;
                        }
                    }
                } else {
                    // row-wise: sum + bias + residual -> X (in place), then the LayerNorm that consumes it -> A
                    const float* bias = blob + (ph == 1 ? p.proj_b : p.proj2_b) + lo;
                    const bool with_ln = ph == 1 || l + 1 < p.L;
                    const float* lw = blob + (ph == 1 ? p.ln2_w + lo : p.ln1_w + lo + p.layer_stride);
                    const float* lb = blob + (ph == 1 ? p.ln2_b + lo : p.ln1_b + lo + p.layer_stride);
                    for (int row = cta; row < M; row += G) {
                        const int n = tid * 4;
                        const bool valid = n < N;
                        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (valid) {
                            const float* src = p.ws + (size_t)row * N + n;
                            for (int z0 = 0; z0 < splits; z0 += 8) {  // eight loads in flight, summed in split order
                                float4 t[8];
#pragma unroll
                                for (int z = 0; z < 8; ++z)
                                    if (z0 + z < splits) t[z] = ldcg4(src + (size_t)(z0 + z) * total);
#pragma unroll
                                for (int z = 0; z < 8; ++z) {
                                    if (z0 + z < splits) {
                                        acc.x += t[z].x; acc.y += t[z].y; acc.z += t[z].z; acc.w += t[z].w;
                                    }
                                }
                            }
                            const float4 b = __ldg(reinterpret_cast<const float4*>(bias + n));
                            acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
                            const float4 r = ldcg4(p.X + (size_t)row * D + n);
                            acc.x += r.x; acc.y += r.y; acc.z += r.z; acc.w += r.w;
                            *reinterpret_cast<float4*>(p.X + (size_t)row * D + n) = acc;
                        }
                        float sum = valid ? ((acc.x + acc.y) + (acc.z + acc.w)) : 0.0f;
                        sum = warp_sum(sum);
                        if (lane == 0) red[warp] = sum;
                        bar_sync(1, PF_WORKERS);
                        float tot = 0.0f;
#pragma unroll
                        for (int w = 0; w < 8; ++w) tot += red[w];
                        const float mean = tot / (float)N;
                        const float d0 = acc.x - mean, d1 = acc.y - mean, d2 = acc.z - mean, d3 = acc.w - mean;
                        float q = valid ? (fmaf(d0, d0, d1 * d1) + fmaf(d2, d2, d3 * d3)) : 0.0f;
                        q = warp_sum(q);
                        if (lane == 0) red[8 + warp] = q;
                        bar_sync(1, PF_WORKERS);
                        float qt = 0.0f;
#pragma unroll
                        for (int w = 0; w < 8; ++w) qt += red[8 + w];
                        const float rstd = 1.0f / sqrtf(qt / (float)N + 1e-5f);
                        if (valid && with_ln) {
                            const float4 ww = __ldg(reinterpret_cast<const float4*>(lw + n));
                            const float4 bb = __ldg(reinterpret_cast<const float4*>(lb + n));
                            *reinterpret_cast<float4*>(p.A + (size_t)row * D + n) =
                                make_float4(d0 * rstd * ww.x + bb.x, d1 * rstd * ww.y + bb.y, d2 * rstd * ww.z + bb.z, d3 * rstd * ww.w + bb.w);
                        }
                        bar_sync(1, PF_WORKERS);  // red[] is rewritten by the next row
                    }
                }
                mark((ph == 0 || ph == 2) ? 5 : 6);
                pf_grid_barrier(p.gbar, epoch, G, tid);
                mark(7);
                // ---------------- causal attention over the rows themselves (after the QKV reduce) ----------------
                if (ph == 0) {
                    // two groups of four warps per CTA; items dealt round-robin over the 2 G groups
                    const int grp = warp >> 2, gw = warp & 3;
                    float* asm_ = att_sm + grp * 4 * (256 + 2);
                    for (int item = grp * G + cta; item < M * H; item += 2 * G) {
                        const int i = item / H, h = item % H;
                        switch (HD) {
                            case 32: pf_attention_item<32>(p.QKV, p.A, D, i, h, lane, gw, asm_, 2 + grp); break;
                            case 64: pf_attention_item<64>(p.QKV, p.A, D, i, h, lane, gw, asm_, 2 + grp); break;
                            case 128: pf_attention_item<128>(p.QKV, p.A, D, i, h, lane, gw, asm_, 2 + grp); break;
                            case 256: pf_attention_item<256>(p.QKV, p.A, D, i, h, lane, gw, asm_, 2 + grp); break;
                            default: break;
                        }
                    }
                    mark(8);
                    pf_grid_barrier(p.gbar, epoch, G, tid);
                    mark(9);
                }
            }
        }
        if (p.prof != nullptr && tid == 0)
            for (int k = 0; k < 16; ++k) p.prof[(size_t)cta * 16 + k] = prof[k];
    }
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(MP) : "memory");
    }
}

}  // namespace tc

// ---------------------------------------------------------------------------------------------
// host side: registry of pre-tiled weights (keyed by the reference-layout device pointer), launcher
// ---------------------------------------------------------------------------------------------
}  // namespace gv

#include <algorithm>
#include <mutex>
#include <unordered_map>

namespace gv {

namespace {
std::mutex g_reg_mu;
std::unordered_map<const float*, const float*> g_reg;
}  // namespace

size_t gemm_tc_packed_floats(int N, int K) {
    if (K % tc::KT) return 0;
    return (size_t)((N + tc::BN - 1) / tc::BN) * (size_t)(K / tc::KT) * tc::W_STAGE_FLOATS;
}

cudaError_t gemm_tc_pack(const float* W, int N, int K, int ldw, int w_nk, float* out, cudaStream_t st) {
    if (K % tc::KT) return cudaErrorInvalidValue;
    dim3 grid((N + tc::BN - 1) / tc::BN, K / tc::KT);
    tc::pack_tc_kernel<<<grid, 256, 0, st>>>(W, N, K, ldw, w_nk, out);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) {
        std::lock_guard<std::mutex> lk(g_reg_mu);
        g_reg[W] = out;
    }
    return e;
}

void gemm_tc_forget(const float* lo, const float* hi) {
    std::lock_guard<std::mutex> lk(g_reg_mu);
    for (auto it = g_reg.begin(); it != g_reg.end();) {
        if (it->first >= lo && it->first < hi) it = g_reg.erase(it);
        else ++it;
    }
}

static const float* lookup(const float* W) {
    std::lock_guard<std::mutex> lk(g_reg_mu);
    auto it = g_reg.find(W);
    return it == g_reg.end() ? nullptr : it->second;
}

bool gemm_tc_eligible(const GemmArgs& a) {
    if (a.M < 16 || a.K < 64 || (a.K % tc::KT) != 0) return false;
    if (a.lda & 3) return false;
    if (reinterpret_cast<uintptr_t>(a.A) & 15) return false;
    return lookup(a.W) != nullptr;
}

// splits: enough K ranges to put >= ~148 CTAs on the machine (the GEMMs are weight-streaming bound)
cudaError_t launch_gemm_tc(const GemmArgs& a, float* ws, size_t ws_floats, const int* skip, cudaStream_t st, int* splits_out) {
    const float* Wt = lookup(a.W);
    if (!Wt) return cudaErrorInvalidValue;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tc::gemm_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Cfg<64>::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(tc::gemm_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Cfg<128>::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int MP = a.M <= 64 ? 64 : 128;
    dim3 grid((a.N + tc::BN - 1) / tc::BN, (a.M + MP - 1) / MP, 1);
    const int tiles = (int)(grid.x * grid.y);
    int splits = 1;
    if (tiles < 120 && ws) {
        splits = 148 / tiles;  // one wave: 1 CTA per SM (shared memory), never more CTAs than SMs
        splits = min(splits, a.K / 64);
        static const int min_chunk = [] { const char* e = getenv("GENVC_TC_MIN_KCHUNK"); return e ? atoi(e) : 0; }();  // tuning knob
        if (min_chunk > 0) splits = min(splits, max(1, a.K / min_chunk));
        while (splits > 1 && (size_t)splits * a.M * a.N > ws_floats) --splits;
        if (splits < 1) splits = 1;
    }
    int k_chunk = a.K;
    if (splits > 1) {
        k_chunk = ((a.K + splits - 1) / splits + tc::KT - 1) / tc::KT * tc::KT;
        splits = (a.K + k_chunk - 1) / k_chunk;
    }
    grid.z = splits;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(tc::THREADS);
    cfg.dynamicSmemBytes = MP == 64 ? tc::Cfg<64>::SMEM_BYTES : tc::Cfg<128>::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    *splits_out = splits;
    if (MP == 64) return cudaLaunchKernelEx(&cfg, tc::gemm_tc_kernel<64>, a, Wt, ws, k_chunk, skip);
    return cudaLaunchKernelEx(&cfg, tc::gemm_tc_kernel<128>, a, Wt, ws, k_chunk, skip);
}

// ---------------------------------------------------------------------------------------------
// persistent fused prefill: host launcher
// ---------------------------------------------------------------------------------------------
bool pc_attention_tc_supported(int n_latents, int hd, int RC) {
    return n_latents == tc::PA_Q && hd == tc::PA_HD && RC >= 1 && tc::pa_smem_bytes(RC) <= 227 * 1024;
}
// Q [B, 32, inner], KV [B, RC, 2 inner] (k | v), O [B, 32, inner]; heads = inner / 64
cudaError_t launch_pc_attention_tc(const float* Q, const float* KV, float* O, int B, int heads, int RC, cudaStream_t st) {
    const size_t smem = tc::pa_smem_bytes(RC);
    cudaError_t e = cudaFuncSetAttribute(tc::pc_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    tc::pc_attention_tc_kernel<<<dim3(heads, B), tc::PA_THREADS, smem, st>>>(Q, KV, O, RC, heads * tc::PA_HD, 1.0f / sqrtf((float)tc::PA_HD));
    return cudaGetLastError();
}

bool prefill_fused_supported(int D, int H, int M, int grid, size_t ws_floats) {
    if (D % 128 || D > 1024 || M < 16 || M > 128 || grid != 148) return false;  // (the split partition assumes 148 SMs)
    const int hd = D / H;
    if (!(hd == 32 || hd == 64 || hd == 128 || hd == 256)) return false;
    // every GEMM phase must be one wave of split-K jobs whose partials fit the workspace
    for (int ph = 0; ph < 4; ++ph) {
        const int N = ph == 0 ? 3 * D : (ph == 2 ? 4 * D : D), K = ph == 3 ? 4 * D : D;
        const int tiles = (N + tc::BN - 1) / tc::BN;
        if (tiles >= 120) return false;
        int splits = std::min(148 / tiles, K / 64);
        if (splits < 2) return false;
        const int k_chunk = ((K + splits - 1) / splits + tc::KT - 1) / tc::KT * tc::KT;
        splits = (K + k_chunk - 1) / k_chunk;
        if (tiles * splits > grid || (size_t)splits * M * N > ws_floats) return false;
    }
    return true;
}

cudaError_t launch_prefill_fused(const PrefillFusedArgs& a, int grid, cudaStream_t st) {
    tc::PrefillParams p;
    p.L = a.L; p.D = a.D; p.H = a.H; p.M = a.M; p.S_max = a.S_max;
    p.blob = a.blob; p.layer_stride = a.layer_stride;
    p.ln1_w = a.ln1_w; p.ln1_b = a.ln1_b; p.attn_b = a.attn_b; p.proj_b = a.proj_b;
    p.ln2_w = a.ln2_w; p.ln2_b = a.ln2_b; p.fc_b = a.fc_b; p.proj2_b = a.proj2_b;
    p.tcw = a.tcw; p.X = a.X; p.A = a.A; p.QKV = a.QKV; p.U = a.U; p.ws = a.ws;
    p.kv = a.kv; p.kv_layer_stride = a.kv_layer_stride; p.gbar = a.gbar; p.prof = a.prof; p.dbg = a.dbg;
    void* args[] = {&p};
    cudaError_t e;
    if (a.M <= 64) {
        const size_t smem = tc::Cfg<64>::SMEM_BYTES + 256 + 2 * 4 * 258 * sizeof(float);
        e = cudaFuncSetAttribute(tc::prefill_fused_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        return cudaLaunchCooperativeKernel((void*)tc::prefill_fused_kernel<64>, dim3(grid), dim3(tc::PF_THREADS), args, smem, st);
    }
    const size_t smem = tc::Cfg<128>::SMEM_BYTES + 256 + 2 * 4 * 258 * sizeof(float);
    e = cudaFuncSetAttribute(tc::prefill_fused_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaLaunchCooperativeKernel((void*)tc::prefill_fused_kernel<128>, dim3(grid), dim3(tc::PF_THREADS), args, smem, st);
}

}  // namespace gv
