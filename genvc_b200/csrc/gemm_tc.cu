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

}  // namespace tc

// ---------------------------------------------------------------------------------------------
// host side: registry of pre-tiled weights (keyed by the reference-layout device pointer), launcher
// ---------------------------------------------------------------------------------------------
}  // namespace gv

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

}  // namespace gv
