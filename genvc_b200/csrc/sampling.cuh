// HF-4.33 sampling chain for one row, executed by a 256-thread group:
//   RepetitionPenalty -> Temperature -> TopK -> TopP -> softmax -> multinomial
// as wired by layers/stream_generator.py:333-344, 412-414 and applied at :837-858
// (bodies: transformers generation/logits_process.py).  multinomial(p,1) is evaluated the way
// ATen's CPU kernel does it: argmax(p / q) with q ~ Exp(1) — the host can pass the very q the
// CPU reference would draw (exp_noise), otherwise q comes from an on-device Philox stream.
//
// Shared by the stand-alone sample kernel (per-op path) and the fused decode kernel, so both
// produce identical tokens.
#pragma once
#include "common.cuh"

namespace gv {

#define GV_SORT_N 2048  // >= n_audio_vocab (1026), power of two
#define GV_SAMPLE_THREADS 256

struct SampleCfg {
    int V;
    int top_k;
    float top_p, top_p_threshold, temperature, rep_penalty;
};

__device__ __forceinline__ uint32_t f2ord(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
__device__ __forceinline__ float key_val(unsigned long long k) { return ord2f((uint32_t)(k >> 32)); }
__device__ __forceinline__ int key_idx(unsigned long long k) { return (int)(0xffffffffu - (uint32_t)k); }

// Scratch layout in shared memory (caller provides): keys[GV_SORT_N] (u64), fscr[16] floats, iscr[8] ints.
// `load_logit(e)` returns raw logit e (L2 or shared memory).  `seen[t] != 0` marks ids present in input_ids.  Returns the token
// (same value in every thread).  `sync` must synchronise exactly the 256 participating threads.
template <class Load, class Sync>
__device__ int sample_token(Load load_logit, const unsigned char* seen, const SampleCfg& c,
                            const float* __restrict__ noise, unsigned long long seed, uint32_t step, uint32_t row,
                            unsigned long long* keys, float* fscr, int* iscr, int tid, Sync sync) {
    const int V = c.V;
    // 1) processors that act element-wise: repetition penalty, temperature
    for (int e = tid; e < GV_SORT_N; e += GV_SAMPLE_THREADS) {
        unsigned long long key = 0ull;
        if (e < V) {
            float s = load_logit(e);
            if (c.rep_penalty != 1.0f && seen[e]) s = (s < 0.0f) ? s * c.rep_penalty : s / c.rep_penalty;
            if (c.temperature != 1.0f) s = s / c.temperature;
            key = ((unsigned long long)f2ord(s) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)e);
        }
        keys[e] = key;
    }
    sync();

    if (c.top_k == 1) {
        // Greedy: the arg-max of the penalised logits.  (An exact tie at the maximum — which the
        // reference would break by sampling — resolves to the lowest index.)
        unsigned long long best = 0ull;
        for (int e = tid; e < GV_SORT_N; e += GV_SAMPLE_THREADS) best = max(best, keys[e]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
        unsigned long long* wbest = reinterpret_cast<unsigned long long*>(fscr);  // 8 warps x u64 = 16 floats
        sync();  // fscr may still be read from a previous call
        if ((tid & 31) == 0) wbest[tid >> 5] = best;
        sync();
        best = wbest[0];
#pragma unroll
        for (int w = 1; w < GV_SAMPLE_THREADS / 32; ++w) best = max(best, wbest[w]);
        sync();
        return key_idx(best);
    }

    // 2) full descending sort (value, then lower index first)
    for (int k = 2; k <= GV_SORT_N; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < GV_SORT_N; i += GV_SAMPLE_THREADS) {
                int ixj = i ^ j;
                if (ixj > i) {
                    unsigned long long a = keys[i], b = keys[ixj];
                    bool desc = ((i & k) == 0);
                    if (desc ? (a < b) : (a > b)) {
                        keys[i] = b;
                        keys[ixj] = a;
                    }
                }
            }
            sync();
        }
    }

    // 3) top-k / top-p survivor count and softmax denominator (serial over the kept prefix)
    if (tid == 0) {
        int n_keep = V;
        if (c.top_k > 0) {
            int k = min(c.top_k, V);
            float kth = key_val(keys[k - 1]);
            n_keep = k;
            while (n_keep < V && key_val(keys[n_keep]) == kth) ++n_keep;  // ties at the threshold survive
        }
        const float maxv = key_val(keys[0]);
        int n_final = n_keep;
        if (c.top_p < 1.0f) {
            // softmax over the kept set (removed entries are -inf -> 0), summed smallest-first
            float sum = 0.0f;
            for (int i = n_keep - 1; i >= 0; --i) sum += expf(key_val(keys[i]) - maxv);
            // ascending cumulative sum accumulated in double like ATen's CPU cumsum, rounded per element
            double cum = 0.0;
            n_final = 1;  // min_tokens_to_keep = 1: the largest is never removed
            for (int i = n_keep - 1; i >= 1; --i) {
                float p = expf(key_val(keys[i]) - maxv) / sum;
                cum += (double)p;
                if (!((float)cum <= c.top_p_threshold)) {
                    n_final = i + 1;
                    break;
                }
            }
        }
        float fsum = 0.0f;
        for (int i = n_final - 1; i >= 0; --i) fsum += expf(key_val(keys[i]) - maxv);
        iscr[0] = n_final;
        fscr[0] = maxv;
        fscr[1] = fsum;
    }
    sync();
    const int n_final = iscr[0];
    const float maxv = fscr[0], fsum = fscr[1];
    if (n_final == 1) {
        int tok = key_idx(keys[0]);
        sync();
        return tok;
    }
    // 4) multinomial(softmax(scores), 1) == argmax(p / q)
    float best_r = -1.0f;
    int best_i = 0x7fffffff;
    for (int i = tid; i < n_final; i += GV_SAMPLE_THREADS) {
        int idx = key_idx(keys[i]);
        float p = expf(key_val(keys[i]) - maxv) / fsum;
        float q = noise ? ldcg(noise + idx) : philox_exponential(seed, step, row, (uint32_t)idx);
        float r = p / q;
        if (r > best_r || (r == best_r && idx < best_i)) {
            best_r = r;
            best_i = idx;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float orr = __shfl_xor_sync(0xffffffffu, best_r, o);
        int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (orr > best_r || (orr == best_r && oi < best_i)) {
            best_r = orr;
            best_i = oi;
        }
    }
    sync();  // everyone is past reading fscr[0..1] / iscr[0]
    if ((tid & 31) == 0) {
        fscr[2 + (tid >> 5)] = best_r;
        iscr[1 + (tid >> 5)] = best_i;
    }
    sync();
    best_r = fscr[2];
    best_i = iscr[1];
#pragma unroll
    for (int w = 1; w < GV_SAMPLE_THREADS / 32; ++w) {
        float orr = fscr[2 + w];
        int oi = iscr[1 + w];
        if (orr > best_r || (orr == best_r && oi < best_i)) {
            best_r = orr;
            best_i = oi;
        }
    }
    sync();
    return best_i;
}

}  // namespace gv
