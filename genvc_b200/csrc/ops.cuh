// Per-op fp32 kernels: the building blocks of prefill, the teacher-forced latent pass, the
// perceiver and the batched (B>1) decode path.  Launchers return cudaError_t and bump a
// launch counter.  All kernels take an optional `skip` flag (device int): when it points at a
// non-zero value the kernel returns immediately — that is how steps enqueued after the
// device-side generation loop has finished (EOS / max length) become no-ops without a host sync.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gv {

enum { ACT_NONE = 0, ACT_GELU_NEW = 1 };

struct GemmArgs {
    const float* A;  // [M,K] row-major, leading dim lda
    int lda;
    const float* W;  // w_nk ? [N,K] (nn.Linear) : [K,N] (HF Conv1D); leading dim ldw
    int ldw;
    int w_nk;
    const float* bias;      // [N] or null
    const float* residual;  // [M,N] (ldr) or null; may alias C
    int ldr;
    float* C;  // [M,N], leading dim ldc
    int ldc;
    int M, N, K;
    int act;
    // optional fused follow-ups (launch_gemm runs them in the split-K epilogue when it can, else as separate launches):
    //   LayerNorm of the finished rows (N == D <= 1024): ln_out[m, :] = LN(C[m, :]; ln_w, ln_b)
    const float* ln_w = nullptr;
    const float* ln_b = nullptr;
    float* ln_out = nullptr;
    int ld_ln = 0;
    //   K/V cache append of a [q|k|v] projection (N == 3 D_kv): columns [D_kv, 3 D_kv) of row (b, i) also go to
    //   kv_{k,v} + b * kv_bs + ((col / hd) * S_max + pos0 + i) * hd + col % hd
    float* kv_k = nullptr;
    float* kv_v = nullptr;
    long kv_bs = 0;
    int kv_rows = 0, kv_H = 0, kv_S_max = 0, kv_pos0 = 0;
};

struct AttnArgs {
    const float* Q;
    long q_bs, q_rs, q_hs;  // element strides: batch, row, head
    const float* K;
    long k_bs, k_rs, k_hs;
    const float* V;
    long v_bs, v_rs, v_hs;
    float* O;
    long o_bs, o_rs, o_hs;
    int B, H, M, hd;
    int n_keys;  // keys available to a non-causal query; causal row i sees keys [0, pos0+i]
    int causal, pos0;
    float scale;
};

// Device-resident generation state shared by the per-op path and the fused decode kernel.
#define GV_MAX_BATCH 64
struct GenState {
    int n_emitted;    // tokens emitted per row so far (rows advance in lock-step)
    int done;         // all rows finished or max length reached
    int has_pending;  // logits/latent of the next step are already computed
    int P, B;
    int finished[GV_MAX_BATCH];
    long long last_tok[GV_MAX_BATCH];
};

struct SampleArgs {
    GenState* st;
    const float* logits;   // pending logits [B,V]
    const float* latent;   // pending latent [B,D]
    unsigned char* seen;   // [B,Vpad] ids already in input_ids (repetition penalty)
    int V, Vpad, D;
    int top_k;
    float top_p, top_p_threshold, temperature, rep_penalty;
    int ignore_eos, stop_token, max_total;  // max_total: stop when n_emitted reaches it
    unsigned long long seed;
    const float* noise;           // [B,V] for this step or null
    const long long* forced;      // [B] for this step or null
    long long* ids_out;           // [B]
    float* latents_out;           // [B,D]
    float* logits_out;            // [B,V] or null
    int* status;                  // {emitted this call, done, out-of-range ids seen, reserved}
    int* bad_ids;                 // workspace flag set by the embedding kernels when they clamp an id (read + cleared here)
    int step_in_call;
};

// tcgen05 3xTF32 GEMM (gemm_tc.cu).  Weights are pre-split / pre-tiled once (gemm_tc_pack registers the tiled copy
// under the reference-layout pointer W); launch_gemm routes a GEMM whose W is registered to the tensor-core
// kernel (it reports the split-K factor it used; the caller runs the split-K reduction epilogue when > 1).
size_t gemm_tc_packed_floats(int N, int K);  // 0 when K is not a multiple of 32 (not eligible)
cudaError_t gemm_tc_pack(const float* W, int N, int K, int ldw, int w_nk, float* out, cudaStream_t st);
void gemm_tc_forget(const float* lo, const float* hi);  // drop registrations of weights inside [lo, hi)
bool gemm_tc_eligible(const GemmArgs& a);
cudaError_t launch_gemm_tc(const GemmArgs& a, float* splitk_ws, size_t splitk_ws_floats, const int* skip, cudaStream_t st,
                           int* splits_out);
// Persistent fused prefill (gemm_tc.cu): the blocks of a batch-1 prefill of <= 128 rows in one cooperative launch.
struct PrefillFusedArgs {
    int L, D, H, M, S_max;
    const float* blob;
    long long layer_stride;
    long long ln1_w, ln1_b, attn_b, proj_b, ln2_w, ln2_b, fc_b, proj2_b;  // blob offsets of layer 0
    const float* const* tcw;  // device array [L][4]: packed tensor-core weights (c_attn, attn c_proj, c_fc, mlp c_proj)
    float *X, *A, *QKV, *U, *ws;
    float* kv;
    long long kv_layer_stride;
    unsigned* gbar;  // zeroed by the caller before the launch
    unsigned long long* prof = nullptr;  // debug phase profile [grid][16] (genvc_debug_trace), or null
    int dbg = 0;                         // debug knobs (results invalid): see gemm_tc.cu
};
// Perceiver cross-attention on tcgen05 (gemm_tc.cu): 32 latent queries, head dim 64, RC = 32 + S context keys.
bool pc_attention_tc_supported(int n_latents, int hd, int RC);
cudaError_t launch_pc_attention_tc(const float* Q, const float* KV, float* O, int B, int heads, int RC, cudaStream_t st);
bool prefill_fused_supported(int D, int H, int M, int grid, size_t ws_floats);
cudaError_t launch_prefill_fused(const PrefillFusedArgs& a, int grid, cudaStream_t st);
cudaError_t launch_gemm(const GemmArgs& a, float* splitk_ws, size_t splitk_ws_floats, const int* skip,
                        cudaStream_t st, unsigned long long* nlaunch);
// rows are addressed as group g = r / rpg, i = r % rpg: X + g * x_gs + i * x_stride
cudaError_t launch_layernorm(const float* X, long x_stride, long x_gs, float* Y, long y_stride, long y_gs, int rows, int rpg,
                             int D, const float* w1, const float* b1, const float* w2, const float* b2, const int* skip,
                             cudaStream_t st, unsigned long long* nlaunch);
cudaError_t launch_rmsnorm(const float* X, float* Y, int rows, int D, const float* gamma, cudaStream_t st,
                           unsigned long long* nlaunch);
cudaError_t launch_attention(const AttnArgs& a, const int* skip, cudaStream_t st, unsigned long long* nlaunch);
cudaError_t launch_kv_scatter(const float* qkv, int B, int M, int D, int H, float* kcache, float* vcache, long batch_stride,
                              int S_max, int pos0, const int* skip, cudaStream_t st, unsigned long long* nlaunch);
cudaError_t launch_embed_prefix(const float* cond, const long long* text_ids, int B, int T, int n_lat, int D,
                                const float* text_emb, const float* text_pos, int start_text, int stop_text, float* out,
                                long out_bs, int n_vocab, int* bad, cudaStream_t st, unsigned long long* nlaunch);
// rows of [tokens] with mel embeddings: out[b, r, :] = mel_emb[tok(b,r)] + mel_pos[pos0 + r]
// tok(b,r): r == 0 ? first_tok : (codes ? (r-1 < M ? codes[b, r-1] : pad_tok) : n/a)
cudaError_t launch_embed_mel_rows(const long long* codes, int B, int R, int M, int first_tok, int pad_tok, int pos0, int D,
                                  const float* mel_emb, const float* mel_pos, float* out, long out_bs, int n_vocab, int* bad,
                                  cudaStream_t st, unsigned long long* nlaunch);
cudaError_t launch_embed_last_token(const GenState* stt, int B, int pos, int D, const float* mel_emb, const float* mel_pos,
                                    float* out, cudaStream_t st, unsigned long long* nlaunch);
cudaError_t launch_copy_rows(const float* src, long src_bs, float* dst, long dst_bs, int B, long row_floats, cudaStream_t st,
                             unsigned long long* nlaunch);
cudaError_t launch_transpose_mel(const float* mel, int B, int C, int Cp, int S, float* out, cudaStream_t st,
                                 unsigned long long* nlaunch);
cudaError_t launch_geglu(const float* h, int rows, int inner, int inner_pad, float* out, cudaStream_t st,
                         unsigned long long* nlaunch);
cudaError_t launch_init_state(GenState* st, unsigned char* seen, int B, int P, int V, int Vpad, int start_audio,
                              cudaStream_t s, unsigned long long* nlaunch);
cudaError_t launch_sample(const SampleArgs& a, int B, cudaStream_t st, unsigned long long* nlaunch);
cudaError_t launch_kv_attention(const float* q, long q_bs, const float* k, const float* v, long kv_bs, int B, int H, int hd,
                                int S, int S_max, float* out, long o_bs, const int* skip, cudaStream_t st,
                                unsigned long long* nlaunch);

}  // namespace gv
