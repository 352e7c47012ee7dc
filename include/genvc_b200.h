/*
 * genvc_b200 — C ABI of the B200-native GenVC codec-token inference path.
 *
 * This is the drop-in boundary: a shared library (libgenvc_b200.so) with plain
 * pointers and sizes in every signature.  The reference is pure Python/PyTorch
 * and has no FFI of its own; each entry point below names the reference
 * function it replaces (file:line under the reference repo).  The Python shim
 * that binds these with ctypes and re-exposes the reference's GPT interface is
 * genvc_b200/gpt.py; INTEGRATION.md shows the stub a maintainer of the
 * reference would add.
 *
 * Ownership: the CALLER (PyTorch) owns every device buffer — weight blob,
 * decode weight stream, KV cache, workspace, inputs and outputs; the library
 * allocates nothing on the device except a 4-byte grid-barrier counter inside
 * the caller's workspace.  Pointers must stay valid while the context uses
 * them.  All device pointers are on the device given to genvc_create.
 *
 * Errors: every function returns 0 on success or a negative code (GENVC_E_*);
 * nothing throws.  genvc_last_error() returns a human-readable message for the
 * most recent failure on that context.
 *
 * Streams: calls enqueue on the CUDA stream passed as `stream` (a cudaStream_t
 * cast to void*; NULL = default stream), are asynchronous with respect to the
 * host, and are not re-entrant per context.
 *
 * All tensors are fp32 and contiguous unless a comment says otherwise; token
 * ids are int64 (the reference's torch.long).
 */
#ifndef GENVC_B200_H
#define GENVC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GENVC_OK 0
#define GENVC_E_INVALID (-1)   /* bad argument / unsupported shape        */
#define GENVC_E_STATE (-2)     /* call out of order (e.g. decode w/o prefill) */
#define GENVC_E_CUDA (-3)      /* a CUDA runtime call failed              */
#define GENVC_E_UNSUPPORTED (-4) /* shape not supported by the fused decode kernel */

typedef struct genvc_ctx genvc_ctx;

/* Model shape.  Mirrors GPT.__init__ (layers/gpt.py:88-188) and the hard-coded
 * PerceiverResampler arguments (layers/gpt.py:179-188). */
typedef struct genvc_config {
    int32_t n_layer;        /* gpt_layers                                   */
    int32_t d_model;        /* gpt_n_model_channels                         */
    int32_t n_head;         /* gpt_n_heads                                  */
    int32_t n_text_vocab;   /* gpt_number_text_tokens (258)                 */
    int32_t n_audio_vocab;  /* gpt_num_audio_tokens (1026)                  */
    int32_t start_text, stop_text;   /* 256 / 257                           */
    int32_t start_audio, stop_audio; /* 1024 / 1025                         */
    int32_t n_mel_pos;      /* rows of mel_pos_embedding (608)              */
    int32_t n_text_pos;     /* rows of text_pos_embedding (404)             */
    int32_t max_gen_mel_tokens; /* 602, layers/gpt.py:131                   */
    int32_t pc_depth, pc_dim_context, pc_latents, pc_dim_head, pc_heads, pc_ff_inner;
    int32_t max_batch;      /* rows of the static KV cache                  */
    int32_t max_seq;        /* positions of the static KV cache (>= P+1+602) */
    int32_t max_mel_frames; /* longest reference mel the perceiver accepts  */
} genvc_config;

/* HF generate() knobs the reference passes (inference/inference_utils.py:55-66,
 * 170-182) in the order HF 4.33 applies them: repetition penalty -> temperature
 * -> top-k -> top-p -> softmax -> multinomial (layers/stream_generator.py:834-858). */
typedef struct genvc_sampling {
    int32_t top_k;             /* 0 disables                                  */
    float top_p;               /* >= 1 disables                               */
    float top_p_threshold;     /* (float)(1 - top_p) as torch computes it     */
    float temperature;
    float repetition_penalty;
    int32_t ignore_eos;        /* bench mode: never finish on stop_audio      */
    int32_t max_new_tokens;    /* <=0: reference cap (max_gen_mel_tokens)     */
    uint64_t seed;             /* on-device Philox when exp_noise == NULL     */
} genvc_sampling;

/* ---- lifetime ------------------------------------------------------------ */
int genvc_create(const genvc_config* cfg, int device, genvc_ctx** out);
void genvc_destroy(genvc_ctx* ctx);
const char* genvc_last_error(const genvc_ctx* ctx);
/* Number of SMs the fused decode kernel will occupy (one persistent CTA each). */
int genvc_decode_grid(const genvc_ctx* ctx);
/* Largest batch the fused persistent decode kernels take in one launch: 0 (shape unsupported: per-op kernels only),
 * 1 (single-row kernel) or up to 8 (batched kernel: rows share one pass of the weight stream; equal-length rows with
 * a shared position index as layers/gpt_inference.py:92-96, finished rows padded as layers/stream_generator.py:860-881). */
int genvc_fused_max_rows(const genvc_ctx* ctx);

/* ---- weights: checkpoint layout -> device blob ---------------------------- *
 * Replaces model.load_state_dict(...).to(device) for the `gpt.*` keys
 * (inference/model_init.py:22-24).  The caller allocates `genvc_blob_floats`
 * floats and copies each state-dict tensor (key WITHOUT the "gpt." prefix, e.g.
 * "gpt.h.3.attn.c_attn.weight") to the place genvc_tensor_info reports:
 * rows x cols elements at float offset `offset` with row stride `row_stride`
 * (>= cols; padding must be zero).  Conv1D weights stay [in,out], nn.Linear
 * weights stay [out,in]. */
uint64_t genvc_blob_floats(const genvc_ctx* ctx);
int genvc_num_tensors(const genvc_ctx* ctx);
int genvc_tensor_name(const genvc_ctx* ctx, int index, char* buf, size_t buf_len);
int genvc_tensor_info(const genvc_ctx* ctx, const char* key, uint64_t* offset,
                      uint64_t* rows, uint64_t* cols, uint64_t* row_stride);
int genvc_bind_weights(genvc_ctx* ctx, const float* blob_dev, uint64_t n_floats);

/* Decode weight stream: the 30 blocks' matrices + mel_head re-tiled so that each
 * persistent CTA of the fused decode kernel reads ONE contiguous byte stream in
 * the order it consumes it (column slices, K-contiguous, bias folded in). */
uint64_t genvc_stream_floats(const genvc_ctx* ctx);
int genvc_pack_stream(genvc_ctx* ctx, float* stream_dev, uint64_t n_floats, void* stream);

/* Tensor-core weight copy: the dense matrices of the batched GEMMs (prefill, latent pass, perceiver projections)
 * pre-split into TF32 hi | lo halves and pre-tiled into 32 KB blocks (128 output columns x 32 k in the UMMA
 * canonical layout) so that one bulk TMA copy feeds one tcgen05 pipeline stage.  Optional: without it those GEMMs
 * run on the fp32 CUDA-core kernel.  The buffer must stay valid while the weights are bound. */
uint64_t genvc_tc_floats(const genvc_ctx* ctx);
int genvc_pack_tc(genvc_ctx* ctx, float* tc_dev, uint64_t n_floats, void* stream);

/* ---- caller-owned scratch -------------------------------------------------- */
uint64_t genvc_kv_floats(const genvc_ctx* ctx);       /* [L][2][max_batch][H][max_seq][hd] */
uint64_t genvc_workspace_bytes(const genvc_ctx* ctx);
int genvc_bind_buffers(genvc_ctx* ctx, float* kv_dev, uint64_t kv_floats,
                       void* workspace_dev, uint64_t workspace_bytes);

/* Optional projected-value cache of the single-row fused decode kernel: [L][grid][max_seq][H][8] floats holding
 * v_j . W_proj per head and position, sliced by the CTA that owns the output columns.  attn c_proj is linear, so the
 * kernel applies it to the new value while the softmax is still being computed and sums cached projected values
 * instead of merging partial attention outputs (HF GPT2Attention: attn_output = c_proj(softmax(q k^T) v); same
 * arithmetic, different association).  genvc_prefill fills the prefix rows (batch 1).  NULL unbinds it.
 * genvc_vw_floats returns 0 when the shape has no such variant. */
uint64_t genvc_vw_floats(const genvc_ctx* ctx);
int genvc_bind_vw(genvc_ctx* ctx, float* vw_dev, uint64_t n_floats);

/* ---- the path -------------------------------------------------------------- */

/* GPT.get_style_emb -> PerceiverResampler.forward (layers/gpt.py:351-373,
 * layers/perceiver_encoder.py:265-276): mel [B,80,S] -> latents [B,32,D]
 * (the transpose to [B,D,32] the reference returns is a view on the host side). */
int genvc_perceiver(genvc_ctx* ctx, const float* mel_dev, int B, int S_mel,
                    float* latents_out_dev, void* stream);

/* GPT.compute_embeddings (layers/gpt.py:572-592): cond [B,32,D], text ids [B,T]
 * -> prefix embeddings [B,P,D], P = 32+T+2. */
int genvc_embed_prefix(genvc_ctx* ctx, const float* cond_dev, const int64_t* text_ids_dev,
                       int B, int T, float* prefix_out_dev, void* stream);

/* First iteration of the generate loop: GPT2InferenceModel.forward with the
 * stored prefix (layers/gpt_inference.py:81-91, 97-112) — P prefix rows plus the
 * start_audio row through all blocks, KV cache filled for positions [0,P],
 * logits and latent of the last row left pending on the device.  Resets the
 * generation state (step 0, nothing emitted, fake ids {1, start_audio} marked
 * for the repetition penalty). */
int genvc_prefill(genvc_ctx* ctx, const float* prefix_dev, int B, int P, void* stream);

/* Up to n_steps iterations of sample()/sample_stream()
 * (layers/stream_generator.py:809-881; HF GenerationMixin.sample): per step
 * [forward of the previous token with the KV cache, unless logits are already
 * pending] -> processors/warpers -> sample -> emit (token, latent) -> EOS /
 * max-length bookkeeping, tokens fed back on the device.
 *   exp_noise_dev : [n_steps,B,V] Exp(1) variates consumed as torch.multinomial
 *                   does (argmax(p/q)), or NULL for on-device Philox noise.
 *   forced_ids_dev: [n_steps,B] tokens to emit instead of the sampled ones
 *                   (teacher forcing for logits parity), or NULL.
 *   ids_out_dev   : [n_steps,B] int64;  latents_out_dev: [n_steps,B,D];
 *   logits_out_dev: [n_steps,B,V] raw logits before processing, or NULL.
 *   status_dev    : int32[4] = {steps emitted by this call, done flag, bad-id flag, reserved}.  The bad-id flag is 1
 *                   when a text id (genvc_embed_prefix) or a forced id since the previous status was outside its
 *                   vocabulary: the kernels clamp such ids (no out-of-bounds read) and report them here.
 *   mode          : 0 auto, 1 per-op kernels, 2 fused persistent kernel (B <= genvc_fused_max_rows). */
int genvc_decode(genvc_ctx* ctx, int n_steps, const genvc_sampling* sp,
                 const float* exp_noise_dev, const int64_t* forced_ids_dev,
                 int64_t* ids_out_dev, float* latents_out_dev, float* logits_out_dev,
                 int32_t* status_dev, int mode, void* stream);

/* GPT.forward(..., return_latent=True) (layers/gpt.py:375-508): teacher-forced
 * uncached pass over [cond(32) | text(T+2) | start, codes, stop x4]; returns
 * final_norm(ln_f(h)) of the M code rows: [B,M,D]. */
int genvc_forward_latents(genvc_ctx* ctx, const float* cond_dev, const int64_t* text_ids_dev, int T,
                          const int64_t* codes_dev, int M, int B, float* latents_out_dev, void* stream);

/* Single-token KV-cache attention microbenchmark (BASELINE.json configs[4]); the
 * arithmetic of HF GPT2Attention._attn for one query per (cache, head):
 * q [N,H,hd]; k,v caches [N,H,S_max,hd]; out [N,H,hd]; S keys are attended.
 * Needs no context. */
int genvc_kv_attention(const float* q_dev, const float* k_dev, const float* v_dev,
                       int N, int H, int hd, int S, int S_max, float* out_dev, void* stream);

/* Number of kernels launched by this context since creation (bench bookkeeping). */
uint64_t genvc_launch_count(const genvc_ctx* ctx);

/* Debug/profiling hook (no reference counterpart): when trace_dev != NULL, thread 0 of
 * every persistent CTA of the fused decode kernel writes %globaltimer (ns) at each phase
 * boundary of step `step` of every following genvc_decode launch into
 * trace_dev[cta * slots_per_cta + slot]; slot = layer*14 + k with k = 0 QKV input ready,
 * 1 LN1 done, 2 QKV done, 3 attention item done, 4 attention output merged, 5 PROJ done, 6 FC
 * input ready, 7 LN2 done, 8 FC done, 9 mlp.c_proj partial done, 10 partial sums gathered
 * (reducer CTAs), 11 reduce done; k = 12, 13 are DURATIONS (ns thread 0 waited for weight tiles
 * in QKV+PROJ and in FC+P2); n_layer*14 + {0,1,2,3,4} = head input ready, head done, logits
 * gathered, sample done, step start.  The buffer must hold grid * slots_per_cta words.  NULL
 * switches it off. */
int genvc_debug_trace(genvc_ctx* ctx, uint64_t* trace_dev, int slots_per_cta, int step);

/* ---- the stage after the path: HiFi-GAN generator building blocks ----
 * Stateless (no context): the kernels are enqueued on `stream` of the CALLER'S CURRENT device, which must own every
 * pointer (the host classes wrap their calls in the device of their tensors); device pointers; fp32; layouts [B, C, T].  The host side (genvc_b200/vocoder.py) strings them
 * together as layers/hifigan.py:210-225 does.  Weights are passed REPACKED to [Cin][K][Cout]
 * (reference: Conv1d weight [Cout][Cin][K] -> permute(1,2,0); ConvTranspose1d weight [Cin][Cout][K] -> permute(0,2,1)),
 * weight norm already folded (w = g * v / ||v||).
 *
 * genvc_conv1d: torch.nn.Conv1d(Cin, Cout, K, stride 1|2, padding, dilation); To = (T + 2*padding - dilation*(K-1) - 1) / stride + 1
 *   v = bias[co] + sum_ci sum_j w[ci][j][co] * lrelu(x[b][ci][t*stride + j*dilation - padding], pre_slope)   (pre_slope 1 = none)
 *   v += residual[b][co][t] if residual (needs To == T);  v += y[b][co][t] if accumulate;  v *= out_scale;
 *   act_out: 0 none, 1 tanh, 2 relu;  y = v
 *   — i.e. `xt = c(leaky_relu(x)); x = xt + x` of ResBlock2 (layers/hifigan.py:147-152) is ONE call, and
 *   `xs += resblock(x)`, `x = xs / num_kernels` (:216-221) ride in the epilogue of each block's last conv.
 * genvc_conv_transpose1d: torch.nn.ConvTranspose1d(Cin, Cout, K, stride, padding) on lrelu(x, pre_slope)
 *   (layers/hifigan.py:213-214); Tout = (Tin - 1) * stride - 2 * padding + K.
 * scratch_dev (optional, scratch_floats floats, caller-owned): layers with too few output tiles to fill the GPU are split
 *   over their input channels into up to scratch_floats / (B*Cout*T) slices that are summed in a fixed order (results do
 *   not depend on timing); NULL keeps every layer in one pass. */
int genvc_conv1d(const float* x_dev, const float* w_dev, const float* bias_dev, const float* residual_dev, float* y_dev,
                 int B, int Cin, int Cout, int T, int K, int dilation, int padding, int stride, float pre_slope,
                 int accumulate, float out_scale, int act_out, float* scratch_dev, uint64_t scratch_floats, void* stream);
int genvc_conv_transpose1d(const float* x_dev, const float* w_dev, const float* bias_dev, float* y_dev, int B, int Cin,
                           int Cout, int Tin, int K, int stride, int padding, float pre_slope, float* scratch_dev,
                           uint64_t scratch_floats, void* stream);

/* ---- the stage before the path: content-DVAE tokeniser ----
 * DiscreteVAE.get_codebook_indices (layers/dvae.py:324-331): encoder = strided convolutions + ReLU, ResBlocks, 1x1
 * convolution (genvc_conv1d above), then Quantize.forward (layers/dvae.py:84-88): for the encoder output x [B, dim, T]
 * and the codebook embed [dim, n_embed]:  codes[b][t] = argmax_n -( |f|^2 - 2 f.E[:,n] + |E[:,n]|^2 ), first index on ties. */
int genvc_codebook_argmin(const float* x_dev, const float* embed_dev, int64_t* codes_dev, int B, int dim, int n_embed, int T,
                          void* stream);

/* Mel front-end of the conditioning path (utils.py:95-158 TorchMelSpectrogram; a1 = trainers/hifigan_trainer.py:438-455 calls
 * it on every 6 s chunk of the reference audio): torchaudio MelSpectrogram (power 2, centre / reflect padding, window centred
 * in n_fft) + log(clamp(., clamp_min)) / norms.  wav [B, N] -> mel [B, n_mels, 1 + N / hop].  Tables are caller-built
 * (genvc_b200/mel.py): window [n_fft] (zeros outside [win_lo, win_hi)), twiddle [n_fft][2] = cos, sin(2 pi k / n_fft),
 * fbank [n_fft / 2 + 1][n_mels] (torchaudio.functional.melscale_fbanks), norms [n_mels] or NULL.  n_fft: power of two <= 4096. */
int genvc_mel_spectrogram(const float* wav_dev, int B, int N, const float* window_dev, const float* twiddle_dev,
                          const float* fbank_dev, const float* norms_dev, float* mel_dev, int n_fft, int hop, int win_lo,
                          int win_hi, int n_mels, float clamp_min, void* stream);

/* Post-mortem aid (tools/hang_dump.py): byte offsets inside the workspace of the fused decode kernel's exchange buffers and
 * arrival counters: out[0..11] = xq, att_o, att_ml, x1, pp, x2, lg, hops, sbuf offsets, then grid, counter stride (words),
 * number of counters.  The buffers hold {value, tag} pairs; reading them from a side stream while a launch is stuck shows
 * which producer of which exchange never stored. */
int genvc_debug_layout(const genvc_ctx* ctx, uint64_t* out, int n);

/* Tuning / debug knobs of the fused decode kernel (no reference counterpart):
 *   window > 0 : weight tiles the per-SM TMA producer keeps in flight (requested, not landed);
 *   nosync != 0: consumers do not wait for exchange data — results are garbage; probes the pure
 *                weight-streaming rate.  Never set outside profiling;
 *   l2_ahead_tiles >= 0: distance (in 16 KB tiles per SM) at which the producer prefetches the
 *                weight stream HBM -> L2 ahead of the shared-memory ring (0 = off; < 0 keeps);
 *   hop_settle_ns >= 0: pause between seeing an exchange counter complete and loading the data (< 0 keeps);
 *   hop_hold 0 / 1: 1 = the producer issues no new bulk copies while the consumers poll / load an exchange;
 *   hop_hold >= 1000: 1000 + near + 100 * near_ao = early-release margins of the exchanges (how many arrivals may be
 *   outstanding when a CTA starts polling the tagged data itself; 1000 = none; default: grid / 37 and items / 3 <= 4). */
int genvc_debug_tune(genvc_ctx* ctx, int window, int nosync, int l2_ahead_tiles, int hop_settle_ns, int hop_hold);

#ifdef __cplusplus
}
#endif
#endif /* GENVC_B200_H */
