// Parameters of the fused persistent decode kernel (decode_mega.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ops.cuh"
#include "stream_layout.h"

namespace gv {

// hop counters (one 128-byte line each, zeroed by the host before every launch)
enum { HC_XQ = 0, HC_AO = 1, HC_X1 = 2, HC_PP = 3, HC_X2 = 4, HC_LG = 5, HC_COUNT = 6 };
#define GV_HOP_STRIDE 32  // uint32 words between two counters
// exchange tags of one forward: tag(layer l, buffer b) = tbase + GV_TAGS_PER_LAYER*l + b; logits = tbase + GV_TAGS_PER_LAYER*L
#define GV_TAGS_PER_LAYER 5
// debug timeline slots per layer (genvc_debug_trace)
#define GV_TRACE_PER_LAYER 28

struct MegaParams {
    int L, D, H, V, Vpad, S_max;
    int P;          // prefix length: cache row of mel position n is P + n
    int n_steps;    // tokens to emit in this launch (at most)
    // weights
    const float* stream;  // per-CTA weight streams (stream_layout.h)
    const float* blob;    // reference-layout blob (LayerNorm params, biases, embeddings)
    long long proj2_b_off, layer_stride, lnf_off, mel_emb_off, mel_pos_off;
    // activations / state (global, L2-resident)
    float* kv;
    long long kv_layer_stride;  // floats per (layer, k|v) plane
    // exchange buffers: every element is an 8-byte {value, tag} pair (decode_mega.cu)
    float* xq;      // [3D] q | k | v of the token being decoded
    float* att_o;   // [items][hd] un-normalised partial attention outputs
    float* att_ml;  // [items][2]  (max, sum)
    float* x1;      // [D]  residual stream after attention
    float* pp;      // [G][D] per-CTA partial sums of mlp.c_proj (reducer-CTA variant)
    unsigned long long* acc;  // [2][L][D] self-counting fixed-point accumulators of mlp.c_proj (zero at launch)
    float* x2;      // [D]  residual stream leaving the block
    float* lg;      // [V]  logits
    unsigned* hops;  // [HC_COUNT * GV_HOP_STRIDE] arrival counters, zero at launch
    unsigned tag0;   // first exchange tag of this launch (never 0; never reused while data with it is live)
    float* pend_logits;  // [V]
    float* pend_latent;  // [D]
    // generation state, double-buffered: a launch READS st / seen and WRITES st_out / seen_out (every CTA reads the
    // state at kernel start with no grid-wide ordering, so the launch must not overwrite what it reads)
    const GenState* st;
    const unsigned char* seen;  // [B][Vpad]
    GenState* st_out;
    unsigned char* seen_out;    // [B][Vpad]
    int B;         // rows (decode_batch.cu: 1..GV_BATCH_ROWS; decode_mega.cu: 1)
    float* tokx;   // [GV_BATCH_ROWS] sampled tokens of the step, tagged (decode_batch.cu: row r is sampled by CTA r)
    float* ao;     // [B][D] merged, normalised attention output rows (decode_batch.cu: written by the last item of a (row, head))
    unsigned* att_cnt;  // [grid * GV_ATTCNT_STRIDE] items finished per (row, head), zero at launch (decode_batch.cu)
    // projected-value variant of the single-row kernel (decode_mega.cu, PVW): null = classic K / V attention items
    float* vw;     // [L][grid][S_max][H][8] cache of v_j . W_proj per head, sliced by the CTA that owns the output columns
    float* sbuf;   // [H][S_max] scaled attention scores of the step, tagged
    int vw_fill;   // 1: the first forward of this launch computes the projected values of all cached positions (after a prefill)
    // sampling
    int top_k;
    float top_p, top_p_threshold, temperature, rep_penalty;
    int ignore_eos, stop_token, max_total;
    unsigned long long seed;
    const float* noise;        // [n_steps, V] or null
    const long long* forced;   // [n_steps] or null
    // outputs
    long long* ids_out;   // [n_steps]
    float* latents_out;   // [n_steps, D]
    float* logits_out;    // [n_steps, V] or null
    int* status;          // {emitted, done, out-of-range ids seen, reserved}
    int* bad_ids;         // workspace flag set by the embedding kernels when they clamp an id (read + cleared by CTA 0)
    // debug timeline: tid 0 of every CTA stamps %globaltimer at phase boundaries of step `trace_step`
    unsigned long long* trace;  // [grid][trace_slots] or null
    int trace_step, trace_slots;
    // tuning / debug knobs (genvc_debug_tune)
    int window;      // producer: max tiles in flight (1..GV_MEGA_NSLOT)
    int l2_ahead_tiles;  // producer: L2 prefetch distance in tiles (0 = off)
    int hop_near, hop_near_ao;  // early-release margins of the hops (-1 = default; decode_mega.cu: hop_wait)
    int hop_hold;        // 1: the producer issues no new bulk copies while the consumers poll / load a hop
    int hop_settle_ns;   // pause between seeing a hop counter complete and loading the data (decode_mega.cu: hop_wait)
    int dbg_nosync;  // consumers do not wait for exchange data (results are garbage; streaming-rate probe)
};

// rows the batched fused kernel decodes per pass of the weight stream (one mma n-tile)
#define GV_BATCH_ROWS 8
#define GV_BATCH_NSLOT 10  // ring depth of the batched kernel (its row buffers take the rest of shared memory)
// exchange tags of one step of the batched kernel: the forward's tags + one for the sampled tokens
#define GV_BATCH_TAGS_EXTRA 1
#define GV_ATTCNT_STRIDE 8  // uint32 words between two (row, head) item counters

size_t mega_smem_bytes(int D, int Vpad);
size_t batch_smem_bytes(int D, int Vpad);
cudaError_t launch_decode_batch(const MegaParams& p, int grid, cudaStream_t st);
cudaError_t launch_decode_mega(const MegaParams& p, int grid, cudaStream_t st);
cudaError_t launch_pack_stream(const StreamDims& s, int layer, int ph, const float* W, const float* bias, int w_nk,
                               const float* lnw, const float* lnb, float* stream, cudaStream_t st);

}  // namespace gv
