// Host-side description of the device weight blob (one contiguous fp32 buffer owned by the caller).
//
// Tensors keep the reference state-dict names (keys under "gpt." in the checkpoint,
// inference/model_init.py:22) and orientations: HF Conv1D weights [in,out], nn.Linear [out,in].
// Every tensor starts on a 128-byte boundary; LayerNorm weight/bias pairs are adjacent so one
// bulk copy fetches both; the perceiver's second FF matrix is stored with its inner dimension
// padded to a multiple of 4 floats so rows stay 16-byte aligned for vector loads.
#pragma once
#include <stdint.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/genvc_b200.h"

namespace gv {

struct TensorEntry {
    std::string name;
    uint64_t off, rows, cols, stride;
};

struct LayerOff {
    uint64_t ln1_w, ln1_b, attn_w, attn_b, proj_w, proj_b, ln2_w, ln2_b, fc_w, fc_b, proj2_w, proj2_b;
};
struct PcLayerOff {
    uint64_t to_q, to_kv, to_out, ff0_w, ff0_b, ff2_w, ff2_b;
};

struct Layout {
    std::vector<TensorEntry> tensors;
    std::unordered_map<std::string, int> index;
    std::vector<LayerOff> layers;
    std::vector<PcLayerOff> pc_layers;
    uint64_t text_emb, mel_emb, mel_pos, text_pos, lnf_w, lnf_b, fn_w, fn_b, mel_head_w, mel_head_b, text_head_w,
        text_head_b;
    uint64_t pc_latents, pc_proj_w, pc_proj_b, pc_gamma;
    uint64_t pc_ff_inner_pad;  // row stride of ff.2.weight: inner padded to a multiple of 32 (k granularity of the tcgen05 GEMM)
    uint64_t pc_ctx_pad;       // row stride of proj_context.weight: dim_context (80) padded to a multiple of 32
    uint64_t total = 0;

    uint64_t add(const std::string& name, uint64_t rows, uint64_t cols, uint64_t stride = 0, bool align = true) {
        if (stride == 0) stride = cols;
        if (align) total = (total + 31) & ~uint64_t(31);
        TensorEntry e{name, total, rows, cols, stride};
        index[name] = (int)tensors.size();
        tensors.push_back(e);
        total += rows * stride;
        return e.off;
    }

    void build(const genvc_config& c) {
        const uint64_t D = c.d_model;
        text_emb = add("text_embedding.weight", c.n_text_vocab, D);
        mel_emb = add("mel_embedding.weight", c.n_audio_vocab, D);
        mel_pos = add("mel_pos_embedding.emb.weight", c.n_mel_pos, D);
        text_pos = add("text_pos_embedding.emb.weight", c.n_text_pos, D);
        layers.resize(c.n_layer);
        for (int i = 0; i < c.n_layer; ++i) {
            std::string p = "gpt.h." + std::to_string(i) + ".";
            LayerOff& l = layers[i];
            l.ln1_w = add(p + "ln_1.weight", 1, D);
            l.ln1_b = add(p + "ln_1.bias", 1, D, 0, false);  // adjacent to the weight
            l.attn_w = add(p + "attn.c_attn.weight", D, 3 * D);
            l.attn_b = add(p + "attn.c_attn.bias", 1, 3 * D);
            l.proj_w = add(p + "attn.c_proj.weight", D, D);
            l.proj_b = add(p + "attn.c_proj.bias", 1, D);
            l.ln2_w = add(p + "ln_2.weight", 1, D);
            l.ln2_b = add(p + "ln_2.bias", 1, D, 0, false);
            l.fc_w = add(p + "mlp.c_fc.weight", D, 4 * D);
            l.fc_b = add(p + "mlp.c_fc.bias", 1, 4 * D);
            l.proj2_w = add(p + "mlp.c_proj.weight", 4 * D, D);
            l.proj2_b = add(p + "mlp.c_proj.bias", 1, D);
        }
        // ln_f and final_norm: four adjacent vectors (one bulk copy in the fused decode kernel)
        lnf_w = add("gpt.ln_f.weight", 1, D);
        lnf_b = add("gpt.ln_f.bias", 1, D, 0, false);
        fn_w = add("final_norm.weight", 1, D, 0, false);
        fn_b = add("final_norm.bias", 1, D, 0, false);
        mel_head_w = add("mel_head.weight", c.n_audio_vocab, D);
        mel_head_b = add("mel_head.bias", 1, c.n_audio_vocab);
        text_head_w = add("text_head.weight", c.n_text_vocab, D);  // loaded for completeness; unused at inference
        text_head_b = add("text_head.bias", 1, c.n_text_vocab);
        const std::string pc = "conditioning_perceiver.";
        const uint64_t inner = (uint64_t)c.pc_dim_head * c.pc_heads;
        const uint64_t ffi = c.pc_ff_inner;
        pc_ff_inner_pad = (ffi + 31) & ~uint64_t(31);
        pc_ctx_pad = ((uint64_t)c.pc_dim_context + 31) & ~uint64_t(31);
        pc_latents = add(pc + "latents", c.pc_latents, D);
        pc_proj_w = add(pc + "proj_context.weight", D, c.pc_dim_context, pc_ctx_pad);
        pc_proj_b = add(pc + "proj_context.bias", 1, D);
        pc_layers.resize(c.pc_depth);
        for (int i = 0; i < c.pc_depth; ++i) {
            std::string a = pc + "layers." + std::to_string(i) + ".0.";
            std::string f = pc + "layers." + std::to_string(i) + ".1.";
            PcLayerOff& l = pc_layers[i];
            l.to_q = add(a + "to_q.weight", inner, D);
            l.to_kv = add(a + "to_kv.weight", 2 * inner, D);
            l.to_out = add(a + "to_out.weight", D, inner);
            l.ff0_w = add(f + "0.weight", 2 * ffi, D);
            l.ff0_b = add(f + "0.bias", 1, 2 * ffi);
            l.ff2_w = add(f + "2.weight", D, ffi, pc_ff_inner_pad);
            l.ff2_b = add(f + "2.bias", 1, D);
        }
        pc_gamma = add(pc + "norm.gamma", 1, D);
        total = (total + 31) & ~uint64_t(31);
    }
};

}  // namespace gv
