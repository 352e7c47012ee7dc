"""ORACLE support — imports the REAL reference modules from ``/root/reference``.

Only usable in the build container (the reference tree does not exist on the GPU
box); used by ``tests/golden/make_golden.py`` to generate fixtures and by
``tests/test_oracle_vs_reference.py`` (skipped when the tree is absent) to pin
``oracle/genvc_oracle.py`` against the reference's own code.  Never imported by
product code, ``-m gpu`` tests, ``smoke()`` or ``bench.py``.

Import recipe: ``torchmetrics`` (training-only metric,
``layers/gpt.py:12, 166-172``) and the reference's top-level ``utils`` module
(drags in librosa/matplotlib; only ``get_mask_from_lengths`` ``utils.py:16-24`` is
needed, and only with ``seq_lens``, never at inference) are stubbed.

What cannot be imported here and is therefore driven by a restated loop:
``layers/stream_generator.py`` (needs ``BeamSearchScorer``, gone in
transformers 5.x) and ``GenerationMixin.generate`` on ``GPT2InferenceModel``
(no longer inherited since transformers 4.50).  ``ref_generate`` below follows
``stream_generator.py:809-881`` using the reference's own
``prepare_inputs_for_generation`` / ``forward`` and HF's own processor classes.
"""
from __future__ import annotations

import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("GENVC_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "layers", "gpt.py"))


_installed = False


def install():
    global _installed
    if _installed:
        return
    import transformers  # noqa: F401  (must be imported before the stubs)

    tm = types.ModuleType("torchmetrics")
    tmc = types.ModuleType("torchmetrics.classification")

    class MulticlassAccuracy(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

        def forward(self, *a, **k):
            return torch.tensor(0.0)

    tmc.MulticlassAccuracy = MulticlassAccuracy
    tm.classification = tmc
    sys.modules.setdefault("torchmetrics", tm)
    sys.modules.setdefault("torchmetrics.classification", tmc)

    if "utils" not in sys.modules:
        u = types.ModuleType("utils")

        def get_mask_from_lengths(lengths, max_len=None):
            if max_len is None:
                max_len = int(lengths.max())
            ids = torch.arange(0, max_len, device=lengths.device)
            return ids < lengths.unsqueeze(1)

        u.get_mask_from_lengths = get_mask_from_lengths
        sys.modules["utils"] = u
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    _installed = True


def build_reference_gpt(ckpt: dict, attn_impl: str = "eager"):
    """Reference ``GPT`` built the way ``HiFiGANTrainer.__init__`` does
    (``trainers/hifigan_trainer.py:31-45``), loaded from ``ckpt["model"]`` with strict=True."""
    install()
    from layers.gpt import GPT

    ma = ckpt["config"]["model_args"]
    g = GPT(
        layers=ma["gpt_layers"],
        model_dim=ma["gpt_n_model_channels"],
        start_text_token=ma["gpt_start_text_token"],
        stop_text_token=ma["gpt_stop_text_token"],
        heads=ma["gpt_n_heads"],
        max_text_tokens=ma["gpt_max_text_tokens"],
        max_mel_tokens=ma["gpt_max_audio_tokens"],
        max_prompt_tokens=ma["gpt_max_prompt_tokens"],
        number_text_tokens=ma["gpt_number_text_tokens"],
        num_audio_tokens=ma["gpt_num_audio_tokens"],
        start_audio_token=ma["gpt_start_audio_token"],
        stop_audio_token=ma["gpt_stop_audio_token"],
        code_stride_len=ma["gpt_code_stride_len"],
    ).eval()
    sd = {k[len("gpt."):]: v for k, v in ckpt["model"].items() if k.startswith("gpt.")}
    g.load_state_dict(sd, strict=True)
    g.gpt.config._attn_implementation = attn_impl
    g.init_gpt_for_inference()
    g.gpt_inference.config._attn_implementation = attn_impl
    return g


@torch.no_grad()
def ref_generate(g, cond_latents, text_inputs, top_k, top_p, temperature, repetition_penalty,
                 generator=None, noise=None, max_new_tokens=None, forced_ids=None, trace=None):
    """The reference hot loop with the reference's own modules.  Returns (ids, latents)."""
    from transformers.generation.logits_process import (
        LogitsProcessorList,
        RepetitionPenaltyLogitsProcessor,
        TemperatureLogitsWarper,
        TopKLogitsWarper,
        TopPLogitsWarper,
    )

    input_ids = g.compute_embeddings(cond_latents, text_inputs)
    max_length = g.max_gen_mel_tokens + input_ids.shape[-1]
    if max_new_tokens is not None:
        max_length = min(max_length, input_ids.shape[-1] + max_new_tokens)
    procs = LogitsProcessorList()
    if repetition_penalty != 1.0:
        procs.append(RepetitionPenaltyLogitsProcessor(penalty=repetition_penalty))
    warp = LogitsProcessorList()
    if temperature != 1.0:
        warp.append(TemperatureLogitsWarper(temperature))
    if top_k:
        warp.append(TopKLogitsWarper(top_k=top_k, min_tokens_to_keep=1))
    if top_p < 1.0:
        warp.append(TopPLogitsWarper(top_p=top_p, min_tokens_to_keep=1))
    eos = g.stop_audio_token
    unfinished = torch.ones(input_ids.shape[0], dtype=torch.long)
    past = None
    attention_mask = torch.ones_like(input_ids)
    toks, lats = [], []
    step = 0
    while True:
        inp = g.gpt_inference.prepare_inputs_for_generation(
            input_ids, past_key_values=past, attention_mask=attention_mask, use_cache=True
        )
        out = g.gpt_inference(**inp, return_dict=True, output_hidden_states=True)
        logits = out.logits[:, -1, :]
        scores = warp(input_ids, procs(input_ids, logits))
        probs = torch.softmax(scores, dim=-1)
        if noise is not None:
            nxt = torch.argmax(probs / noise[step], dim=-1)
        else:
            nxt = torch.multinomial(probs, num_samples=1, generator=generator).squeeze(1)
        if forced_ids is not None:
            nxt = forced_ids[:, step]
        nxt = nxt * unfinished + eos * (1 - unfinished)
        latent = g.final_norm(out.hidden_states[-1][:, -1])
        if trace is not None:
            trace.setdefault("logits", []).append(logits.clone())
            trace.setdefault("scores", []).append(scores.clone())
        toks.append(nxt)
        lats.append(latent)
        input_ids = torch.cat([input_ids, nxt[:, None]], dim=-1)
        attention_mask = torch.cat([attention_mask, attention_mask.new_ones((attention_mask.shape[0], 1))], dim=-1)
        past = out.past_key_values
        unfinished = unfinished.mul((nxt != eos).long())
        step += 1
        if unfinished.max() == 0 or input_ids.shape[-1] >= max_length:
            break
    return torch.stack(toks, 1), torch.stack(lats, 1)
