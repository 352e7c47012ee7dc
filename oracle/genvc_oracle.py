"""ORACLE — test infrastructure, not product code.

CPU (PyTorch fp32) restatement of GenVC's autoregressive codec-token inference
path, used ONLY by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` as the checker and
the timed CPU arm.  Nothing under ``genvc_b200/`` imports it; the product path
fails loudly when the CUDA library is missing.

Parity status: the reference ships no tests, golden vectors or fixtures for this
path — **"parity unpinned" by the reference's own tests**.
The pin used instead: this restatement is checked against the reference's own
modules (``layers/gpt.py``, ``layers/gpt_inference.py``,
``layers/perceiver_encoder.py``) imported from ``/root/reference`` in the build
container (``tests/test_oracle_vs_reference.py``; skipped where the reference is
absent) and against fixtures those modules generated
(``tests/golden/*.pt`` via ``tests/golden/make_golden.py``).

Third-party arithmetic that is not under ``/root/reference``: ``transformers``
(reference pins ``==4.33.0``, ``README.md:43``) — ``GPT2Model`` / ``GPT2Block`` /
``GPT2Attention`` / ``GPT2MLP`` / ``Conv1D`` / ``NewGELUActivation`` and
``GenerationMixin.sample`` with ``RepetitionPenaltyLogitsProcessor``,
``TemperatureLogitsWarper``, ``TopKLogitsWarper``, ``TopPLogitsWarper``.  Their
published algorithms are restated below; the call sites that anchor them are
``layers/gpt.py:54-66, 199-217, 290-295, 601-608``,
``layers/gpt_inference.py:97-112`` and ``layers/stream_generator.py:769-881``.

Every function cites the reference lines it follows.  Weights are taken from a
checkpoint ``state_dict`` in the reference layout (keys under ``gpt.``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Iterator, List, Optional, Tuple

import torch
import torch.nn.functional as F

PFX = "gpt."


@dataclass
class OracleDims:
    n_layer: int
    d_model: int
    n_head: int
    n_text_vocab: int = 258
    n_audio_vocab: int = 1026
    start_text: int = 256
    stop_text: int = 257
    start_audio: int = 1024
    stop_audio: int = 1025
    max_audio_tokens: int = 605
    pc_depth: int = 4
    pc_heads: int = 8
    pc_dim_head: int = 64

    @property
    def max_gen_mel_tokens(self) -> int:  # layers/gpt.py:131
        return self.max_audio_tokens - 1 - 2

    @staticmethod
    def from_config(cfg: dict) -> "OracleDims":
        ma = cfg.get("model_args", {})
        g = lambda k, dflt: dflt if ma.get(k) is None else ma[k]
        return OracleDims(
            n_layer=g("gpt_layers", 30),
            d_model=g("gpt_n_model_channels", 1024),
            n_head=g("gpt_n_heads", 16),
            n_text_vocab=g("gpt_number_text_tokens", 258),
            n_audio_vocab=g("gpt_num_audio_tokens", 1026),
            start_text=g("gpt_start_text_token", 256),
            stop_text=g("gpt_stop_text_token", 257),
            start_audio=g("gpt_start_audio_token", 1024),
            stop_audio=g("gpt_stop_audio_token", 1025),
            max_audio_tokens=g("gpt_max_audio_tokens", 605),
        )


@dataclass
class SamplingParams:
    """The HF ``generate`` kwargs the reference passes
    (``inference/inference_utils.py:55-66, 170-182``)."""

    top_k: int = 15
    top_p: float = 0.85
    temperature: float = 0.85
    repetition_penalty: float = 2.0


# ----------------------------------------------------------------------------------------
# element-wise pieces
# ----------------------------------------------------------------------------------------
def gelu_new(x: torch.Tensor) -> torch.Tensor:
    """HF ``NewGELUActivation`` (transformers ``activations.py``; GPT-2's ``gelu_new``)."""
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


def conv1d(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """HF ``Conv1D.forward``: ``addmm(bias, x.view(-1, in), weight)`` with weight ``[in, out]``."""
    out_shape = x.shape[:-1] + (w.shape[1],)
    return torch.addmm(b, x.reshape(-1, x.shape[-1]), w).view(out_shape)


# ----------------------------------------------------------------------------------------
# sampling chain (A4)
# ----------------------------------------------------------------------------------------
def process_logits(input_ids: torch.Tensor, logits: torch.Tensor, sp: SamplingParams) -> torch.Tensor:
    """Processor + warper chain in HF-4.33 order, as wired by
    ``layers/stream_generator.py:333-344`` (``_get_logits_processor``) and ``:412-414``
    (``_get_logits_warper``), applied at ``:837-838``.

    ``input_ids`` is the WHOLE row so far — fake prefix ids ``[1]*P + [1024]`` included
    (``layers/gpt.py:582-592``) — so ids 1 and 1024 are always penalised.
    """
    scores = logits.clone()
    # RepetitionPenaltyLogitsProcessor: gather / where / scatter (once per distinct id)
    if sp.repetition_penalty != 1.0:
        score = torch.gather(scores, 1, input_ids)
        score = torch.where(score < 0, score * sp.repetition_penalty, score / sp.repetition_penalty)
        scores = scores.scatter(1, input_ids, score)
    # TemperatureLogitsWarper
    if sp.temperature != 1.0:
        scores = scores / sp.temperature
    # TopKLogitsWarper: remove everything strictly below the k-th largest (ties survive)
    if sp.top_k is not None and sp.top_k != 0:
        k = min(int(sp.top_k), scores.size(-1))
        kth = torch.topk(scores, k)[0][..., -1, None]
        scores = scores.masked_fill(scores < kth, -float("inf"))
    # TopPLogitsWarper: ascending sort, drop the low tail whose cumulative mass <= 1 - top_p
    if sp.top_p is not None and sp.top_p < 1.0:
        sorted_logits, sorted_indices = torch.sort(scores, descending=False)
        cumulative_probs = sorted_logits.softmax(dim=-1).cumsum(dim=-1)
        sorted_remove = cumulative_probs <= (1 - sp.top_p)
        sorted_remove[..., -1:] = 0  # min_tokens_to_keep = 1
        remove = sorted_remove.scatter(1, sorted_indices, sorted_remove)
        scores = scores.masked_fill(remove, -float("inf"))
    return scores


def draw_exponential_noise(shape, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """The noise ``torch.multinomial(probs, 1)`` consumes on CPU: ``q ~ Exp(1)`` of the
    shape of ``probs``; the sample is ``argmax(probs / q)`` (ATen ``multinomial``,
    n_sample == 1 branch).  Drawing it explicitly lets the CUDA path be fed the same
    noise and reproduce the CPU reference's sampled ids (``stream_generator.py:857-858``)."""
    return torch.empty(shape, dtype=torch.float32).exponential_(1, generator=generator)


def sample_from_scores(scores: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
    probs = F.softmax(scores, dim=-1)
    return torch.argmax(probs / noise, dim=-1)


# ----------------------------------------------------------------------------------------
# the model
# ----------------------------------------------------------------------------------------
@dataclass
class KVCache:
    k: List[Optional[torch.Tensor]] = field(default_factory=list)  # per layer [B,H,S,hd]
    v: List[Optional[torch.Tensor]] = field(default_factory=list)

    @property
    def length(self) -> int:
        return 0 if not self.k or self.k[0] is None else self.k[0].shape[2]


class GenVCOracle:
    def __init__(self, state_dict: Dict[str, torch.Tensor], dims: OracleDims):
        self.d = dims
        self.w = {k[len(PFX):]: v.detach().to(torch.float32) for k, v in state_dict.items() if k.startswith(PFX)}

    # ---------------- A5: perceiver (layers/perceiver_encoder.py:265-276, 305-319, 108-151) -----
    def perceiver(self, mel: torch.Tensor) -> torch.Tensor:
        """mel [B, 80, S] -> speaker latents [B, 32, D]  (mask=None, as at inference)."""
        w, d = self.w, self.d
        pc = "conditioning_perceiver."
        x = mel.permute(0, 2, 1)  # layers/gpt.py:369
        x = F.linear(x, w[pc + "proj_context.weight"], w[pc + "proj_context.bias"])
        B = x.shape[0]
        lat = w[pc + "latents"].unsqueeze(0).expand(B, -1, -1)
        H, hd = d.pc_heads, d.pc_dim_head
        for i in range(d.pc_depth):
            a = f"{pc}layers.{i}.0."
            f = f"{pc}layers.{i}.1."
            ctx = torch.cat((lat, x), dim=-2)  # latents FIRST (perceiver_encoder.py:310-311)
            q = F.linear(lat, w[a + "to_q.weight"])
            kv = F.linear(ctx, w[a + "to_kv.weight"])
            k, v = kv.chunk(2, dim=-1)
            sh = lambda t: t.reshape(B, t.shape[1], H, hd).permute(0, 2, 1, 3)
            q, k, v = sh(q), sh(k), sh(v)
            sim = torch.einsum("bhid,bhjd->bhij", q, k) * (hd ** -0.5)
            attn = sim.softmax(dim=-1)
            out = torch.einsum("bhij,bhjd->bhid", attn, v)
            out = out.permute(0, 2, 1, 3).reshape(B, -1, H * hd)
            lat = F.linear(out, w[a + "to_out.weight"]) + lat
            h = F.linear(lat, w[f + "0.weight"], w[f + "0.bias"])
            xx, gate = h.chunk(2, dim=-1)  # GEGLU: second half is the gate (:206-208)
            h = F.gelu(gate) * xx
            lat = F.linear(h, w[f + "2.weight"], w[f + "2.bias"]) + lat
        # RMSNorm (:177-179): F.normalize(x, dim=-1) * sqrt(D) * gamma
        return F.normalize(lat, dim=-1) * (d.d_model ** 0.5) * w[pc + "norm.gamma"]

    def get_style_emb(self, cond_input: torch.Tensor) -> torch.Tensor:
        """layers/gpt.py:351-373 (return_latent=False, seq_lens=None): -> [B, D, 32]."""
        if cond_input.ndim == 4:
            cond_input = cond_input.squeeze(1)
        return self.perceiver(cond_input).transpose(1, 2)

    def get_gpt_cond_latents(self, mel_chunks: List[torch.Tensor]) -> torch.Tensor:
        """trainers/hifigan_trainer.py:438-455 after the mel front-end: per-chunk style
        embedding, mean over chunks, transpose -> [B, 32, D]."""
        embs = [self.get_style_emb(m) for m in mel_chunks]
        return torch.stack(embs).mean(dim=0).transpose(1, 2)

    # ---------------- A1: embeddings (layers/gpt.py:572-592) -------------------------------------
    def prefix_embeddings(self, cond_latents: torch.Tensor, text_inputs: torch.Tensor) -> torch.Tensor:
        w, d = self.w, self.d
        t = F.pad(text_inputs, (0, 1), value=d.stop_text)
        t = F.pad(t, (1, 0), value=d.start_text)
        emb = w["text_embedding.weight"][t] + w["text_pos_embedding.emb.weight"][: t.shape[1]]
        return torch.cat([cond_latents, emb], dim=1)

    def fake_inputs(self, prefix: torch.Tensor) -> torch.Tensor:
        ids = torch.full((prefix.shape[0], prefix.shape[1] + 1), 1, dtype=torch.long)
        ids[:, -1] = self.d.start_audio
        return ids

    def mel_token_embedding(self, tok: torch.Tensor, pos: int) -> torch.Tensor:
        """layers/gpt_inference.py:92-96: mel_embedding[tok] + mel_pos_embedding[pos]; tok [B] -> [B,1,D]."""
        w = self.w
        return (w["mel_embedding.weight"][tok] + w["mel_pos_embedding.emb.weight"][pos]).unsqueeze(1)

    # ---------------- A2: GPT-2 blocks (HF modeling_gpt2.py, eager attention of 4.33) -------------
    def forward_rows(self, emb: torch.Tensor, cache: Optional[KVCache]) -> torch.Tensor:
        """emb [B,M,D] -> ln_f(hidden) [B,M,D]; appends K/V to ``cache`` when given."""
        w, d = self.w, self.d
        B, M, D = emb.shape
        H, hd = d.n_head, d.d_model // d.n_head
        x = emb
        past = cache.length if cache is not None else 0
        if cache is not None and not cache.k:
            cache.k = [None] * d.n_layer
            cache.v = [None] * d.n_layer
        for i in range(d.n_layer):
            p = f"gpt.h.{i}."
            a = F.layer_norm(x, (D,), w[p + "ln_1.weight"], w[p + "ln_1.bias"], 1e-5)
            qkv = conv1d(a, w[p + "attn.c_attn.weight"], w[p + "attn.c_attn.bias"])
            q, k, v = qkv.split(D, dim=2)
            sh = lambda t: t.view(B, M, H, hd).permute(0, 2, 1, 3)
            q, k, v = sh(q), sh(k), sh(v)
            if cache is not None:
                if cache.k[i] is not None:
                    k = torch.cat((cache.k[i], k), dim=-2)
                    v = torch.cat((cache.v[i], v), dim=-2)
                cache.k[i], cache.v[i] = k, v
            S = k.shape[-2]
            att = torch.matmul(q, k.transpose(-1, -2))
            att = att / torch.full([], hd ** 0.5, dtype=att.dtype)
            causal = torch.tril(torch.ones((S, S), dtype=torch.bool))[S - M : S, :S]
            att = torch.where(causal, att, torch.full([], torch.finfo(att.dtype).min, dtype=att.dtype))
            att = F.softmax(att, dim=-1)
            o = torch.matmul(att, v).permute(0, 2, 1, 3).reshape(B, M, D)
            x = x + conv1d(o, w[p + "attn.c_proj.weight"], w[p + "attn.c_proj.bias"])
            m = F.layer_norm(x, (D,), w[p + "ln_2.weight"], w[p + "ln_2.bias"], 1e-5)
            u = gelu_new(conv1d(m, w[p + "mlp.c_fc.weight"], w[p + "mlp.c_fc.bias"]))
            x = x + conv1d(u, w[p + "mlp.c_proj.weight"], w[p + "mlp.c_proj.bias"])
        assert past + M == (cache.length if cache is not None else M)
        return F.layer_norm(x, (D,), w["gpt.ln_f.weight"], w["gpt.ln_f.bias"], 1e-5)

    # ---------------- A3: head (layers/gpt_inference.py:18, 111-112) ------------------------------
    def final_norm(self, h: torch.Tensor) -> torch.Tensor:
        w = self.w
        return F.layer_norm(h, (h.shape[-1],), w["final_norm.weight"], w["final_norm.bias"], 1e-5)

    def mel_logits(self, z: torch.Tensor) -> torch.Tensor:
        return F.linear(z, self.w["mel_head.weight"], self.w["mel_head.bias"])

    # ---------------- a8/a9: generation loop (layers/stream_generator.py:769-881) -----------------
    def stream(
        self,
        cond_latents: torch.Tensor,
        text_inputs: torch.Tensor,
        sp: SamplingParams,
        generator: Optional[torch.Generator] = None,
        noise: Optional[torch.Tensor] = None,
        max_new_tokens: Optional[int] = None,
        trace: Optional[dict] = None,
        forced_ids: Optional[torch.Tensor] = None,
        ignore_eos: bool = False,
    ) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        """Yields ``(next_tokens [B] i64, latent [B,D])`` per step exactly as
        ``sample_stream`` does: the yield happens BEFORE the EOS test (``:865`` vs ``:873-881``).

        ``noise`` [steps, B, V] overrides the generator draw (Exp(1) variates);
        ``forced_ids`` [B, n] teacher-forces the emitted tokens (logits parity runs);
        ``ignore_eos`` keeps generating past EOS (bench mode: fixed work per segment).
        """
        d = self.d
        prefix = self.prefix_embeddings(cond_latents, text_inputs)  # gpt.compute_embeddings
        B, P, _ = prefix.shape
        input_ids = self.fake_inputs(prefix)
        max_length = d.max_gen_mel_tokens + input_ids.shape[1]  # layers/gpt.py:606
        if max_new_tokens is not None:
            max_length = min(max_length, input_ids.shape[1] + max_new_tokens)
        cache = KVCache()
        unfinished = torch.ones(B, dtype=torch.long)
        step = 0
        while True:
            if step == 0:  # layers/gpt_inference.py:81-91
                tok = input_ids[:, -1]
                emb = torch.cat([prefix, self.mel_token_embedding(tok, 0)], dim=1)
            else:  # :92-96, position = attention_mask_len - (P + 1) = step
                emb = self.mel_token_embedding(input_ids[:, -1], step)
            h = self.forward_rows(emb, cache)  # post-ln_f hidden states
            z = self.final_norm(h[:, -1])  # the "latent" (stream_generator.py:865)
            logits = self.mel_logits(z)  # lm_head = Sequential(final_norm, mel_head)
            scores = process_logits(input_ids, logits, sp)
            if noise is not None:
                q = noise[step]
            else:
                q = draw_exponential_noise(scores.shape, generator)
            next_tokens = sample_from_scores(scores, q)
            if forced_ids is not None:
                next_tokens = forced_ids[:, step]
            if not ignore_eos:
                next_tokens = next_tokens * unfinished + d.stop_audio * (1 - unfinished)
            if trace is not None:
                trace.setdefault("logits", []).append(logits.clone())
                trace.setdefault("scores", []).append(scores.clone())
            yield next_tokens, z
            input_ids = torch.cat([input_ids, next_tokens[:, None]], dim=-1)
            if not ignore_eos:
                unfinished = unfinished.mul((next_tokens != d.stop_audio).long())
            step += 1
            if unfinished.max() == 0 or input_ids.shape[-1] >= max_length:
                break

    def generate(self, cond_latents, text_inputs, sp: SamplingParams, **kw) -> Tuple[torch.Tensor, torch.Tensor]:
        """``GPT.generate`` (layers/gpt.py:594-609): returns (ids [B, n] i64, latents [B, n, D])."""
        toks, lats = [], []
        for t, z in self.stream(cond_latents, text_inputs, sp, **kw):
            toks.append(t)
            lats.append(z)
        return torch.stack(toks, dim=1), torch.stack(lats, dim=1)

    # ---------------- A6: second (latent) pass (layers/gpt.py:375-508, return_latent=True) --------
    def forward_latents(self, text_inputs: torch.Tensor, audio_codes: torch.Tensor, cond_latents: torch.Tensor) -> torch.Tensor:
        """B == 1 (or equal lengths) inference use: text [B,T], codes [B,M] -> latents [B,M,D].

        With ``wav_lengths = M * code_stride_len`` (``inference_utils.py:69``) the reference pads
        the codes to ``M + 3`` with zeros, overwrites the padding with stop tokens
        (``set_mel_padding``), appends one more stop and prepends start: ``[1024, g.., 1025 x4]``;
        the returned slice drops the last 5 rows (``sub = -5``)."""
        w, d = self.w, self.d
        t = F.pad(text_inputs, (0, 1), value=d.stop_text)
        t = F.pad(t, (1, 0), value=d.start_text)
        text_emb = w["text_embedding.weight"][t] + w["text_pos_embedding.emb.weight"][: t.shape[1]]
        a = F.pad(audio_codes, (0, 4), value=d.stop_audio)
        a = F.pad(a, (1, 0), value=d.start_audio)
        mel_emb = w["mel_embedding.weight"][a] + w["mel_pos_embedding.emb.weight"][: a.shape[1]]
        emb = torch.cat([cond_latents, text_emb, mel_emb], dim=1)
        h = self.forward_rows(emb, None)
        enc = self.final_norm(h[:, cond_latents.shape[1] :])
        mel = enc[:, -a.shape[1] :]
        return mel[:, :-5]


def load_oracle(ckpt: dict) -> GenVCOracle:
    return GenVCOracle(ckpt["model"], OracleDims.from_config(ckpt["config"]))
