"""Drop-in replacement for the reference ``layers.gpt.GPT`` on its inference path.

Same method names, argument meaning, tensor shapes/dtypes and generator protocol as the
reference object reachable as ``model.gpt``:

* ``get_style_emb``            layers/gpt.py:351-373  -> perceiver kernels
* ``compute_embeddings``       layers/gpt.py:572-592
* ``generate``                 layers/gpt.py:594-609  (HF ``generate``/``sample``)
* ``get_generator``            layers/gpt.py:612-621  (layers/stream_generator.py ``sample_stream``)
* ``__call__(..., return_latent=True)``  layers/gpt.py:375-508
* ``init_gpt_for_inference``, ``eval``, ``to`` and the token-id attributes

All compute runs in ``libgenvc_b200.so``; there is no PyTorch fallback.  Training-only
branches of the reference (losses, masked perceiver, prompts) raise ``NotImplementedError``.
"""
from __future__ import annotations

import math
from typing import Dict, Iterator, Optional, Tuple

import torch

from .config import GenVCDims
from .engine import DecodeChunk, Engine, Sampling

_SUPPORTED_KW = {
    "do_sample", "top_p", "top_k", "temperature", "num_beams", "length_penalty", "repetition_penalty",
    "output_attentions", "output_hidden_states", "num_return_sequences", "return_dict_in_generate",
    # extensions (not HF): reproducible-noise / teacher-forcing / bench hooks
    "exp_noise", "forced_ids", "max_new_tokens", "ignore_eos", "decode_mode", "seed", "stream_chunk_size", "run_ahead",
}


class GPT:
    def __init__(self, dims: GenVCDims, device="cuda", max_batch: int = 1, max_mel_frames: int = 576,
                 stream_chunk_size: int = 8):
        self.dims = dims
        # attributes the reference exposes and its callers read
        self.layers = dims.n_layer
        self.model_dim = dims.d_model
        self.heads = dims.n_head
        self.start_text_token = dims.start_text
        self.stop_text_token = dims.stop_text
        self.start_audio_token = dims.start_audio
        self.stop_audio_token = dims.stop_audio
        self.number_text_tokens = dims.n_text_vocab
        self.num_audio_tokens = dims.n_audio_vocab
        self.max_gen_mel_tokens = dims.max_gen_mel_tokens
        self.max_mel_tokens = dims.max_audio_tokens + 2  # layers/gpt.py:132 (max_conditioning_inputs = 1)
        self.max_text_tokens = dims.max_text_tokens + 2
        self.max_prompt_tokens = dims.max_prompt_tokens
        self.code_stride_len = dims.code_stride_len
        self.training = False
        self.stream_chunk_size = int(stream_chunk_size)
        self._max_batch = int(max_batch)
        self._max_mel_frames = int(max_mel_frames)
        self._engine: Optional[Engine] = None
        self._state_dict: Optional[Dict[str, torch.Tensor]] = None
        self._blob: Optional[torch.Tensor] = None
        self._device = torch.device(device)
        self._prefix: Optional[torch.Tensor] = None  # the reference's gpt_inference.cached_prefix_emb
        self.last_latents: Optional[torch.Tensor] = None  # per-step latents of the last generate() call
        self._gen_stream: Optional[torch.cuda.Stream] = None  # stream of the streaming generator's device loop

    # ------------------------------------------------------------------ nn.Module-like surface
    def load_state_dict(self, state_dict: Dict[str, torch.Tensor], strict: bool = False):
        """Accepts the checkpoint's flat state dict (keys under ``gpt.``), as
        ``inference/model_init.py:22`` passes it; tensors outside the path are ignored.  ``strict=False`` (the
        reference's call) zero-fills ``gpt.*`` tensors the checkpoint lacks; ``strict=True`` raises on them."""
        self._state_dict = state_dict
        self._strict = bool(strict)
        if self._engine is not None:
            self._engine.load_state_dict(state_dict, strict=self._strict)
        return self

    def load_blob(self, blob: torch.Tensor):
        """Packed weight blob (e.g. received by an NCCL broadcast from rank 0)."""
        self._blob = blob
        if self._engine is not None:
            self._engine.load_blob(blob)
        return self

    def eval(self):
        self.training = False
        return self

    def train(self, mode: bool = True):
        if mode:
            raise NotImplementedError("genvc_b200.GPT implements the inference path only")
        return self

    def to(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError(f"genvc_b200.GPT runs on CUDA devices only (no CPU fallback); got '{device}'")
        if self._engine is not None and self._engine.device != torch.device("cuda", device.index or 0):
            self._engine = None
        self._device = device
        return self

    def init_gpt_for_inference(self, kv_cache: bool = True, use_deepspeed: bool = False):
        """layers/gpt.py:197-230: here this is where the device engine is built and the weights uploaded."""
        if use_deepspeed:
            raise NotImplementedError("DeepSpeed kernel injection is not part of this path")
        if not kv_cache:
            raise NotImplementedError("the KV cache cannot be disabled")
        if self._engine is None:
            self._engine = Engine(self.dims, self._device, max_batch=self._max_batch, max_mel_frames=self._max_mel_frames)
            if self._blob is not None:
                self._engine.load_blob(self._blob)
            elif self._state_dict is not None:
                self._engine.load_state_dict(self._state_dict, strict=getattr(self, "_strict", False))
        return self

    @property
    def engine(self) -> Engine:
        if self._engine is None:
            self.init_gpt_for_inference()
        if self._engine.blob is None:
            raise RuntimeError("no weights loaded: call load_state_dict() first")
        return self._engine

    @property
    def device(self) -> torch.device:
        return self._engine.device if self._engine is not None else self._device

    # ------------------------------------------------------------------ a2: perceiver
    def get_style_emb(self, cond_input: torch.Tensor, return_latent: bool = False, seq_lens=None) -> torch.Tensor:
        """cond_input (b, 80, s) or (b, 1, 80, s) -> (b, D, 32)   (layers/gpt.py:351-373)."""
        if return_latent:
            return cond_input.unsqueeze(1)
        if seq_lens is not None:
            raise NotImplementedError("masked perceiver (seq_lens) is a training-only branch")
        if cond_input.ndim == 4:
            cond_input = cond_input.squeeze(1)
        return self.engine.perceiver(cond_input).transpose(1, 2)

    # ------------------------------------------------------------------ a4: prefix
    def compute_embeddings(self, cond_latents: torch.Tensor, text_inputs: torch.Tensor) -> torch.Tensor:
        """Stores the prefix embeddings [B, P, D] and returns the fake ids ``[1]*P + [start_audio]``
        (layers/gpt.py:572-592)."""
        eng = self.engine
        self._prefix = eng.embed_prefix(cond_latents, text_inputs)
        B, P, _ = self._prefix.shape
        gpt_inputs = torch.full((B, P + 1), 1, dtype=torch.long, device=eng.device)
        gpt_inputs[:, -1] = self.start_audio_token
        return gpt_inputs

    # ------------------------------------------------------------------ sampling knobs
    def _sampling(self, kw: dict) -> Tuple[Sampling, dict]:
        unknown = set(kw) - _SUPPORTED_KW
        if unknown:
            raise TypeError(f"unsupported generate() arguments: {sorted(unknown)}")
        if kw.get("num_beams", 1) not in (None, 1):
            raise NotImplementedError("beam search is not part of the GenVC path (num_beams must be 1)")
        if kw.get("num_return_sequences", 1) not in (None, 1):
            raise NotImplementedError("num_return_sequences must be 1")
        # HF GenerationConfig defaults: do_sample=False (greedy) -- a bare ``generate(cond, text)`` like the reference's
        # own ``self.generate(cond_latent_cv, text_input_cv)`` is greedy; the inference drivers pass do_sample=True
        do_sample = kw.get("do_sample", False)
        top_k = kw.get("top_k", 50)
        top_p = kw.get("top_p", 1.0)
        temperature = kw.get("temperature", 1.0)
        if not do_sample:
            # greedy search: processors (repetition penalty) apply, warpers do not; argmax
            top_k, top_p, temperature = 1, 1.0, 1.0
        seed = kw.get("seed")
        if seed is None:
            # tie the on-device Philox stream to torch's global generator so torch.manual_seed() reproduces
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        sp = Sampling(
            top_k=int(top_k or 0), top_p=1.0 if top_p is None else float(top_p),
            temperature=1.0 if temperature is None else float(temperature),
            repetition_penalty=1.0 if kw.get("repetition_penalty") is None else float(kw["repetition_penalty"]),
            ignore_eos=bool(kw.get("ignore_eos", False)), max_new_tokens=int(kw.get("max_new_tokens") or 0), seed=seed,
        )
        return sp, kw

    def _cap(self, sp: Sampling) -> int:
        cap = self.max_gen_mel_tokens
        return min(cap, sp.max_new_tokens) if sp.max_new_tokens > 0 else cap

    # ------------------------------------------------------------------ a8: generate
    @torch.no_grad()
    def generate(self, cond_latents: torch.Tensor, text_inputs: torch.Tensor, **generate_kwargs) -> torch.Tensor:
        """Returns the new token ids [B, n] (int64): EOS included if emitted, finished rows padded
        with ``stop_audio_token`` (layers/gpt.py:594-609 + HF ``sample``)."""
        sp, kw = self._sampling(generate_kwargs)
        eng = self.engine
        noise, forced = kw.get("exp_noise"), kw.get("forced_ids")
        mode = int(kw.get("decode_mode", 0))
        B = int(text_inputs.shape[0])
        self.compute_embeddings(cond_latents, text_inputs)
        eng.prefill(self._prefix)
        cap = self._cap(sp)
        fused = mode == 2 or (mode == 0 and eng.fused_rows(B))
        # the fused kernels (one row: decode_mega; 2..8 rows: decode_batch) run the whole loop in one launch; the per-op
        # path is enqueued in slices so an early EOS does not leave hundreds of skipped launches behind
        step = cap if fused else 32
        ids, lats, done, n = [], [], False, 0
        while not done and n < cap:
            k = min(step, cap - n)
            ch = eng.decode(k, sp, None if noise is None else noise[n:n + k], None if forced is None else forced[n:n + k],
                            mode=mode)
            emitted, done_flag, bad = ch.status.tolist()[:3]  # host sync
            if bad:
                raise IndexError("a text or forced token id was outside its vocabulary (clamped on the device)")
            ids.append(ch.ids[:emitted])
            lats.append(ch.latents[:emitted])
            n += emitted
            done = bool(done_flag) or emitted < k
        out = torch.cat(ids, 0).transpose(0, 1).contiguous()
        self.last_latents = torch.cat(lats, 0).transpose(0, 1).contiguous()
        return out

    def inference(self, cond_latents, text_inputs, **generate_kwargs):
        return self.generate(cond_latents, text_inputs, **generate_kwargs)

    # ------------------------------------------------------------------ a9: streaming generator
    @torch.no_grad()
    def get_generator(self, fake_inputs: torch.Tensor, **generate_kwargs) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        """Generator of ``(next_tokens [B] int64, latent [B, D] fp32)`` per step; the EOS step's pair IS
        delivered before the generator stops (layers/stream_generator.py:865 vs :873-881).

        The device loop runs ``stream_chunk_size`` steps per launch and one launch ahead of the consumer;
        the host synchronises once per chunk (the cadence at which the reference flushes latents to the
        vocoder, inference/inference_utils.py:195), not once per token."""
        if self._prefix is None:
            raise RuntimeError("compute_embeddings() must be called before get_generator()")
        if fake_inputs.shape[-1] != self._prefix.shape[1] + 1 or fake_inputs.shape[0] != self._prefix.shape[0]:
            raise ValueError("fake_inputs does not match the stored prefix embeddings")
        sp, kw = self._sampling(generate_kwargs)
        eng = self.engine
        # The generation loop runs on its own stream (the caller's stream stays free for what it does with a delivered
        # chunk).  By default the next chunk is launched when the consumer ASKS for its first token, not ahead of time:
        # the persistent decode kernel occupies every SM (216 KB of shared memory and the whole register file per SM), so
        # a run-ahead launch would delay the consumer's own GPU work on the delivered chunk (device-to-host copy kernels,
        # the vocoder) by a whole chunk of decoding — measured 12.5 ms vs 7.8 ms to the first 8 tokens on the host.
        # `run_ahead=True` restores the overlap for pure throughput runs.
        caller = torch.cuda.current_stream(eng.device)
        if self._gen_stream is None:
            self._gen_stream = torch.cuda.Stream(device=eng.device)
        self._gen_stream.wait_stream(caller)  # the prefix embeddings (and any earlier engine call) come first
        with torch.cuda.stream(self._gen_stream):
            eng.prefill(self._prefix)
        return self._stream(eng, sp, kw, caller)

    def _stream(self, eng: Engine, sp: Sampling, kw: dict, caller: "torch.cuda.Stream"):
        cap = self._cap(sp)
        chunk = int(kw.get("stream_chunk_size") or self.stream_chunk_size)
        noise, forced = kw.get("exp_noise"), kw.get("forced_ids")
        mode = int(kw.get("decode_mode", 0))
        gs = self._gen_stream

        def launch(n0: int) -> Optional[Tuple[DecodeChunk, torch.Tensor, torch.cuda.Event]]:
            if n0 >= cap:
                return None
            k = min(chunk, cap - n0)
            with torch.cuda.stream(gs):
                ch = eng.decode(k, sp, None if noise is None else noise[n0:n0 + k],
                                None if forced is None else forced[n0:n0 + k], mode=mode)
                host_status = torch.empty(4, dtype=torch.int32, pin_memory=True)
                host_status.copy_(ch.status, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(gs)
            return ch, host_status, ev

        run_ahead = bool(kw.get("run_ahead", False))
        try:
            n = 0
            cur = launch(0)
            while cur is not None:
                k = cur[0].ids.shape[0]
                nxt = launch(n + k) if run_ahead else None  # (a finished device loop turns a run-ahead launch into a no-op)
                ch, host_status, ev = cur
                ev.synchronize()  # the chunk is complete: its tensors are safe on any stream from here on
                for t in (ch.ids, ch.latents):
                    t.record_stream(caller)
                emitted, done = int(host_status[0]), int(host_status[1])
                if int(host_status[2]):
                    raise IndexError("a text or forced token id was outside its vocabulary (clamped on the device)")
                for i in range(emitted):
                    yield ch.ids[i], ch.latents[i]
                n += emitted
                if done or emitted < k:
                    return
                cur = nxt if run_ahead else launch(n)
        finally:
            # whatever the caller enqueues next on its stream (e.g. the next segment's engine calls) comes after the
            # generation stream's remaining work (a run-ahead launch that found the loop finished)
            caller.wait_stream(gs)

    # ------------------------------------------------------------------ a11: teacher-forced latent pass
    @torch.no_grad()
    def __call__(self, text_inputs, text_lengths, audio_codes, wav_lengths, cond_mels=None, cond_lens=None,
                 cond_latents=None, return_attentions=False, return_latent=False):
        return self.forward(text_inputs, text_lengths, audio_codes, wav_lengths, cond_mels, cond_lens, cond_latents,
                            return_attentions, return_latent)

    def forward(self, text_inputs, text_lengths, audio_codes, wav_lengths, cond_mels=None, cond_lens=None,
                cond_latents=None, return_attentions=False, return_latent=False):
        """``GPT.forward(..., return_latent=True)`` (layers/gpt.py:375-508): latents [B, M', D] with
        M' = max(ceil(wav_lengths / code_stride_len))."""
        if not return_latent or return_attentions:
            raise NotImplementedError("only the return_latent=True inference branch is implemented")
        if cond_latents is None:
            if cond_mels is None:
                raise ValueError("cond_latents or cond_mels is required")
            if cond_lens is not None:
                raise NotImplementedError("masked perceiver (cond_lens) is a training-only branch")
            cond_latents = self.get_style_emb(cond_mels).transpose(1, 2)
        text_lengths = torch.as_tensor(text_lengths).reshape(-1).cpu()
        wav_lengths = torch.as_tensor(wav_lengths).reshape(-1).cpu()
        B = text_inputs.shape[0]
        if text_lengths.numel() != B or wav_lengths.numel() != B or audio_codes.shape[0] != B:
            raise ValueError("batch mismatch between inputs and lengths")
        max_text_len = int(text_lengths.max())
        code_lens = torch.ceil(wav_lengths.to(torch.float64) / self.code_stride_len).long()  # without start/stop
        M = int(code_lens.max())
        assert max_text_len <= text_inputs.shape[-1], \
            f" max_text_len ({max_text_len}) > text_inputs.shape[-1] ({text_inputs.shape[-1]})"
        if M <= 0:
            raise ValueError("wav_lengths imply no audio codes")
        text = text_inputs[:, :max_text_len].clone()
        codes = audio_codes[:, :M].clone()
        if codes.shape[1] < M:  # the reference zero-pads, then overwrites the padding with stop tokens
            codes = torch.nn.functional.pad(codes, (0, M - codes.shape[1]), value=self.stop_audio_token)
        for b in range(B):  # set_text_padding / set_mel_padding
            if int(text_lengths[b]) < max_text_len:
                text[b, int(text_lengths[b]):] = self.stop_text_token
            if int(code_lens[b]) < M:
                codes[b, int(code_lens[b]):] = self.stop_audio_token
        return self.engine.forward_latents(cond_latents, text, codes)
