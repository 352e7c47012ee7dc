"""Pipeline drivers around the codec-token path, re-hosted with the reference's call contract
(``inference/inference_utils.py:5-217``): ``handle_chunks``, ``synthesize_utt``, ``synthesize_utt_chunked``,
``synthesize_utt_streaming`` — same arguments, same results for the same model object.

What is different from the reference's drivers:

* the segment plan (6 s windows, last window zero-padded up to 0.32 s) is computed once by ``plan_segments``
  and shared by the three drivers;
* ``synthesize_utt(..., reuse_decode_latents=True)`` skips the teacher-forced second pass
  (``inference_utils.py:71-76``): the fused decode kernel already hands out ``final_norm(ln_f(h))`` of every
  generated position, which is what the second pass recomputes (equal up to fp32 rounding, so it is opt-in);
* ``synthesize_utt(..., batch_segments=True)`` decodes the equal-length segments of the utterance together: the segments
  are independent given ``cond_latent`` (no GPT state crosses them, ``inference_utils.py:43-77``), so up to 8 of them
  share one pass of the weight stream per generated token in the batched fused kernel (rows finish independently and are
  padded with the stop token, as ``layers/stream_generator.py:860-881`` does for a batch); greedy results are identical to
  the serial loop row for row, sampled results draw from a different random stream;
* the streaming driver hands every waveform chunk to an optional ``on_chunk`` callback as soon as it exists and
  records first-chunk latency / real-time factor on the model (``last_latency_s``, ``last_rtf``) besides printing
  them like the reference does.

The stages outside the path (``content_extractor``, ``content_dvae``, ``hifigan``, the mel front-end inside
``get_gpt_cond_latents``) are whatever modules are attached to the model.
"""
from __future__ import annotations

import time
from typing import Callable, List, Optional, Tuple

import torch
import torch.nn.functional as F

OVERLAP = 1024  # samples cross-faded between consecutive waveform chunks (inference_utils.py:127, 203)


def plan_segments(total: int, seg: int, min_len: int) -> List[Tuple[int, int, int]]:
    """``(start, end, pad)`` of every source window: ``seg`` samples each, the last one takes the rest and is
    zero-padded to ``min_len`` samples if shorter (inference_utils.py:43-50)."""
    plan = []
    for start in range(0, total, seg):
        if start + seg < total:
            plan.append((start, start + seg, 0))
        else:
            plan.append((start, total, max(0, min_len - (total - start))))
    return plan


def _segment(src_wav: torch.Tensor, start: int, end: int, pad: int) -> torch.Tensor:
    piece = src_wav[:, start:end]
    return F.pad(piece, (0, pad), "constant", 0) if pad else piece


def _sampling_kwargs(cfg) -> dict:
    return dict(top_p=cfg.top_p, top_k=cfg.top_k, temperature=cfg.temperature, length_penalty=cfg.length_penalty,
                repetition_penalty=cfg.repetition_penalty)


def _content_codes(model, wav_seg: torch.Tensor) -> torch.Tensor:
    feat = model.content_extractor.extract_content_features(wav_seg)
    return model.content_dvae.get_codebook_indices(feat.transpose(1, 2))


def _vocode(model, latents: torch.Tensor) -> torch.Tensor:
    """[1, M, D] latents -> waveform: x``hifigan_scale_factor`` linear interpolation, then the vocoder."""
    mel_input = F.interpolate(latents.transpose(1, 2), scale_factor=[model.hifigan_scale_factor], mode="linear").squeeze(1)
    return model.hifigan(mel_input)


@torch.inference_mode()
def handle_chunks(wav_gen, wav_gen_prev, wav_overlap, overlap_len=OVERLAP):
    """Cross-fade consecutive waveform chunks (inference_utils.py:5-21).  Returns
    ``(chunk to emit, wav_gen, tail kept for the next call)``: the emitted chunk is ``wav_gen`` without its last
    ``overlap_len`` samples, its head blended linearly with the previous tail.  A chunk too short to hold the
    blend is emitted as its own tail and drops the pending overlap (the reference's short-chunk branch)."""
    body = wav_gen[:-overlap_len]
    if wav_overlap is not None:
        if overlap_len > len(body):
            return wav_gen[-overlap_len:], wav_gen, None
        ramp = torch.linspace(0.0, 1.0, overlap_len)
        fade_in = body[:overlap_len] * ramp.to(body.device)
        # in place, like the reference: `body` is a view of wav_gen
        body[:overlap_len] = wav_overlap * torch.linspace(1.0, 0.0, overlap_len).to(wav_overlap.device)
        body[:overlap_len] += fade_in
    return body, wav_gen, wav_gen[-overlap_len:]


def _prepare(model, src_wav, tgt_audio, seg_len):
    src_wav = src_wav.to(model.device)
    seg = int(seg_len * model.content_sample_rate)
    min_len = int(0.32 * model.content_sample_rate)
    cond_latent = model.get_gpt_cond_latents(tgt_audio.to(model.device), model.config.audio.sample_rate)
    return src_wav, cond_latent, plan_segments(src_wav.shape[-1], seg, min_len)


def _segment_latents(m, cond_latent, codes, gen_rows, reuse_decode_latents):
    """Latents of the kept (non-stop) tokens of every row of one generate() call -> list of [1, n_b, D]."""
    B = codes.shape[0]
    keeps = [(gen_rows[b] != m.gpt.stop_audio_token).nonzero().squeeze(-1) for b in range(B)]
    if reuse_decode_latents and getattr(m.gpt, "last_latents", None) is not None:
        # straight from the decode kernel (no second forward)
        return [m.gpt.last_latents[b][keeps[b]].reshape(1, -1, m.gpt.last_latents.shape[-1]) for b in range(B)]
    stride = m.config.model_args.gpt_code_stride_len
    n_max = max(int(k.numel()) for k in keeps)
    gen = torch.full((B, n_max), m.gpt.stop_audio_token, dtype=gen_rows.dtype, device=gen_rows.device)
    for b in range(B):
        gen[b, : keeps[b].numel()] = gen_rows[b][keeps[b]]
    out_len = torch.tensor([int(k.numel()) * stride for k in keeps], device=m.device)
    content_len = torch.full((B,), codes.shape[-1], device=m.device)
    lat = m.gpt(codes, content_len, gen, out_len, cond_latents=cond_latent.expand(B, -1, -1).contiguous(), return_latent=True)
    return [lat[b: b + 1, : keeps[b].numel()] for b in range(B)]


@torch.inference_mode()
def synthesize_utt(genVC_mdl, src_wav, tgt_audio, seg_len=6.0, reuse_decode_latents: bool = False, batch_segments: bool = False,
                   max_rows: int = 8):
    """Non-streaming conversion, segments joined at the latent level (inference_utils.py:23-87).
    ``batch_segments``: equal-length segments are decoded together, up to ``max_rows`` rows per ``generate`` call."""
    m = genVC_mdl
    src_wav, cond_latent, plan = _prepare(m, src_wav, tgt_audio, seg_len)
    kw = dict(do_sample=True, num_beams=1, output_attentions=False, **_sampling_kwargs(m.config))
    all_codes = [_content_codes(m, _segment(src_wav, start, end, pad)) for start, end, pad in plan]
    latents = [None] * len(plan)
    if batch_segments:
        # bucket by code length (the reference batches only equal-length rows: layers/gpt_inference.py:92-96), keep order
        buckets = {}
        for i, c in enumerate(all_codes):
            buckets.setdefault(int(c.shape[-1]), []).append(i)
        groups = [idx[k: k + max_rows] for idx in buckets.values() for k in range(0, len(idx), max_rows)]
    else:
        groups = [[i] for i in range(len(plan))]
    for grp in groups:
        codes = torch.cat([all_codes[i] for i in grp], dim=0)
        cond = cond_latent.expand(len(grp), -1, -1).contiguous()
        gen = m.gpt.generate(cond, codes, **kw)
        for i, lat in zip(grp, _segment_latents(m, cond_latent, codes, gen, reuse_decode_latents)):
            latents[i] = lat
    return _vocode(m, torch.cat(latents, dim=1))[0].squeeze()


@torch.inference_mode()
def synthesize_utt_chunked(genVC_mdl, src_wav, tgt_audio, seg_len=6.0):
    """Non-streaming conversion, segments vocoded separately and cross-faded (inference_utils.py:89-134)."""
    m = genVC_mdl
    src_wav, cond_latent, plan = _prepare(m, src_wav, tgt_audio, seg_len)
    prev, overlap, pieces = None, None, []
    for start, end, pad in plan:
        audio = m.inference(_segment(src_wav, start, end, pad), cond_latent, **_sampling_kwargs(m.config))
        piece, prev, overlap = handle_chunks(audio.squeeze(), prev, overlap, OVERLAP)
        pieces.append(piece)
    return torch.cat(pieces, dim=-1)


@torch.inference_mode()
def synthesize_utt_streaming(genVC_mdl, src_wav, tgt_audio, seg_len=6.0, stream_chunk_size=8,
                             on_chunk: Optional[Callable[[torch.Tensor], None]] = None):
    """Streaming conversion (inference_utils.py:136-217): every ``stream_chunk_size`` tokens (and at the end of a
    segment) the latents collected so far go through the vocoder and the cross-fade."""
    m = genVC_mdl
    t0 = time.time()
    total = src_wav.shape[-1]
    src_wav, cond_latent, plan = _prepare(m, src_wav, tgt_audio, seg_len)
    prev, overlap, pieces = None, None, []
    first = True
    for start, end, pad in plan:
        codes = _content_codes(m, _segment(src_wav, start, end, pad))
        gpt_inputs = m.gpt.compute_embeddings(cond_latent, codes)
        stream = m.gpt.get_generator(fake_inputs=gpt_inputs, do_sample=True, num_beams=1, num_return_sequences=1,
                                     output_attentions=False, output_hidden_states=True, **_sampling_kwargs(m.config))
        pending: List[torch.Tensor] = []
        n_tokens = 0
        finished = False
        while not finished:
            try:
                _, latent = next(stream)
                pending.append(latent)
                n_tokens += 1
            except StopIteration:
                finished = True
            if finished or (stream_chunk_size > 0 and n_tokens >= stream_chunk_size):
                if not pending:
                    # token count a multiple of the chunk size: nothing left to flush (the reference's torch.cat
                    # raises on the empty list here, inference_utils.py:196)
                    continue
                audio = _vocode(m, torch.cat(pending, dim=0)[None, :])
                piece, prev, overlap = handle_chunks(audio.squeeze(), prev, overlap, OVERLAP)
                pieces.append(piece)
                if on_chunk is not None:
                    on_chunk(piece)
                pending, n_tokens = [], 0
                if first:
                    first = False
                    m.last_latency_s = time.time() - t0
                    print(f"Latency: {m.last_latency_s:.3f}s")
    out = torch.cat(pieces, dim=-1)
    m.last_rtf = (time.time() - t0) / (total / m.content_sample_rate)
    print(f"Real-time factor: {m.last_rtf:.3f}")
    return out
