"""Checkpoint -> packed fp32 weight blob (host side; needs the shared library but no GPU).

The blob layout is owned by the C library (``genvc_tensor_info``): reference state-dict names and
orientations (HF Conv1D ``[in, out]``, nn.Linear ``[out, in]``), every tensor 128-byte aligned.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Tuple

import torch

from .config import GenVCDims
from .lib import GenvcConfig, GenvcError, load_library

PREFIX = "gpt."  # state-dict prefix of the GPT inside a GenVC checkpoint (trainers/hifigan_trainer.py:31)


def c_config(d: GenVCDims, max_batch: int = 1, max_seq: int = 0, max_mel_frames: int = 576) -> GenvcConfig:
    if max_seq <= 0:
        max_seq = (d.max_seq + 7) // 8 * 8
    return GenvcConfig(
        n_layer=d.n_layer, d_model=d.d_model, n_head=d.n_head,
        n_text_vocab=d.n_text_vocab, n_audio_vocab=d.n_audio_vocab,
        start_text=d.start_text, stop_text=d.stop_text, start_audio=d.start_audio, stop_audio=d.stop_audio,
        n_mel_pos=d.n_mel_pos, n_text_pos=d.n_text_pos, max_gen_mel_tokens=d.max_gen_mel_tokens,
        pc_depth=d.pc_depth, pc_dim_context=d.pc_dim_context, pc_latents=d.pc_latents,
        pc_dim_head=d.pc_dim_head, pc_heads=d.pc_heads, pc_ff_inner=d.pc_ff_inner,
        max_batch=max_batch, max_seq=max_seq, max_mel_frames=max_mel_frames,
    )


def tensor_table(lib, ctx) -> List[Tuple[str, int, int, int, int]]:
    """(key, float offset, rows, cols, row stride) of every tensor of the blob."""
    out = []
    buf = C.create_string_buffer(256)
    o, r, c, s = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
    for i in range(lib.genvc_num_tensors(ctx)):
        if lib.genvc_tensor_name(ctx, i, buf, 256) != 0:
            raise GenvcError(-1, "genvc_tensor_name failed")
        if lib.genvc_tensor_info(ctx, buf.value, C.byref(o), C.byref(r), C.byref(c), C.byref(s)) != 0:
            raise GenvcError(-1, lib.genvc_last_error(ctx).decode())
        out.append((buf.value.decode(), o.value, r.value, c.value, s.value))
    return out


def blob_layout(dims: GenVCDims) -> Tuple[int, List[Tuple[str, int, int, int, int]]]:
    """(total floats, tensor table) — a layout-only context, no device needed."""
    lib = load_library()
    cfg = c_config(dims)
    ctx = C.c_void_p()
    rc = lib.genvc_create(C.byref(cfg), 0, C.byref(ctx))
    try:
        if rc != 0:
            raise GenvcError(rc, lib.genvc_last_error(ctx).decode() if ctx else "genvc_create failed")
        return int(lib.genvc_blob_floats(ctx)), tensor_table(lib, ctx)
    finally:
        if ctx:
            lib.genvc_destroy(ctx)


def pack_state_dict(dims: GenVCDims, state_dict: Dict[str, torch.Tensor], strict: bool = True) -> torch.Tensor:
    """Pack the checkpoint's ``gpt.*`` tensors into one fp32 host blob (replaces
    ``model.load_state_dict`` for this path, inference/model_init.py:22).  Keys outside the path are
    ignored (the checkpoint also holds the DVAEs and HiFi-GAN); buffers such as ``attn.bias`` /
    ``attn.masked_bias`` written by old transformers versions are ignored too."""
    total, table = blob_layout(dims)
    blob = torch.zeros(total, dtype=torch.float32)
    missing = []
    for key, off, rows, cols, stride in table:
        t = state_dict.get(PREFIX + key)
        if t is None:
            if not key.startswith("text_head."):  # unused at inference; tolerate pruned checkpoints
                missing.append(PREFIX + key)
            continue
        if t.numel() != rows * cols:
            raise ValueError(f"checkpoint tensor {PREFIX + key} has shape {tuple(t.shape)}; expected {rows}x{cols}")
        blob[off: off + rows * stride].view(rows, stride)[:, :cols].copy_(t.detach().to(torch.float32).reshape(rows, cols))
    if missing and strict:
        raise KeyError(f"checkpoint is missing {len(missing)} tensors of the GPT path, e.g. {missing[:3]}")
    return blob
