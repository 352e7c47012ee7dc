"""Multi-GPU: replicas only.

The autoregressive step does not shard (30 sequential layers, 1.5 GB of weights); utterances — and
the 6 s segments of one utterance — are independent.  So: one process per GPU, rank 0 reads and packs
the checkpoint, ONE broadcast of the packed blob at init (NCCL over NVLink on GPUs; gloo in the CPU
tests), units dealt round-robin to ranks, result ids gathered at the end.  Nothing is exchanged per step.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist

from .config import GenVCDims
from .weights import blob_layout, pack_state_dict


def shard_units(n_units: int, rank: int, world: int) -> List[int]:
    """Indices of the utterances/segments rank ``rank`` processes."""
    if not (0 <= rank < world):
        raise ValueError("rank outside the world")
    return list(range(rank, n_units, world))


def plan_batches(n_units: int, max_rows: int) -> List[List[int]]:
    """Group ``n_units`` equal-shape units (utterances whose segments have the same lengths) into batches of at most
    ``max_rows`` rows, as even as possible (the reference batches only equal-length rows: layers/gpt_inference.py:92-96)."""
    if n_units <= 0:
        return []
    if max_rows <= 0:
        raise ValueError("max_rows must be positive")
    n_batches = (n_units + max_rows - 1) // max_rows
    base, extra = divmod(n_units, n_batches)
    out, start = [], 0
    for b in range(n_batches):
        size = base + (1 if b < extra else 0)
        out.append(list(range(start, start + size)))
        start += size
    return out


def broadcast_blob(blob: Optional[torch.Tensor], n_floats: int, rank: int, world: int, device) -> torch.Tensor:
    """Rank 0 passes the packed host blob; every rank returns it on ``device``."""
    device = torch.device(device)
    if rank == 0:
        if blob is None or blob.numel() != n_floats:
            raise ValueError("rank 0 must provide the packed blob")
        t = blob.to(device)
    else:
        t = torch.empty(n_floats, dtype=torch.float32, device=device)
    if world > 1:
        dist.broadcast(t, src=0)
    return t


def gather_ids(local_ids: torch.Tensor, world: int, pad: int) -> List[torch.Tensor]:
    """all-gather of padded id matrices [n_local, max_len] (rows padded with ``pad``); ranks may hold
    different row counts."""
    if world == 1:
        return [local_ids]
    shape = torch.tensor(list(local_ids.shape), dtype=torch.int64, device=local_ids.device)
    shapes = [torch.zeros_like(shape) for _ in range(world)]
    dist.all_gather(shapes, shape)
    rows = max(int(s[0]) for s in shapes)
    cols = max(int(s[1]) for s in shapes)
    buf = torch.full((rows, cols), pad, dtype=torch.int64, device=local_ids.device)
    buf[: local_ids.shape[0], : local_ids.shape[1]] = local_ids
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return [o[: int(s[0]), : int(s[1])] for o, s in zip(out, shapes)]


def init_replica(ckpt: dict, device, rank: int, world: int, max_batch: int = 1, max_mel_frames: int = 576):
    """Build this rank's model replica.  ``ckpt["model"]`` is only read on rank 0."""
    from .inference.model_init import model_from_checkpoint

    dims = GenVCDims.from_config(ckpt["config"])
    n_floats, _ = blob_layout(dims)
    blob = pack_state_dict(dims, ckpt["model"]) if rank == 0 else None
    blob = broadcast_blob(blob, n_floats, rank, world, device)
    model, _ = model_from_checkpoint(ckpt, device, max_batch=max_batch, max_mel_frames=max_mel_frames, blob=blob)
    return model
