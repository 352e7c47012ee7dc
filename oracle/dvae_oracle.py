"""ORACLE (test infrastructure only) — CPU restatement of the content-DVAE tokeniser, the stage right before the codec-token
path: ``DiscreteVAE.get_codebook_indices`` for ``positional_dims=1``, ``normalization=None``.

    layers/dvae.py:251-292   encoder = num_layers x [Conv1d(k, stride 2, pad (k-1)//2), ReLU], num_resnet_blocks x ResBlock,
                             Conv1d(innermost, codebook_dim, 1)
    layers/dvae.py:171-184   ResBlock: conv3 - act - conv3 - act - conv1, plus the input
    layers/dvae.py:84-88     Quantize.forward: dist = |f|^2 - 2 f @ embed + |embed|^2 ; code = argmax(-dist)
    layers/dvae.py:324-331   get_codebook_indices
    trainers/hifigan_trainer.py:149-160   the content DVAE is built from config.content_dvae_config

Pinned against the reference module: ``tests/golden/make_golden_dvae.py`` loads the same synthetic state dict into
``layers.dvae.DiscreteVAE`` and stores its codes.  Only ``tests/`` may import this file.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F


def encoder_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, num_layers: int, num_resnet_blocks: int, kernel_size: int,
                    stride: int = 2) -> torch.Tensor:
    """x [B, channels, T] -> encoder output [B, codebook_dim, T']."""
    pad = (kernel_size - 1) // 2  # layers/dvae.py:263
    i = 0
    for _ in range(num_layers):  # nn.Sequential(conv, act()) (:265)
        x = F.relu(F.conv1d(x, sd[f"encoder.{i}.0.weight"], sd[f"encoder.{i}.0.bias"], stride=stride, padding=pad))
        i += 1
    for _ in range(num_resnet_blocks):  # ResBlock (:171-184, :279)
        h = F.relu(F.conv1d(x, sd[f"encoder.{i}.net.0.weight"], sd[f"encoder.{i}.net.0.bias"], padding=1))
        h = F.relu(F.conv1d(h, sd[f"encoder.{i}.net.2.weight"], sd[f"encoder.{i}.net.2.bias"], padding=1))
        x = F.conv1d(h, sd[f"encoder.{i}.net.4.weight"], sd[f"encoder.{i}.net.4.bias"]) + x
        i += 1
    return F.conv1d(x, sd[f"encoder.{i}.weight"], sd[f"encoder.{i}.bias"])  # :284


def get_codebook_indices(sd: Dict[str, torch.Tensor], x: torch.Tensor, **arch) -> torch.Tensor:
    logits = encoder_forward(sd, x.float(), **arch).permute(0, 2, 1)  # :327
    embed = sd["codebook.embed"].float()  # [dim, n_embed]
    flatten = logits.reshape(-1, embed.shape[0])
    dist = flatten.pow(2).sum(1, keepdim=True) - 2 * flatten @ embed + embed.pow(2).sum(0, keepdim=True)  # :85
    _, ind = (-dist).max(1)  # :86-87
    return ind.view(*logits.shape[:-1])
