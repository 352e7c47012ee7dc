"""ORACLE (test infrastructure only) — CPU restatement of the reference's HiFi-GAN generator, the stage right after the
codec-token path.  Plain torch fp32 functional ops; every function cites the reference lines it follows.

    layers/hifigan.py:118-153   ResBlock2 (resblock_type "2", the vocoder config's default: configs/vocoder_configs.py:20)
    layers/hifigan.py:28-116    ResBlock1
    layers/hifigan.py:156-232   HiFiGAN.__init__ / forward
    inference/inference_utils.py:81-85  x``hifigan_scale_factor`` linear interpolation in front of the vocoder

Pinned against the reference module itself: ``tests/golden/make_golden_hifigan.py`` (committed) loads the same synthetic
state dict (``genvc_b200/synth.py::synth_hifigan_state``) into ``layers.hifigan.HiFiGAN`` and stores its output;
``tests/test_hifigan.py`` compares this oracle and the CUDA path with that fixture.
Only ``tests/`` and the CPU leg of ``tools/vocoder_bench.py`` may import this file.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.1  # layers/hifigan.py:22


def get_padding(kernel_size: int, dilation: int = 1) -> int:  # utils.py:174-175
    return int((kernel_size * dilation - dilation) / 2)


def fold_weight_norm(sd: Dict[str, torch.Tensor], prefix: str) -> torch.Tensor:
    """torch.nn.utils.weight_norm (dim=0): w = g * v / ||v||, the norm taken over every dim but the first
    (layers/hifigan.py:32-42 wraps every conv).  A state dict saved after remove_weight_norm has plain ``weight``."""
    if prefix + ".weight" in sd:
        return sd[prefix + ".weight"].float()
    g, v = sd[prefix + ".weight_g"].float(), sd[prefix + ".weight_v"].float()
    norm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, *([1] * (v.dim() - 1)))
    return g * v / norm


def hifigan_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, resblock_kernel_sizes: Sequence[int] = (3, 5, 7),
                    resblock_dilation_sizes: Sequence[Sequence[int]] = ((1, 2), (2, 6), (3, 12)),
                    upsample_rates: Sequence[int] = (8, 8, 4), upsample_kernel_sizes: Sequence[int] = (16, 16, 8),
                    resblock_type: str = "2") -> torch.Tensor:
    """x [B, input_feat_dim, T] -> waveform [B, 1, T * prod(upsample_rates)]   (layers/hifigan.py:210-225)."""
    nk = len(resblock_kernel_sizes)
    x = F.conv1d(x.float(), fold_weight_norm(sd, "conv_pre"), sd["conv_pre.bias"].float(), padding=3)  # :166-174, :211
    for i, (u, k) in enumerate(zip(upsample_rates, upsample_kernel_sizes)):
        x = F.leaky_relu(x, LRELU_SLOPE)  # :213
        x = F.conv_transpose1d(x, fold_weight_norm(sd, f"ups.{i}"), sd[f"ups.{i}.bias"].float(), stride=u,
                               padding=(k - u) // 2)  # :178-196, :214
        xs = None
        for j, (rk, rd) in enumerate(zip(resblock_kernel_sizes, resblock_dilation_sizes)):
            r = _resblock(sd, f"resblocks.{i * nk + j}", x, rk, rd, resblock_type)
            xs = r if xs is None else xs + r  # :216-220
        x = xs / nk  # :221
    x = F.leaky_relu(x)  # default slope 0.01 (:222)
    x = F.conv1d(x, fold_weight_norm(sd, "conv_post"), sd["conv_post.bias"].float(), padding=3)  # :207, :223
    return torch.tanh(x)  # :224


def _resblock(sd, prefix: str, x: torch.Tensor, k: int, dil: Sequence[int], kind: str) -> torch.Tensor:
    if kind == "1":  # layers/hifigan.py:98-105
        for m, d in enumerate(dil):
            xt = F.leaky_relu(x, LRELU_SLOPE)
            xt = F.conv1d(xt, fold_weight_norm(sd, f"{prefix}.convs1.{m}"), sd[f"{prefix}.convs1.{m}.bias"].float(),
                          dilation=d, padding=get_padding(k, d))
            xt = F.leaky_relu(xt, LRELU_SLOPE)
            xt = F.conv1d(xt, fold_weight_norm(sd, f"{prefix}.convs2.{m}"), sd[f"{prefix}.convs2.{m}.bias"].float(),
                          dilation=1, padding=get_padding(k, 1))
            x = xt + x
        return x
    for m, d in enumerate(dil):  # layers/hifigan.py:147-152
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = F.conv1d(xt, fold_weight_norm(sd, f"{prefix}.convs.{m}"), sd[f"{prefix}.convs.{m}.bias"].float(), dilation=d,
                      padding=get_padding(k, d))
        x = xt + x
    return x


def vocode(sd, latents: torch.Tensor, scale_factor: float = 4.0, **cfg) -> torch.Tensor:
    """[B, M, D] GPT latents -> waveform (inference/inference_utils.py:81-85)."""
    mel_input = F.interpolate(latents.transpose(1, 2), scale_factor=[scale_factor], mode="linear").squeeze(1)
    return hifigan_forward(sd, mel_input, **cfg)
