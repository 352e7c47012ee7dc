"""Seeded synthetic GenVC checkpoints in the reference's on-disk layout.

No pretrained checkpoint is reachable offline, so parity and benchmarks run on
random weights of the right architecture.  The layout is exactly what
``inference/model_init.py:11-22`` consumes: ``{"config": nested dict, "model":
flat state_dict}`` with the GPT's tensors under the ``gpt.`` prefix (the
``HiFiGANTrainer.gpt`` attribute, ``trainers/hifigan_trainer.py:31``).  Key names
and shapes were read off the reference ``GPT.state_dict()``;
``tests/golden/make_golden.py`` loads these state dicts into the real reference
module with ``strict=True`` to keep the two in sync.

Weights are N(0, 0.02) like GPT-2's init, but biases and LayerNorm parameters are
randomised too (a trained checkpoint has non-trivial ones and zero biases would
hide indexing bugs).  Every tensor is drawn from its own generator seeded by
(seed, key) so the values do not depend on creation order or on which subset of
tensors is requested.
"""
from __future__ import annotations

import hashlib
from typing import Dict, Iterator, Tuple

import torch

from .config import GenVCDims, make_config_dict

PREFIX = "gpt."


def gpt_state_shapes(d: GenVCDims) -> Iterator[Tuple[str, Tuple[int, ...], str]]:
    """(key without prefix, shape, kind) for every tensor of the reference GPT module."""
    D = d.d_model
    yield "text_embedding.weight", (d.n_text_vocab, D), "emb"
    yield "mel_embedding.weight", (d.n_audio_vocab, D), "emb"
    for i in range(d.n_layer):
        p = f"gpt.h.{i}."
        yield p + "ln_1.weight", (D,), "ln_w"
        yield p + "ln_1.bias", (D,), "ln_b"
        yield p + "attn.c_attn.weight", (D, 3 * D), "w"  # Conv1D: [in, out]
        yield p + "attn.c_attn.bias", (3 * D,), "b"
        yield p + "attn.c_proj.weight", (D, D), "w"
        yield p + "attn.c_proj.bias", (D,), "b"
        yield p + "ln_2.weight", (D,), "ln_w"
        yield p + "ln_2.bias", (D,), "ln_b"
        yield p + "mlp.c_fc.weight", (D, 4 * D), "w"
        yield p + "mlp.c_fc.bias", (4 * D,), "b"
        yield p + "mlp.c_proj.weight", (4 * D, D), "w"
        yield p + "mlp.c_proj.bias", (D,), "b"
    yield "gpt.ln_f.weight", (D,), "ln_w"
    yield "gpt.ln_f.bias", (D,), "ln_b"
    yield "mel_pos_embedding.emb.weight", (d.n_mel_pos, D), "emb"
    yield "text_pos_embedding.emb.weight", (d.n_text_pos, D), "emb"
    yield "final_norm.weight", (D,), "ln_w"
    yield "final_norm.bias", (D,), "ln_b"
    yield "text_head.weight", (d.n_text_vocab, D), "lin"  # nn.Linear: [out, in]
    yield "text_head.bias", (d.n_text_vocab,), "b"
    yield "mel_head.weight", (d.n_audio_vocab, D), "lin"
    yield "mel_head.bias", (d.n_audio_vocab,), "b"
    pc = "conditioning_perceiver."
    yield pc + "latents", (d.pc_latents, D), "emb"
    yield pc + "proj_context.weight", (D, d.pc_dim_context), "lin"
    yield pc + "proj_context.bias", (D,), "b"
    for i in range(d.pc_depth):
        a = f"{pc}layers.{i}.0."
        f = f"{pc}layers.{i}.1."
        yield a + "to_q.weight", (d.pc_inner, D), "lin"
        yield a + "to_kv.weight", (2 * d.pc_inner, D), "lin"
        yield a + "to_out.weight", (D, d.pc_inner), "lin"
        yield f + "0.weight", (2 * d.pc_ff_inner, D), "lin"
        yield f + "0.bias", (2 * d.pc_ff_inner,), "b"
        yield f + "2.weight", (D, d.pc_ff_inner), "lin"
        yield f + "2.bias", (D,), "b"
    yield pc + "norm.gamma", (D,), "ln_w"


def _key_seed(seed: int, key: str) -> int:
    h = hashlib.sha256(f"{seed}:{key}".encode()).digest()
    return int.from_bytes(h[:7], "little")


def _draw(kind: str, shape, gen: torch.Generator) -> torch.Tensor:
    if kind in ("w", "emb"):
        return torch.randn(shape, generator=gen) * 0.02
    if kind == "lin":
        bound = 1.0 / (shape[-1] ** 0.5)
        return (torch.rand(shape, generator=gen) * 2 - 1) * bound
    if kind == "b":
        return torch.randn(shape, generator=gen) * 0.02
    if kind == "ln_w":
        return 1.0 + torch.randn(shape, generator=gen) * 0.1
    if kind == "ln_b":
        return torch.randn(shape, generator=gen) * 0.05
    raise ValueError(kind)


def synth_state_dict(d: GenVCDims, seed: int = 1234, eos_bias: float = 0.0) -> Dict[str, torch.Tensor]:
    """State dict (keys carry the ``gpt.`` prefix).  ``eos_bias`` is added to
    ``mel_head.bias[stop_audio]`` so a checkpoint can be made to emit EOS early."""
    sd: Dict[str, torch.Tensor] = {}
    for key, shape, kind in gpt_state_shapes(d):
        gen = torch.Generator(device="cpu")
        gen.manual_seed(_key_seed(seed, key))
        sd[PREFIX + key] = _draw(kind, shape, gen).to(torch.float32).contiguous()
    if eos_bias:
        sd[PREFIX + "mel_head.bias"][d.stop_audio] += eos_bias
    return sd


def synth_checkpoint(n_layer=30, d_model=1024, n_head=4, seed=1234, eos_bias=0.0, **cfg_overrides) -> dict:
    cfg = make_config_dict(n_layer, d_model, n_head, **cfg_overrides)
    d = GenVCDims.from_config(cfg)
    return {"config": cfg, "model": synth_state_dict(d, seed, eos_bias)}


def write_checkpoint(path: str, **kw) -> dict:
    ck = synth_checkpoint(**kw)
    torch.save(ck, path)
    return ck


def state_dict_digest(sd: Dict[str, torch.Tensor], keys=None) -> str:
    """Short fingerprint of a state dict (fixtures store it to detect RNG drift)."""
    h = hashlib.sha256()
    for k in sorted(keys or sd.keys()):
        t = sd[k]
        h.update(k.encode())
        # a strided sample keeps this cheap at full size while still catching any drift
        flat = t.reshape(-1)
        step = max(1, flat.numel() // 4096)
        h.update(flat[::step].contiguous().numpy().tobytes())
    return h.hexdigest()[:16]


# ------------------------------------------------------------------------------------------------------------------
# HiFi-GAN generator (the stage after the path): synthetic weights under the reference's state-dict names
# ------------------------------------------------------------------------------------------------------------------
HIFIGAN_DEFAULTS = dict(input_feat_dim=1024, upsample_initial_channel=256, resblock_kernel_sizes=(3, 5, 7),
                        resblock_dilation_sizes=((1, 2), (2, 6), (3, 12)), upsample_rates=(8, 8, 4),
                        upsample_kernel_sizes=(16, 16, 8), resblock_type="2")  # configs/vocoder_configs.py:7-20


def hifigan_conv_shapes(cfg: dict):
    """(name, weight shape, transposed?) of every conv of layers/hifigan.py::HiFiGAN in construction order (:166-207)."""
    c0 = cfg["upsample_initial_channel"]
    out = [("conv_pre", (c0, cfg["input_feat_dim"], 7), False)]
    for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        out.append((f"ups.{i}", (c0 // 2 ** i, c0 // 2 ** (i + 1), k), True))  # ConvTranspose1d weight: [in, out, k]
    nk = len(cfg["resblock_kernel_sizes"])
    ch = c0
    for i in range(len(cfg["upsample_rates"])):
        ch = c0 // 2 ** (i + 1)
        for j, (k, d) in enumerate(zip(cfg["resblock_kernel_sizes"], cfg["resblock_dilation_sizes"])):
            names = [f"convs.{m}" for m in range(len(d))] if cfg["resblock_type"] != "1" else \
                [f"convs{a}.{m}" for m in range(len(d)) for a in (1, 2)]
            for nm in names:
                out.append((f"resblocks.{i * nk + j}.{nm}", (ch, ch, k), False))
    out.append(("conv_post", (1, ch, 7), False))
    return out


def synth_hifigan_state(seed: int = 77, weight_norm: bool = True, **overrides) -> Dict[str, torch.Tensor]:
    """Random generator weights, keys as ``HiFiGAN.state_dict()`` with (old-style) weight norm: ``<conv>.weight_g``
    [out,1,1], ``<conv>.weight_v``, ``<conv>.bias``.  Scales keep the activations O(1) through the stack so that a
    parity test sees every layer."""
    cfg = dict(HIFIGAN_DEFAULTS, **overrides)
    gen = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for name, shape, transposed in hifigan_conv_shapes(cfg):
        fan_in = (shape[0] if transposed else shape[1]) * shape[2]
        taps = shape[2] / (shape[2] // 2 if transposed else 1)  # a transposed conv with k = 2 * stride sees 2 taps per output
        v = torch.randn(shape, generator=gen) * (1.3 / (fan_in / (taps if transposed else 1)) ** 0.5)
        n_bias = shape[1] if transposed else shape[0]
        bias = torch.randn(n_bias, generator=gen) * 0.05
        if weight_norm:
            norm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, 1, 1)
            sd[name + ".weight_g"] = norm * (0.8 + 0.4 * torch.rand(norm.shape, generator=gen))
            sd[name + ".weight_v"] = v
        else:
            sd[name + ".weight"] = v
        sd[name + ".bias"] = bias
    return sd


# ------------------------------------------------------------------------------------------------------------------
# content-DVAE tokeniser (the stage before the path): synthetic weights under the reference's state-dict names
# ------------------------------------------------------------------------------------------------------------------
CONTENT_DVAE_DEFAULTS = dict(channels=256, num_tokens=256, codebook_dim=512, hidden_dim=512, num_resnet_blocks=3, kernel_size=3,
                             num_layers=2)  # train_content_dvae.py:32-38


def synth_dvae_state(seed: int = 55, **overrides) -> Dict[str, torch.Tensor]:
    """Encoder + codebook of ``layers/dvae.py::DiscreteVAE(positional_dims=1)`` (keys ``encoder.N...``, ``codebook.embed``);
    the decoder is not part of inference and is left out (load with strict=False)."""
    cfg = dict(CONTENT_DVAE_DEFAULTS, **overrides)
    gen = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def conv(name, cout, cin, k, gain=1.4):
        sd[name + ".weight"] = torch.randn(cout, cin, k, generator=gen) * (gain / (cin * k) ** 0.5)
        sd[name + ".bias"] = torch.randn(cout, generator=gen) * 0.05

    chans = [cfg["channels"]] + [cfg["hidden_dim"] * 2 ** i for i in range(cfg["num_layers"])]
    i = 0
    for cin, cout in zip(chans[:-1], chans[1:]):
        conv(f"encoder.{i}.0", cout, cin, cfg["kernel_size"])
        i += 1
    inner = chans[-1]
    for _ in range(cfg["num_resnet_blocks"]):
        conv(f"encoder.{i}.net.0", inner, inner, 3)
        conv(f"encoder.{i}.net.2", inner, inner, 3)
        conv(f"encoder.{i}.net.4", inner, inner, 1, gain=0.5)  # keeps the residual stream O(1) through the blocks
        i += 1
    conv(f"encoder.{i}", cfg["codebook_dim"], inner, 1, gain=1.0)
    sd["codebook.embed"] = torch.randn(cfg["codebook_dim"], cfg["num_tokens"], generator=gen)
    return sd
