#!/usr/bin/env python
"""Generates tests/golden/dvae_*.pt FROM THE REFERENCE MODULE (layers/dvae.py::DiscreteVAE.get_codebook_indices) — build
container only.  Weights are rebuilt on both sides from ``genvc_b200.synth.synth_dvae_state(seed)``; the fixture holds the
seed, a digest, the input features, the reference's codes and the top-2 distance gap of every position (a parity test can
tell a rounding flip at a near-tie from an error).

    python tests/golden/make_golden_dvae.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = os.environ.get("GENVC_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, REF)

from genvc_b200.synth import CONTENT_DVAE_DEFAULTS, state_dict_digest, synth_dvae_state  # noqa: E402
from oracle.dvae_oracle import encoder_forward  # noqa: E402


def main():
    from layers.dvae import DiscreteVAE
    out_dir = os.path.dirname(os.path.abspath(__file__))
    cases = {
        "dvae_content_t50": dict(seed=55, T=50, B=1, cfg={}),        # 1 s of 50 Hz content features -> 13 codes
        "dvae_content_t300_b2": dict(seed=56, T=300, B=2, cfg={}),   # 6 s, batch 2 -> 75 codes
        "dvae_small": dict(seed=57, T=37, B=1, cfg=dict(channels=24, num_tokens=40, codebook_dim=32, hidden_dim=16, num_resnet_blocks=1)),
    }
    for name, c in cases.items():
        cfg = dict(CONTENT_DVAE_DEFAULTS, **c["cfg"])
        sd = synth_dvae_state(c["seed"], **c["cfg"])
        m = DiscreteVAE(channels=cfg["channels"], normalization=None, positional_dims=1, num_tokens=cfg["num_tokens"],
                        codebook_dim=cfg["codebook_dim"], hidden_dim=cfg["hidden_dim"], num_resnet_blocks=cfg["num_resnet_blocks"],
                        kernel_size=cfg["kernel_size"], num_layers=cfg["num_layers"], use_transposed_convs=False)
        res = m.load_state_dict(sd, strict=False)
        assert not res.unexpected_keys and all(k.startswith(("decoder.", "codebook.", "discrete_loss.")) for k in res.missing_keys), res
        x = torch.randn(c["B"], cfg["channels"], c["T"], generator=torch.Generator().manual_seed(c["seed"] + 1000))
        codes = m.get_codebook_indices(x)
        arch = dict(num_layers=cfg["num_layers"], num_resnet_blocks=cfg["num_resnet_blocks"], kernel_size=cfg["kernel_size"])
        f = encoder_forward(sd, x, **arch).permute(0, 2, 1).reshape(-1, cfg["codebook_dim"])
        e = sd["codebook.embed"]
        dist = f.pow(2).sum(1, keepdim=True) - 2 * f @ e + e.pow(2).sum(0, keepdim=True)
        top2 = (-dist).topk(2, dim=1).values
        gap = (top2[:, 0] - top2[:, 1]).view(codes.shape)
        print(name, tuple(codes.shape), "distinct codes", int(codes.unique().numel()), "min top-2 gap", float(gap.min()),
              "|dist| ~", float(dist.abs().mean()))
        torch.save({"seed": c["seed"], "cfg": c["cfg"], "digest": state_dict_digest(sd), "x": x, "codes": codes.clone(), "gap": gap.clone()},
                   os.path.join(out_dir, name + ".pt"))


if __name__ == "__main__":
    main()
