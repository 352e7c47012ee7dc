#!/usr/bin/env python
"""Generates tests/golden/hifigan_*.pt FROM THE REFERENCE MODULE (layers/hifigan.py::HiFiGAN) — build container only.

The weights are not stored: both sides rebuild them from ``genvc_b200.synth.synth_hifigan_state(seed)`` (reference key names,
weight norm on); the fixture holds the seed, a digest of the state dict, the input and the reference's output.

    python tests/golden/make_golden_hifigan.py
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = os.environ.get("GENVC_REFERENCE_ROOT", "/root/reference")

from genvc_b200.synth import HIFIGAN_DEFAULTS, state_dict_digest, synth_hifigan_state  # noqa: E402


def import_reference_hifigan():
    """layers/hifigan.py imports nnAudio (absent; discriminators only) and the reference's top-level utils (librosa,
    matplotlib): both stubbed, the generator needs get_padding / init_weights only (utils.py:174-188)."""
    na = types.ModuleType("nnAudio")
    na.features = types.ModuleType("nnAudio.features")
    sys.modules.setdefault("nnAudio", na)
    sys.modules.setdefault("nnAudio.features", na.features)
    ut = types.ModuleType("utils")
    ut.get_padding = lambda k, d=1: int((k * d - d) / 2)
    ut.get_2d_padding = lambda k, d=(1, 1): (((k[0] - 1) * d[0]) // 2, ((k[1] - 1) * d[1]) // 2)

    def init_weights(m, mean=0.0, std=0.01):
        if m.__class__.__name__.find("Conv") != -1:
            m.weight.data.normal_(mean, std)

    ut.init_weights = init_weights
    ut.NormConv2d = torch.nn.Conv2d
    sys.modules["utils"] = ut
    sys.path.insert(0, REF)
    from layers.hifigan import HiFiGAN
    return HiFiGAN


def main():
    HiFiGAN = import_reference_hifigan()
    out_dir = os.path.dirname(os.path.abspath(__file__))
    cases = {
        "hifigan_default_t32": dict(seed=77, T=32, B=1, cfg={}),            # one streaming chunk: 8 tokens x 4 frames
        "hifigan_default_t94_b2": dict(seed=78, T=94, B=2, cfg={}),         # ~1 s of audio, batch 2
        "hifigan_rb1_small": dict(seed=79, T=20, B=1, cfg=dict(input_feat_dim=64, upsample_initial_channel=64, resblock_type="1",
                                                               resblock_dilation_sizes=((1, 3, 5), (1, 3, 5), (1, 3, 5)),
                                                               upsample_rates=(4, 2), upsample_kernel_sizes=(8, 4))),
    }
    for name, c in cases.items():
        cfg = dict(HIFIGAN_DEFAULTS, **c["cfg"])
        sd = synth_hifigan_state(c["seed"], **c["cfg"])
        m = HiFiGAN(cfg["input_feat_dim"], cfg["upsample_initial_channel"], list(cfg["resblock_kernel_sizes"]),
                    [list(d) for d in cfg["resblock_dilation_sizes"]], list(cfg["upsample_rates"]),
                    list(cfg["upsample_kernel_sizes"]), cfg["resblock_type"])
        missing = m.load_state_dict(sd, strict=True)
        m.eval()
        x = torch.randn(c["B"], cfg["input_feat_dim"], c["T"], generator=torch.Generator().manual_seed(c["seed"] + 1000))
        with torch.inference_mode():
            y = m(x)
        print(name, tuple(y.shape), "abs max", float(y.abs().max()), "std", float(y.std()), missing)
        torch.save({"seed": c["seed"], "cfg": c["cfg"], "digest": state_dict_digest(sd), "x": x, "y": y.clone()},
                   os.path.join(out_dir, name + ".pt"))


if __name__ == "__main__":
    main()
