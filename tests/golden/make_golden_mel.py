#!/usr/bin/env python
"""Generates tests/golden/mel_*.pt — build container only.

``utils.py:95-158 TorchMelSpectrogram`` cannot be imported here (the reference's ``utils`` module needs librosa, which is
absent); the class is a thin wrapper: ``torchaudio.transforms.MelSpectrogram(n_fft, hop_length, win_length, power=2,
normalized=False, sample_rate, f_min, f_max, n_mels, norm="slaney")`` (``utils.py:117-128``) followed by
``log(clamp(mel, min=1e-5))`` and ``mel / mel_norms[None, :, None]`` (``utils.py:152-157``).  The fixtures are produced by that
torchaudio transform itself with the two parameter sets the reference instantiates (``trainers/hifigan_trainer.py:105-115,
142-144``), i.e. by the third-party code the reference calls, not by our restatement.

    python tests/golden/make_golden_mel.py
"""
import os

import torch
import torchaudio

OUT = os.path.dirname(os.path.abspath(__file__))
CASES = {
    # style-encoder front-end (hifigan_trainer.py:105-115), 1.5 s of audio at 24 kHz
    "mel_style_24k": dict(filter_length=2048, hop_length=256, win_length=1024, n_mel_channels=80, mel_fmin=0, mel_fmax=8000,
                          sampling_rate=24000, n=36000, batch=2, seed=11, norms=True),
    # DVAE front-end defaults (utils.py:98-105), odd length
    "mel_default_22k": dict(filter_length=1024, hop_length=256, win_length=1024, n_mel_channels=80, mel_fmin=0, mel_fmax=8000,
                            sampling_rate=22050, n=9999, batch=1, seed=12, norms=False),
}


def signal(batch, n, seed):
    """White noise, a speech-like 1/f^2 process and a few tones (60 dB of dynamic range across the spectrum)."""
    g = torch.Generator().manual_seed(seed)
    white = torch.randn(batch, n, generator=g) * 0.05
    brown = torch.cumsum(torch.randn(batch, n, generator=g), -1)
    brown = (brown - brown.mean(-1, keepdim=True)) / brown.abs().amax(-1, keepdim=True) * 0.5
    t = torch.arange(n) / 24000.0
    tones = 0.2 * torch.sin(2 * torch.pi * 220.0 * t) + 0.05 * torch.sin(2 * torch.pi * 3100.0 * t)
    return white + brown + tones


for name, c in CASES.items():
    tr = torchaudio.transforms.MelSpectrogram(n_fft=c["filter_length"], hop_length=c["hop_length"], win_length=c["win_length"], power=2,
                                              normalized=False, sample_rate=c["sampling_rate"], f_min=c["mel_fmin"], f_max=c["mel_fmax"],
                                              n_mels=c["n_mel_channels"], norm="slaney")
    x = signal(c["batch"], c["n"], c["seed"])
    mel = torch.log(torch.clamp(tr(x), min=1e-5))
    norms = None
    if c["norms"]:
        norms = 0.5 + torch.rand(c["n_mel_channels"], generator=torch.Generator().manual_seed(c["seed"] + 1)) * 4.0
        mel = mel / norms.unsqueeze(0).unsqueeze(-1)
    print(name, tuple(mel.shape), float(mel.min()), float(mel.max()))
    torch.save({"args": {k: c[k] for k in ("filter_length", "hop_length", "win_length", "n_mel_channels", "mel_fmin", "mel_fmax", "sampling_rate")},
                "x": x, "mel": mel.clone(), "norms": norms}, os.path.join(OUT, name + ".pt"))
