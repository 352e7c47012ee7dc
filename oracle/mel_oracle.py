"""ORACLE (test infrastructure only) — CPU restatement of the mel front-end in front of the perceiver (SURVEY.md §8 row a1,
§8f #3): ``utils.py:95-158  TorchMelSpectrogram`` = ``torchaudio.transforms.MelSpectrogram`` (power 2, centre / reflect
padding, periodic Hann window of ``win_length`` centred in ``n_fft``, HTK mel scale with Slaney area normalisation) followed by
``log(clamp(mel, 1e-5))`` and the optional division by per-channel ``mel_norms``.

Third-party arithmetic absent from /root/reference: torchaudio (``MelSpectrogram``, ``functional.melscale_fbanks``, ``spectrogram``),
restated here with torch.fft; pinned in ``tests/test_mel_frontend.py`` against torchaudio itself where it is importable and
against fixtures generated with it (``tests/golden/make_golden_mel.py``).
The style-encoder instance is ``TorchMelSpectrogram(filter_length=2048, hop_length=256, win_length=1024, sampling_rate=24000,
mel_fmin=0, mel_fmax=8000, n_mel_channels=80)`` (``trainers/hifigan_trainer.py:105-115``).
"""
from __future__ import annotations

import math
from typing import Optional

import torch


def hz_to_mel_htk(f):
    return 2595.0 * math.log10(1.0 + f / 700.0)


def melscale_fbanks(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int) -> torch.Tensor:
    """torchaudio.functional.melscale_fbanks(norm="slaney", mel_scale="htk") -> [n_freqs, n_mels]."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs, dtype=torch.float64)
    m_min, m_max = hz_to_mel_htk(f_min), hz_to_mel_htk(f_max)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2, dtype=torch.float64)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = torch.clamp(torch.min(down, up), min=0.0)
    enorm = 2.0 / (f_pts[2: n_mels + 2] - f_pts[:n_mels])  # slaney
    return (fb * enorm.unsqueeze(0)).float()


def padded_window(win_length: int, n_fft: int) -> torch.Tensor:
    """torch.stft centres a window shorter than n_fft."""
    w = torch.hann_window(win_length, periodic=True, dtype=torch.float64)
    left = (n_fft - win_length) // 2
    out = torch.zeros(n_fft, dtype=torch.float64)
    out[left: left + win_length] = w
    return out


def log_mel(wav: torch.Tensor, n_fft=1024, hop=256, win=1024, n_mels=80, f_min=0.0, f_max=8000.0, sample_rate=22050,
            mel_norms: Optional[torch.Tensor] = None) -> torch.Tensor:
    """wav [B, N] -> [B, n_mels, 1 + N // hop]   (utils.py:143-158)."""
    if wav.dim() == 3:
        wav = wav.squeeze(1)
    x = wav.float()
    pad = n_fft // 2
    xp = torch.nn.functional.pad(x.unsqueeze(1), (pad, pad), mode="reflect").squeeze(1)
    frames = xp.unfold(-1, n_fft, hop)  # [B, T, n_fft]
    spec = torch.fft.rfft(frames * padded_window(win, n_fft).float(), dim=-1)
    power = spec.real.pow(2) + spec.imag.pow(2)  # [B, T, n_fft // 2 + 1]
    mel = torch.matmul(power, melscale_fbanks(n_fft // 2 + 1, f_min, f_max, n_mels, sample_rate)).transpose(1, 2)
    mel = torch.log(torch.clamp(mel, min=1e-5))
    if mel_norms is not None:
        mel = mel / mel_norms.float().unsqueeze(0).unsqueeze(-1)
    return mel
