"""Mel front-end of the conditioning path on the CUDA library.

Host-side mirror of ``utils.py:95-158 TorchMelSpectrogram`` (same constructor arguments, ``forward(wav[B, N] or [B, 1, N]) ->
[B, n_mel_channels, 1 + N // hop]``): the tables torchaudio builds (periodic Hann window centred in ``filter_length``, HTK mel
filterbank with Slaney normalisation) are built once in float64, the arithmetic is one kernel (``genvc_mel_spectrogram``).
``mel_norm_file`` may be a path (``torch.load``), a tensor of per-channel norms, or None.  No CPU fallback.
"""
from __future__ import annotations

import math
from typing import Optional, Union

import torch

from .lib import GenvcError, load_library


def _melscale_fbanks(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int) -> torch.Tensor:
    """torchaudio.functional.melscale_fbanks(norm="slaney", mel_scale="htk"), in float64."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs, dtype=torch.float64)
    to_mel = lambda f: 2595.0 * math.log10(1.0 + f / 700.0)  # noqa: E731
    m_pts = torch.linspace(to_mel(f_min), to_mel(f_max), n_mels + 2, dtype=torch.float64)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    fb = torch.clamp(torch.min((-1.0 * slopes[:, :-2]) / f_diff[:-1], slopes[:, 2:] / f_diff[1:]), min=0.0)
    return fb * (2.0 / (f_pts[2: n_mels + 2] - f_pts[:n_mels])).unsqueeze(0)


class TorchMelSpectrogram:
    def __init__(self, filter_length=1024, hop_length=256, win_length=1024, n_mel_channels=80, mel_fmin=0, mel_fmax=8000,
                 sampling_rate=22050, normalize=False, mel_norm_file: Union[None, str, torch.Tensor] = None, device="cuda"):
        if normalize:
            raise NotImplementedError("normalized=True is not used by GenVC (trainers/hifigan_trainer.py:109)")
        if filter_length & (filter_length - 1) or not (64 <= filter_length <= 4096) or win_length > filter_length:
            raise ValueError("filter_length must be a power of two in [64, 4096] and >= win_length")
        self.filter_length, self.hop_length, self.win_length = int(filter_length), int(hop_length), int(win_length)
        self.n_mel_channels, self.sampling_rate = int(n_mel_channels), int(sampling_rate)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("genvc_b200.mel.TorchMelSpectrogram runs on a CUDA device only (no CPU fallback)")
        self.lib = load_library()
        n = self.filter_length
        self.win_lo = (n - self.win_length) // 2
        self.win_hi = self.win_lo + self.win_length
        window = torch.zeros(n, dtype=torch.float64)
        window[self.win_lo: self.win_hi] = torch.hann_window(self.win_length, periodic=True, dtype=torch.float64)
        k = torch.arange(n, dtype=torch.float64) * (2.0 * math.pi / n)
        self._window = window.float().to(self.device)
        self._twiddle = torch.stack([torch.cos(k), torch.sin(k)], dim=1).float().contiguous().to(self.device)
        self._fbank = _melscale_fbanks(n // 2 + 1, float(mel_fmin), float(mel_fmax), self.n_mel_channels,
                                       self.sampling_rate).float().contiguous().to(self.device)
        if isinstance(mel_norm_file, str):
            mel_norm_file = torch.load(mel_norm_file, map_location="cpu")
        self.mel_norms: Optional[torch.Tensor] = None if mel_norm_file is None else mel_norm_file.float().contiguous().to(self.device)

    def to(self, device):
        if torch.device(device) != self.device:
            raise RuntimeError("construct the front-end on the target device")
        return self

    @torch.inference_mode()
    def forward(self, inp: torch.Tensor) -> torch.Tensor:
        if inp.dim() == 3:  # mono audio with a channel dimension (utils.py:145-148)
            inp = inp.squeeze(1)
        assert inp.dim() == 2
        x = inp.to(self.device, torch.float32).contiguous()
        B, N = (int(v) for v in x.shape)
        T = 1 + N // self.hop_length
        mel = torch.empty((B, self.n_mel_channels, T), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.genvc_mel_spectrogram(x.data_ptr(), B, N, self._window.data_ptr(), self._twiddle.data_ptr(), self._fbank.data_ptr(),
                                                self.mel_norms.data_ptr() if self.mel_norms is not None else None, mel.data_ptr(),
                                                self.filter_length, self.hop_length, self.win_lo, self.win_hi, self.n_mel_channels,
                                                1e-5, torch.cuda.current_stream(self.device).cuda_stream)
        if rc != 0:
            raise GenvcError(rc, "genvc_mel_spectrogram")
        return mel

    __call__ = forward
