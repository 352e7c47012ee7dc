"""Mel front-end of the conditioning path (utils.py:95-158 TorchMelSpectrogram): oracle vs torchaudio-generated fixtures
(tests/golden/make_golden_mel.py), CUDA kernel vs the same fixtures."""
import os

import pytest
import torch

from oracle.mel_oracle import log_mel

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["mel_style_24k", "mel_default_22k"]
# log-mel values (range about -11 .. +9).  fp32 FFT (fixture) vs fp32 direct summation (kernel) differ in the last bits of the
# power spectrum; bins far below the strongest one are noise in both, hence an absolute bar on the LOG values.
LOGMEL_ATOL = 2e-3


def load(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"))


def oracle_args(a):
    return dict(n_fft=a["filter_length"], hop=a["hop_length"], win=a["win_length"], n_mels=a["n_mel_channels"], f_min=a["mel_fmin"],
                f_max=a["mel_fmax"], sample_rate=a["sampling_rate"])


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_torchaudio_fixture(name):
    fx = load(name)
    mel = log_mel(fx["x"], mel_norms=fx["norms"], **oracle_args(fx["args"]))
    assert mel.shape == fx["mel"].shape
    assert float((mel - fx["mel"]).abs().max()) < 2e-4


def test_oracle_matches_torchaudio_live():
    torchaudio = pytest.importorskip("torchaudio")
    x = torch.randn(1, 5000, generator=torch.Generator().manual_seed(0)) * 0.1
    tr = torchaudio.transforms.MelSpectrogram(n_fft=2048, hop_length=256, win_length=1024, power=2, normalized=False, sample_rate=24000,
                                              f_min=0, f_max=8000, n_mels=80, norm="slaney")
    ref = torch.log(torch.clamp(tr(x), min=1e-5))
    assert float((log_mel(x, 2048, 256, 1024, 80, 0, 8000, 24000) - ref).abs().max()) < 2e-4


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_mel_matches_fixture(name, cuda_device):
    from genvc_b200.mel import TorchMelSpectrogram
    fx = load(name)
    m = TorchMelSpectrogram(mel_norm_file=fx["norms"], device=cuda_device, **fx["args"])
    mel = m(fx["x"].to(cuda_device))
    assert mel.shape == fx["mel"].shape
    err = (mel.cpu() - fx["mel"]).abs()
    assert float(err.max()) < LOGMEL_ATOL, f"log-mel max err {float(err.max())} (mean {float(err.mean())})"
    assert torch.equal(m(fx["x"].unsqueeze(1).to(cuda_device)), mel)  # [B, 1, N] is squeezed (utils.py:145-148)


@pytest.mark.gpu
def test_cond_latents_with_cuda_mel_frontend(cuda_device):
    """a1 end to end on the device: audio -> CUDA mel front-end -> perceiver, against the oracle fed with the oracle's mels."""
    from genvc_b200.mel import TorchMelSpectrogram
    from oracle.genvc_oracle import load_oracle
    from test_gpu_pipeline import _model, load_golden, reference_chunks, SR
    fx = load_golden("full_h4_cfg1")
    model, _, ck = _model(fx, cuda_device)
    model.torch_mel_spectrogram_style_encoder = TorchMelSpectrogram(filter_length=2048, hop_length=256, win_length=1024, sampling_rate=SR,
                                                                    mel_fmin=0, mel_fmax=8000, n_mel_channels=80, device=cuda_device)
    audio = torch.randn(1, int(7.3 * SR), generator=torch.Generator().manual_seed(4)) * 0.2
    got = model.get_gpt_cond_latents(audio.to(cuda_device), SR)
    mels = [log_mel(ch.unsqueeze(0), 2048, 256, 1024, 80, 0, 8000, SR) for ch in reference_chunks(audio, SR)]
    ref = load_oracle(ck).get_gpt_cond_latents(mels)
    err = (got.cpu() - ref).abs()
    assert bool((err <= 5e-4 + 1e-3 * ref.abs()).all()), f"max err {err.max().item()}"
