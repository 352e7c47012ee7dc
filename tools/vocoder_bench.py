#!/usr/bin/env python
"""HiFi-GAN generator (csrc/vocoder.cu) on one GPU: ms per streaming chunk (8 tokens = 32 frames = 8192 samples) and per 1 s /
6 s segment and achieved GFLOP/s, then the content-DVAE tokeniser and the mel front-end.  (Parity against the oracles is
the tests' job: tests/test_hifigan.py, test_content_dvae.py, test_mel_frontend.py; this tool does not touch oracle/.)
    python tools/vocoder_bench.py"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from genvc_b200.synth import HIFIGAN_DEFAULTS, hifigan_conv_shapes, synth_hifigan_state
from genvc_b200.vocoder import HiFiGAN

ap = argparse.ArgumentParser()
ap.add_argument("--no-cpu", action="store_true", help="accepted for symmetry with bench.py; the tool never runs a CPU arm")
ap.add_argument("--reps", type=int, default=50)
a = ap.parse_args()
dev = torch.device("cuda:0")
cfg = dict(HIFIGAN_DEFAULTS)
sd = synth_hifigan_state(77)
v = HiFiGAN.from_config({}, device=dev).load_state_dict(sd)


def flops(T):
    """2 * MACs of one forward over T input frames."""
    total, t = 0, T
    shapes = {n: s for n, s, _ in hifigan_conv_shapes(cfg)}
    total += 2 * shapes["conv_pre"][0] * shapes["conv_pre"][1] * 7 * t
    for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        cin, cout, _ = shapes[f"ups.{i}"]
        t *= u
        total += 2 * cin * cout * (k // u) * t
        for n, s in shapes.items():
            if n.startswith("resblocks.") and int(n.split(".")[1]) // 3 == i:
                total += 2 * s[0] * s[1] * s[2] * t
    total += 2 * shapes["conv_post"][1] * 7 * t
    return total


out = {}
for label, T in (("chunk_8_tokens", 32), ("segment_1s", 94), ("segment_6s", 563)):
    x = torch.randn(1, 1024, T, device=dev)
    for _ in range(3):
        v(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        v(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    r = {"frames": T, "samples": T * 256, "ms": round(ms, 4), "gflops": round(flops(T) / ms / 1e6, 1),
         "audio_s": round(T * 256 / 24000, 3), "rtf": round(ms / 1e3 / (T * 256 / 24000), 6)}
    out[label] = r
out["launches_per_forward"] = 1 + 3 + 18 + 1
print(json.dumps(out))

# ---- content-DVAE tokeniser (the stage before the path)
from genvc_b200.content_dvae import DiscreteVAE
from genvc_b200.synth import synth_dvae_state
dv = DiscreteVAE.from_config({}, device=dev).load_state_dict(synth_dvae_state(55))
dres = {}
for label, T in (("segment_1s", 50), ("segment_6s", 300)):
    x = torch.randn(1, 256, T, device=dev)
    for _ in range(3):
        dv.get_codebook_indices(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        dv.get_codebook_indices(x)
    e1.record()
    torch.cuda.synchronize()
    dres[label] = {"frames": T, "codes": (T + 3) // 4, "ms": round(e0.elapsed_time(e1) / a.reps, 4)}
print(json.dumps({"content_dvae": dres}))

# ---- mel front-end of the conditioning path (style encoder: n_fft 2048, hop 256, win 1024, 24 kHz)
from genvc_b200.mel import TorchMelSpectrogram
mf = TorchMelSpectrogram(filter_length=2048, hop_length=256, win_length=1024, sampling_rate=24000, n_mel_channels=80, device=dev)
mres = {}
for label, n in (("audio_6s", 144000), ("audio_30s", 720000)):
    w = torch.randn(1, n, device=dev) * 0.1
    for _ in range(3):
        mf(w)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        mf(w)
    e1.record()
    torch.cuda.synchronize()
    r = {"samples": n, "frames": 1 + n // 256, "ms": round(e0.elapsed_time(e1) / a.reps, 4)}
    mres[label] = r
print(json.dumps({"mel_frontend": mres}))
