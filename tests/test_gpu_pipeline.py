"""Parity of the pieces either side of the token loop:

* a1 ``get_gpt_cond_latents`` (trainers/hifigan_trainer.py:438-455): 30 s clip, 6 s chunks, chunks under 0.33 s
  dropped, mean over the per-chunk perceiver outputs -- host logic on CPU against a line-by-line restatement, and on
  the GPU (full-size perceiver, 2..5 chunks incl. the 563-frame chunk a 6 s reference produces) against the oracle;
* the on-device Philox draw (``exp_noise`` = NULL) follows the post-warper distribution (chi-square over 20 k draws);
* the pipeline drivers (inference/inference_utils.py:23-89, 135-217) on a ``GenVCModel`` whose ``.gpt`` is the CUDA ``GPT``
  and whose content / vocoder stages are deterministic stubs: ids and latents equal to the oracle's for the same codes.
"""
import math

import pytest
import torch

from conftest import golden_checkpoint, load_golden

SR = 24000


class StubMel:
    """Deterministic stand-in for the mel front-end (utils.py TorchMelSpectrogram is outside the path): hop 256,
    80 fixed random projections of the frame, O(1) values -- [1, 1, n] audio -> [1, 80, n // 256 + 1]."""

    def __init__(self):
        g = torch.Generator().manual_seed(5)
        self.proj = torch.randn(80, 256, generator=g) / 16.0

    def __call__(self, audio):
        x = audio.reshape(-1).cpu()
        n = x.numel() // 256 + 1
        x = torch.nn.functional.pad(x, (0, n * 256 - x.numel()))
        return torch.tanh(self.proj @ x.view(n, 256).t())[None]


def reference_chunks(audio, sr, length=30, chunk_length=6):
    """The chunk plan of trainers/hifigan_trainer.py:440-449, restated: list of audio chunks that reach the perceiver."""
    out = []
    if audio.shape[1] > sr * length:
        audio = audio[:, : sr * length]
    for i in range(0, audio.shape[1], sr * chunk_length):
        chunk = audio[:, i: i + sr * chunk_length]
        if chunk.size(-1) < sr * 0.33:
            continue
        out.append(chunk)
    return out


@pytest.mark.parametrize("seconds,n_chunks", [(3.0, 1), (6.0, 1), (12.2, 2), (12.4, 3), (17.0, 3), (31.5, 5), (30.1, 5)])
def test_cond_latent_chunk_plan_matches_reference(seconds, n_chunks):
    """CPU: clip / chunk / drop / mean of GenVCModel.get_gpt_cond_latents with a recording stand-in for the perceiver."""
    from genvc_b200.inference.model_init import GenVCModel

    audio = torch.randn(1, int(seconds * SR), generator=torch.Generator().manual_seed(1))
    m = GenVCModel.__new__(GenVCModel)  # host logic only: no engine
    m.device = torch.device("cpu")
    m.torch_mel_spectrogram_style_encoder = StubMel()
    seen = []

    class RecordingGPT:
        def get_style_emb(self, mel, seq_lens=None):
            seen.append(mel.clone())
            return mel[:, :4, :3] * 1.0 + len(seen)  # [1, 4, 3], different per chunk

    m.gpt = RecordingGPT()
    out = m.get_gpt_cond_latents(audio, SR)
    chunks = reference_chunks(audio, SR)
    assert len(chunks) == n_chunks == len(seen)
    stub = StubMel()
    expect = []
    for k, ch in enumerate(chunks):
        mel = stub(ch.unsqueeze(0))
        assert torch.equal(seen[k], mel)
        expect.append(mel[:, :4, :3] + (k + 1))
    assert torch.equal(out, torch.stack(expect).mean(dim=0).transpose(1, 2))


def test_cond_latents_need_one_chunk():
    from genvc_b200.inference.model_init import GenVCModel

    m = GenVCModel.__new__(GenVCModel)
    m.device = torch.device("cpu")
    m.torch_mel_spectrogram_style_encoder = StubMel()
    m.gpt = None
    with pytest.raises(ValueError):
        m.get_gpt_cond_latents(torch.zeros(1, int(0.2 * SR)), SR)  # the only chunk is under 0.33 s


# ------------------------------------------------------------------------------------------------------- GPU
def _model(fx, device, max_batch=1):
    from genvc_b200.inference.model_init import model_from_checkpoint

    ck = golden_checkpoint(fx)
    model, cfg = model_from_checkpoint(ck, device, max_batch=max_batch)
    return model, cfg, ck


@pytest.mark.gpu
@pytest.mark.parametrize("seconds", [12.4, 17.0, 31.5])
def test_cond_latents_multi_chunk_matches_oracle(seconds, cuda_device):
    """a1 on the GPU: 3 / 3 / 5 chunks (6 s chunks are 563 mel frames; tails 0.4 s, 5 s, clipped) through the full-size
    perceiver; mean over chunks against the oracle's restatement on the same mels (2e-4 abs + 1e-4 rel, as the
    single-chunk perceiver tests)."""
    from oracle.genvc_oracle import load_oracle

    fx = load_golden("full_h4_cfg1")
    model, _, ck = _model(fx, cuda_device)
    model.torch_mel_spectrogram_style_encoder = StubMel()
    audio = torch.randn(1, int(seconds * SR), generator=torch.Generator().manual_seed(2)) * 0.3
    got = model.get_gpt_cond_latents(audio.to(cuda_device), SR)
    stub = StubMel()
    mels = [stub(ch.unsqueeze(0)) for ch in reference_chunks(audio, SR)]
    assert max(m.shape[-1] for m in mels) == 563
    ref = load_oracle(ck).get_gpt_cond_latents(mels)
    assert got.shape == ref.shape == (1, 32, 1024)
    err = (got.cpu() - ref).abs()
    assert bool((err <= 2e-4 + 1e-4 * ref.abs()).all()), f"max err {err.max().item()}"


@pytest.mark.gpu
@pytest.mark.parametrize("top_k,top_p", [(15, 0.85), (0, 1.0)])
def test_philox_draw_follows_post_warper_distribution(top_k, top_p, cuda_device):
    """The default product path samples with on-device Philox noise (exp_noise = NULL).  20 000 first-step draws
    (different seeds, same logits) against the post-warper probabilities the oracle computes from the same logits:
    no draw outside the kept set, chi-square over the kept tokens below the 99.9 % quantile."""
    from genvc_b200.engine import Sampling
    from oracle.genvc_oracle import SamplingParams, process_logits

    fx = load_golden("toy_d128_topk20")
    from test_gpu_parity import make_gpt

    g = make_gpt(fx, cuda_device)
    eng = g.engine
    cond = fx["style_emb"].transpose(1, 2).contiguous().to(cuda_device)
    g.compute_embeddings(cond, fx["codes"].to(cuda_device))
    P = g._prefix.shape[1]
    n_draws = 20000
    draws = torch.empty(n_draws, dtype=torch.int64, device=cuda_device)
    logits0 = None
    # every draw re-runs prefill (cheap at toy size) and takes the first token from the prefill's logits
    for i in range(n_draws):
        eng.prefill(g._prefix)
        sp = Sampling(top_k=top_k, top_p=top_p, temperature=0.85, repetition_penalty=2.0, max_new_tokens=1, seed=1000003 * i + 17)
        ch = eng.decode(1, sp, want_logits=(i == 0), mode=2)
        if i == 0:
            logits0 = ch.logits[0, 0].cpu()
        draws[i] = ch.ids[0, 0]  # stays on the device queue; one sync at the end
    torch.cuda.synchronize()
    draws = draws.cpu()
    fake = torch.full((1, P + 1), 1, dtype=torch.long)
    fake[:, -1] = 1024
    scores = process_logits(fake, logits0[None], SamplingParams(top_k=top_k, top_p=top_p, temperature=0.85, repetition_penalty=2.0))
    p = torch.softmax(scores, -1)[0].double()
    kept = p > 0
    counts = torch.bincount(draws, minlength=p.numel()).double()
    assert counts[~kept].sum() == 0, "a token outside the top-k / top-p set was drawn"
    # chi-square over the kept tokens, cells with expectation < 5 pooled
    exp = p[kept] * n_draws
    obs = counts[kept]
    big = exp >= 5
    chi = float((((obs[big] - exp[big]) ** 2) / exp[big]).sum())
    dof = int(big.sum()) - 1
    if (~big).any():
        e_small, o_small = float(exp[~big].sum()), float(obs[~big].sum())
        if e_small > 0:
            chi += (o_small - e_small) ** 2 / e_small
            dof += 1
    # Wilson-Hilferty 99.9 % quantile of chi-square(dof)
    z = 3.0902
    q = dof * (1 - 2 / (9 * dof) + z * math.sqrt(2 / (9 * dof))) ** 3 if dof > 0 else 10.83
    assert chi < q, f"chi-square {chi:.1f} over {dof} dof exceeds {q:.1f}"
    assert int(kept.sum()) > 1


class _StubStages:
    """Deterministic content / vocoder stand-ins (ContentVec, content-DVAE and HiFi-GAN are outside the path)."""

    @staticmethod
    def extract_content_features(wav):
        # 50 Hz frames of the 16 kHz source, 3 features
        return wav[:, ::320][:, :, None].repeat(1, 1, 3)

    @staticmethod
    def get_codebook_indices(feat):
        x = (feat[:, 0, :].abs() * 1000).long() % 256
        return x[:, ::4].contiguous()  # 12.5 codes per second

    @staticmethod
    def hifigan(mel):
        return torch.tanh(mel.repeat_interleave(64, dim=-1).sum(dim=1, keepdim=True) * 0.05)


@pytest.mark.gpu
def test_drivers_on_cuda_gpt_match_oracle(cuda_device):
    """SURVEY §4 layer 5: synthesize_utt / synthesize_utt_streaming (the re-hosted drivers, equal to the reference's
    on a stand-in model: tests/test_pipeline_drivers.py) driving the REAL CUDA GPT with stub content / vocoder stages.
    Greedy, so the result is deterministic: the waveform built from the oracle's latents for the same segments must match."""
    from genvc_b200.inference import inference_utils as drv
    from oracle.genvc_oracle import SamplingParams, load_oracle

    fx = load_golden("toy_d128_eos")  # EOS fires after a few dozen tokens: short, tie-free greedy runs
    model, cfg, ck = _model(fx, cuda_device)
    stages = _StubStages()
    model.content_extractor = stages
    model.content_dvae = stages
    model.hifigan = stages.hifigan
    model.torch_mel_spectrogram_style_encoder = StubMel()
    cfg.top_k, cfg.top_p, cfg.temperature, cfg.repetition_penalty = 1, 0.85, 0.85, 2.0
    g = torch.Generator().manual_seed(3)
    src = torch.randn(1, int(7.3 * 16000), generator=g)  # two segments: 6 s + 1.3 s
    tgt = torch.randn(1, int(7.0 * SR), generator=g) * 0.3  # two reference chunks
    wav = drv.synthesize_utt(model, src.clone(), tgt.clone())
    wav_fold = drv.synthesize_utt(model, src.clone(), tgt.clone(), reuse_decode_latents=True)
    pieces = []
    wav_stream = drv.synthesize_utt_streaming(model, src.clone(), tgt.clone(), stream_chunk_size=8, on_chunk=pieces.append)

    # oracle: same plan, same stub stages, CPU
    o = load_oracle(ck)
    stub = StubMel()
    cond = o.get_gpt_cond_latents([stub(ch.unsqueeze(0)) for ch in reference_chunks(tgt, SR)])
    sp = SamplingParams(top_k=1, top_p=0.85, temperature=0.85, repetition_penalty=2.0)
    lat2, lat1, all_ids = [], [], []
    for start, end, pad in drv.plan_segments(src.shape[-1], 96000, 5120):
        seg = torch.nn.functional.pad(src[:, start:end], (0, pad))
        codes = stages.get_codebook_indices(stages.extract_content_features(seg).transpose(1, 2))
        ids, lats = o.generate(cond, codes, sp)
        keep = ids[0] != 1025
        all_ids.append(ids[0][keep])
        lat1.append(lats[:, keep])
        lat2.append(o.forward_latents(codes, ids[0][keep][None], cond))

    def vocode(lat):
        mel = torch.nn.functional.interpolate(lat.transpose(1, 2), scale_factor=[4], mode="linear").squeeze(1)
        return stages.hifigan(mel)[0].squeeze()

    ref_wav = vocode(torch.cat(lat2, dim=1))
    assert wav.shape == ref_wav.shape
    assert (wav.cpu() - ref_wav).abs().max() < 2e-3  # tanh of sums of ~1e-5-accurate latents
    ref_fold = vocode(torch.cat(lat1, dim=1))
    assert wav_fold.shape == ref_fold.shape
    assert (wav_fold.cpu() - ref_fold).abs().max() < 2e-3
    assert wav_stream.ndim == 1 and len(pieces) >= 2 and wav_stream.numel() == sum(p.numel() for p in pieces)
    assert model.last_latency_s > 0 and model.last_rtf > 0


@pytest.mark.gpu
def test_batched_segments_driver_matches_serial(cuda_device):
    """the equal-length segments of one utterance decoded together (rows share one pass of the weight stream
    in the batched fused kernel) give the serial driver's result: greedy ids are identical row for row, so the waveforms
    built from the latents agree to rounding; with and without the teacher-forced second pass."""
    from genvc_b200.inference import inference_utils as drv

    fx = load_golden("toy_d128_eos")
    model, cfg, ck = _model(fx, cuda_device, max_batch=4)
    stages = _StubStages()
    model.content_extractor = stages
    model.content_dvae = stages
    model.hifigan = stages.hifigan
    model.torch_mel_spectrogram_style_encoder = StubMel()
    cfg.top_k, cfg.top_p, cfg.temperature, cfg.repetition_penalty = 1, 0.85, 0.85, 2.0
    g = torch.Generator().manual_seed(4)
    src = torch.randn(1, int(20.5 * 16000), generator=g)  # three 6 s segments (decoded as one batch of 3) + 2.5 s
    tgt = torch.randn(1, int(4.0 * SR), generator=g) * 0.3
    for fold in (False, True):
        a = drv.synthesize_utt(model, src.clone(), tgt.clone(), reuse_decode_latents=fold)
        b = drv.synthesize_utt(model, src.clone(), tgt.clone(), reuse_decode_latents=fold, batch_segments=True)
        assert a.shape == b.shape, (a.shape, b.shape)
        assert (a - b).abs().max() < 2e-3


@pytest.mark.gpu
def test_pipeline_with_every_cuda_stage_matches_oracle(cuda_device):
    """The whole conversion on the device except ContentVec (fairseq, absent: a stub): CUDA mel front-end -> perceiver ->
    content-DVAE tokeniser -> GPT prefill / fused decode -> x4 interpolation -> CUDA HiFi-GAN generator, driven by
    synthesize_utt / synthesize_utt_streaming, against the CPU oracles of every stage chained the same way (greedy)."""
    from genvc_b200.content_dvae import DiscreteVAE
    from genvc_b200.inference import inference_utils as drv
    from genvc_b200.mel import TorchMelSpectrogram
    from genvc_b200.synth import HIFIGAN_DEFAULTS, synth_dvae_state, synth_hifigan_state
    from genvc_b200.vocoder import HiFiGAN
    from oracle.dvae_oracle import get_codebook_indices
    from oracle.genvc_oracle import SamplingParams, load_oracle
    from oracle.hifigan_oracle import vocode
    from oracle.mel_oracle import log_mel

    fx = load_golden("toy_d128_eos")
    model, cfg, ck = _model(fx, cuda_device)
    dv_cfg = dict(channels=3, num_tokens=256, codebook_dim=32, hidden_dim=16, num_resnet_blocks=1)
    dv_sd = synth_dvae_state(21, **dv_cfg)
    hf_cfg = dict(input_feat_dim=128, upsample_initial_channel=32)
    hf_sd = synth_hifigan_state(22, **hf_cfg)
    model.content_extractor = _StubStages()  # ContentVec stand-in: [B, T50Hz, 3]
    model.content_dvae = DiscreteVAE(positional_dims=1, kernel_size=3, num_layers=2, use_transposed_convs=False, device=cuda_device,
                                     **dv_cfg).load_state_dict(dv_sd)
    arch = dict(HIFIGAN_DEFAULTS, **hf_cfg)
    model.hifigan = HiFiGAN(arch["input_feat_dim"], arch["upsample_initial_channel"], arch["resblock_kernel_sizes"],
                            arch["resblock_dilation_sizes"], arch["upsample_rates"], arch["upsample_kernel_sizes"], arch["resblock_type"],
                            device=cuda_device).load_state_dict(hf_sd)
    model.torch_mel_spectrogram_style_encoder = TorchMelSpectrogram(filter_length=2048, hop_length=256, win_length=1024, sampling_rate=SR,
                                                                    n_mel_channels=80, device=cuda_device)
    cfg.top_k, cfg.top_p, cfg.temperature, cfg.repetition_penalty = 1, 0.85, 0.85, 2.0
    g = torch.Generator().manual_seed(9)
    src = torch.randn(1, int(7.1 * 16000), generator=g)
    tgt = torch.randn(1, int(6.6 * SR), generator=g) * 0.3
    wav = drv.synthesize_utt(model, src.clone(), tgt.clone())
    pieces = []
    wav_stream = drv.synthesize_utt_streaming(model, src.clone(), tgt.clone(), stream_chunk_size=8, on_chunk=pieces.append)

    o = load_oracle(ck)
    cond = o.get_gpt_cond_latents([log_mel(ch.unsqueeze(0), 2048, 256, 1024, 80, 0, 8000, SR) for ch in reference_chunks(tgt, SR)])
    sp = SamplingParams(top_k=1, top_p=0.85, temperature=0.85, repetition_penalty=2.0)
    lat2 = []
    hf_arch = {k: arch[k] for k in ("resblock_kernel_sizes", "resblock_dilation_sizes", "upsample_rates", "upsample_kernel_sizes", "resblock_type")}
    for start, end, pad in drv.plan_segments(src.shape[-1], 96000, 5120):
        seg = torch.nn.functional.pad(src[:, start:end], (0, pad))
        feat = _StubStages.extract_content_features(seg)
        codes = get_codebook_indices(dv_sd, feat.transpose(1, 2), num_layers=2, num_resnet_blocks=1, kernel_size=3)
        ids, _ = o.generate(cond, codes, sp)
        keep = ids[0] != 1025
        lat2.append(o.forward_latents(codes, ids[0][keep][None], cond))
    ref_wav = vocode(hf_sd, torch.cat(lat2, dim=1), 4.0, **hf_arch)[0].squeeze()
    assert wav.shape == ref_wav.shape
    assert (wav.cpu() - ref_wav).abs().max() < 2e-3
    assert wav_stream.ndim == 1 and len(pieces) >= 2
