"""Pipeline drivers (genvc_b200/inference/inference_utils.py) against the reference's own drivers
(inference/inference_utils.py) on the same duck-typed model: segment plan, padding, cross-fade, streaming flushes.

CPU only.  The GPT and the out-of-path stages are deterministic stand-ins, so what is compared is exactly the
host logic the drivers add.  The reference half needs /root/reference (absent on the GPU box: skipped there)."""
import importlib.util
import os
import types

import pytest
import torch

from genvc_b200.inference import inference_utils as mine

REF = "/root/reference/inference/inference_utils.py"


def _ref_module():
    if not os.path.exists(REF):
        pytest.skip("reference tree not present")
    spec = importlib.util.spec_from_file_location("ref_inference_utils", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _NS(types.SimpleNamespace):
    pass


class FakeGPT:
    """Deterministic stand-in with the path's interface: token count and values depend only on the content codes."""
    stop_audio_token = 1025
    D = 16

    def _tokens(self, codes):
        n = int(codes.shape[-1]) * 2 + 1  # includes one trailing EOS
        g = torch.Generator().manual_seed(int(codes.sum()) % 9973)
        t = torch.randint(0, 1024, (n,), generator=g)
        t[-1] = self.stop_audio_token
        return t

    def _latents(self, toks):
        return torch.sin(toks.float()[:, None] * 0.01 + torch.arange(self.D)[None, :] * 0.3)

    def generate(self, cond_latents, text_inputs, **kw):
        toks = self._tokens(text_inputs)
        self.last_latents = self._latents(toks)[None]
        return toks[None]

    def __call__(self, text_inputs, text_lengths, audio_codes, wav_lengths, cond_latents=None, return_latent=False):
        assert return_latent
        return self._latents(audio_codes[0])[None]

    def compute_embeddings(self, cond_latents, text_inputs):
        self._pending = text_inputs
        return torch.ones(1, 4, dtype=torch.long)

    def get_generator(self, fake_inputs, **kw):
        toks = self._tokens(self._pending)
        lats = self._latents(toks)
        for i in range(toks.shape[0]):
            yield toks[i:i + 1], lats[i:i + 1]


class FakeModel:
    def __init__(self):
        self.device = torch.device("cpu")
        self.content_sample_rate = 16000
        self.hifigan_scale_factor = 4
        self.config = _NS(top_p=0.85, top_k=15, temperature=0.75, length_penalty=1.0, repetition_penalty=2.0,
                          audio=_NS(sample_rate=24000), model_args=_NS(gpt_code_stride_len=1024))
        self.gpt = FakeGPT()
        self.content_extractor = _NS(extract_content_features=lambda wav: wav[:, ::320][:, :, None].repeat(1, 1, 3))
        self.content_dvae = _NS(get_codebook_indices=lambda f: (f[:, 0, :].abs() * 1000).long() % 256)
        hif = lambda mel: torch.tanh(mel.repeat_interleave(64, dim=-1).sum(dim=1, keepdim=True) * 0.1)  # noqa: E731
        self.hifigan = _Callable(hif)

    def get_gpt_cond_latents(self, audio, sr):
        return audio[:, :32 * 16].reshape(1, 32, 16)

    def inference(self, src_audio, cond_latent, **kw):
        feat = self.content_extractor.extract_content_features(src_audio)
        codes = self.content_dvae.get_codebook_indices(feat.transpose(1, 2))
        gen = self.gpt.generate(cond_latent, codes, **kw)[0]
        gen = gen[(gen != self.gpt.stop_audio_token).nonzero().squeeze()]
        lat = self.gpt(codes, None, gen.unsqueeze(0), None, cond_latents=cond_latent, return_latent=True)
        mel = torch.nn.functional.interpolate(lat.transpose(1, 2), scale_factor=[self.hifigan_scale_factor], mode="linear").squeeze(1)
        return self.hifigan(mel)


class _Callable:
    def __init__(self, f):
        self._f = f

    def __call__(self, x):
        return self._f(x)

    def forward(self, x):
        return self._f(x)


def _inputs(seconds, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(1, int(seconds * 16000), generator=g), torch.randn(1, 24000, generator=g)


def test_plan_segments_edges():
    seg, mn = 96000, 5120
    assert mine.plan_segments(96000, seg, mn) == [(0, 96000, 0)]                      # exactly one window
    assert mine.plan_segments(96001, seg, mn) == [(0, 96000, 0), (96000, 96001, 5119)]  # 1-sample tail, padded
    assert mine.plan_segments(100, seg, mn) == [(0, 100, 5020)]                       # shorter than the minimum
    assert mine.plan_segments(0, seg, mn) == []
    p = mine.plan_segments(3 * 96000 + 7000, seg, mn)
    assert [e - s for s, e, _ in p] == [96000, 96000, 96000, 7000] and p[-1][2] == 0


@pytest.mark.parametrize("n_prev,n", [(None, 5000), (3000, 5000), (3000, 1500), (3000, 2048)])
def test_handle_chunks_matches_reference(n_prev, n):
    ref = _ref_module()
    g = torch.Generator().manual_seed(3)
    wav = torch.randn(n, generator=g)
    prev = None if n_prev is None else torch.randn(n_prev, generator=g)
    ov = None if prev is None else prev[-1024:].clone()
    a = ref.handle_chunks(wav.clone(), prev, None if ov is None else ov.clone(), 1024)
    b = mine.handle_chunks(wav.clone(), prev, None if ov is None else ov.clone(), 1024)
    for x, y in zip(a, b):
        assert (x is None) == (y is None)
        if x is not None:
            assert torch.equal(x, y)


@pytest.mark.parametrize("seconds", [1.0, 6.0, 6.2, 13.3])
def test_synthesize_utt_matches_reference(seconds):
    ref = _ref_module()
    src, tgt = _inputs(seconds)
    a = ref.synthesize_utt(FakeModel(), src.clone(), tgt.clone())
    b = mine.synthesize_utt(FakeModel(), src.clone(), tgt.clone())
    assert torch.equal(a, b)
    # latents straight from decode: the stand-in's two sources are identical, so is the waveform
    c = mine.synthesize_utt(FakeModel(), src.clone(), tgt.clone(), reuse_decode_latents=True)
    assert torch.equal(a, c)


@pytest.mark.parametrize("seconds", [1.0, 6.2, 13.3])
def test_synthesize_utt_chunked_matches_reference(seconds):
    ref = _ref_module()
    src, tgt = _inputs(seconds, seed=1)
    a = ref.synthesize_utt_chunked(FakeModel(), src.clone(), tgt.clone())
    b = mine.synthesize_utt_chunked(FakeModel(), src.clone(), tgt.clone())
    assert torch.equal(a, b)


@pytest.mark.parametrize("seconds,chunk", [(1.0, 8), (6.2, 8), (13.3, 5), (2.0, 0)])
def test_synthesize_utt_streaming_matches_reference(seconds, chunk, capsys):
    ref = _ref_module()
    src, tgt = _inputs(seconds, seed=2)
    a = ref.synthesize_utt_streaming(FakeModel(), src.clone(), tgt.clone(), stream_chunk_size=chunk)
    got = []
    m = FakeModel()
    b = mine.synthesize_utt_streaming(m, src.clone(), tgt.clone(), stream_chunk_size=chunk, on_chunk=got.append)
    assert torch.equal(a, b)
    assert torch.equal(torch.cat(got, dim=-1), b)
    assert m.last_latency_s >= 0.0 and m.last_rtf > 0.0
    out = capsys.readouterr().out
    assert out.count("Latency:") == 2 and out.count("Real-time factor:") == 2


def test_streaming_token_count_multiple_of_chunk():
    """Token count a multiple of the chunk size: the reference's end-of-segment flush raises on the empty list
    (inference_utils.py:196); ours just has nothing left to flush."""
    m = FakeModel()
    src, tgt = _inputs(1.0, seed=4)
    n_tok = int(m.gpt._tokens(m.content_dvae.get_codebook_indices(
        m.content_extractor.extract_content_features(src).transpose(1, 2))).shape[0])
    out = mine.synthesize_utt_streaming(m, src, tgt, stream_chunk_size=n_tok)
    assert out.numel() > 0
