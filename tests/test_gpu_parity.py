"""GPU parity: the CUDA path (through the C ABI, via the drop-in ``GPT`` object) against
(a) the committed golden fixtures the REAL reference generated and (b) the CPU oracle on
the same seeded inputs.

Bars: token ids bit-exact; teacher-forced logits within 1e-3 abs + 1e-4 rel;
latents within 1e-4 abs + 1e-4 rel (two fp32 evaluations with different summation orders;
observed ~1e-5).
"""
import pytest
import torch

from conftest import golden_checkpoint, load_golden

pytestmark = pytest.mark.gpu

LOGIT_ATOL, LOGIT_RTOL = 1e-3, 1e-4
LAT_ATOL, LAT_RTOL = 1e-4, 1e-4
TOY = ["toy_d128_greedy", "toy_d128_topk20", "toy_d128_topk0_topp1", "toy_d128_eos", "toy_d256_h4_greedy",
       "toy_d512_h2_greedy"]
FULL = ["full_h4_cfg1", "full_h16_greedy", "full_h4_topk20", "full_h4_topk15_mel563"]
TOY_BATCH = ["toy_d128_batch3", "toy_d128_batch8_eos", "toy_d256_batch5_topk20"]
FULL_BATCH = ["full_h4_batch4_greedy", "full_h4_batch8_topk20", "full_h16_batch8_greedy", "full_large_seed4321"]

_GPT_CACHE = {}


def make_gpt(fx, device, max_batch=1):
    from genvc_b200.config import GenVCDims
    from genvc_b200.gpt import GPT

    key = (tuple(sorted(fx["model"].items())), max_batch)
    if key not in _GPT_CACHE:
        if len(_GPT_CACHE) >= 2:
            _GPT_CACHE.clear()
            torch.cuda.empty_cache()
        ck = golden_checkpoint(fx)
        g = GPT(GenVCDims.from_config(ck["config"]), device=device, max_batch=max_batch)
        g.load_state_dict(ck["model"])
        g.eval().to(device).init_gpt_for_inference()
        _GPT_CACHE[key] = g
    return _GPT_CACHE[key]


def fixture_noise(fx, V=1026):
    if fx["noise_seed"] is None:
        return None
    cap = fx["new_tokens"] if fx["new_tokens"] is not None else 602
    g = torch.Generator().manual_seed(fx["noise_seed"])
    return torch.empty((cap, fx["ids"].shape[0], V)).exponential_(1, generator=g)


def close(a, b, atol, rtol):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    err = (a - b).abs()
    ok = bool((err <= atol + rtol * b.abs()).all())
    return ok, float(err.max())


def gen_kwargs(fx, **extra):
    s = fx["sampling"]
    kw = dict(do_sample=True, top_p=s["top_p"], top_k=s["top_k"], temperature=s["temperature"], num_beams=1,
              length_penalty=1.0, repetition_penalty=s["repetition_penalty"], output_attentions=False)
    if fx["new_tokens"] is not None:
        kw["max_new_tokens"] = fx["new_tokens"]
    kw.update(extra)
    return kw


# ------------------------------------------------------------------------------------ perceiver
@pytest.mark.parametrize("name", TOY + FULL)
def test_perceiver_matches_reference_fixture(name, cuda_device):
    fx = load_golden(name)
    g = make_gpt(fx, cuda_device)
    out = g.get_style_emb(fx["mel"].to(cuda_device))
    assert out.shape == fx["style_emb"].shape
    ok, err = close(out, fx["style_emb"], 2e-4, 1e-4)
    assert ok, f"perceiver max err {err}"
    out4 = g.get_style_emb(fx["mel"].unsqueeze(1).to(cuda_device))
    assert torch.equal(out4, out)


# ------------------------------------------------------------------------------------ generation
def _run_generate(fx, g, device, mode):
    cond = fx["style_emb"].transpose(1, 2).contiguous().to(device)
    noise = fixture_noise(fx)
    kw = gen_kwargs(fx, decode_mode=mode)
    if noise is not None:
        kw["exp_noise"] = noise.to(device)
    ids = g.generate(cond, fx["codes"].to(device), **kw)
    return ids, g.last_latents


@pytest.mark.parametrize("mode", [1, 2], ids=["per_op", "fused"])
@pytest.mark.parametrize("name", TOY)
def test_toy_token_ids_bit_exact(name, mode, cuda_device):
    fx = load_golden(name)
    g = make_gpt(fx, cuda_device)
    ids, lats = _run_generate(fx, g, cuda_device, mode)
    assert ids.dtype == torch.int64
    assert torch.equal(ids.cpu(), fx["ids"]), f"first mismatch at {(ids.cpu() != fx['ids']).nonzero()[:1].tolist()}"
    ok, err = close(lats[:, fx["steps"].to(lats.device)], fx["latents"], LAT_ATOL, LAT_RTOL)
    assert ok, f"latent max err {err}"


@pytest.mark.parametrize("mode", [1, 2], ids=["per_op", "fused"])
@pytest.mark.parametrize("name", TOY + FULL)
def test_teacher_forced_logits(name, mode, cuda_device):
    """Logits of every step with the reference's tokens forced: isolates numerics from sampling."""
    fx = load_golden(name)
    n = min(fx["ids"].shape[1], 96 if name in FULL else 10 ** 6)
    g = make_gpt(fx, cuda_device)
    eng = g.engine
    cond = fx["style_emb"].transpose(1, 2).contiguous().to(cuda_device)
    g.compute_embeddings(cond, fx["codes"].to(cuda_device))
    eng.prefill(g._prefix)
    from genvc_b200.engine import Sampling

    sp = Sampling(**fx["sampling"], max_new_tokens=n)
    forced = fx["ids"][:, :n].transpose(0, 1).contiguous().to(cuda_device)
    ch = eng.decode(n, sp, forced=forced, want_logits=True, mode=mode)
    emitted = ch.status.tolist()[0]
    assert emitted == n
    assert torch.equal(ch.ids.cpu(), forced.cpu())
    steps = fx["steps"][fx["steps"] < n]
    got = ch.logits.transpose(0, 1)[:, steps.to(cuda_device)]
    ok, err = close(got, fx["logits"][:, : len(steps)], LOGIT_ATOL, LOGIT_RTOL)
    assert ok, f"logit max err {err}"
    ok, err = close(ch.latents.transpose(0, 1)[:, steps.to(cuda_device)], fx["latents"][:, : len(steps)], LAT_ATOL, LAT_RTOL)
    assert ok, f"latent max err {err}"


@pytest.mark.parametrize("name,T,n", [("toy_d128_greedy", 400, 600), ("toy_d256_h4_greedy", 400, 600),
                                      ("toy_d512_h2_greedy", 300, 200), ("full_h4_cfg1", 240, 40), ("full_h16_greedy", 300, 24)])
def test_fused_long_context_matches_per_op(name, T, n, cuda_device):
    """Long prefixes (S up to ~1035 of S_max = 1088): the fused kernel's attention items span several 32-key blocks
    and up to 8 key ranges per head.  Forced random tokens; fused logits / latents against the per-op path (whose
    attention kernel is checked against torch at S = 333 and whose GEMMs are pinned by the fixtures)."""
    fx = load_golden(name)
    g = make_gpt(fx, cuda_device)
    eng = g.engine
    from genvc_b200.engine import Sampling

    gen = torch.Generator().manual_seed(99)
    codes = torch.randint(0, 256, (1, T), generator=gen).to(cuda_device)
    cond = fx["style_emb"].transpose(1, 2).contiguous().to(cuda_device)
    forced = torch.randint(0, 1024, (n, 1), generator=gen).to(cuda_device)
    sp = Sampling(**fx["sampling"], max_new_tokens=n)
    out = {}
    for mode in (1, 2):
        g.compute_embeddings(cond, codes)
        eng.prefill(g._prefix)
        ch = eng.decode(n, sp, forced=forced, want_logits=True, mode=mode)
        emitted = ch.status.tolist()[0]
        assert emitted == n
        assert torch.equal(ch.ids.cpu(), forced.cpu())
        out[mode] = (ch.logits.clone(), ch.latents.clone())
    ok, err = close(out[2][0], out[1][0], LOGIT_ATOL, LOGIT_RTOL)
    assert ok, f"logit max err {err}"
    ok, err = close(out[2][1], out[1][1], LAT_ATOL, LAT_RTOL)
    assert ok, f"latent max err {err}"


@pytest.mark.parametrize("name", TOY + ["full_h4_topk20", "full_h16_greedy"])
def test_projected_value_variant_matches_fixtures(name, cuda_device, monkeypatch):
    """The single-row fused kernel has two attention arrangements: K / V items, and (for generations of >= 96 tokens) a
    cache of PROJECTED values v_j . W_proj with scores-only items (attn c_proj is linear, so it commutes with the
    softmax-weighted sum).  Forced on for every fixture here: ids bit-exact, latents / teacher-forced logits in tolerance,
    including the lazy projection of the prefix rows at the first forward and a switch per-op -> fused mid-sequence."""
    from genvc_b200.config import GenVCDims
    from genvc_b200.engine import Sampling
    from genvc_b200.gpt import GPT

    monkeypatch.setenv("GENVC_VW", "1")  # the variant is opt-in
    monkeypatch.setenv("GENVC_VW_MIN_TOKENS", "1")
    fx = load_golden(name)
    ck = golden_checkpoint(fx)
    g = GPT(GenVCDims.from_config(ck["config"]), device=cuda_device)
    g.load_state_dict(ck["model"])
    g.eval().to(cuda_device).init_gpt_for_inference()
    assert g.engine.vw is not None
    ids, lats = _run_generate(fx, g, cuda_device, 2)
    assert torch.equal(ids.cpu(), fx["ids"]), f"first mismatch at {(ids.cpu() != fx['ids']).nonzero()[:1].tolist()}"
    ok, err = close(lats[:, fx["steps"].to(lats.device)], fx["latents"], LAT_ATOL, LAT_RTOL)
    assert ok, f"latent max err {err}"
    # teacher-forced: the first 5 steps through the per-op kernels (K / V caches only), the rest fused (the projected values
    # of everything cached so far are computed at its first forward)
    n = min(fx["ids"].shape[1], 48)
    eng = g.engine
    cond = fx["style_emb"].transpose(1, 2).contiguous().to(cuda_device)
    g.compute_embeddings(cond, fx["codes"].to(cuda_device))
    eng.prefill(g._prefix)
    sp = Sampling(**fx["sampling"], max_new_tokens=n)
    forced = fx["ids"][:, :n].transpose(0, 1).contiguous().to(cuda_device)
    k = min(5, n - 1)
    ch1 = eng.decode(k, sp, forced=forced[:k], want_logits=True, mode=1)
    ch2 = eng.decode(n - k, sp, forced=forced[k:], want_logits=True, mode=2)
    assert ch1.status.tolist()[0] == k and ch2.status.tolist()[0] == n - k
    logits = torch.cat([ch1.logits, ch2.logits], 0).transpose(0, 1)
    steps = fx["steps"][fx["steps"] < n]
    ok, err = close(logits[:, steps.to(cuda_device)], fx["logits"][:, : len(steps)], LOGIT_ATOL, LOGIT_RTOL)
    assert ok, f"logit max err {err}"
    del g
    torch.cuda.empty_cache()


@pytest.mark.parametrize("name", FULL)
def test_full_size_token_ids_bit_exact(name, cuda_device):
    """BASELINE configs[0] and friends at L=30, D=1024: free-running ids through the fused kernel."""
    fx = load_golden(name)
    g = make_gpt(fx, cuda_device)
    ids, _ = _run_generate(fx, g, cuda_device, 2)
    assert ids.shape == fx["ids"].shape
    assert torch.equal(ids.cpu(), fx["ids"]), f"first mismatch at {(ids.cpu() != fx['ids']).nonzero()[:1].tolist()}"


def test_full_size_per_op_prefix_matches(cuda_device):
    fx = load_golden("full_h4_cfg1")
    g = make_gpt(fx, cuda_device)
    cond = fx["style_emb"].transpose(1, 2).contiguous().to(cuda_device)
    ids = g.generate(cond, fx["codes"].to(cuda_device), **gen_kwargs(fx, decode_mode=1, max_new_tokens=40))
    assert torch.equal(ids.cpu(), fx["ids"][:, :40])


@pytest.mark.parametrize("mode", [1, 2, 0], ids=["per_op", "fused_batch", "auto"])
@pytest.mark.parametrize("name", TOY_BATCH)
def test_toy_batched_rows_equal_reference(name, mode, cuda_device):
    """A7 / layers/stream_generator.py:860-881: equal-T batches; rows finish at different steps and are padded with
    1025.  Fixtures come from the reference modules run on the whole batch; per-op kernels and the batched fused
    kernel (rows share one pass of the weight stream) must both reproduce ids bit-exactly."""
    fx = load_golden(name)
    B = fx["ids"].shape[0]
    g = make_gpt(fx, cuda_device, max_batch=B)
    out = g.get_style_emb(fx["mel"].to(cuda_device))
    ok, err = close(out, fx["style_emb"], 2e-4, 1e-4)
    assert ok, err
    assert g.engine.fused_rows(B)
    ids, lats = _run_generate(fx, g, cuda_device, mode)
    assert torch.equal(ids.cpu(), fx["ids"]), f"mode {mode}: first mismatch at {(ids.cpu() != fx['ids']).nonzero()[:1].tolist()}"
    ok, err = close(lats[:, fx["steps"].to(lats.device)], fx["latents"], LAT_ATOL, LAT_RTOL)
    assert ok, f"latent max err {err}"


@pytest.mark.parametrize("name", FULL_BATCH)
def test_full_size_batched_fused_ids_bit_exact(name, cuda_device):
    """BASELINE configs[2]/[3] shapes at L=30, D=1024 (B = 2, 4, 8; H = 4 and 16; greedy and top-k 20 with injected
    multinomial noise; second-seed "large" checkpoint): free-running ids of the batched fused kernel against
    fixtures generated by the reference modules on the whole batch."""
    fx = load_golden(name)
    B = fx["ids"].shape[0]
    g = make_gpt(fx, cuda_device, max_batch=B)
    ids, lats = _run_generate(fx, g, cuda_device, 2)
    assert ids.shape == fx["ids"].shape
    assert torch.equal(ids.cpu(), fx["ids"]), f"first mismatch at {(ids.cpu() != fx['ids']).nonzero()[:1].tolist()}"
    ok, err = close(lats[:, fx["steps"].to(lats.device)], fx["latents"], LAT_ATOL, LAT_RTOL)
    assert ok, f"latent max err {err}"


@pytest.mark.parametrize("mode", [1, 2], ids=["per_op", "fused_batch"])
@pytest.mark.parametrize("name", TOY_BATCH + FULL_BATCH)
def test_batched_teacher_forced_logits(name, mode, cuda_device):
    """Logits and latents of every row at every step with the reference's tokens forced."""
    fx = load_golden(name)
    B, n = fx["ids"].shape
    g = make_gpt(fx, cuda_device, max_batch=B)
    eng = g.engine
    cond = fx["style_emb"].transpose(1, 2).contiguous().to(cuda_device)
    g.compute_embeddings(cond, fx["codes"].to(cuda_device))
    eng.prefill(g._prefix)
    from genvc_b200.engine import Sampling

    # ignore_eos: rows that finished keep being fed the fixture's (pad) tokens, as the reference's loop does
    sp = Sampling(**fx["sampling"], max_new_tokens=n, ignore_eos=True)
    forced = fx["ids"].transpose(0, 1).contiguous().to(cuda_device)
    ch = eng.decode(n, sp, forced=forced, want_logits=True, mode=mode)
    emitted = ch.status.tolist()[0]
    assert emitted == n
    assert torch.equal(ch.ids.cpu(), forced.cpu())
    steps = fx["steps"]
    got = ch.logits.transpose(0, 1)[:, steps.to(cuda_device)]
    ok, err = close(got, fx["logits"], LOGIT_ATOL, LOGIT_RTOL)
    assert ok, f"logit max err {err}"
    ok, err = close(ch.latents.transpose(0, 1)[:, steps.to(cuda_device)], fx["latents"], LAT_ATOL, LAT_RTOL)
    assert ok, f"latent max err {err}"


def test_batched_fused_long_context_matches_per_op(cuda_device):
    """Batched fused kernel at long contexts (S up to ~760: several key ranges per (row, head) item, capped by
    grid / (rows * heads)) against the per-op path, forced random tokens, B = 3 and 8."""
    from genvc_b200.engine import Sampling

    for name, B, T, n in (("toy_d256_h4_greedy", 8, 400, 330), ("full_h4_cfg1", 3, 240, 24)):
        fx = load_golden(name)
        g = make_gpt(fx, cuda_device, max_batch=B)
        eng = g.engine
        gen = torch.Generator().manual_seed(98)
        codes = torch.randint(0, 256, (B, T), generator=gen).to(cuda_device)
        cond = fx["style_emb"].transpose(1, 2).contiguous().to(cuda_device).expand(B, -1, -1).contiguous()
        forced = torch.randint(0, 1024, (n, B), generator=gen).to(cuda_device)
        sp = Sampling(**fx["sampling"], max_new_tokens=n)
        out = {}
        for mode in (1, 2):
            g.compute_embeddings(cond, codes)
            eng.prefill(g._prefix)
            ch = eng.decode(n, sp, forced=forced, want_logits=True, mode=mode)
            assert ch.status.tolist()[0] == n
            out[mode] = (ch.logits.clone(), ch.latents.clone())
        ok, err = close(out[2][0], out[1][0], LOGIT_ATOL, LOGIT_RTOL)
        assert ok, f"{name}: logit max err {err}"
        ok, err = close(out[2][1], out[1][1], LAT_ATOL, LAT_RTOL)
        assert ok, f"{name}: latent max err {err}"


@pytest.mark.parametrize("name", ["toy_d128_greedy", "toy_d128_batch3"])
def test_chunk_size_one_greedy(name, cuda_device):
    """stream_chunk_size = 1: the first launch after a prefill runs no forward at all (it samples from the prefill's
    logits), so CTA 0 can finish while other CTAs are still reading the generation state -- the state is
    double-buffered for exactly this (a launch never writes what it reads)."""
    fx = load_golden(name)
    B = fx["ids"].shape[0]
    g = make_gpt(fx, cuda_device, max_batch=B)
    cond = fx["style_emb"].transpose(1, 2).contiguous().to(cuda_device)
    for _ in range(3):
        fake = g.compute_embeddings(cond, fx["codes"].to(cuda_device))
        toks = [t for t, _ in g.get_generator(fake_inputs=fake, **gen_kwargs(fx, stream_chunk_size=1))]
        ids = torch.stack(toks, 1).cpu()
        assert torch.equal(ids, fx["ids"])


def test_out_of_range_ids_are_flagged_not_read(cuda_device):
    """C-ABI safety: ids outside the vocabulary never index a table out of bounds; device tensors are not synchronised
    on, the kernels clamp and the flag comes back with the decode status."""
    fx = load_golden("toy_d128_greedy")
    g = make_gpt(fx, cuda_device)
    cond = fx["style_emb"].transpose(1, 2).contiguous().to(cuda_device)
    bad = fx["codes"].clone()
    bad[0, 3] = 100000
    with pytest.raises(IndexError):  # host tensor: checked before any launch
        g.compute_embeddings(cond, bad)
    with pytest.raises(IndexError):  # device tensor: flagged by the kernel, raised with the first chunk's status
        g.generate(cond, bad.to(cuda_device), **gen_kwargs(fx))
    from genvc_b200.engine import Sampling

    g.compute_embeddings(cond, fx["codes"].to(cuda_device))
    g.engine.prefill(g._prefix)
    forced = torch.full((4, 1), 5000, dtype=torch.int64, device=cuda_device)
    ch = g.engine.decode(4, Sampling(**fx["sampling"]), forced=forced, mode=2)
    assert ch.status.tolist()[2] == 1
    ids = g.generate(cond, fx["codes"].to(cuda_device), **gen_kwargs(fx))  # the engine is fine afterwards
    assert torch.equal(ids.cpu(), fx["ids"])


# ------------------------------------------------------------------------------------ latent pass
@pytest.mark.parametrize("name", ["toy_d128_greedy", "toy_d256_h4_greedy", "toy_d128_eos", "full_h4_cfg1"])
def test_latent_pass(name, cuda_device):
    fx = load_golden(name)
    g = make_gpt(fx, cuda_device)
    cond = fx["style_emb"].transpose(1, 2).contiguous().to(cuda_device)[:1]
    g0 = fx["ids"][0][fx["ids"][0] != 1025]
    codes = fx["codes"][:1].to(cuda_device)
    lat = g(codes, torch.tensor([codes.shape[1]]), g0[None].to(cuda_device), torch.tensor([g0.numel() * 1024]),
            cond_latents=cond, return_latent=True)
    assert lat.shape == (1, g0.numel(), g.model_dim)
    if fx["latent_pass_steps"] is not None:
        lat = lat[:, fx["latent_pass_steps"].to(cuda_device)]
    ok, err = close(lat, fx["latent_pass"], LAT_ATOL, LAT_RTOL)
    assert ok, f"latent-pass max err {err}"


# ------------------------------------------------------------------------------------ streaming protocol
@pytest.mark.parametrize("name,chunk", [("toy_d128_eos", 8), ("toy_d128_greedy", 5), ("toy_d128_topk20", 8)])
def test_streaming_generator_protocol(name, chunk, cuda_device):
    """(token, latent) per step, EOS pair delivered, StopIteration afterwards; ids equal generate()'s."""
    fx = load_golden(name)
    g = make_gpt(fx, cuda_device)
    cond = fx["style_emb"].transpose(1, 2).contiguous().to(cuda_device)
    fake = g.compute_embeddings(cond, fx["codes"].to(cuda_device))
    P = 32 + fx["codes"].shape[1] + 2
    assert fake.shape == (1, P + 1) and fake.dtype == torch.int64
    assert fake[0, :-1].eq(1).all() and fake[0, -1].item() == 1024
    kw = gen_kwargs(fx, num_return_sequences=1, output_hidden_states=True, stream_chunk_size=chunk)
    noise = fixture_noise(fx)
    if noise is not None:
        kw["exp_noise"] = noise.to(cuda_device)
    gen = g.get_generator(fake_inputs=fake, **kw)
    toks, lats = [], []
    while True:
        try:
            x, latent = next(gen)
        except StopIteration:
            break
        assert x.shape == (1,) and x.dtype == torch.int64
        assert latent.shape == (1, g.model_dim) and latent.dtype == torch.float32
        toks.append(x)
        lats.append(latent)
    ids = torch.stack(toks, 1).cpu()
    assert torch.equal(ids, fx["ids"])
    if name == "toy_d128_eos":
        assert ids[0, -1].item() == 1025  # the EOS step's pair is delivered
    ok, err = close(torch.stack(lats, 1)[:, fx["steps"].to(cuda_device)], fx["latents"], LAT_ATOL, LAT_RTOL)
    assert ok, err


def test_philox_sampling_is_seed_reproducible(cuda_device):
    fx = load_golden("toy_d128_topk20")
    g = make_gpt(fx, cuda_device)
    cond = fx["style_emb"].transpose(1, 2).contiguous().to(cuda_device)
    codes = fx["codes"].to(cuda_device)
    kw = gen_kwargs(fx)
    torch.manual_seed(5)
    a = g.generate(cond, codes, **kw)
    torch.manual_seed(5)
    b = g.generate(cond, codes, **kw)
    torch.manual_seed(6)
    c = g.generate(cond, codes, **kw)
    assert torch.equal(a, b)
    assert not torch.equal(a, c)
    assert (a >= 0).all() and (a < 1026).all()


# ------------------------------------------------------------------------------------ KV-cache attention microbenchmark
@pytest.mark.parametrize("H,hd", [(4, 256), (16, 64), (4, 32), (2, 128)])
@pytest.mark.parametrize("S", [1, 7, 64, 333])
def test_kv_attention_matches_torch(H, hd, S, cuda_device):
    """genvc_kv_attention (BASELINE configs[4]) against a plain fp32 torch evaluation of HF GPT2Attention._attn for one
    query: softmax(q k^T / sqrt(hd)) v.  Tolerance: two fp32 summation orders (online softmax vs torch): 2e-5 abs."""
    from genvc_b200.config import GenVCDims, make_config_dict
    from genvc_b200.engine import Engine

    eng = Engine(GenVCDims.from_config(make_config_dict(2, 128, 4)), cuda_device)
    g = torch.Generator().manual_seed(S * 131 + hd)
    N, S_max = 5, S + 3
    k = torch.randn((N, H, S_max, hd), generator=g).to(cuda_device)
    v = torch.randn((N, H, S_max, hd), generator=g).to(cuda_device)
    q = torch.randn((N, H, hd), generator=g).to(cuda_device)
    out = eng.kv_attention(q, k, v, S)
    p = torch.softmax(torch.einsum("nhd,nhsd->nhs", q, k[:, :, :S]) / hd ** 0.5, -1)
    ref = torch.einsum("nhs,nhsd->nhd", p, v[:, :, :S])
    ok, err = close(out, ref, 2e-5, 1e-4)
    assert ok, f"kv attention max err {err}"
