"""CPU: the oracle restatement reproduces the fixtures the REAL reference generated
(tests/golden/make_golden.py).  Integer outputs (token ids) bit-exact; logits/latents
within 2e-5 abs (two fp32 CPU evaluations of the same graph; observed ~1e-6)."""
import pytest
import torch

from conftest import golden_checkpoint, load_golden
from oracle.genvc_oracle import SamplingParams, draw_exponential_noise, load_oracle

TOY = ["toy_d128_greedy", "toy_d128_topk20", "toy_d128_topk0_topp1", "toy_d128_eos", "toy_d128_batch3",
       "toy_d256_h4_greedy", "toy_d512_h2_greedy", "toy_d128_batch8_eos", "toy_d256_batch5_topk20"]
TOL = 2e-5


def _noise(fx, n, B, V):
    if fx["noise_seed"] is None:
        return None
    cap = fx["new_tokens"] if fx["new_tokens"] is not None else 602
    g = torch.Generator().manual_seed(fx["noise_seed"])
    return torch.empty((cap, B, V)).exponential_(1, generator=g)[:n]


def _run(fx, max_new=None):
    ck = golden_checkpoint(fx)
    o = load_oracle(ck)
    style = o.get_style_emb(fx["mel"])
    assert (style - fx["style_emb"]).abs().max() < TOL
    cond = style.transpose(1, 2).contiguous()
    sp = SamplingParams(**fx["sampling"])
    n = fx["ids"].shape[1] if max_new is None else max_new
    trace = {}
    ids, lats = o.generate(cond, fx["codes"], sp, noise=_noise(fx, fx["ids"].shape[1], *fx["ids"].shape[:1], 1026),
                           max_new_tokens=fx["new_tokens"] if max_new is None else max_new, trace=trace)
    return o, cond, ids, lats, torch.stack(trace["logits"], 1)


@pytest.mark.parametrize("name", TOY)
def test_oracle_matches_reference_fixture(name):
    fx = load_golden(name)
    o, cond, ids, lats, logits = _run(fx)
    assert torch.equal(ids, fx["ids"])
    steps = fx["steps"]
    assert (logits[:, steps] - fx["logits"]).abs().max() < TOL
    assert (lats[:, steps] - fx["latents"]).abs().max() < TOL
    if fx["latent_pass"] is not None:
        g0 = ids[0][ids[0] != 1025]
        lp = o.forward_latents(fx["codes"][:1], g0[None], cond[:1])
        assert lp.shape == fx["latent_pass"].shape
        assert (lp - fx["latent_pass"]).abs().max() < TOL


def test_eos_fixture_ends_with_stop_token():
    fx = load_golden("toy_d128_eos")
    assert fx["ids"][0, -1].item() == 1025 and fx["ids"].shape[1] < 602


def test_batch_rows_equal_single_runs():
    """A7: row r of an equal-T batch == the B=1 run on row r, finished rows padded with 1025."""
    fx = load_golden("toy_d128_batch3")
    ck = golden_checkpoint(fx)
    o = load_oracle(ck)
    cond = o.get_style_emb(fx["mel"]).transpose(1, 2).contiguous()
    sp = SamplingParams(**fx["sampling"])
    n = fx["ids"].shape[1]
    for r in range(fx["ids"].shape[0]):
        ids, _ = o.generate(cond[r : r + 1], fx["codes"][r : r + 1], sp, max_new_tokens=fx["new_tokens"])
        m = ids.shape[1]
        assert torch.equal(fx["ids"][r, :m], ids[0])
        assert (fx["ids"][r, m:] == 1025).all()
        assert m <= n


def test_full_size_prefix_of_cfg1():
    """BASELINE configs[0] (L=30, D=1024, H=4, 3 s/3 s, greedy): first 12 tokens on CPU."""
    fx = load_golden("full_h4_cfg1")
    o, cond, ids, lats, logits = _run(fx, max_new=12)
    assert torch.equal(ids, fx["ids"][:, :12])
    assert (logits - fx["logits"][:, :12]).abs().max() < 1e-4
    assert fx["ids"].shape[1] == 602  # random weights never emit EOS: cap reached


def test_multinomial_is_argmax_of_p_over_exponential_noise():
    """The noise contract of genvc_decode(exp_noise): torch.multinomial(p, 1) on CPU ==
    argmax(p / q), q = empty_like(p).exponential_(1) from the same generator state."""
    g = torch.Generator().manual_seed(1)
    for _ in range(50):
        p = torch.softmax(torch.randn(2, 1026, generator=g) * 3, -1)
        st = g.get_state()
        a = torch.multinomial(p, 1, generator=g).squeeze(1)
        g.set_state(st)
        q = draw_exponential_noise(p.shape, g)
        assert torch.equal(a, torch.argmax(p / q, -1))
