"""CPU, build container only: pins oracle/genvc_oracle.py against the REAL reference modules
imported from /root/reference (skipped on the GPU box where the tree does not exist)."""
import pytest
import torch

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def pair():
    from genvc_b200.synth import synth_checkpoint
    from oracle.genvc_oracle import load_oracle

    ck = synth_checkpoint(n_layer=2, d_model=128, n_head=4, seed=21)
    return ref_shim.build_reference_gpt(ck), load_oracle(ck)


def test_state_dict_layout_loads_strict(pair):
    pass  # build_reference_gpt uses strict=True: key names and shapes match the reference


@pytest.mark.parametrize("S_mel", [28, 282])
def test_perceiver(pair, S_mel):
    ref, o = pair
    mel = torch.randn(2, 80, S_mel, generator=torch.Generator().manual_seed(S_mel))
    with torch.no_grad():
        a = ref.get_style_emb(mel)
        a4 = ref.get_style_emb(mel.unsqueeze(1))
    b = o.get_style_emb(mel)
    assert a.shape == b.shape == (2, 128, 32)
    assert (a - b).abs().max() < 1e-5 and torch.equal(a, a4)


@pytest.mark.parametrize("top_k,top_p,seed", [(1, 0.85, 0), (15, 0.85, 1), (0, 0.6, 2), (50, 1.0, 3)])
def test_generation_loop(pair, top_k, top_p, seed):
    from oracle.genvc_oracle import SamplingParams

    ref, o = pair
    g = torch.Generator().manual_seed(seed)
    mel = torch.randn(1, 80, 64, generator=g)
    codes = torch.randint(0, 256, (1, 9 + seed), generator=g)
    cond = o.get_style_emb(mel).transpose(1, 2).contiguous()
    t1, t2 = {}, {}
    ids_r, lat_r = ref_shim.ref_generate(ref, cond, codes, top_k, top_p, 0.85, 2.0,
                                         generator=torch.Generator().manual_seed(5), max_new_tokens=30, trace=t1)
    ids_o, lat_o = o.generate(cond, codes, SamplingParams(top_k=top_k, top_p=top_p),
                              generator=torch.Generator().manual_seed(5), max_new_tokens=30, trace=t2)
    assert torch.equal(ids_r, ids_o)
    assert (lat_r - lat_o).abs().max() < 1e-5
    for a, b in zip(t1["scores"], t2["scores"]):
        assert torch.equal(torch.isinf(a), torch.isinf(b))


def test_compute_embeddings_and_latent_pass(pair):
    ref, o = pair
    g = torch.Generator().manual_seed(9)
    cond = torch.randn(1, 32, 128, generator=g)
    codes = torch.randint(0, 256, (1, 17), generator=g)
    gen = torch.randint(0, 1024, (1, 23), generator=g)
    with torch.no_grad():
        fake = ref.compute_embeddings(cond, codes)
        pre = ref.gpt_inference.cached_prefix_emb
        lp = ref(codes, torch.tensor([17]), gen, torch.tensor([23 * 1024]), cond_latents=cond, return_latent=True)
    assert torch.equal(fake, o.fake_inputs(o.prefix_embeddings(cond, codes)))
    assert torch.equal(pre, o.prefix_embeddings(cond, codes))
    assert (lp - o.forward_latents(codes, gen, cond)).abs().max() < 1e-5
