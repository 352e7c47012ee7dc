"""Content-DVAE tokeniser (the stage before the path): oracle vs the reference's own module (fixtures from
tests/golden/make_golden_dvae.py), CUDA path vs the same fixtures — codes are indices: bit-exact."""
import os

import pytest
import torch

from genvc_b200.synth import CONTENT_DVAE_DEFAULTS, state_dict_digest, synth_dvae_state
from oracle.dvae_oracle import get_codebook_indices

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["dvae_content_t50", "dvae_content_t300_b2", "dvae_small"]


def load(name):
    fx = torch.load(os.path.join(GOLDEN, name + ".pt"))
    sd = synth_dvae_state(fx["seed"], **fx["cfg"])
    assert state_dict_digest(sd) == fx["digest"]
    cfg = dict(CONTENT_DVAE_DEFAULTS, **fx["cfg"])
    return fx, sd, cfg


@pytest.mark.parametrize("name", CASES)
def test_oracle_codes_equal_reference(name):
    fx, sd, cfg = load(name)
    codes = get_codebook_indices(sd, fx["x"], num_layers=cfg["num_layers"], num_resnet_blocks=cfg["num_resnet_blocks"],
                                 kernel_size=cfg["kernel_size"])
    assert torch.equal(codes, fx["codes"])
    assert float(fx["gap"].min()) > 0.05  # the fixtures stay clear of near-ties (fp32 rounding of |dist| ~ 1e3 is ~1e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_codes_equal_reference(name, cuda_device):
    from genvc_b200.content_dvae import DiscreteVAE
    fx, sd, cfg = load(name)
    m = DiscreteVAE(positional_dims=1, channels=cfg["channels"], num_tokens=cfg["num_tokens"], codebook_dim=cfg["codebook_dim"],
                    hidden_dim=cfg["hidden_dim"], num_resnet_blocks=cfg["num_resnet_blocks"], kernel_size=cfg["kernel_size"],
                    num_layers=cfg["num_layers"], use_transposed_convs=False, device=cuda_device).load_state_dict(sd)
    codes = m.get_codebook_indices(fx["x"].to(cuda_device))
    assert codes.dtype == torch.int64 and codes.shape == fx["codes"].shape
    assert torch.equal(codes.cpu(), fx["codes"]), f"first mismatch at {(codes.cpu() != fx['codes']).nonzero()[:1].tolist()}"
    assert m.launches > 0
    # second and third call: the captured graph is replayed (new input, then the fixture's again)
    other = get_codebook_indices(sd, fx["x"].flip(-1).contiguous(), num_layers=cfg["num_layers"], num_resnet_blocks=cfg["num_resnet_blocks"],
                                 kernel_size=cfg["kernel_size"])
    assert torch.equal(m.get_codebook_indices(fx["x"].flip(-1).contiguous().to(cuda_device)).cpu(), other)
    assert torch.equal(m.get_codebook_indices(fx["x"].to(cuda_device)).cpu(), fx["codes"])


@pytest.mark.gpu
def test_codebook_argmin_ties_take_the_first_index(cuda_device):
    """torch.max semantics of layers/dvae.py:87 on exact ties: duplicate codebook columns -> the lower index wins."""
    from genvc_b200.lib import load_library
    lib = load_library()
    dim, n, T = 16, 300, 5
    g = torch.Generator().manual_seed(3)
    embed = torch.randn(dim, n, generator=g)
    embed[:, 200] = embed[:, 17]
    embed[:, 299] = embed[:, 17]
    x = embed[:, [17, 5, 299, 200, 120]].clone().unsqueeze(0)  # [1, dim, T]: exact codebook entries
    codes = torch.empty((1, T), dtype=torch.int64, device=cuda_device)
    xe, ee = x.contiguous().to(cuda_device), embed.contiguous().to(cuda_device)
    assert lib.genvc_codebook_argmin(xe.data_ptr(), ee.data_ptr(), codes.data_ptr(), 1, dim, n, T, None) == 0
    assert codes.cpu().tolist() == [[17, 5, 17, 17, 120]]
    assert lib.genvc_codebook_argmin(None, ee.data_ptr(), codes.data_ptr(), 1, dim, n, T, None) < 0


@pytest.mark.gpu
def test_checkpoint_with_content_dvae_weights_attaches_cuda_tokeniser(cuda_device):
    from genvc_b200.content_dvae import DiscreteVAE
    from genvc_b200.inference.model_init import model_from_checkpoint
    from genvc_b200.synth import synth_checkpoint
    small = dict(channels=24, num_tokens=40, codebook_dim=32, hidden_dim=16, num_resnet_blocks=1)
    ck = synth_checkpoint(n_layer=2, d_model=128, n_head=2, seed=3)
    sd = synth_dvae_state(57, **small)
    ck["model"].update({"content_dvae." + k: v for k, v in sd.items()})
    ck["config"]["content_dvae_config"] = dict(num_channels=24, num_tokens=40, codebook_dim=32, hidden_dim=16, num_resnet_blocks=1,
                                               kernel_size=3, num_layers=2)
    model, _ = model_from_checkpoint(ck, cuda_device)
    assert isinstance(model.content_dvae, DiscreteVAE)
    fx, _, cfg = load("dvae_small")
    assert torch.equal(model.content_dvae.get_codebook_indices(fx["x"].to(cuda_device)).cpu(), fx["codes"])
