"""HiFi-GAN generator (the stage after the path): oracle vs the reference's own module (fixtures generated
by tests/golden/make_golden_hifigan.py from layers/hifigan.py::HiFiGAN), CUDA path vs the same fixtures."""
import os

import pytest
import torch

from genvc_b200.synth import HIFIGAN_DEFAULTS, hifigan_conv_shapes, state_dict_digest, synth_hifigan_state
from oracle.hifigan_oracle import hifigan_forward, vocode

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["hifigan_default_t32", "hifigan_default_t94_b2", "hifigan_rb1_small"]
WAV_ATOL = 1e-4  # waveform after tanh, |y| <= 1: fp32 FFMA in a different summation order than cuDNN / MKL


def load(name):
    fx = torch.load(os.path.join(GOLDEN, name + ".pt"))
    sd = synth_hifigan_state(fx["seed"], **fx["cfg"])
    assert state_dict_digest(sd) == fx["digest"], "synthetic vocoder weights drifted from the ones the fixture was made with"
    cfg = dict(HIFIGAN_DEFAULTS, **fx["cfg"])
    return fx, sd, cfg


def arch(cfg):
    return {k: cfg[k] for k in ("resblock_kernel_sizes", "resblock_dilation_sizes", "upsample_rates", "upsample_kernel_sizes",
                                "resblock_type")}


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_fixture(name):
    fx, sd, cfg = load(name)
    y = hifigan_forward(sd, fx["x"], **arch(cfg))
    assert y.shape == fx["y"].shape
    assert float((y - fx["y"]).abs().max()) < 1e-5


def test_weight_norm_removed_state_dict_is_equivalent():
    fx, sd, cfg = load("hifigan_rb1_small")
    from oracle.hifigan_oracle import fold_weight_norm
    plain = {}
    for name, _, _ in hifigan_conv_shapes(cfg):
        plain[name + ".weight"] = fold_weight_norm(sd, name)
        plain[name + ".bias"] = sd[name + ".bias"]
    y = hifigan_forward(plain, fx["x"], **arch(cfg))
    assert float((y - fx["y"]).abs().max()) < 1e-5


def test_conv_inventory_matches_host_module_order():
    """The host module walks the convolutions in the reference's construction order with the reference's paddings."""
    from genvc_b200.vocoder import HiFiGAN, _get_padding
    cfg = dict(HIFIGAN_DEFAULTS)
    v = HiFiGAN.__new__(HiFiGAN)
    v.rks, v.rds = list(cfg["resblock_kernel_sizes"]), [list(d) for d in cfg["resblock_dilation_sizes"]]
    v.rates, v.uks = list(cfg["upsample_rates"]), list(cfg["upsample_kernel_sizes"])
    v.resblock_type, v.num_kernels, v.num_upsamples = cfg["resblock_type"], 3, 3
    names = [n for n, *_ in v.conv_names()]
    assert sorted(names) == sorted(n for n, _, _ in hifigan_conv_shapes(cfg))
    assert _get_padding(7, 12) == 36 and _get_padding(3, 1) == 1


# ------------------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_generator_matches_reference_fixture(name, cuda_device):
    from genvc_b200.vocoder import HiFiGAN
    fx, sd, cfg = load(name)
    v = HiFiGAN(cfg["input_feat_dim"], cfg["upsample_initial_channel"], cfg["resblock_kernel_sizes"], cfg["resblock_dilation_sizes"],
                cfg["upsample_rates"], cfg["upsample_kernel_sizes"], cfg["resblock_type"], device=cuda_device)
    v.load_state_dict(sd)
    y = v(fx["x"].to(cuda_device))
    assert y.shape == fx["y"].shape
    err = float((y.cpu() - fx["y"]).abs().max())
    assert err < WAV_ATOL, f"waveform max err {err}"
    # second call: the captured graph is replayed on new input
    x2 = fx["x"].flip(-1).contiguous()
    y2 = v(x2.to(cuda_device))
    ref2 = hifigan_forward(sd, x2, **arch(cfg))
    err2 = float((y2.cpu() - ref2).abs().max())
    assert err2 < WAV_ATOL, f"graph replay: waveform max err {err2}"
    assert v.launches > 0


@pytest.mark.gpu
def test_vocode_driver_step_on_cuda_generator(cuda_device):
    """inference_utils._vocode (x4 linear interpolation + vocoder, inference/inference_utils.py:81-85) with the CUDA generator
    attached as model.hifigan."""
    import types
    from genvc_b200.inference.inference_utils import _vocode
    from genvc_b200.vocoder import HiFiGAN
    sd = synth_hifigan_state(91)
    v = HiFiGAN.from_config({}, device=cuda_device).load_state_dict(sd)
    lat = torch.randn(1, 11, 1024, generator=torch.Generator().manual_seed(5))
    model = types.SimpleNamespace(hifigan=v, hifigan_scale_factor=4.0)
    wav = _vocode(model, lat.to(cuda_device))
    ref = vocode(sd, lat, 4.0, **arch(dict(HIFIGAN_DEFAULTS)))
    assert wav.shape == ref.shape == (1, 1, 11 * 4 * 256)
    assert float((wav.cpu() - ref).abs().max()) < WAV_ATOL


@pytest.mark.gpu
def test_conv_entry_points_reject_bad_arguments(cuda_device):
    from genvc_b200.lib import load_library
    lib = load_library()
    x = torch.zeros(1, 8, 16, device=cuda_device)
    w = torch.zeros(8, 3, 8, device=cuda_device)
    y = torch.zeros(1, 8, 16, device=cuda_device)
    # unsupported stride
    assert lib.genvc_conv1d(x.data_ptr(), w.data_ptr(), None, None, y.data_ptr(), 1, 8, 8, 16, 3, 1, 0, 3, 1.0, 0, 1.0, 0, None, 0, None) < 0
    assert lib.genvc_conv1d(None, w.data_ptr(), None, None, y.data_ptr(), 1, 8, 8, 16, 3, 1, 1, 1, 1.0, 0, 1.0, 0, None, 0, None) < 0
    assert lib.genvc_conv_transpose1d(x.data_ptr(), w.data_ptr(), None, y.data_ptr(), 1, 8, 8, 0, 3, 2, 0, 1.0, None, 0, None) < 0


@pytest.mark.gpu
def test_checkpoint_with_hifigan_weights_attaches_cuda_generator(cuda_device):
    """model_init on a checkpoint that carries ``hifigan.*`` (as trainers/hifigan_trainer.py saves it): model.hifigan is the
    CUDA generator built from config.vocoder_config; a checkpoint without them keeps the attachment point."""
    from genvc_b200.inference.model_init import model_from_checkpoint
    from genvc_b200.synth import synth_checkpoint
    from genvc_b200.vocoder import HiFiGAN
    small = dict(input_feat_dim=128, upsample_initial_channel=64, upsample_rates=(4, 2), upsample_kernel_sizes=(8, 4))
    ck = synth_checkpoint(n_layer=2, d_model=128, n_head=2, seed=3)
    model, _ = model_from_checkpoint(ck, cuda_device)
    with pytest.raises(Exception):
        model.hifigan(torch.zeros(1, 128, 8, device=cuda_device))  # not attached
    sd = synth_hifigan_state(12, **small)
    ck["model"].update({"hifigan." + k: v for k, v in sd.items()})
    ck["config"]["vocoder_config"] = dict(input_feat_dim=128, upsample_initial_channel=64, upsample_rates=[4, 2],
                                          upsample_kernal_sizes=[8, 4], resblock_kernel_sizes=[3, 5, 7],
                                          resblock_dilation_sizes=[[1, 2], [2, 6], [3, 12]], resblock_type="2", hop_length=256)
    model, _ = model_from_checkpoint(ck, cuda_device)
    assert isinstance(model.hifigan, HiFiGAN)
    x = torch.randn(1, 128, 9, generator=torch.Generator().manual_seed(1))
    y = model.hifigan(x.to(cuda_device))
    ref = hifigan_forward(sd, x, **arch(dict(HIFIGAN_DEFAULTS, **small)))
    assert y.shape == ref.shape == (1, 1, 72)
    assert float((y.cpu() - ref).abs().max()) < WAV_ATOL
