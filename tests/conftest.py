import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    import torch

    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


_CKPT_CACHE = {}


def golden_checkpoint(fx):
    """Regenerate the fixture's seeded synthetic checkpoint and check its digest."""
    from genvc_b200.synth import state_dict_digest, synth_checkpoint

    key = tuple(sorted(fx["model"].items()))
    if key not in _CKPT_CACHE:
        if len(_CKPT_CACHE) > 2:  # full-size checkpoints are 1.7 GB each
            _CKPT_CACHE.clear()
        ck = synth_checkpoint(**fx["model"])
        assert state_dict_digest(ck["model"]) == fx["digest"], "synthetic checkpoint drifted from the fixture"
        _CKPT_CACHE[key] = ck
    return _CKPT_CACHE[key]


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
