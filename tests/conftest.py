import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def tiny_cfg():
    from sam_audio_b200.config import stand_in_config
    return stand_in_config("sam-audio-tiny")


@pytest.fixture(scope="session")
def tiny_sd(tiny_cfg):
    from sam_audio_b200.synthetic import make_state_dict
    return make_state_dict(tiny_cfg, seed=0)


@pytest.fixture(scope="session")
def tiny_model(tiny_cfg):
    """Session-wide tiny model on cuda:0 (gpu tests only)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import __graft_entry__ as g
    g.build()
    from sam_audio_b200.model import build_synthetic_model
    return build_synthetic_model("sam-audio-tiny", seed=0, device="cuda:0")
