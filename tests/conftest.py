import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "3d-super-resolution-face-reconstruction_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def built_lib():
    """Path of libb200sr3.so, building it in-tree with nvcc when it is missing or stale."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("b200sr3_build", os.path.join(PKG, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()


_WEIGHTS = {}


def synthetic_weights(seed, gain, n_timestep=10):
    """oracle.weights.make_state_dict, cached per session (92.5 M parameters)."""
    from oracle.weights import make_state_dict
    import b200sr3
    key = (seed, gain)
    if key not in _WEIGHTS:
        _WEIGHTS[key] = make_state_dict(b200sr3.configs.model_opt(n_timestep), seed=seed, gain=gain)
    return _WEIGHTS[key]


def build_net(n_timestep, seed=0, gain=1.0, device="cuda"):
    """define_G drop-in with synthetic weights and the val schedule installed on `device`."""
    import torch
    import b200sr3
    opt = {"phase": "val", "sr": {"model": b200sr3.configs.model_opt(n_timestep)}}
    net = b200sr3.define_G(opt)
    net.load_state_dict(synthetic_weights(seed, gain), strict=True)
    net = net.to(device).eval()
    net.set_new_noise_schedule(opt["sr"]["model"]["beta_schedule"]["val"], [torch.device(device)])
    return net, opt["sr"]["model"]
