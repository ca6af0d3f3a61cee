"""Checkpoint ingestion (b200sr3/checkpoint.py) against the reference's two on-disk formats
(model/sr/model.py:146-195 `<prefix>_gen.pth`; lib/trainer_temp.py:165-188 combined `sr_model_state`)."""
import os

import pytest
import torch

import b200sr3
from b200sr3 import checkpoint
from conftest import synthetic_weights


def _net(T=10):
    opt = {"phase": "val", "sr": {"model": b200sr3.configs.model_opt(T)}}
    net = b200sr3.define_G(opt)
    net.set_new_noise_schedule(opt["sr"]["model"]["beta_schedule"]["val"], [torch.device("cpu")])
    return net


def test_gen_pth_and_combined_checkpoint_round_trip(tmp_path):
    sd = synthetic_weights(0, 1.0)                      # reference-keyed: strict-loads into the reference module
    net = _net()
    # (1) `<prefix>_gen.pth` as save_network writes it
    prefix = os.path.join(tmp_path, "I100_E3")
    torch.save({k: v.clone() for k, v in sd.items()}, prefix + "_gen.pth")
    missing, unexpected = checkpoint.load_network(net, prefix, strict=True)
    assert not missing and not unexpected
    own = net.state_dict()
    for k, v in sd.items():
        assert torch.equal(own[k], v), k
    # (2) combined checkpoint with the DistributedDataParallel prefix and a schedule saved at another T
    other = _net(T=25)
    combined = {"sr_model_state": {"module." + k: v for k, v in other.state_dict().items()},
                "mica_model_state": {}, "epoch": 3, "global_step": 100}
    for k, v in sd.items():
        combined["sr_model_state"]["module." + k] = v.clone() * 0.5
    path = os.path.join(tmp_path, "combined.pth")
    torch.save(combined, path)
    betas_before = net.betas.clone()
    missing, unexpected = checkpoint.load_combined(net, path)
    assert not missing and not unexpected
    assert torch.equal(net.betas, betas_before) and net.betas.shape[0] == 10      # the T=25 buffers were not taken
    for k, v in sd.items():
        assert torch.equal(net.state_dict()[k], v * 0.5), k
    # (3) save_network writes what load_network reads
    out = checkpoint.save_network(net, os.path.join(tmp_path, "again"))
    again = torch.load(out, weights_only=True)
    assert set(again) == set(net.state_dict()) and all(t.device.type == "cpu" for t in again.values())


def test_strict_load_reports_a_foreign_checkpoint(tmp_path):
    net = _net()
    prefix = os.path.join(tmp_path, "bad")
    torch.save({"denoise_fn.not_a_layer.weight": torch.zeros(3)}, prefix + "_gen.pth")
    with pytest.raises(RuntimeError):
        checkpoint.load_network(net, prefix, strict=True)
    with pytest.raises(FileNotFoundError):
        checkpoint.load_network(net, os.path.join(tmp_path, "nowhere"))
