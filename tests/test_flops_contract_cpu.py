"""The denominator of every roofline figure: bench.py's GFLOP_PER_IMG_STEP must equal a recount of 2 x MAC over every
Conv2d and Linear the reference graph executes per image and sampling step (un-optimised graph, Upsample
convs at full output resolution, no credit for folding; the attention bmm's 8.4 MFLOP are not part of the figure). The
recount runs the oracle's UNet forward (= the reference's ops, oracle/sr3_oracle.py) with F.conv2d / F.linear counted."""
import importlib.util
import os

import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("R", [32, 64, 128])
def test_contract_flops_match_a_recount_of_the_reference_graph(R, monkeypatch):
    from oracle import sr3_oracle as O
    from oracle import make_golden as G
    from oracle.weights import make_state_dict

    mopt = G.model_opt(10)
    sd = make_state_dict(mopt, seed=0, gain=1.0)
    macs = [0]
    conv2d, linear = F.conv2d, F.linear

    def counted_conv2d(x, w, *a, **k):
        y = conv2d(x, w, *a, **k)
        macs[0] += y.numel() * w.shape[1] * w.shape[2] * w.shape[3]      # outputs x (Cin * kh * kw)
        return y

    def counted_linear(x, w, *a, **k):
        y = linear(x, w, *a, **k)
        macs[0] += y.numel() * w.shape[1]
        return y

    monkeypatch.setattr(O.F, "conv2d", counted_conv2d)
    monkeypatch.setattr(O.F, "linear", counted_linear)
    with torch.no_grad():
        O.unet_forward(sd, mopt, torch.zeros(1, 6, R, R), torch.full((1, 1), 0.5))
    gflop = 2.0 * macs[0] / 1e9
    assert abs(gflop - _bench().GFLOP_PER_IMG_STEP[R]) <= 5e-5 * gflop, (R, gflop)
