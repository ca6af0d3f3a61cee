"""The fused posterior update (diffusion.py:144-187) against torch fp32, op for op."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from conftest import build_net

pytestmark = pytest.mark.gpu


def test_update_through_step_with_known_eps(golden_dir):
    """With every conv weight zero the UNet returns eps = final bias, so p_sample reduces to the
    update formula and must match torch fp32 to rounding (1 ulp of exp)."""
    import b200sr3
    T = 600
    opt = {"phase": "val", "sr": {"model": b200sr3.configs.model_opt(T)}}
    net = b200sr3.define_G(opt)
    with torch.no_grad():
        for p in net.parameters():
            p.zero_()
        net.denoise_fn.state_dict()["final_conv.block.3.bias"].copy_(torch.tensor([0.3, -0.7, 1.1]))
    net = net.cuda().eval()
    net.set_new_noise_schedule(opt["sr"]["model"]["beta_schedule"]["val"], [torch.device("cuda")])
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 3, 16, 16, generator=g).cuda()
    z = torch.randn(2, 3, 16, 16, generator=g).cuda()
    cond = torch.zeros_like(x)
    eps = torch.tensor([0.3, -0.7, 1.1], device="cuda").view(1, 3, 1, 1).expand_as(x)
    for t in (599, 300, 1, 0):
        out = net.p_sample(x, t, condition_x=cond, noise=z)
        x0 = (net.sqrt_recip_alphas_cumprod[t] * x - net.sqrt_recipm1_alphas_cumprod[t] * eps).clamp(-1, 1)
        mean = net.posterior_mean_coef1[t] * x0 + net.posterior_mean_coef2[t] * x
        ref = mean + (z if t > 0 else 0) * (0.5 * net.posterior_log_variance_clipped[t]).exp()
        assert float((out - ref).abs().max()) <= 2e-6, t
