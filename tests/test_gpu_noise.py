"""The sampler's own noise stream (Philox4x32-10 + Box-Muller with __logf / __sincosf, csrc/common.cuh) replaces the
reference's torch.randn / torch.randn_like draws (diffusion.py:186,205) in every throughput run, so its statistics are
tested, not assumed: moments, Kolmogorov-Smirnov distance to N(0,1), independence across channels / pixels / rows /
timesteps / seeds, that x_T and the per-step draws never share a key, and that the in-kernel draws of the sampling
chain ARE this stream (recovered through a network whose weights are all zero). All through the C ABI."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _net(T, zero=False, device="cuda"):
    import b200sr3
    opt = {"phase": "val", "sr": {"model": b200sr3.configs.model_opt(T)}}
    net = b200sr3.define_G(opt)
    if zero:
        with torch.no_grad():
            for p in net.parameters():
                p.zero_()
    net = net.to(device).eval()
    net.set_new_noise_schedule(opt["sr"]["model"]["beta_schedule"]["val"], [torch.device(device)])
    return net


@pytest.fixture(scope="module")
def zero_net():
    return _net(10, zero=True)


def _corr(a, b):
    a = a.double().flatten() - a.double().mean()
    b = b.double().flatten() - b.double().mean()
    return float((a * b).sum() / (a.norm() * b.norm()))


def test_moments_and_ks(zero_net):
    """1.26e7 draws: |mean| < 1e-3 (3 sigma), |var - 1| < 2e-3 (4 sigma), skewness, kurtosis, tails, KS distance."""
    from scipy import stats
    z = zero_net.philox_normal((64, 3, 256, 256), t=7, seed=12345)
    n = z.numel()
    assert n >= 10_000_000 and torch.isfinite(z).all()
    zd = z.double()
    mean, var = float(zd.mean()), float(zd.var())
    skew = float(((zd - mean) ** 3).mean() / var ** 1.5)
    kurt = float(((zd - mean) ** 4).mean() / var ** 2)
    assert abs(mean) < 1e-3, mean
    assert abs(var - 1.0) < 2e-3, var
    assert abs(skew) < 3e-3, skew                   # sigma = sqrt(6/n) = 6.9e-4
    assert abs(kurt - 3.0) < 8e-3, kurt             # sigma = sqrt(24/n) = 1.4e-3
    # tails: P(|z| > 3) = 2.6998e-3, P(|z| > 4) = 6.334e-5 (binomial sigma 1.5e-5 / 2.2e-6)
    assert abs(float((zd.abs() > 3).double().mean()) - 2.6998e-3) < 8e-5
    assert abs(float((zd.abs() > 4).double().mean()) - 6.334e-5) < 1.2e-5
    assert float(zd.abs().max()) < 6.7              # u1 >= 2^-32: |z| <= sqrt(-2 ln 2^-32) = 6.66
    # KS on 2e6 draws taken with a stride co-prime to every tensor dimension
    sub = z.flatten()[::6 + 1][:2_000_000].cpu().numpy().astype(np.float64)
    d = stats.kstest(sub, "norm").statistic
    assert d < 1.95 / math.sqrt(sub.size), d        # p ~ 1e-3
    # per-channel moments (one Philox block feeds the 3 channels of a pixel; channel 3 of the block is unused)
    for c in range(3):
        zc = zd[:, c]
        assert abs(float(zc.mean())) < 2e-3 and abs(float(zc.var()) - 1.0) < 4e-3, c


def test_independence(zero_net):
    """No linear correlation between channels of a pixel (the Box-Muller pair!), neighbouring pixels, batch rows,
    timesteps, seeds; also none between squares (variance coupling of a Box-Muller pair shows up there)."""
    net = zero_net
    shape = (16, 3, 128, 128)
    z = net.philox_normal(shape, t=3, seed=99)
    tol = 5.0 / math.sqrt(z[:, 0].numel())          # 5 sigma for n = 262144: 9.8e-3
    for a, b in ((0, 1), (0, 2), (1, 2)):
        assert abs(_corr(z[:, a], z[:, b])) < tol
        assert abs(_corr(z[:, a] ** 2, z[:, b] ** 2)) < tol
    assert abs(_corr(z[..., :, :-1], z[..., :, 1:])) < tol          # x neighbours
    assert abs(_corr(z[..., :-1, :], z[..., 1:, :])) < tol          # y neighbours
    assert abs(_corr(z[:-1], z[1:])) < tol                          # batch rows
    other_t = net.philox_normal(shape, t=4, seed=99)
    other_seed = net.philox_normal(shape, t=3, seed=100)
    x_T = net.philox_normal(shape, t=net.num_timesteps, seed=99)    # the key x_T is drawn with
    for o in (other_t, other_seed, x_T):
        assert abs(_corr(z, o)) < tol / 1.7 and not torch.equal(z, o)
    assert torch.equal(z, net.philox_normal(shape, t=3, seed=99))   # deterministic
    # a 64-bit seed is used in full
    assert not torch.equal(z, net.philox_normal(shape, t=3, seed=99 + 2 ** 32))


def test_rows_are_keyed_globally(zero_net):
    """Row j of a call with row_offset = k is row j + k of the unsplit stream: shards and chunks see their own rows."""
    net = zero_net
    full = net.philox_normal((6, 3, 32, 32), t=5, seed=7)
    assert torch.equal(net.philox_normal((2, 3, 32, 32), t=5, seed=7, row_offset=4), full[4:6])
    assert torch.equal(net.philox_normal((3, 3, 32, 32), t=5, seed=7, row_offset=1), full[1:4])
    assert not torch.equal(full[0], full[1])


def test_chain_draws_are_this_stream(zero_net):
    """With all weights zero eps = 0, so x_{t-1} = c1*clamp(A x_t) + c2*x_t + sigma_t z_t: the z_t the tail kernel drew
    inside the CUDA-graph chain is recovered from consecutive snapshots (T = 10: one snapshot per step) and must be the
    philox_normal draw at key t; x_T must be the draw at key T; nothing is drawn at t = 0 (diffusion.py:186)."""
    net = zero_net
    T, B, R, seed, row0 = net.num_timesteps, 3, 32, 4242, 5
    cond = torch.zeros(B, 3, R, R, device="cuda")
    out, snaps, x_T = net.sample_batched(cond, seed=seed, return_snapshots=True, return_x_T=True, row_offset=row0)
    assert snaps.shape[0] == T
    assert torch.equal(x_T, net.philox_normal((B, 3, R, R), t=T, seed=seed, row_offset=row0))
    xs = [x_T] + [snaps[i] for i in range(T)]            # xs[k] = x after k steps
    for t in range(T - 1, -1, -1):
        x_t, x_tm1 = xs[T - 1 - t].double(), xs[T - t].double()
        a, c1, c2 = (float(net.sqrt_recip_alphas_cumprod[t]), float(net.posterior_mean_coef1[t]),
                     float(net.posterior_mean_coef2[t]))
        sigma = math.exp(0.5 * float(net.posterior_log_variance_clipped[t]))
        mean = c1 * (a * x_t).clamp(-1, 1) + c2 * x_t
        if t == 0:
            assert float((x_tm1 - mean).abs().max()) < 1e-6
            continue
        z = (x_tm1 - mean) / sigma
        want = net.philox_normal((B, 3, R, R), t=t, seed=seed, row_offset=row0).double()
        assert float((z - want).abs().max()) < 2e-6 * 8 / sigma + 1e-5, t      # fp32 rounding of x, divided by sigma
    assert torch.equal(out, snaps[-1])


def test_sharded_and_chunked_calls_draw_the_unsplit_noise():
    """every rank used to draw the same noise for its shard. With the global row offset, the shards of a
    batch (and the chunks of super_resolution_samples) reproduce the unsplit call bit for bit, and two shards differ."""
    from conftest import build_net
    from oracle.weights import make_inputs
    net, _ = build_net(10)
    cond, _ = make_inputs(4, 32, 1, seed=3)
    cond = cond.cuda()
    full = net.super_resolution_batched(cond, seed=77)
    lo = net.super_resolution_batched(cond[:2].contiguous(), seed=77, row_offset=0)
    hi = net.super_resolution_batched(cond[2:].contiguous(), seed=77, row_offset=2)
    assert torch.equal(torch.cat([lo, hi]), full)
    same_cond = cond[:1].repeat(4, 1, 1, 1)
    rows = net.super_resolution_batched(same_cond, seed=77)
    assert not torch.equal(rows[0], rows[1]) and not torch.equal(rows[1], rows[3])      # rows draw different noise
    a = net.super_resolution_samples(cond[:2], 3, seed=5)
    b = net.super_resolution_samples(cond[:2], 3, seed=5, max_batch=4)                  # ragged chunks 4 + 2
    assert torch.equal(a, b)
    # default seed: taken from torch's global generator, so manual_seed controls it
    torch.manual_seed(123)
    u = net.super_resolution_batched(cond)
    v = net.super_resolution_batched(cond)
    torch.manual_seed(123)
    assert torch.equal(u, net.super_resolution_batched(cond)) and not torch.equal(u, v)
