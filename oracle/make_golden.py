"""TEST INFRASTRUCTURE ONLY (oracle/): generate tests/golden/*.npz from the REAL reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python -m oracle.make_golden

For every case it (1) builds the unmodified reference module through its own factory
(model/sr/networks.py:83-116), (2) loads oracle.weights.make_state_dict with strict=True —
proving the key/shape contract, (3) runs the reference sampler with the noise list injected
through torch.randn / torch.randn_like, (4) asserts that
oracle.sr3_oracle reproduces every x_{t-1} of the reference, and (5) stores inputs and the
reference's outputs as small fixtures. Weights are NOT stored (353 MB); their sha256 is.
"""
import copy
import os
import sys

import numpy as np
import torch

from . import sr3_oracle as O
from .weights import make_inputs, make_state_dict, state_dict_digest

REF = os.environ.get("B200SR3_REF", "/root/reference")
# The oracle is bit-identical to the reference everywhere except the mid-block attention, where
# it uses bmm instead of the reference's einsum (unet.py:132-139); torch's CPU GEMM blocks those
# differently, so results differ at the fp32 reduction-order level (1-2 ulp per step).
TOL_STEP = 5e-6
TOL_CHAIN = 5e-5
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def model_opt(n_timestep):
    """The `sr.model` block shared by all 22 reference YAMLs (config/*.yml:33-62)."""
    sched = {"schedule": "linear", "n_timestep": n_timestep, "linear_start": 1e-6, "linear_end": 1e-2}
    return {
        "which_model_G": "sr3", "finetune_norm": False,
        "unet": {"in_channel": 6, "out_channel": 3, "inner_channel": 64,
                 "channel_multiplier": [1, 2, 4, 8, 8], "attn_res": [16], "res_blocks": 2, "dropout": 0.2},
        "beta_schedule": {"train": dict(sched), "val": dict(sched)},
        "diffusion": {"image_size": 224, "channels": 3, "conditional": True},
    }


def build_reference(mopt, sd):
    sys.path.insert(0, REF)
    from model.sr.networks import define_G  # noqa: the reference's own factory
    cfg = {"sr": {"model": copy.deepcopy(mopt)}, "phase": "val"}
    net = define_G(cfg)
    missing = net.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    net.set_new_noise_schedule(mopt["beta_schedule"]["val"], [torch.device("cpu")])
    net.eval()
    return net


class inject_noise:
    """Patch torch.randn / randn_like to pop from [x_T, z_{T-1}, ..., z_1]."""

    def __init__(self, noise):
        self.q = [n.clone() for n in noise]

    def __enter__(self):
        self.saved = (torch.randn, torch.randn_like)
        torch.randn = lambda *a, **k: self.q.pop(0)
        torch.randn_like = lambda *a, **k: self.q.pop(0)
        return self

    def __exit__(self, *exc):
        torch.randn, torch.randn_like = self.saved


def reference_chain(net, cond, noise):
    """Run the reference p_sample_loop step by step, recording every x."""
    T = net.num_timesteps
    xs = []
    with inject_noise(noise):
        img = torch.randn(cond.shape)
        xs.append(img)
        for i in reversed(range(T)):
            img = net.p_sample(img, i, condition_x=cond)
            xs.append(img)
    return xs  # xs[k] = x after k steps; xs[0] = x_T


def main():
    torch.set_num_threads(os.cpu_count())
    os.makedirs(OUT, exist_ok=True)
    report = []

    # ---- case A: free-running short chain, R=32, B=2, T=10 (config 1 "short T") -------------
    mopt = model_opt(10)
    sd = make_state_dict(mopt, seed=0, gain=1.0)
    digest0 = state_dict_digest(sd)
    cond, noise = make_inputs(2, 32, 10, seed=123)
    net = build_reference(mopt, sd)
    xs = reference_chain(net, cond, noise)
    with inject_noise(noise):
        snaps = net.super_resolution(cond, continous=True)
    with inject_noise(noise):
        last = net.super_resolution(cond, continous=False)
    tabs = O.schedule_tables(mopt["beta_schedule"]["val"])
    rec = []
    fin, osnaps = O.sample_loop(sd, mopt, tabs, cond, noise, record=lambda t, a, b: rec.append(b))
    dmax = max(float((a - b).abs().max()) for a, b in zip(rec, xs[1:]))
    assert dmax <= TOL_CHAIN, f"oracle != reference on chain A: {dmax}"
    assert float((osnaps - snaps).abs().max()) <= TOL_CHAIN and float((fin[-1] - last).abs().max()) <= TOL_CHAIN
    report.append(f"A r32 B2 T10 free-running chain: max|oracle-reference| = {dmax:.2e} over {len(rec)} steps")
    np.savez_compressed(os.path.join(OUT, "chain_r32_T10.npz"), cond=cond.numpy(), noise=noise.numpy(),
                        xs=torch.stack(xs).numpy(), snapshots=snaps.numpy(), last=last.numpy(),
                        weight_seed=0, weight_gain=1.0, weight_sha256=digest0)

    # ---- case B: teacher-forced steps out of the full T=400 chain of config 1 ---------------
    mopt = model_opt(400)
    cond, noise = make_inputs(2, 32, 400, seed=321)
    net = build_reference(mopt, sd)
    xs = reference_chain(net, cond, noise)
    tabs = O.schedule_tables(mopt["beta_schedule"]["val"])
    keep = [399, 398, 300, 200, 100, 10, 1, 0]
    x_t, z_t, x_tm1 = [], [], []
    dB = 0.0
    for t in keep:
        xin = xs[400 - 1 - t]
        z = noise[400 - t] if t > 0 else torch.zeros_like(xin)
        out = O.p_sample(sd, mopt, tabs, xin, t, cond, z)
        dB = max(dB, float((out - xs[400 - t]).abs().max()))
        assert dB <= TOL_STEP, f"oracle != reference at t={t}: {dB}"
        x_t.append(xin); z_t.append(z); x_tm1.append(xs[400 - t])
    report.append(f"B r32 B2 T400 teacher-forced t={keep}: max|oracle-reference| = {dB:.2e}")
    np.savez_compressed(os.path.join(OUT, "steps_r32_T400.npz"), cond=cond.numpy(), t=np.array(keep),
                        x_t=torch.stack(x_t).numpy(), z_t=torch.stack(z_t).numpy(),
                        x_tm1=torch.stack(x_tm1).numpy(), final=xs[-1].numpy(),
                        weight_seed=0, weight_gain=1.0, weight_sha256=digest0)

    # ---- case C: one UNet forward at R=64, harsher weights (gain 1.7) ------------------------
    mopt = model_opt(200)
    sd2 = make_state_dict(mopt, seed=7, gain=1.7)
    net = build_reference(mopt, sd2)
    cond, noise = make_inputs(1, 64, 2, seed=55)
    x6 = torch.cat([cond, noise[0]], dim=1)
    nl = torch.full((1, 1), 0.37, dtype=torch.float32)
    with torch.no_grad():
        eps_ref = net.denoise_fn(x6, nl)
        eps_or = O.unet_forward(sd2, mopt, x6, nl)
    dC = float((eps_ref - eps_or).abs().max())
    assert dC <= TOL_STEP, dC
    report.append(f"C r64 B1 unet forward gain1.7: max|oracle-reference| = {dC:.2e}, eps std {float(eps_ref.std()):.3f}")
    np.savez_compressed(os.path.join(OUT, "unet_r64.npz"), x6=x6.numpy(), noise_level=nl.numpy(),
                        eps=eps_ref.numpy(), weight_seed=7, weight_gain=1.7,
                        weight_sha256=state_dict_digest(sd2))

    # ---- case D: schedule tables for every T the YAMLs use ----------------------------------
    tab_out = {}
    for T in (100, 200, 400, 600, 1000):
        mopt = model_opt(T)
        net.set_new_noise_schedule(mopt["beta_schedule"]["val"], [torch.device("cpu")])
        tabs = O.schedule_tables(mopt["beta_schedule"]["val"])
        pairs = [("sqrt_recip_ac", net.sqrt_recip_alphas_cumprod), ("sqrt_recipm1_ac", net.sqrt_recipm1_alphas_cumprod),
                 ("coef1", net.posterior_mean_coef1), ("coef2", net.posterior_mean_coef2),
                 ("post_logvar", net.posterior_log_variance_clipped), ("betas", net.betas)]
        for name, ref in pairs:
            assert torch.equal(tabs[name], ref), (T, name)
            tab_out[f"T{T}_{name}"] = ref.numpy()
        assert np.array_equal(tabs["sqrt_ac_prev"], net.sqrt_alphas_cumprod_prev)
        tab_out[f"T{T}_sqrt_ac_prev"] = net.sqrt_alphas_cumprod_prev
    report.append("D schedule tables T in {100,200,400,600,1000}: bit-exact")
    np.savez_compressed(os.path.join(OUT, "schedules.npz"), **tab_out)

    with open(os.path.join(OUT, "README.md"), "w") as f:
        f.write("# Golden vectors\n\nGenerated by `python -m oracle.make_golden` from the unmodified reference at\n"
                "`/root/reference` (CPU fp32, torch " + torch.__version__ + "). Weights are regenerated from\n"
                "`oracle.weights.make_state_dict(seed, gain)` and pinned by sha256.\n\n"
                + "\n".join("* " + r for r in report) + "\n")
    print("\n".join(report))


if __name__ == "__main__":
    main()
