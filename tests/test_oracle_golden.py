"""The CPU oracle against the vectors the REAL reference produced (oracle/make_golden.py)."""
import hashlib
import os

import numpy as np
import pytest
import torch

from conftest import synthetic_weights
from oracle import sr3_oracle as O
from oracle.make_golden import TOL_CHAIN, TOL_STEP, model_opt
from oracle.weights import make_inputs, state_dict_digest


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_weight_stream_is_pinned(golden_dir):
    g = _load(golden_dir, "chain_r32_T10.npz")
    sd = synthetic_weights(int(g["weight_seed"]), float(g["weight_gain"]))
    assert state_dict_digest(sd) == str(g["weight_sha256"])
    assert len(sd) == 337 and sum(v.numel() for v in sd.values()) == 92556931


def test_schedule_tables_match_reference(golden_dir):
    g = _load(golden_dir, "schedules.npz")
    for T in (100, 200, 400, 600, 1000):
        tabs = O.schedule_tables(model_opt(T)["beta_schedule"]["val"])
        for k in ("sqrt_recip_ac", "sqrt_recipm1_ac", "coef1", "coef2", "post_logvar", "betas"):
            assert np.array_equal(tabs[k].numpy(), g[f"T{T}_{k}"]), (T, k)
        assert np.array_equal(tabs["sqrt_ac_prev"], g[f"T{T}_sqrt_ac_prev"])
    # SURVEY.md appendix A check values
    t600 = O.schedule_tables(model_opt(600)["beta_schedule"]["val"])
    assert abs(float(t600["sqrt_recip_ac"][599]) - 4.50496) < 1e-4
    assert abs(float(t600["coef2"][599]) - 0.99446654) < 1e-6
    assert float(t600["coef1"][0]) == 1.0 and float(t600["coef2"][0]) == 0.0


def test_unet_forward_matches_reference(golden_dir):
    g = _load(golden_dir, "unet_r64.npz")
    sd = synthetic_weights(int(g["weight_seed"]), float(g["weight_gain"]))
    assert state_dict_digest(sd) == str(g["weight_sha256"])
    with torch.no_grad():
        eps = O.unet_forward(sd, model_opt(200), torch.from_numpy(g["x6"]), torch.from_numpy(g["noise_level"]))
    assert float((eps - torch.from_numpy(g["eps"])).abs().max()) <= TOL_STEP


def test_teacher_forced_steps_match_reference(golden_dir):
    g = _load(golden_dir, "steps_r32_T400.npz")
    sd = synthetic_weights(int(g["weight_seed"]), float(g["weight_gain"]))
    mopt = model_opt(400)
    tabs = O.schedule_tables(mopt["beta_schedule"]["val"])
    cond = torch.from_numpy(g["cond"])
    for i, t in enumerate(g["t"].tolist()):
        if t not in (399, 100, 0):          # three of the eight stored steps keep the CPU suite short
            continue
        with torch.no_grad():
            out = O.p_sample(sd, mopt, tabs, torch.from_numpy(g["x_t"][i]), t, cond, torch.from_numpy(g["z_t"][i]))
        assert float((out - torch.from_numpy(g["x_tm1"][i])).abs().max()) <= TOL_STEP, t


def test_free_running_chain_matches_reference(golden_dir):
    g = _load(golden_dir, "chain_r32_T10.npz")
    sd = synthetic_weights(int(g["weight_seed"]), float(g["weight_gain"]))
    mopt = model_opt(10)
    tabs = O.schedule_tables(mopt["beta_schedule"]["val"])
    fin, snaps = O.sample_loop(sd, mopt, tabs, torch.from_numpy(g["cond"]), torch.from_numpy(g["noise"]))
    assert float((fin - torch.from_numpy(g["xs"][-1])).abs().max()) <= TOL_CHAIN
    assert snaps.shape == g["snapshots"].shape == (2 * 11, 3, 32, 32)
    assert float((snaps - torch.from_numpy(g["snapshots"])).abs().max()) <= TOL_CHAIN
    assert float((fin[-1] - torch.from_numpy(g["last"])).abs().max()) <= TOL_CHAIN   # continous=False quirk
    assert O.psnr_uint8(fin[0], torch.from_numpy(g["xs"][-1][0])) > 80.0


@pytest.mark.parametrize("fname", ["chain_r64_T200.npz", "chain_r128_T600.npz"])
def test_headline_chain_steps_match_reference(golden_dir, fname):
    """Cases E / F (oracle/make_golden_headline.py): the oracle reproduces x_{t-1} of the unmodified reference's own
    full-T chains of the headline configs (16->128 T=600, 16->64 T=200) at the first step and at t = 0."""
    g = _load(golden_dir, fname)
    T, R, B = int(g["T"]), int(g["R"]), int(g["B"])
    sd = synthetic_weights(int(g["weight_seed"]), float(g["weight_gain"]))
    assert state_dict_digest(sd) == str(g["weight_sha256"])
    cond, noise = make_inputs(B, R, T, seed=int(g["input_seed"]))
    assert np.array_equal(cond.numpy(), g["cond"])
    assert hashlib.sha256(noise.numpy().tobytes()).hexdigest() == str(g["noise_sha256"])     # PCG64 stream is portable
    mopt = model_opt(T)
    tabs = O.schedule_tables(mopt["beta_schedule"]["val"])
    state = {0: noise[0]}
    state.update({int(k): torch.from_numpy(x) for k, x in zip(g["keep_k"], g["xs"])})
    assert np.array_equal(g["xs"][-1], g["final"])
    for t in (T - 1, 0):
        z = noise[T - t] if t > 0 else torch.zeros_like(cond)
        with torch.no_grad():
            out = O.p_sample(sd, mopt, tabs, state[T - 1 - t], t, cond, z)
        assert float((out - state[T - t]).abs().max()) <= TOL_STEP, (fname, t)


def test_arcface_oracle_matches_reference(golden_dir):
    """Case I (oracle/make_golden_arcface.py): iResNet-100 + F.normalize + MappingNetwork against the reference modules."""
    import torch.nn.functional as F
    from oracle import arcface_oracle as A
    g = _load(golden_dir, "arcface_b2.npz")
    arc, mp = A.make_arcface_state_dict(int(g["arcface_weight_seed"])), A.make_mapping_state_dict(int(g["mapping_weight_seed"]))
    assert A.digest(arc) == str(g["arcface_sha256"]) and A.digest(mp) == str(g["mapping_sha256"])
    blob = A.make_blob(2, seed=int(g["blob_seed"]))
    assert hashlib.sha256(blob.numpy().tobytes()).hexdigest() == str(g["blob_sha256"])
    with torch.no_grad():
        emb = A.arcface_forward(arc, blob)
        ident, shape = A.mica_encode(arc, mp, blob)
    assert float((emb - torch.from_numpy(g["embedding"])).abs().max()) <= 1e-4
    assert float((ident - torch.from_numpy(g["identity"])).abs().max()) <= 1e-6
    assert float((shape - torch.from_numpy(g["shape_code"])).abs().max()) <= 1e-6
    assert len(arc) == 925 and sum(v.numel() for k, v in arc.items() if "running" not in k and "tracked" not in k) == 65156160
