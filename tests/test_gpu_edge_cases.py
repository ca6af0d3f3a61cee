"""Reference behaviours off the headline path: a checkpoint saved by the reference module,
attention on a non-mid level and with more tokens than the tensor-core attention kernel takes (unet.py:192-197,211-220),
the unconditional branch (diffusion.py:193-200,217-221), clip_denoised=False (diffusion.py:175-176), and the host-side
caches (weights, workspaces). Goldens: oracle/make_golden_ckpt.py (the unmodified reference)."""
import copy
import ctypes as C
import os
import shutil

import numpy as np
import pytest
import torch

import b200sr3
from b200sr3 import checkpoint
from oracle import sr3_oracle as O
from oracle.make_golden_ckpt import SMALL, UNCOND
from oracle.weights import make_inputs, make_state_dict, state_dict_digest

pytestmark = pytest.mark.gpu

TOL_EPS_RMS, TOL_EPS_MAX = 0.01, 0.06      # as tests/test_gpu_parity.py


def _build(mopt, sd=None, device="cuda"):
    mopt = copy.deepcopy(mopt)
    net = b200sr3.define_G({"phase": "val", "sr": {"model": mopt}})
    if sd is not None:
        net.load_state_dict(sd, strict=True)
    net = net.to(device).eval()
    net.set_new_noise_schedule(mopt["beta_schedule"]["val"], [torch.device(device)])
    return net, mopt


def _close(got, ref, rms_tol=TOL_EPS_RMS, max_tol=TOL_EPS_MAX):
    rms = float(ref.pow(2).mean().sqrt())
    assert float((got - ref).pow(2).mean().sqrt()) <= rms_tol * rms
    assert float((got - ref).abs().max()) <= max_tol * rms


def test_checkpoint_saved_by_the_reference_module(golden_dir, tmp_path):
    """`I100_E3_gen.pth` written by the reference's own module and save code -> checkpoint.load_network -> engine:
    bit-identical to loading the same tensors directly, and the eps the REFERENCE computed from that file."""
    io = np.load(os.path.join(golden_dir, "ref_small_ckpt_io.npz"))
    prefix = os.path.join(tmp_path, "I100_E3")
    shutil.copy(os.path.join(golden_dir, "ref_small_I100_E3_gen.pth"), prefix + "_gen.pth")
    net, mopt = _build(SMALL)
    missing, unexpected = checkpoint.load_network(net, prefix, strict=True)
    assert not missing and not unexpected
    assert sum(p.numel() for p in net.parameters()) == int(io["n_params"])
    x6 = torch.from_numpy(io["x6"]).cuda()
    nl = float(io["noise_level"][0, 0])
    eps = net.unet_eps(x6[:, :3], x6[:, 3:], nl)
    _close(eps.cpu(), torch.from_numpy(io["eps"]))
    # the same tensors loaded without the checkpoint module
    state = torch.load(prefix + "_gen.pth", map_location="cpu", weights_only=True)
    direct, _ = _build(SMALL, {k: v for k, v in state.items() if k.startswith("denoise_fn.")})
    assert torch.equal(direct.unet_eps(x6[:, :3], x6[:, 3:], nl), eps)
    # the file carries the reference's T=20 schedule buffers; sampling with them works end to end
    assert net.num_timesteps == 20 and torch.equal(net.betas.cpu(), state["betas"])
    out = net.super_resolution_batched(x6[:, :3], seed=1)
    assert torch.isfinite(out).all() and float(out.abs().max()) <= 1.0 + 1e-6
    # every module output (attention sits on downs.3 / ups.0 / ups.1 here, not only on mid.0)
    taps = {}
    sd = {k: v for k, v in state.items() if k.startswith("denoise_fn.")}
    with torch.no_grad():
        O.unet_forward(sd, mopt, x6.cpu(), torch.from_numpy(io["noise_level"]), taps)
    net.unet_eps(x6[:, :3], x6[:, 3:], nl)
    eng = net._engine()
    for name, t in taps.items():
        if name == "final_conv":
            continue
        buf = torch.empty(t.shape, device="cuda")
        c, h, w = C.c_int(), C.c_int(), C.c_int()
        assert eng.lib.b200sr3_layer_output(eng.handle, name.encode(), C.c_void_p(buf.data_ptr()), C.byref(c),
                                            C.byref(h), C.byref(w), C.c_void_p(0)) == 0, name
        lrms = float(t.pow(2).mean().sqrt())
        assert float((buf.cpu() - t).pow(2).mean().sqrt()) <= 0.02 * lrms, name


def test_attention_with_256_tokens_on_a_down_and_up_level():
    """attn_res hits the 16x16 level of a 32 px model: 256 tokens per image (the mma attention kernel takes <= 64)."""
    mopt = copy.deepcopy(SMALL)
    mopt["unet"]["attn_res"] = [16]
    mopt["diffusion"]["image_size"] = 32
    sd = make_state_dict(mopt, seed=11, gain=1.2)
    net, mopt = _build(mopt, sd)
    cond, noise = make_inputs(3, 32, 2, seed=12)
    with torch.no_grad():
        ref = O.unet_forward(sd, mopt, torch.cat([cond, noise[0]], 1), torch.full((3, 1), 0.5))
    _close(net.unet_eps(cond.cuda(), noise[0].cuda(), 0.5).cpu(), ref)


@pytest.mark.parametrize("inner,mults,groups,res_blocks,R", [(128, [1, 2], 16, 1, 32), (64, [1, 2, 2], 32, 3, 64), (192, [1], 32, 2, 16)])
def test_other_unet_hyper_parameters(inner, mults, groups, res_blocks, R):
    """define_G reads inner_channel / channel_multiplier / norm_groups / res_blocks from the YAML (networks.py:91-101);
    every shipped config uses 64 / [1,2,4,8,8] / 32 / 2, the engine must not depend on that."""
    mopt = copy.deepcopy(SMALL)
    mopt["unet"].update(inner_channel=inner, channel_multiplier=mults, norm_groups=groups, res_blocks=res_blocks, attn_res=[])
    mopt["diffusion"]["image_size"] = R
    sd = make_state_dict(mopt, seed=inner, gain=1.1)
    net, mopt = _build(mopt, sd)
    cond, noise = make_inputs(3, R, 2, seed=inner + 1)
    with torch.no_grad():
        ref = O.unet_forward(sd, mopt, torch.cat([cond, noise[0]], 1), torch.full((3, 1), 0.4))
    _close(net.unet_eps(cond.cuda(), noise[0].cuda(), 0.4).cpu(), ref)
    tabs = O.schedule_tables(mopt["beta_schedule"]["val"])
    with torch.no_grad():
        want = O.p_sample(sd, mopt, tabs, noise[0], 7, cond, noise[1])
    got = net.p_sample(noise[0].cuda(), 7, condition_x=cond.cuda(), noise=noise[1].cuda()).cpu()
    assert float((got - want).abs().max()) <= 2e-3


def test_unconditional_sample_and_clip_flag(golden_dir):
    """diffusion.py:193-200: sample(batch_size, continous) on a conditional=False model; the list starts with x_T."""
    g = np.load(os.path.join(golden_dir, "uncond_r16_T20.npz"))
    sd = make_state_dict(UNCOND, seed=int(g["weight_seed"]), gain=float(g["weight_gain"]))
    assert state_dict_digest(sd) == str(g["weight_sha256"])
    net, mopt = _build(UNCOND, sd)
    noise = torch.from_numpy(g["noise"]).cuda()
    B = noise.shape[1]
    snaps = net.p_sample_loop((B, 3, 16, 16), continous=True, noise=noise).cpu()
    ref = torch.from_numpy(g["snapshots"])
    assert snaps.shape == ref.shape and torch.equal(snaps[:B], noise[0].cpu())
    assert float((snaps - ref).abs().max()) <= 5e-3
    last = net.p_sample_loop((B, 3, 16, 16), continous=False, noise=noise).cpu()
    assert last.shape == (3, 16, 16) and float((last - torch.from_numpy(g["last"])).abs().max()) <= 5e-3
    # the public entry point with the library's own noise: x_T comes back first, and is the stream's key-T draw
    torch.manual_seed(3)
    own = net.sample(batch_size=B, continous=True)
    assert own.shape == ref.shape and torch.isfinite(own).all()
    own2 = net.p_sample_loop((B, 3, 16, 16), continous=True, seed=41)
    assert torch.equal(own2[:B], net.philox_normal((B, 3, 16, 16), t=20, seed=41))
    assert net.sample(batch_size=1).shape == (3, 16, 16)
    # clip_denoised=False (never used by the reference's sampler, but part of p_sample's signature)
    x_t, z, t = torch.from_numpy(g["x_t"]).cuda(), torch.from_numpy(g["z"]).cuda(), int(g["t"])
    noclip = net.p_sample(x_t, t, clip_denoised=False, noise=z).cpu()
    clip = net.p_sample(x_t, t, clip_denoised=True, noise=z).cpu()
    assert float((noclip - torch.from_numpy(g["noclip"])).abs().max()) <= 2e-2      # |x0| reaches ~13 unclamped
    assert float((clip - torch.from_numpy(g["clip"])).abs().max()) <= 1e-3
    assert float((noclip - clip).abs().max()) > 0.1


def test_weight_edits_reach_the_engine():
    """edits through `p.data` are invisible to torch's version counter; invalidate_weights() covers them,
    everything else (in-place ops on the parameter, load_state_dict, init_orthogonal) is picked up automatically."""
    net, _ = _build(SMALL, make_state_dict(SMALL, seed=1, gain=1.0))
    cond, noise = make_inputs(1, 16, 1, seed=2)
    cond, x = cond.cuda(), noise[0].cuda()
    base = net.unet_eps(cond, x, 0.7)
    w = net.denoise_fn.get_parameter("final_conv.block.3.bias")
    with torch.no_grad():
        w.add_(0.25)                                             # in place on the parameter: detected
    moved = net.unet_eps(cond, x, 0.7)
    assert float((moved - base - 0.25).abs().max()) < 1e-6
    w.data.add_(0.25)                                            # through .data: needs the explicit call
    net.invalidate_weights()
    assert float((net.unet_eps(cond, x, 0.7) - base - 0.5).abs().max()) < 1e-6
    net.load_state_dict(net.state_dict())                        # load paths invalidate by themselves
    net.denoise_fn.init_orthogonal()
    after = net.unet_eps(cond, x, 0.7)
    assert float((after - base).abs().max()) > 1e-3 and torch.isfinite(after).all()


def test_workspace_cache_is_bounded():
    """one plan per (B,R) forever. The cache now keeps the 3 most recent plans; evicted shapes are rebuilt
    on demand and give the same bits."""
    net, _ = _build(SMALL, make_state_dict(SMALL, seed=1, gain=1.0))
    cond, noise = make_inputs(6, 16, 1, seed=4)
    cond, x = cond.cuda(), noise[0].cuda()
    first = {b: net.unet_eps(cond[:b].contiguous(), x[:b].contiguous(), 0.3) for b in (1, 2, 3, 4, 5, 6)}
    free_after_six = torch.cuda.mem_get_info()[0]
    for b in (1, 6, 3):
        assert torch.equal(net.unet_eps(cond[:b].contiguous(), x[:b].contiguous(), 0.3), first[b])
    assert torch.cuda.mem_get_info()[0] >= free_after_six - (64 << 20)        # nothing accumulated
