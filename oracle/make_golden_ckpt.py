"""TEST INFRASTRUCTURE ONLY (oracle/): a `_gen.pth` checkpoint SAVED BY THE REFERENCE MODULE.

Run in the build container only (needs /root/reference):

    python -m oracle.make_golden_ckpt

The reference ships no checkpoint and its full generator is 353 MB, so the fixture is the reference's own module at a
small UNet config (same code path: model/sr/networks.py:83-116 define_G with phase='train', i.e. the reference's own
orthogonal init, networks.py:44-57,110-112), its val schedule installed (diffusion.py:93-142, so the 12 schedule
buffers are in the file as in a real checkpoint), saved exactly as model/sr/model.py:146-155 save_network does
(`state_dict()` -> `.cpu()` -> torch.save to `I{iter}_E{epoch}_gen.pth`). Next to it: an input and the eps the
reference module itself computes from the file's weights, so the GPU test checks ingestion end to end.
The config has attention on a NON-mid level (attn_res = [8] with image_size = 16): the reference places attention by
comparing attn_res with the level's resolution counted down from image_size (unet.py:192-197,211-220).
"""
import copy
import os
import sys

import numpy as np
import torch

from . import sr3_oracle as O
from .make_golden import OUT, REF, TOL_STEP

SMALL = {
    "which_model_G": "sr3", "finetune_norm": False,
    "unet": {"in_channel": 6, "out_channel": 3, "inner_channel": 64, "channel_multiplier": [1, 1],
             "attn_res": [8], "res_blocks": 1, "dropout": 0.2},
    "beta_schedule": {k: {"schedule": "linear", "n_timestep": 20, "linear_start": 1e-6, "linear_end": 1e-2}
                      for k in ("train", "val")},
    "diffusion": {"image_size": 16, "channels": 3, "conditional": True},
}


def main():
    sys.path.insert(0, REF)
    from model.sr.networks import define_G      # noqa: the reference's own factory
    torch.manual_seed(20260)
    mopt = copy.deepcopy(SMALL)
    net = define_G({"sr": {"model": mopt}, "phase": "train"})       # orthogonal init, as in training
    # GroupNorm affine parameters start at (1, 0); move them so the file exercises them
    with torch.no_grad():
        for name, p in net.named_parameters():
            if p.dim() == 1 and ("block.0" in name or "norm" in name):
                p.add_(0.1 * torch.randn_like(p))
    net.set_new_noise_schedule(mopt["beta_schedule"]["val"], [torch.device("cpu")])
    net.eval()
    # model/sr/model.py:146-155
    state_dict = net.state_dict()
    for key, param in state_dict.items():
        state_dict[key] = param.cpu()
    path = os.path.join(OUT, "ref_small_I100_E3_gen.pth")
    torch.save(state_dict, path)

    g = torch.Generator().manual_seed(7)
    x6 = torch.rand(2, 6, 16, 16, generator=g) * 2 - 1
    nl = torch.full((2, 1), 0.61)
    with torch.no_grad():
        eps_ref = net.denoise_fn(x6, nl)
        sd = {k: v for k, v in state_dict.items() if k.startswith("denoise_fn.")}
        eps_or = O.unet_forward(sd, mopt, x6, nl)
    d = float((eps_ref - eps_or).abs().max())
    assert d <= TOL_STEP, d
    np.savez_compressed(os.path.join(OUT, "ref_small_ckpt_io.npz"), x6=x6.numpy(), noise_level=nl.numpy(),
                        eps=eps_ref.numpy(), n_params=sum(p.numel() for p in net.parameters()))
    line = (f"G reference-saved checkpoint ref_small_I100_E3_gen.pth ({os.path.getsize(path) / 1e6:.1f} MB, "
            f"{len(state_dict)} tensors, mults [1,1], attn_res [8] @ image_size 16): "
            f"max|oracle-reference| on its eps = {d:.2e}")
    print(line)
    with open(os.path.join(OUT, "README.md"), "a") as f:
        f.write("* " + line + " (`python -m oracle.make_golden_ckpt`)\n")


UNCOND = {
    "which_model_G": "sr3", "finetune_norm": False,
    "unet": {"in_channel": 3, "out_channel": 3, "inner_channel": 64, "channel_multiplier": [1, 1],
             "attn_res": [8], "res_blocks": 1, "dropout": 0.0},
    "beta_schedule": {k: {"schedule": "linear", "n_timestep": 20, "linear_start": 1e-6, "linear_end": 1e-2}
                      for k in ("train", "val")},
    "diffusion": {"image_size": 16, "channels": 3, "conditional": False},
}


def main_unconditional():
    """Case H: the reference's UNCONDITIONAL branch (diffusion.py:193-200: `sample(batch_size, continous)`, the list
    starts with x_T) and p_sample(clip_denoised=False) (diffusion.py:175-176 skipped), small config, weights from
    oracle.weights.make_state_dict (regenerated wherever the tests run, pinned by sha256)."""
    from .make_golden import build_reference, inject_noise
    from .weights import make_inputs, make_state_dict, state_dict_digest
    mopt = copy.deepcopy(UNCOND)
    sd = make_state_dict(mopt, seed=3, gain=1.3)
    net = build_reference(mopt, sd)
    T, B, R = 20, 2, 16
    _, noise = make_inputs(B, R, T, seed=808)
    with inject_noise(noise):
        snaps = net.sample(batch_size=B, continous=True)          # [B * (1 + 10), 3, R, R], x_T first
    with inject_noise(noise):
        last = net.sample(batch_size=B, continous=False)
    tabs = O.schedule_tables(mopt["beta_schedule"]["val"])
    fin, osnaps = O.sample_loop(sd, mopt, tabs, None, noise)
    d = float((osnaps - snaps).abs().max())
    assert d <= 5e-5 and float((fin[-1] - last).abs().max()) <= 5e-5, d
    # one step without the x0 clamp, at a t where the clamp matters (large |x|)
    x_t, z = 3.0 * noise[1], noise[2]
    with inject_noise([z]):
        noclip = net.p_sample(x_t, 15, clip_denoised=False)
    with inject_noise([z]):
        clip = net.p_sample(x_t, 15, clip_denoised=True)
    assert float((noclip - clip).abs().max()) > 0.1               # the flag matters on this input
    d2 = float((O.p_sample(sd, mopt, tabs, x_t, 15, None, z, clip_denoised=False) - noclip).abs().max())
    assert d2 <= TOL_STEP, d2
    np.savez_compressed(os.path.join(OUT, "uncond_r16_T20.npz"), noise=noise.numpy(), snapshots=snaps.numpy(),
                        last=last.numpy(), x_t=x_t.numpy(), z=z.numpy(), t=15, noclip=noclip.numpy(), clip=clip.numpy(),
                        weight_seed=3, weight_gain=1.3, weight_sha256=state_dict_digest(sd))
    line = (f"H unconditional r16 B2 T20 sample(continous=True) + p_sample(clip_denoised=False): "
            f"max|oracle-reference| = {d:.2e} / {d2:.2e}")
    print(line)
    with open(os.path.join(OUT, "README.md"), "a") as f:
        f.write("* " + line + " (`python -m oracle.make_golden_ckpt`)\n")


if __name__ == "__main__":
    if "H" in sys.argv[1:] or not sys.argv[1:]:
        main_unconditional()
    if "G" in sys.argv[1:] or not sys.argv[1:]:
        main()
