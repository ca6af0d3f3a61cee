"""TEST INFRASTRUCTURE ONLY (oracle/): CPU restatement of the SR3 sampling hot path.

Not product code. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this; the product (b200sr3) never does.

What it restates (reference = zouiner/3d-super-resolution-Face-reconstruction):
  * model/sr/sr3_modules/diffusion.py:20-50,93-142   schedule tables (float64 -> fp32)
  * model/sr/sr3_modules/diffusion.py:144-215        p_mean_variance / p_sample / p_sample_loop
  * model/sr/sr3_modules/unet.py:18-265              the UNet denoiser
as a flat, functional fp32 program over a reference-keyed state_dict. The arithmetic itself
(conv2d, group_norm, sigmoid, softmax, matmul) is torch's CPU fp32 — the same third-party
library the reference calls (requirements.txt pins torch==2.3.1; this image has 2.11), so it
is not re-derived here.

Parity pin: the reference has NO tests, golden vectors or fixtures for this path (SURVEY.md
section 4), so by its own tests parity is unpinned. This oracle is instead pinned against the
reference ITSELF, imported from /root/reference in the build container by
oracle/make_golden.py: identical weights (strict load), identical injected noise, every
x_{t-1} compared (bit-identical except for the attention bmm-vs-einsum reduction order,
max |diff| 1.4e-6; tolerances in make_golden.py); the reference's outputs are committed
under tests/golden/ and tests/test_oracle_golden.py re-checks the oracle against them
wherever the suite runs.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from .weights import PREFIX, unet_layout


# ----------------------------------------------------------------------------- schedule
def beta_schedule(schedule, n_timestep, linear_start, linear_end):
    """diffusion.py:20-50. Only the branches the YAMLs can name without torch tensors."""
    if schedule == "linear":
        return np.linspace(linear_start, linear_end, n_timestep, dtype=np.float64)
    if schedule == "quad":
        return np.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=np.float64) ** 2
    if schedule == "const":
        return linear_end * np.ones(n_timestep, dtype=np.float64)
    if schedule in ("warmup10", "warmup50"):
        frac = 0.1 if schedule == "warmup10" else 0.5
        betas = linear_end * np.ones(n_timestep, dtype=np.float64)
        w = int(n_timestep * frac)
        betas[:w] = np.linspace(linear_start, linear_end, w, dtype=np.float64)
        return betas
    if schedule == "jsd":
        return 1.0 / np.linspace(n_timestep, 1, n_timestep, dtype=np.float64)
    if schedule == "cosine":
        s = 8e-3
        ts = np.arange(n_timestep + 1, dtype=np.float64) / n_timestep + s
        a = np.cos(ts / (1 + s) * math.pi / 2) ** 2
        a = a / a[0]
        return np.minimum(1 - a[1:] / a[:-1], 0.999)
    raise NotImplementedError(schedule)


def schedule_tables(schedule_opt):
    """diffusion.py:93-142: the per-t scalars the sampler reads, fp64 math, fp32 storage."""
    betas = beta_schedule(schedule_opt["schedule"], schedule_opt["n_timestep"],
                          schedule_opt["linear_start"], schedule_opt["linear_end"])
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
    f32 = lambda a: torch.tensor(a, dtype=torch.float32)
    return {
        "T": int(betas.shape[0]),
        "betas": f32(betas),
        "sqrt_recip_ac": f32(np.sqrt(1.0 / ac)),
        "sqrt_recipm1_ac": f32(np.sqrt(1.0 / ac - 1)),
        "post_logvar": f32(np.log(np.maximum(post_var, 1e-20))),
        "coef1": f32(betas * np.sqrt(ac_prev) / (1.0 - ac)),
        "coef2": f32((1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac)),
        # diffusion.py:108-109 keeps this one as float64 numpy of length T+1
        "sqrt_ac_prev": np.sqrt(np.append(1.0, ac)),
    }


# ----------------------------------------------------------------------------- UNet pieces
def _swish(x):  # unet.py:53-55
    return x * torch.sigmoid(x)


def _gn(sd, key, x, groups):
    return F.group_norm(x, groups, sd[key + ".weight"], sd[key + ".bias"], eps=1e-5)


def noise_embedding(sd, noise_level, inner):
    """unet.py:18-31 + 179-184: [B,1] noise level -> [B,1,inner]."""
    count = inner // 2
    step = torch.arange(count, dtype=noise_level.dtype, device=noise_level.device) / count
    enc = noise_level.unsqueeze(1) * torch.exp(-math.log(1e4) * step.unsqueeze(0))
    enc = torch.cat([torch.sin(enc), torch.cos(enc)], dim=-1)
    h = F.linear(enc, sd[PREFIX + "noise_level_mlp.1.weight"], sd[PREFIX + "noise_level_mlp.1.bias"])
    return F.linear(_swish(h), sd[PREFIX + "noise_level_mlp.3.weight"], sd[PREFIX + "noise_level_mlp.3.bias"])


def _block(sd, key, x, groups):
    """unet.py:80-91 in eval mode (Dropout is the identity)."""
    h = _swish(_gn(sd, key + ".block.0", x, groups))
    return F.conv2d(h, sd[key + ".block.3.weight"], sd[key + ".block.3.bias"], padding=1)


def _resnet(sd, key, x, emb, groups):
    """unet.py:94-110 with use_affine_level=False (unet.py:34-50)."""
    h = _block(sd, key + ".block1", x, groups)
    nb = F.linear(emb, sd[key + ".noise_func.noise_func.0.weight"], sd[key + ".noise_func.noise_func.0.bias"])
    h = h + nb.view(x.shape[0], -1, 1, 1)
    h = _block(sd, key + ".block2", h, groups)
    if key + ".res_conv.weight" in sd:
        x = F.conv2d(x, sd[key + ".res_conv.weight"], sd[key + ".res_conv.bias"])
    return h + x


def _attention(sd, key, x, groups):
    """unet.py:113-142, n_head = 1."""
    b, c, hh, ww = x.shape
    n = _gn(sd, key + ".norm", x, groups)
    qkv = F.conv2d(n, sd[key + ".qkv.weight"])
    q, k, v = qkv.view(b, 3, c, hh * ww).unbind(1)            # q = ch 0..C-1, k, v follow
    att = torch.softmax(torch.bmm(q.transpose(1, 2), k) / math.sqrt(c), dim=-1)   # [b, hw_q, hw_k]
    out = torch.bmm(v, att.transpose(1, 2)).view(b, c, hh, ww)
    out = F.conv2d(out, sd[key + ".out.weight"], sd[key + ".out.bias"])
    return out + x


def unet_forward(sd, model_opt, x, noise_level, taps=None):
    """unet.py:235-265. x = cat([cond, x_t], 1) fp32 NCHW, noise_level [B,1]. Returns eps.

    `taps`, if a dict, receives named intermediate activations for layer-level parity tests.
    """
    u = model_opt["unet"]
    groups = u.get("norm_groups") or 32
    emb = noise_embedding(sd, noise_level, u["inner_channel"])
    feats = []
    for prefix, kind, cin, cout, attn in unet_layout(u, model_opt["diffusion"]["image_size"]):
        key = PREFIX + prefix
        if prefix.startswith("ups.") and kind == "res":
            x = torch.cat((x, feats.pop()), dim=1)            # unet.py:261, x first
        if kind == "conv3":
            x = F.conv2d(x, sd[key + ".weight"], sd[key + ".bias"], padding=1)
        elif kind == "down":
            x = F.conv2d(x, sd[key + ".conv.weight"], sd[key + ".conv.bias"], stride=2, padding=1)
        elif kind == "up":
            x = F.interpolate(x, scale_factor=2, mode="nearest")
            x = F.conv2d(x, sd[key + ".conv.weight"], sd[key + ".conv.bias"], padding=1)
        elif kind == "res":
            x = _resnet(sd, key + ".res_block", x, emb, groups)
            if attn:
                x = _attention(sd, key + ".attn", x, groups)
        elif kind == "final":
            x = _block(sd, key, x, groups)
        if prefix.startswith("downs."):
            feats.append(x)
        if taps is not None:
            taps[prefix] = x
    return x


# ----------------------------------------------------------------------------- sampler
def p_sample(sd, model_opt, tabs, x, t, cond, noise, clip_denoised=True):
    """One reverse step, diffusion.py:164-187. `noise` is z_t (ignored at t == 0); cond None = the unconditional
    branch (diffusion.py:172-173: the UNet sees x alone)."""
    b = x.shape[0]
    nl = torch.FloatTensor([tabs["sqrt_ac_prev"][t + 1]]).repeat(b, 1).to(x.device)
    eps = unet_forward(sd, model_opt, torch.cat([cond, x], dim=1) if cond is not None else x, nl)
    x0 = tabs["sqrt_recip_ac"][t] * x - tabs["sqrt_recipm1_ac"][t] * eps
    if clip_denoised:
        x0.clamp_(-1.0, 1.0)
    mean = tabs["coef1"][t] * x0 + tabs["coef2"][t] * x
    z = noise if t > 0 else torch.zeros_like(x)
    return mean + z * (0.5 * tabs["post_logvar"][t]).exp()


@torch.no_grad()
def sample_loop(sd, model_opt, tabs, cond, noise, record=None, t_stop=0):
    """diffusion.py:189-215 with the noise list injected.

    noise[0] is x_T, noise[T - t] is z_t for t >= 1. Returns (final x [B,3,R,R], snapshots
    as in `continous=True`: cat([cond, x at every t % (1 | T//10) == 0]); with cond None (unconditional branch,
    diffusion.py:193-200) the list starts with x_T instead).
    `record(t, x_t, x_tm1)` is called after every step; t_stop > 0 truncates the chain.
    """
    T = tabs["T"]
    inter = 1 | (T // 10)
    x = noise[0]
    ret = cond if cond is not None else x
    for t in reversed(range(t_stop, T)):
        x_new = p_sample(sd, model_opt, tabs, x, t, cond, noise[T - t] if t > 0 else None)
        if record is not None:
            record(t, x, x_new)
        x = x_new
        if t % inter == 0:
            ret = torch.cat([ret, x], dim=0)
    return x, ret


# ----------------------------------------------------------------------------- metric
def psnr_uint8(a, b):
    """core/metrics.py:16-42 (tensor2img, 3-D branch) + 74-81 (calculate_psnr), per image."""
    def to_img(t):
        t = t.float().clamp(-1, 1)
        return ((t + 1) / 2 * 255.0).numpy().round().astype(np.uint8).astype(np.float64)
    mse = np.mean((to_img(a) - to_img(b)) ** 2)
    return float("inf") if mse == 0 else 20 * math.log10(255.0 / math.sqrt(mse))
