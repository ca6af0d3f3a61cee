"""TEST INFRASTRUCTURE ONLY (oracle/): seeded synthetic SR3 weights.

Not product code. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this.

The reference ships no checkpoint, so parity runs use synthetic weights.
They must be reproducible on the GPU box, where /root/reference does not exist, so they are
drawn from numpy's PCG64 stream (platform-independent) rather than from torch's module-init
order. The key set and shapes follow the reference state_dict contract exactly
(model/sr/sr3_modules/unet.py:161-233, enumerated in SURVEY.md section 8a); `make_golden.py`
proves that by loading the result into the reference module with strict=True.
"""
import hashlib
import math

import numpy as np
import torch

PREFIX = "denoise_fn."


def unet_layout(unet_cfg, image_size):
    """Walk the reference constructor (unet.py:175-233) and return the layer list.

    Each entry is (state-dict prefix, kind, cin, cout, with_attn) with kind in
    {"conv3", "res", "down", "up", "final"}.
    """
    inner = unet_cfg["inner_channel"]
    mults = list(unet_cfg["channel_multiplier"])
    attn_res = unet_cfg["attn_res"]
    attn_res = list(attn_res) if isinstance(attn_res, (list, tuple)) else [attn_res]
    nres = unet_cfg["res_blocks"]
    layers = [("downs.0", "conv3", unet_cfg["in_channel"], inner, False)]
    pre = inner
    feat = [pre]
    now_res = image_size
    idx = 1
    for lvl, m in enumerate(mults):
        last = lvl == len(mults) - 1
        use_attn = now_res in attn_res
        ch = inner * m
        for _ in range(nres):
            layers.append((f"downs.{idx}", "res", pre, ch, use_attn))
            idx += 1
            feat.append(ch)
            pre = ch
        if not last:
            layers.append((f"downs.{idx}", "down", pre, pre, False))
            idx += 1
            feat.append(pre)
            now_res //= 2
    layers.append(("mid.0", "res", pre, pre, True))
    layers.append(("mid.1", "res", pre, pre, False))
    idx = 0
    for lvl in reversed(range(len(mults))):
        last = lvl < 1
        use_attn = now_res in attn_res
        ch = inner * mults[lvl]
        for _ in range(nres + 1):
            layers.append((f"ups.{idx}", "res", pre + feat.pop(), ch, use_attn))
            idx += 1
            pre = ch
        if not last:
            layers.append((f"ups.{idx}", "up", pre, pre, False))
            idx += 1
            now_res *= 2
    out_ch = unet_cfg["out_channel"] if unet_cfg.get("out_channel") is not None else unet_cfg["in_channel"]
    layers.append(("final_conv", "final", pre, out_ch, False))
    return layers


def make_state_dict(model_opt, seed=0, gain=1.0):
    """Synthetic fp32 weights for the `sr.model` block of a reference YAML.

    conv / linear: U(-b, b), b = gain / sqrt(fan_in) (torch's default bound when gain == 1);
    GroupNorm: weight 1 + 0.1 N(0,1), bias 0.1 N(0,1) so the affine part is exercised.
    gain > 1 mimics the larger activations of the reference's orthogonal init
    (networks.py:110-112), the harsher parity case.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    sd = {}

    def put(name, arr):
        sd[PREFIX + name] = torch.from_numpy(np.ascontiguousarray(arr.astype(np.float32)))

    def dense(name, shape, fan_in, bias=True, g=gain):
        b = g / math.sqrt(fan_in)
        put(name + ".weight", rng.uniform(-b, b, size=shape))
        if bias:
            bb = 1.0 / math.sqrt(fan_in)
            put(name + ".bias", rng.uniform(-bb, bb, size=(shape[0],)))

    def gnorm(name, c):
        put(name + ".weight", 1.0 + 0.1 * rng.standard_normal(c))
        put(name + ".bias", 0.1 * rng.standard_normal(c))

    u = model_opt["unet"]
    inner = u["inner_channel"]
    dense("noise_level_mlp.1", (inner * 4, inner), inner)
    dense("noise_level_mlp.3", (inner, inner * 4), inner * 4)
    for prefix, kind, cin, cout, attn in unet_layout(u, model_opt["diffusion"]["image_size"]):
        if kind == "conv3":
            dense(prefix, (cout, cin, 3, 3), cin * 9)
        elif kind in ("down", "up"):
            dense(prefix + ".conv", (cout, cin, 3, 3), cin * 9)
        elif kind == "final":
            gnorm(prefix + ".block.0", cin)
            dense(prefix + ".block.3", (cout, cin, 3, 3), cin * 9)
        else:
            rb = prefix + ".res_block"
            gnorm(rb + ".block1.block.0", cin)
            dense(rb + ".block1.block.3", (cout, cin, 3, 3), cin * 9)
            dense(rb + ".noise_func.noise_func.0", (cout, inner), inner)
            gnorm(rb + ".block2.block.0", cout)
            dense(rb + ".block2.block.3", (cout, cout, 3, 3), cout * 9)
            if cin != cout:
                dense(rb + ".res_conv", (cout, cin, 1, 1), cin)
            if attn:
                gnorm(prefix + ".attn.norm", cout)
                dense(prefix + ".attn.qkv", (cout * 3, cout, 1, 1), cout, bias=False)
                dense(prefix + ".attn.out", (cout, cout, 1, 1), cout)
    return sd


def state_dict_digest(sd):
    """sha256 over the tensors in key order: pins the weight stream across machines."""
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].numpy().tobytes())
    return h.hexdigest()


def make_inputs(batch, res, n_timestep, seed=123):
    """cond in [-1,1] and the injected-noise list [x_T, z_{T-1}, ..., z_1]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    cond = torch.from_numpy((rng.random((batch, 3, res, res)) * 2.0 - 1.0).astype(np.float32))
    noise = torch.from_numpy(rng.standard_normal((n_timestep, batch, 3, res, res)).astype(np.float32))
    return cond, noise
