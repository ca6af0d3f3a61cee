"""Host-side mirror of the reference denoiser (model/sr/sr3_modules/unet.py:161-265).

This module only OWNS the parameters, under exactly the reference's state_dict names, so that
checkpoints, optimizers and `named_parameters()` callers keep working. Sampling never runs
this module: GaussianDiffusion hands the parameters to libb200sr3 (sm_100a kernels). `forward`
is the differentiable path the reference's *training* loss needs (diffusion.py:284-313) and is
plain autograd-capable torch — training is outside the accelerated hot path.
"""
import math

import torch
import torch.nn.functional as F
from torch import nn


class _Node(nn.Module):
    """Anonymous container; children are attached by name so keys match the reference."""


def _child(root, dotted):
    node = root
    for part in dotted.split("."):
        if part not in node._modules:
            node.add_module(part, _Node())
        node = node._modules[part]
    return node


def _attr(root, dotted):
    node = root
    for part in dotted.split("."):
        node = node._modules[part]
    return node


def layer_plan(in_channel, out_channel, inner_channel, channel_mults, attn_res, res_blocks, image_size):
    """The constructor walk of unet.py:185-233 as data: dicts with name/kind/cx/cskip/cout/attn."""
    attn_res = list(attn_res) if isinstance(attn_res, (list, tuple)) else [attn_res]
    plan = [dict(name="downs.0", kind="head", cx=in_channel, cskip=0, cout=inner_channel, attn=False)]
    pre, res, idx = inner_channel, image_size, 1
    feats = [pre]
    last_level = len(channel_mults) - 1
    for lvl, mult in enumerate(channel_mults):
        ch = inner_channel * mult
        for _ in range(res_blocks):
            plan.append(dict(name=f"downs.{idx}", kind="res", cx=pre, cskip=0, cout=ch, attn=res in attn_res))
            idx += 1
            feats.append(ch)
            pre = ch
        if lvl != last_level:
            plan.append(dict(name=f"downs.{idx}", kind="down", cx=pre, cskip=0, cout=pre, attn=False))
            idx += 1
            feats.append(pre)
            res //= 2
    plan.append(dict(name="mid.0", kind="res", cx=pre, cskip=0, cout=pre, attn=True))
    plan.append(dict(name="mid.1", kind="res", cx=pre, cskip=0, cout=pre, attn=False))
    idx = 0
    for lvl in range(last_level, -1, -1):
        ch = inner_channel * channel_mults[lvl]
        for _ in range(res_blocks + 1):
            plan.append(dict(name=f"ups.{idx}", kind="res", cx=pre, cskip=feats.pop(), cout=ch, attn=res in attn_res))
            idx += 1
            pre = ch
        if lvl >= 1:
            plan.append(dict(name=f"ups.{idx}", kind="up", cx=pre, cskip=0, cout=pre, attn=False))
            idx += 1
            res *= 2
    plan.append(dict(name="final_conv", kind="final", cx=pre, cskip=0, cout=out_channel, attn=False))
    return plan


class UNet(nn.Module):
    """Same constructor signature as the reference UNet (unet.py:162-174)."""

    def __init__(self, in_channel=6, out_channel=3, inner_channel=32, norm_groups=32, channel_mults=(1, 2, 4, 8, 8),
                 attn_res=(8,), res_blocks=3, dropout=0, with_noise_level_emb=True, image_size=128):
        super().__init__()
        if not with_noise_level_emb:
            raise NotImplementedError("b200sr3 supports the noise-level-conditioned UNet only")
        out_channel = out_channel if out_channel is not None else in_channel
        self.inner_channel = inner_channel
        self.norm_groups = norm_groups
        self.dropout = float(dropout)
        self.plan = layer_plan(in_channel, out_channel, inner_channel, list(channel_mults), attn_res, res_blocks, image_size)

        self._dense("noise_level_mlp.1", (inner_channel * 4, inner_channel))
        self._dense("noise_level_mlp.3", (inner_channel, inner_channel * 4))
        for l in self.plan:
            n, cin, cout = l["name"], l["cx"] + l["cskip"], l["cout"]
            if l["kind"] == "head":
                self._dense(n, (cout, cin, 3, 3))
            elif l["kind"] in ("down", "up"):
                self._dense(n + ".conv", (cout, cin, 3, 3))
            elif l["kind"] == "final":
                self._norm(n + ".block.0", cin)
                self._dense(n + ".block.3", (cout, cin, 3, 3))
            else:
                rb = n + ".res_block"
                self._norm(rb + ".block1.block.0", cin)
                self._dense(rb + ".block1.block.3", (cout, cin, 3, 3))
                self._dense(rb + ".noise_func.noise_func.0", (cout, inner_channel))
                self._norm(rb + ".block2.block.0", cout)
                self._dense(rb + ".block2.block.3", (cout, cout, 3, 3))
                if cin != cout:
                    self._dense(rb + ".res_conv", (cout, cin, 1, 1))
                if l["attn"]:
                    self._norm(n + ".attn.norm", cout)
                    self._dense(n + ".attn.qkv", (cout * 3, cout, 1, 1), bias=False)
                    self._dense(n + ".attn.out", (cout, cout, 1, 1))

    # -- parameter registration (torch's default init for Conv2d / Linear / GroupNorm) ----------
    def _dense(self, path, shape, bias=True):
        node = _child(self, path)
        w = torch.empty(*shape)
        nn.init.kaiming_uniform_(w, a=math.sqrt(5))
        node.register_parameter("weight", nn.Parameter(w))
        if bias:
            fan_in = int(torch.tensor(shape[1:]).prod())
            bound = 1.0 / math.sqrt(fan_in)
            node.register_parameter("bias", nn.Parameter(torch.empty(shape[0]).uniform_(-bound, bound)))

    def _norm(self, path, channels):
        node = _child(self, path)
        node.register_parameter("weight", nn.Parameter(torch.ones(channels)))
        node.register_parameter("bias", nn.Parameter(torch.zeros(channels)))

    def init_orthogonal(self):
        """What define_G applies when opt['phase'] == 'train' (networks.py:44-57, 110-112). Writes through the
        parameters themselves (not `.data`), so their version counters move and a sampling engine that already packed
        the old values repacks (GaussianDiffusion._engine)."""
        with torch.no_grad():
            for name, p in self.named_parameters():
                if name.endswith(".weight") and p.dim() >= 2:
                    nn.init.orthogonal_(p, gain=1)
                elif name.endswith(".bias") and ("block.0" not in name and "norm" not in name):
                    p.zero_()

    def tensors(self):
        """(key, tensor) pairs in the naming libb200sr3 expects (no 'denoise_fn.' prefix)."""
        return list(self.named_parameters())

    # -- differentiable torch path (training only) ----------------------------------------------
    def _p(self, path):
        node = _attr(self, path)
        return node._parameters["weight"], node._parameters.get("bias")

    def _block(self, path, x, drop):
        gw, gb = self._p(path + ".block.0")
        h = F.group_norm(x, self.norm_groups, gw, gb, 1e-5)
        h = h * torch.sigmoid(h)
        if drop and self.dropout > 0:
            h = F.dropout(h, self.dropout, self.training)
        w, b = self._p(path + ".block.3")
        return F.conv2d(h, w, b, padding=1)

    def forward(self, x, time):
        inner = self.inner_channel
        step = torch.arange(inner // 2, dtype=time.dtype, device=time.device) / (inner // 2)
        enc = time.unsqueeze(1) * torch.exp(-math.log(1e4) * step.unsqueeze(0))
        enc = torch.cat([torch.sin(enc), torch.cos(enc)], dim=-1)
        t = F.linear(enc, *self._p("noise_level_mlp.1"))
        t = F.linear(t * torch.sigmoid(t), *self._p("noise_level_mlp.3"))
        feats = []
        for l in self.plan:
            n, kind = l["name"], l["kind"]
            if kind == "head":
                x = F.conv2d(x, *self._p(n), padding=1)
            elif kind == "down":
                x = F.conv2d(x, *self._p(n + ".conv"), stride=2, padding=1)
            elif kind == "up":
                x = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), *self._p(n + ".conv"), padding=1)
            elif kind == "final":
                x = self._block(n, x, drop=False)
            else:
                if l["cskip"]:
                    x = torch.cat((x, feats.pop()), dim=1)
                rb = n + ".res_block"
                h = self._block(rb + ".block1", x, drop=False)
                h = h + F.linear(t, *self._p(rb + ".noise_func.noise_func.0")).view(x.shape[0], -1, 1, 1)
                h = self._block(rb + ".block2", h, drop=True)
                if l["cx"] + l["cskip"] != l["cout"]:
                    x = F.conv2d(x, *self._p(rb + ".res_conv"))
                x = h + x
                if l["attn"]:
                    b, c, hh, ww = x.shape
                    gw, gb = self._p(n + ".attn.norm")
                    qkv = F.conv2d(F.group_norm(x, self.norm_groups, gw, gb, 1e-5), self._p(n + ".attn.qkv")[0])
                    q, k, v = qkv.view(b, 3, c, hh * ww).unbind(1)
                    att = torch.softmax(torch.bmm(q.transpose(1, 2), k) / math.sqrt(c), dim=-1)
                    o = torch.bmm(v, att.transpose(1, 2)).view(b, c, hh, ww)
                    x = F.conv2d(o, *self._p(n + ".attn.out")) + x
            if n.startswith("downs."):
                feats.append(x)
        return x
