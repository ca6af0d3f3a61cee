"""Seeded synthetic weights and inputs for benchmarks and smoke runs (the reference ships no
checkpoint and there is no network). numpy PCG64 streams, so every machine gets the same bits.
tests/test_host_cpu.py checks this generator against the oracle's independent copy by sha256."""
import math

import numpy as np
import torch


def state_dict(net, seed=0, gain=1.0):
    """Weights for every parameter of a define_G module, in the reference's registration order:
    conv/linear U(-gain/sqrt(fan_in), +), their biases U(-1/sqrt(fan_in), +), GroupNorm weight
    1 + 0.1 N(0,1) and bias 0.1 N(0,1)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = {}
    params = dict(net.named_parameters())
    names = list(params)
    for name in names:
        if not name.endswith(".weight"):
            continue
        p = params[name]
        bias = name[:-6] + "bias"
        if p.dim() == 1:                                   # GroupNorm
            out[name] = 1.0 + 0.1 * rng.standard_normal(p.shape[0])
            out[bias] = 0.1 * rng.standard_normal(p.shape[0])
        else:
            fan_in = int(np.prod(p.shape[1:]))
            b = gain / math.sqrt(fan_in)
            out[name] = rng.uniform(-b, b, size=tuple(p.shape))
            if bias in params:
                bb = 1.0 / math.sqrt(fan_in)
                out[bias] = rng.uniform(-bb, bb, size=(p.shape[0],))
    return {k: torch.from_numpy(np.ascontiguousarray(v.astype(np.float32))) for k, v in out.items()}


def inputs(batch, res, n_noise=0, seed=123):
    """cond in [-1,1] (datasets/LRHR_dataset.py:93-98 range) and optionally a noise list."""
    rng = np.random.Generator(np.random.PCG64(seed))
    cond = torch.from_numpy((rng.random((batch, 3, res, res)) * 2.0 - 1.0).astype(np.float32))
    if not n_noise:
        return cond
    noise = torch.from_numpy(rng.standard_normal((n_noise, batch, 3, res, res)).astype(np.float32))
    return cond, noise


def mica_state_dict(module, seed=0):
    """Seeded weights for b200sr3.Arcface / MappingNetwork (or any module made of conv / linear / BatchNorm / PReLU):
    conv and linear weights U(-b, b) with b = sqrt(3 / fan_in) (unit gain), BatchNorm with non-trivial affine
    parameters and running statistics (the last BatchNorm of a residual branch, `bn3`, at 0.3 so that the residual
    stream stays O(1) over 49 blocks), PReLU slopes around 0.25. The reference's own init (conv ~ N(0, 0.1),
    arcface.py:112-118) is meant for training with batch statistics and overflows fp32 in eval mode."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = {}
    for name, t in module.state_dict().items():
        shape = tuple(t.shape)
        leaf = name.rsplit(".", 1)[-1]
        owner = name.rsplit(".", 2)[-2] if name.count(".") else ""
        if leaf == "num_batches_tracked":
            v = np.asarray(1000)
        elif leaf == "running_mean":
            v = 0.1 * rng.standard_normal(shape)
        elif leaf == "running_var":
            v = rng.uniform(0.5, 1.5, size=shape)
        elif len(shape) >= 2:
            v = rng.uniform(-1, 1, size=shape) * math.sqrt(3.0 / int(np.prod(shape[1:])))
        elif owner.startswith("prelu"):
            v = 0.25 + 0.05 * rng.standard_normal(shape)
        elif leaf == "weight":
            v = (0.3 if owner == "bn3" else 1.0) * (1.0 + 0.1 * rng.standard_normal(shape))
        else:
            v = 0.1 * rng.standard_normal(shape)
        out[name] = torch.from_numpy(np.ascontiguousarray(np.asarray(v).astype(np.int64 if leaf == "num_batches_tracked" else np.float32))).reshape(shape)
    return out
