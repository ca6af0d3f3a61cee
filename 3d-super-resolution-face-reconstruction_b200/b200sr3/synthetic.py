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
