"""TEST INFRASTRUCTURE ONLY (oracle/): CPU restatement of the MICA identity encoder.

Not product code. Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may import this.

What it restates (reference = zouiner/3d-super-resolution-Face-reconstruction):
  * model/mica/arcface.py:40-160   IBasicBlock / IResNet (iResNet-100: layers [3, 13, 30, 3], arcface.py:167)
  * model/mica/arcface.py:182-200  Arcface.forward_arcface (eval mode: BatchNorm uses its running statistics,
                                   Dropout is the identity)
  * model/sr3d/model.py:164-170    encode_mica: F.normalize(arcface(arcface_imgs))
  * model/mica/generator.py:31-60  MappingNetwork (z_dim 512, hidden 300, `mapping_layers` = 3, n_shape 300)
as a flat functional fp32 program over reference-keyed state_dicts. The arithmetic (conv2d, batch_norm, prelu, linear)
is torch's CPU fp32, the library the reference itself calls.

Parity pin: the reference has no tests or fixtures for this path; oracle/make_golden_arcface.py imports the unmodified
reference modules, loads identical weights (strict) and compares (max |oracle - reference| in
tests/golden/README.md); the reference's outputs are committed as tests/golden/arcface_b2.npz.
"""
import hashlib
import math

import numpy as np
import torch
import torch.nn.functional as F

LAYERS = (3, 13, 30, 3)          # arcface.py:167
PLANES = (64, 128, 256, 512)
EPS = 1e-5


def _bn(sd, key, x):
    return F.batch_norm(x, sd[key + ".running_mean"], sd[key + ".running_var"], sd[key + ".weight"], sd[key + ".bias"],
                        training=False, eps=EPS)


def block_list():
    """(prefix, inplanes, planes, stride, has_downsample) in forward order (arcface.py:128-155)."""
    out, inplanes = [], 64
    for li, (n, planes) in enumerate(zip(LAYERS, PLANES), start=1):
        for bi in range(n):
            first = bi == 0
            out.append((f"layer{li}.{bi}", inplanes, planes, 2 if first else 1, first))
            inplanes = planes
    return out


def arcface_forward(sd, x, taps=None):
    """arcface.py:182-200 in eval mode. x: fp32 [B,3,112,112] ArcFace blob -> [B,512] (NOT normalised)."""
    x = F.conv2d(x, sd["conv1.weight"], padding=1)
    x = F.prelu(_bn(sd, "bn1", x), sd["prelu.weight"])
    if taps is not None:
        taps["stem"] = x
    for prefix, inplanes, planes, stride, down in block_list():          # arcface.py:58-69
        out = _bn(sd, prefix + ".bn1", x)
        out = F.conv2d(out, sd[prefix + ".conv1.weight"], padding=1)
        out = F.prelu(_bn(sd, prefix + ".bn2", out), sd[prefix + ".prelu.weight"])
        out = F.conv2d(out, sd[prefix + ".conv2.weight"], stride=stride, padding=1)
        out = _bn(sd, prefix + ".bn3", out)
        if down:
            x = _bn(sd, prefix + ".downsample.1", F.conv2d(x, sd[prefix + ".downsample.0.weight"], stride=stride))
        x = out + x
        if taps is not None:
            taps[prefix] = x
    x = torch.flatten(_bn(sd, "bn2", x), 1)
    x = F.linear(x, sd["fc.weight"], sd["fc.bias"])
    return F.batch_norm(x, sd["features.running_mean"], sd["features.running_var"], sd["features.weight"],
                        sd["features.bias"], training=False, eps=EPS)


def mapping_forward(sd, z, hidden=3):
    """generator.py:50-60 (hidden <= 5: no skip connections)."""
    h = z
    for i in range(hidden + 1):
        h = F.leaky_relu(F.linear(h, sd[f"network.{i}.weight"], sd[f"network.{i}.bias"]), negative_slope=0.2)
    return F.linear(h, sd["output.weight"], sd["output.bias"])


def mica_encode(arc_sd, map_sd, blob):
    """model/sr3d/model.py:164-170 + generator.py:88-90 up to the FLAME decoder: identity code and shape code."""
    ident = F.normalize(arcface_forward(arc_sd, blob))
    return ident, mapping_forward(map_sd, ident)


# ----------------------------------------------------------------------------- seeded synthetic weights
def make_arcface_state_dict(seed=0):
    """Reference-keyed iResNet-100 weights from numpy's PCG64 stream (identical on every machine).

    The reference's own init (conv ~ N(0, 0.1), BatchNorm at running_mean 0 / running_var 1, arcface.py:112-118) is
    meant for training with batch statistics; in eval mode it multiplies the activations by ~2.4 per conv and
    overflows fp32 within the 100 layers. So: conv ~ U(-b, b) with b = sqrt(3 / fan_in) (unit gain), every BatchNorm
    with non-trivial affine parameters AND running statistics, PReLU slopes around 0.25, so that each folded term
    (scale, shift, slope) is exercised. bn3 scales are small (0.3): the residual stream stays O(1) over 49 blocks.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    sd = {}

    def put(name, arr, dtype=np.float32):
        a = np.asarray(arr).astype(dtype)
        sd[name] = torch.from_numpy(np.ascontiguousarray(a)).reshape(a.shape)      # (keeps 0-dim entries 0-dim)

    def conv(name, cout, cin, k):
        b = math.sqrt(3.0 / (cin * k * k))
        put(name + ".weight", rng.uniform(-b, b, size=(cout, cin, k, k)))

    def bn(name, c, gain=1.0):
        put(name + ".weight", gain * (1.0 + 0.1 * rng.standard_normal(c)))
        put(name + ".bias", 0.1 * rng.standard_normal(c))
        put(name + ".running_mean", 0.1 * rng.standard_normal(c))
        put(name + ".running_var", rng.uniform(0.5, 1.5, size=c))
        put(name + ".num_batches_tracked", 1000, dtype=np.int64)

    conv("conv1", 64, 3, 3)
    bn("bn1", 64)
    put("prelu.weight", 0.25 + 0.05 * rng.standard_normal(64))
    for prefix, inplanes, planes, stride, down in block_list():
        bn(prefix + ".bn1", inplanes)
        conv(prefix + ".conv1", planes, inplanes, 3)
        bn(prefix + ".bn2", planes)
        put(prefix + ".prelu.weight", 0.25 + 0.05 * rng.standard_normal(planes))
        conv(prefix + ".conv2", planes, planes, 3)
        bn(prefix + ".bn3", planes, gain=0.3)
        if down:
            conv(prefix + ".downsample.0", planes, inplanes, 1)
            bn(prefix + ".downsample.1", planes)
    bn("bn2", 512)
    b = 1.0 / math.sqrt(512 * 49)
    put("fc.weight", rng.uniform(-b, b, size=(512, 512 * 49)))
    put("fc.bias", rng.uniform(-b, b, size=512))
    bn("features", 512)
    return sd


def make_mapping_state_dict(seed=0, z_dim=512, hidden_dim=300, out_dim=300, hidden=3):
    rng = np.random.Generator(np.random.PCG64(seed + 1000))
    sd = {}
    dims = [(hidden_dim, z_dim)] + [(hidden_dim, hidden_dim)] * hidden
    for i, (o, k) in enumerate(dims):
        std = math.sqrt(2.0 / (1 + 0.2 ** 2) / k)                     # kaiming_normal_(a=0.2), generator.py:25-28
        sd[f"network.{i}.weight"] = torch.from_numpy((std * rng.standard_normal((o, k))).astype(np.float32))
        sd[f"network.{i}.bias"] = torch.from_numpy(rng.uniform(-1, 1, size=o).astype(np.float32) / math.sqrt(k))
    b = 1.0 / math.sqrt(hidden_dim)
    sd["output.weight"] = torch.from_numpy((0.25 * rng.uniform(-b, b, size=(out_dim, hidden_dim))).astype(np.float32))
    sd["output.bias"] = torch.from_numpy(rng.uniform(-b, b, size=out_dim).astype(np.float32))
    return sd


def make_blob(batch, seed=0):
    """An ArcFace blob as cv2.dnn.blobFromImages(1/127.5, mean 127.5) leaves it: fp32 [B,3,112,112] in [-1, 1]."""
    rng = np.random.Generator(np.random.PCG64(seed + 2000))
    img = rng.integers(0, 256, size=(batch, 3, 112, 112)).astype(np.float32)
    return torch.from_numpy((img - 127.5) / 127.5)


def digest(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].numpy().tobytes())
    return h.hexdigest()
