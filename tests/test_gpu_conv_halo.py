"""Kernel-level parity of the halo-resident conv with fused GroupNorm+Swish (csrc/conv_halo.cuh,
through the C ABI entry b200sr3_conv_block) against torch fp32 on the same bf16-rounded operands:
one reference Block (unet.py:80-91) with the channel concat of unet.py:261 as two sources and the
ResnetBlock shortcut (unet.py:103-110) folded in as extra K segments.

Tolerance: the bf16 rounding of the output (|y|max * 2^-8) plus, when the GroupNorm is fused, one
bf16 ulp of the normalised activations propagated through the conv (another |y|max * 2^-8) and
2e-3 of accumulation-order / fast-exp noise."""
import ctypes as C
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CASES = [
    # B, C0, C1, Cr0, Cr1, H, W, Cout, gn, up
    (1, 64, 0, 0, 0, 16, 16, 64, 0, 0),
    (1, 64, 0, 0, 0, 16, 16, 64, 1, 0),
    (2, 64, 0, 64, 0, 32, 32, 64, 1, 0),
    (2, 128, 64, 128, 64, 16, 16, 128, 1, 0),       # 192 channels: groups of 6 straddle the 64-blocks and the seam
    (1, 256, 0, 0, 0, 32, 32, 256, 1, 0),
    (2, 512, 256, 512, 256, 16, 16, 512, 1, 0),
    (1, 64, 0, 0, 0, 128, 128, 64, 1, 0),
    (3, 128, 0, 128, 0, 64, 64, 128, 1, 0),
    (2, 128, 0, 0, 0, 16, 16, 128, 0, 1),           # folded nearest-2x upsample, 16 -> 32
    (1, 64, 0, 0, 0, 32, 16, 128, 1, 0),            # non-square
    (4, 512, 0, 512, 0, 8, 8, 512, 1, 0),           # 8x8 images: two images per tile
    (3, 128, 64, 128, 64, 8, 8, 128, 1, 0),         # ... odd batch: the last pair is half empty
    (5, 64, 0, 0, 0, 8, 8, 64, 0, 0),
    (3, 128, 0, 0, 0, 8, 8, 128, 0, 1),             # folded upsample 8 -> 16
    (6, 512, 0, 512, 0, 4, 4, 512, 1, 0),           # 4x4 images: five images per tile (linear 5x5-grid rows)
    (7, 512, 512, 512, 512, 4, 4, 512, 1, 0),       # ... concat + shortcut, batch not a multiple of 5
    (1, 64, 0, 0, 0, 4, 4, 64, 0, 0),
    (5, 128, 64, 0, 0, 4, 4, 128, 1, 0),
    (4, 128, 0, 0, 0, 4, 4, 128, 0, 1),             # folded upsample 4 -> 8
    (11, 64, 0, 64, 0, 4, 4, 64, 1, 0),             # three tiles, the last one holds a single image
    # stride-2 Downsample (unet.py:68-74): `up` = 2, H x W is the INPUT size; four input-parity views, 4 + 2 + 2 + 1 taps
    (2, 64, 0, 0, 0, 128, 128, 64, 0, 2),           # downs.3 of the headline: 128 -> 64
    (3, 128, 0, 0, 0, 32, 32, 128, 0, 2),           # 32 -> 16: one 8x16 tile row pair per image
    (1, 256, 0, 0, 0, 64, 32, 256, 0, 2),           # non-square
    (5, 512, 0, 0, 0, 16, 16, 512, 0, 2),           # 16 -> 8: two output images per tile, odd batch
    (7, 512, 0, 0, 0, 8, 8, 512, 0, 2),             # 8 -> 4: five output images per tile
    (1, 64, 0, 0, 0, 32, 32, 128, 0, 2),            # Cin != Cout
]
SHAPES = [(0, 0), (64, 1), (64, 2), (128, 1), (128, 2), (256, 1)]


def _block(x0, x1, gamma, beta, w, b, r0, r1, wres, up, want_stats=True, iters=0):
    from b200sr3 import _lib
    lib = _lib.load()
    B, C0, H, W = x0.shape
    Cout = w.shape[0]
    y = torch.empty((B, Cout, H // 2, W // 2) if up == 2 else (B, Cout, H * (1 + up), W * (1 + up)), device="cuda")
    st = torch.empty(B, Cout, 2, device="cuda") if want_stats else None
    P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p()
    ch = lambda t: t.shape[1] if t is not None else 0
    ms = C.c_float(0)
    _lib.check(lib.b200sr3_conv_block(0, P(x0), C0, P(x1), ch(x1), P(gamma), P(beta), 32, 1, P(w), P(b), P(r0), ch(r0),
                                      P(r1), ch(r1), P(wres), B, H, W, Cout, int(up), P(y), P(st), iters,
                                      C.byref(ms), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return y, st, ms.value


def _reference(x0, x1, gamma, beta, w, b, r0, r1, wres, up):
    r = lambda t: t.bfloat16().float()
    xin = r(x0) if x1 is None else torch.cat([r(x0), r(x1)], 1)
    if gamma is not None:
        xin = F.group_norm(xin, 32, gamma, beta, eps=1e-5)
        xin = r(xin * torch.sigmoid(xin))
    if up == 1:
        xin = F.interpolate(xin, scale_factor=2, mode="nearest")
    y = F.conv2d(xin.double(), r(w).double(), b.double(), padding=1, stride=2 if up == 2 else 1)
    if r0 is not None:
        rin = r(r0) if r1 is None else torch.cat([r(r0), r(r1)], 1)
        y = y + F.conv2d(rin.double(), r(wres).double())
    return y.float()


def _make(case, seed=0):
    B, C0, C1, Cr0, Cr1, H, W, Cout, gn, up = case
    g = torch.Generator().manual_seed(seed + sum(case))
    rn = lambda *s: torch.randn(*s, generator=g)
    x0 = rn(B, C0, H, W) * 1.5 + 0.3
    x1 = rn(B, C1, H, W) * 0.7 - 0.2 if C1 else None
    gamma = 1 + 0.3 * rn(C0 + C1) if gn else None
    beta = 0.2 * rn(C0 + C1) if gn else None
    w = rn(Cout, C0 + C1, 3, 3) / (9 * (C0 + C1)) ** 0.5
    b = rn(Cout)
    r0 = rn(B, Cr0, H, W) if Cr0 else None
    r1 = rn(B, Cr1, H, W) if Cr1 else None
    wres = rn(Cout, Cr0 + Cr1, 1, 1) / (Cr0 + Cr1) ** 0.5 if Cr0 + Cr1 else None
    return x0, x1, gamma, beta, w, b, r0, r1, wres, up


@pytest.mark.parametrize("case", CASES, ids=lambda c: "x".join(map(str, c)))
def test_block_matches_torch(case):
    args = _make(case)
    ref = _reference(*args)
    cu = [a.cuda() if torch.is_tensor(a) else a for a in args]
    y, st, _ = _block(*cu)
    tol = float(ref.abs().max()) * 2 ** -8 * (2 if case[8] else 1) + 2e-3
    assert float((y.cpu() - ref).abs().max()) <= tol
    # fused GroupNorm statistics of the output: per-(image, channel) sum and sum of squares
    n = ref.shape[2] * ref.shape[3]
    s_ref = torch.stack([ref.sum((2, 3)), (ref * ref).sum((2, 3))], -1)
    rms = float(ref.pow(2).mean().sqrt())
    err = (st.cpu() - s_ref).abs()
    # (a folded upsample pre-sums the 3x3 weights in fp32 and rounds once, the reference rounds each
    # weight: a per-channel systematic difference that grows with n, not sqrt(n))
    assert float(err[..., 0].max()) <= 4e-3 * rms * n ** 0.5 * 4 + 1e-2 + (3e-3 * rms * n if case[9] == 1 else 0)
    # sum of squares of the bf16-ROUNDED output: each y carries <= 2^-9 relative rounding error, i.e. 2^-8 on y^2 (this
    # term does not average out over the 16 pixels of a 4x4 image), plus the accumulation noise bound used above
    assert bool((err[..., 1] <= 2 ** -7 * s_ref[..., 1] + 1e-2 * rms * rms * n).all())


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "bn%d_mt%d" % s)
def test_every_tile_shape(shape):
    """Each (BLOCK_N, MT) instantiation on one layer that all of them can run."""
    case = (3, 128, 128, 128, 128, 32, 32, 256, 1, 0)
    args = _make(case, seed=5)
    ref = _reference(*args)
    cu = [a.cuda() if torch.is_tensor(a) else a for a in args]
    old = {k: os.environ.get(k) for k in ("B200SR3_HALO_BN", "B200SR3_HALO_MT", "B200SR3_HALO_128X2")}
    try:
        if shape[0]:
            os.environ["B200SR3_HALO_BN"], os.environ["B200SR3_HALO_MT"] = str(shape[0]), str(shape[1])
            os.environ["B200SR3_HALO_128X2"] = "1"          # the (128, 2) shape is opt-in
        y, st, _ = _block(*cu)
    finally:
        for k, v in old.items():
            os.environ.pop(k, None)
            if v is not None:
                os.environ[k] = v
    tol = float(ref.abs().max()) * 2 ** -7 + 2e-3
    assert float((y.cpu() - ref).abs().max()) <= tol


def test_block_properties_at_baseline_size():
    """Size-independent properties at a BASELINE-sized layer (64 ch @128x128, B=8)."""
    g = torch.Generator().manual_seed(11)
    x = (torch.randn(8, 64, 128, 128, generator=g)).bfloat16().float().cuda()
    w = (torch.randn(64, 64, 3, 3, generator=g) / 24).cuda()
    b = torch.randn(64, generator=g).cuda()
    y1, s1, _ = _block(x, None, None, None, w, b, None, None, None, 0)
    y2, s2, _ = _block(x, None, None, None, w, b, None, None, None, 0)
    assert torch.equal(y1, y2) and torch.equal(s1, s2)                      # replay is bit-exact
    # batch invariance: an image's output and statistics do not depend on its neighbours
    y3, s3, _ = _block(x[5:6].contiguous(), None, None, None, w, b, None, None, None, 0)
    assert torch.equal(y3, y1[5:6]) and torch.equal(s3, s1[5:6])
    # without a GroupNorm the conv is linear: doubling the input doubles (y - bias) exactly
    z = torch.zeros_like(b)
    ya, _, _ = _block(x, None, None, None, w, z, None, None, None, 0)
    yb, _, _ = _block(x * 2, None, None, None, w, z, None, None, None, 0)
    assert torch.equal(yb, (ya * 2).bfloat16().float())
    # zero input -> exactly the bias (the zero padding stays zero), also through the fused GroupNorm:
    # GN(0) = beta, so with beta = 0 and Swish(0) = 0 the conv sees zeros again
    gamma, beta = torch.ones(64, device="cuda"), torch.zeros(64, device="cuda")
    y0, _, _ = _block(torch.zeros_like(x[:1]), None, gamma, beta, w, b, None, None, None, 0)
    assert torch.equal(y0, b.bfloat16().float().view(1, 64, 1, 1).expand_as(y0))


def test_block_rejects_bad_shapes():
    from b200sr3 import _lib
    x = torch.zeros(1, 64, 2, 2, device="cuda")         # below every tile geometry
    w = torch.zeros(64, 64, 3, 3, device="cuda")
    b = torch.zeros(64, device="cuda")
    with pytest.raises(_lib.B200Error):
        _block(x, None, None, None, w, b, None, None, None, 0)
