"""Kernel-level parity of the tcgen05 implicit-GEMM conv (through the C ABI) against
torch fp32 conv2d on the same bf16-rounded operands. Tolerance: the bf16 rounding of the
output, |y| * 2^-8, plus 1e-3 of fp32 accumulation-order noise."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CASES = [
    # B, Cin, H, W, Cout, k, stride, upsample, residual
    (1, 64, 16, 16, 64, 3, 1, 0, 0),
    (1, 64, 16, 16, 64, 1, 1, 0, 0),
    (2, 128, 32, 32, 128, 3, 1, 0, 1),
    (1, 256, 16, 16, 256, 3, 1, 0, 0),
    (4, 512, 8, 8, 512, 3, 1, 0, 0),
    (2, 64, 32, 32, 64, 3, 2, 0, 0),
    (2, 256, 16, 16, 256, 3, 2, 0, 0),
    (2, 128, 8, 8, 128, 3, 1, 1, 0),
    (1, 64, 128, 128, 64, 3, 1, 0, 0),
    (3, 512, 2, 2, 512, 3, 1, 0, 0),     # ragged batch inside one tile
    (5, 512, 1, 1, 512, 3, 1, 0, 0),     # 1x1 spatial: everything but the centre tap is padding
    (2, 192, 16, 16, 64, 3, 1, 0, 0),
    (1, 1024, 8, 8, 512, 3, 1, 0, 0),
    (2, 512, 8, 8, 1536, 1, 1, 0, 0),
    (2, 64, 64, 64, 128, 3, 1, 0, 1),
]


def _conv(x, w, b, r, k, stride, up):
    from b200sr3 import _lib
    lib = _lib.load()
    B, Cin, H, W = x.shape
    Cout = w.shape[0]
    Ho, Wo = H * (2 if up else 1) // stride, W * (2 if up else 1) // stride
    y = torch.empty(B, Cout, Ho, Wo, device="cuda")
    P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p()
    ms = C.c_float(0)
    _lib.check(lib.b200sr3_conv2d(0, P(x), P(w), P(b), P(r), B, Cin, H, W, Cout, k, stride, int(up), P(y), 0,
                                  C.byref(ms), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return y


@pytest.mark.parametrize("case", CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv_matches_torch(case):
    B, Cin, H, W, Cout, k, stride, up, res = case
    g = torch.Generator().manual_seed(hash(case) % (2 ** 31))
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    b = torch.randn(Cout, generator=g)
    Ho, Wo = H * (2 if up else 1) // stride, W * (2 if up else 1) // stride
    r = torch.randn(B, Cout, Ho, Wo, generator=g) if res else None
    xin = x.bfloat16().float()
    if up:
        xin = F.interpolate(xin, scale_factor=2, mode="nearest")
    ref = F.conv2d(xin, w.bfloat16().float(), b, stride=stride, padding=k // 2)
    if res:
        ref = ref + r.bfloat16().float()
    y = _conv(x.cuda(), w.cuda(), b.cuda(), r.cuda() if res else None, k, stride, up).cpu()
    tol = float(ref.abs().max()) * 2 ** -8 + 1e-3
    assert float((y - ref).abs().max()) <= tol


def test_conv_is_linear_and_deterministic():
    """Size-independent properties at a BASELINE-sized layer (64ch @128x128, B=8)."""
    g = torch.Generator().manual_seed(3)
    x1 = torch.randn(8, 64, 128, 128, generator=g).bfloat16().float().cuda()
    w = (torch.randn(64, 64, 3, 3, generator=g) / 24).cuda()
    y1 = _conv(x1, w, None, None, 3, 1, 0)
    assert torch.equal(y1, _conv(x1, w, None, None, 3, 1, 0))                  # replay is bit-exact
    y2 = _conv(x1 * 2, w, None, None, 3, 1, 0)                                  # exact in bf16/fp32
    assert torch.equal(y2, (y1 * 2).bfloat16().float())
    # batch invariance: each image's result does not depend on its neighbours in the tile
    assert torch.equal(_conv(x1[3:4].contiguous(), w, None, None, 3, 1, 0), y1[3:4])
    # zero input -> exactly the bias
    b = torch.randn(64, generator=g).cuda()
    y0 = _conv(torch.zeros_like(x1[:1]), w, b, None, 3, 1, 0)
    assert torch.equal(y0, b.bfloat16().float().view(1, 64, 1, 1).expand_as(y0))


def test_conv_rejects_bad_arguments():
    from b200sr3 import _lib
    x = torch.zeros(1, 64, 12, 12, device="cuda")       # not a power of two
    w = torch.zeros(64, 64, 3, 3, device="cuda")
    with pytest.raises(_lib.B200Error):
        _conv(x, w, None, None, 3, 1, 0)
