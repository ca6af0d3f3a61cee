"""Times single tcgen05 conv launches (through b200sr3_conv2d) at the UNet's shapes; no CPU
reference, so it is cheap at BASELINE sizes. Env: B200SR3_CONV2D_STATS=1 fuses the GroupNorm
statistics, B200SR3_CONV_TIMING=1 prints per-role cycle counters, B200SR3_BLOCK_N forces a tile.

    python tools/conv_bench.py [B] [iters]
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-super-resolution-face-reconstruction_b200"))

# name, Cin, H, Cout, k, stride, up, res   (R = 128 level shapes)
SHAPES = [
    ("64->64 @128", 64, 128, 64, 3, 1, 0, 1),
    ("128->64 @128", 128, 128, 64, 3, 1, 0, 0),
    ("192->64 @128", 192, 128, 64, 3, 1, 0, 0),
    ("128->128 @64", 128, 64, 128, 3, 1, 0, 1),
    ("384->128 @64", 384, 64, 128, 3, 1, 0, 0),
    ("256->256 @32", 256, 32, 256, 3, 1, 0, 1),
    ("768->256 @32", 768, 32, 256, 3, 1, 0, 0),
    ("512->512 @16", 512, 16, 512, 3, 1, 0, 1),
    ("1024->512 @16", 1024, 16, 512, 3, 1, 0, 0),
    ("512->512 @8", 512, 8, 512, 3, 1, 0, 1),
    ("up 128->128 @64->128", 128, 64, 128, 3, 1, 1, 0),
    ("up 512->512 @16->32", 512, 16, 512, 3, 1, 1, 0),
    ("down 64->64 @128->64", 64, 128, 64, 3, 2, 0, 0),
]


def main():
    import torch
    from b200sr3 import _lib
    lib = _lib.load()
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    only = sys.argv[3] if len(sys.argv) > 3 else None
    P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p()
    for name, cin, h, cout, k, stride, up, res in SHAPES:
        if only and only not in name:
            continue
        x = torch.randn(B, cin, h, h, device="cuda")
        w = torch.randn(cout, cin, k, k, device="cuda") / (cin * k * k) ** 0.5
        b = torch.randn(cout, device="cuda")
        ho = h * (2 if up else 1) // stride
        r = torch.randn(B, cout, ho, ho, device="cuda") if res else None
        y = torch.empty(B, cout, ho, ho, device="cuda")
        ms = C.c_float(0)
        _lib.check(lib.b200sr3_conv2d(0, P(x), P(w), P(b), P(r), B, cin, h, h, cout, k, stride, up, P(y), iters,
                                      C.byref(ms), C.c_void_p(0)))
        flops = 2.0 * B * ho * ho * cout * cin * k * k
        print(f"{name:24s} B={B:3d} {ms.value * 1e3:8.1f} us  {flops / ms.value / 1e9:8.1f} TF/s (reference-graph FLOPs)",
              flush=True)
        del x, w, r, y
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
