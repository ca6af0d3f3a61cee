"""Times single halo-resident conv launches (through b200sr3_conv_block) at the ResnetBlock shapes
of the R=128 UNet, with the fused GroupNorm+Swish and the folded shortcut as the engine runs them.
Env: B200SR3_HALO_BN / B200SR3_HALO_MT force a tile shape.

    python tools/halo_bench.py [B] [iters] [name filter] [gn 0/1]
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-super-resolution-face-reconstruction_b200"))

# name, C0, C1 (concat), Cr0, Cr1 (1x1 shortcut sources), H, Cout, up
SHAPES = [
    ("c1 64->64 @128", 64, 0, 0, 0, 128, 64, 0),
    ("c2 64->64+id @128", 64, 0, 64, 0, 128, 64, 0),
    ("c1 64+64->64 @128", 64, 64, 0, 0, 128, 64, 0),
    ("c1 128+64->64 @128", 128, 64, 0, 0, 128, 64, 0),
    ("c2 64->64+res192 @128", 64, 0, 128, 64, 128, 64, 0),
    ("c2 64->64+res128 @128", 64, 0, 64, 64, 128, 64, 0),
    ("c2 128->128+res384 @64", 128, 0, 256, 128, 64, 128, 0),
    ("c2 128->128+res192 @64", 128, 0, 128, 64, 64, 128, 0),
    ("c2 256->256+res768 @32", 256, 0, 512, 256, 32, 256, 0),
    ("c1 128->128 @64", 128, 0, 0, 0, 64, 128, 0),
    ("c2 128->128+id @64", 128, 0, 128, 0, 64, 128, 0),
    ("c1 256+128->128 @64", 256, 128, 0, 0, 64, 128, 0),
    ("c1 256->256 @32", 256, 0, 0, 0, 32, 256, 0),
    ("c2 256->256+id @32", 256, 0, 256, 0, 32, 256, 0),
    ("c1 512+256->256 @32", 512, 256, 0, 0, 32, 256, 0),
    ("c1 512->512 @16", 512, 0, 0, 0, 16, 512, 0),
    ("c1 512+512->512 @16", 512, 512, 0, 0, 16, 512, 0),
    ("c1 512->512 @8", 512, 0, 0, 0, 8, 512, 0),
    ("c2 512->512+id @8", 512, 0, 512, 0, 8, 512, 0),
    ("c1 512+512->512 @8", 512, 512, 0, 0, 8, 512, 0),
    ("c2 512->512+res1024 @8", 512, 0, 512, 512, 8, 512, 0),
    ("up 512->512 @8->16", 512, 0, 0, 0, 8, 512, 1),
    ("up 128->128 @64->128", 128, 0, 0, 0, 64, 128, 1),
    ("up 256->256 @32->64", 256, 0, 0, 0, 32, 256, 1),
    ("up 512->512 @16->32", 512, 0, 0, 0, 16, 512, 1),
]


def main():
    import torch
    from b200sr3 import _lib
    lib = _lib.load()
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    only = sys.argv[3] if len(sys.argv) > 3 and sys.argv[3] != "-" else None
    gn = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p()
    for name, c0, c1, cr0, cr1, h, cout, up in SHAPES:
        if only and only not in name:
            continue
        rn = lambda *s: torch.randn(*s, device="cuda")
        x0 = rn(B, c0, h, h)
        x1 = rn(B, c1, h, h) if c1 else None
        r0 = rn(B, cr0, h, h) if cr0 else None
        r1 = rn(B, cr1, h, h) if cr1 else None
        use_gn = gn and not up
        gamma = torch.ones(c0 + c1, device="cuda") if use_gn else None
        beta = torch.zeros(c0 + c1, device="cuda") if use_gn else None
        w = rn(cout, c0 + c1, 3, 3) / (9 * (c0 + c1)) ** 0.5
        wres = rn(cout, cr0 + cr1, 1, 1) if cr0 + cr1 else None
        b = rn(cout)
        ho = h * (2 if up else 1)
        y = torch.empty(B, cout, ho, ho, device="cuda")
        st = torch.empty(B, cout, 2, device="cuda")
        ms = C.c_float(0)
        _lib.check(lib.b200sr3_conv_block(0, P(x0), c0, P(x1), c1, P(gamma), P(beta), 32, 1, P(w), P(b), P(r0), cr0,
                                          P(r1), cr1, P(wres), B, h, h, cout, up, P(y), P(st), iters, C.byref(ms),
                                          C.c_void_p(0)))
        flops = 2.0 * B * ho * ho * cout * (9 * (c0 + c1) + cr0 + cr1)
        print(f"{name:26s} B={B:3d} gn={int(bool(use_gn))} {ms.value * 1e3:8.1f} us  {flops / ms.value / 1e9:8.1f} TF/s "
              f"(reference-graph FLOPs)", flush=True)
        del x0, x1, r0, r1, w, y
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
