"""Shapes off the benchmark path: odd batches, every power-of-two resolution from 16 to 256, tiny T.
One teacher-forced step each (finite, bounded, batch-invariant) plus a short Philox chain.

    python tools/robustness_check.py
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-super-resolution-face-reconstruction_b200"))
import torch
import b200sr3
from b200sr3 import synthetic


def main():
    opt = {"phase": "val", "sr": {"model": b200sr3.configs.model_opt(6)}}
    net = b200sr3.define_G(opt)
    net.load_state_dict(synthetic.state_dict(net, 0, 1.0), strict=True)
    net = net.cuda().eval()
    net.set_new_noise_schedule(opt["sr"]["model"]["beta_schedule"]["val"], [torch.device("cuda")])
    ok = True
    for R in (16, 32, 64, 128, 256):
        for B in (1, 3, 7, 37) if R <= 128 else (1, 3):
            cond, noise = synthetic.inputs(B, R, 2, seed=R + B)
            cond, x, z = cond.cuda(), noise[0].cuda(), noise[1].cuda()
            t0 = time.time()
            full = net.p_sample(x, 5, condition_x=cond, noise=z)
            one = net.p_sample(x[B - 1:].contiguous(), 5, condition_x=cond[B - 1:].contiguous(), noise=z[B - 1:].contiguous())
            chain = net.super_resolution_batched(cond, seed=3)
            good = (bool(torch.isfinite(full).all()) and torch.equal(one, full[B - 1:]) and bool(torch.isfinite(chain).all())
                    and float(chain.abs().max()) <= 1.0 + 1e-5)
            ok &= good
            print(f"R={R:3d} B={B:2d}: step finite+batch-invariant, T=6 chain in [-1,1]: {good}  ({time.time() - t0:.2f}s)", flush=True)
    print("ALL OK" if ok else "FAILURES")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
