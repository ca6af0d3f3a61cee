"""Prints the measured parity errors of the CUDA path against the golden vectors of the reference
(the quantities tests/test_gpu_parity.py asserts on), so tolerances and headroom can be read off.

    python tools/parity_report.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "3d-super-resolution-face-reconstruction_b200"))
sys.path.insert(0, ROOT)

from conftest import build_net, synthetic_weights          # noqa: E402
from oracle import sr3_oracle as O                          # noqa: E402
from oracle.weights import make_inputs                      # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    g = np.load(os.path.join(GOLD, "unet_r64.npz"))
    net, _ = build_net(200, seed=int(g["weight_seed"]), gain=float(g["weight_gain"]))
    x6 = torch.from_numpy(g["x6"]).cuda()
    eps = net.unet_eps(x6[:, :3], x6[:, 3:], float(g["noise_level"][0, 0])).cpu()
    ref = torch.from_numpy(g["eps"])
    rms = float(ref.pow(2).mean().sqrt())
    print(f"unet R=64 golden: rms err {float((eps - ref).pow(2).mean().sqrt()) / rms:.4%} of rms(eps) (tol 1%), "
          f"max err {float((eps - ref).abs().max()) / rms:.3%} (tol 6%)")

    for R, B in [(16, 3), (32, 2), (128, 1)]:
        net, mopt = build_net(10)
        sd = synthetic_weights(0, 1.0)
        cond, noise = make_inputs(B, R, 1, seed=R)
        with torch.no_grad():
            ref = O.unet_forward(sd, mopt, torch.cat([cond, noise[0]], 1), torch.full((B, 1), 0.8), {})
        eps = net.unet_eps(cond.cuda(), noise[0].cuda(), 0.8).cpu()
        rms = float(ref.pow(2).mean().sqrt())
        print(f"unet R={R} vs oracle: rms err {float((eps - ref).pow(2).mean().sqrt()) / rms:.4%}, "
              f"max {float((eps - ref).abs().max()) / rms:.3%}")

    g = np.load(os.path.join(GOLD, "steps_r32_T400.npz"))
    net, _ = build_net(400)
    cond = torch.from_numpy(g["cond"]).cuda()
    worst = 0.0
    for i, t in enumerate(g["t"].tolist()):
        out = net.p_sample(torch.from_numpy(g["x_t"][i]).cuda(), t, condition_x=cond,
                           noise=torch.from_numpy(g["z_t"][i]).cuda()).cpu()
        worst = max(worst, float((out - torch.from_numpy(g["x_tm1"][i])).abs().max()))
    print(f"teacher-forced steps (T=400, R=32): worst max|dx| {worst:.2e} (tol 1e-3)")
    cond, noise = make_inputs(2, 32, 400, seed=321)
    out = net.super_resolution_batched(cond.cuda(), noise=noise.cuda()).cpu()
    ref = torch.from_numpy(g["final"])
    psnr = min(O.psnr_uint8(out[b], ref[b]) for b in range(2))
    print(f"free-running T=400 chain: PSNR {psnr:.2f} dB (tol >= 40), max|dx| {float((out - ref).abs().max()):.2e} (tol 0.05)")
    g = np.load(os.path.join(GOLD, "chain_r32_T10.npz"))
    net, _ = build_net(10)
    out = net.super_resolution_batched(torch.from_numpy(g["cond"]).cuda(), noise=torch.from_numpy(g["noise"]).cuda()).cpu()
    print(f"free-running T=10 chain: max|dx| {float((out - torch.from_numpy(g['xs'][-1])).abs().max()):.2e} (tol 5e-3)")


if __name__ == "__main__":
    main()
