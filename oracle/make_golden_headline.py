"""TEST INFRASTRUCTURE ONLY (oracle/): headline-config goldens from the REAL reference.

Run in the build container only (needs /root/reference):

    python -m oracle.make_golden_headline [E] [F]

Case E = BASELINE config 3 (sr_sr3_VGGF2_16_128_model3): R=128, T=600, B=1 — the full chain of
the unmodified reference (model/sr/sr3_modules/diffusion.py:182-215) with the noise list
injected, ~600 UNet evaluations on the CPU. Case F = config 2 (16->64, T=200) at B=2.

Stored per case (small: the noise list is NOT stored, it is regenerated from
oracle.weights.make_inputs(seed), a PCG64 stream that is identical on every machine):
  cond, the chain states x after k steps for the k in `keep_k` (teacher-forced pairs
  x_t -> x_{t-1} for t in `t_steps`; k = 0 is noise[0] and is not stored), the final image, the
  snapshots the reference returns with continous=True (case F only), and the sha256 of the
  weights and of the noise list.
The oracle restatement is re-pinned on the teacher-forced steps (max |oracle - reference|).
"""
import hashlib
import os
import sys
import time

import numpy as np
import torch

from . import sr3_oracle as O
from .make_golden import OUT, TOL_STEP, build_reference, inject_noise, model_opt
from .weights import make_inputs, make_state_dict, state_dict_digest

CASES = {
    # name: (file, R, T, B, input seed, teacher-forced t's, store the continous=True snapshots)
    "E": ("chain_r128_T600.npz", 128, 600, 1, 2024, [599, 300, 1, 0], False),
    "F": ("chain_r64_T200.npz", 64, 200, 2, 2025, [199, 100, 1, 0], True),
}


def run_case(name):
    fname, R, T, B, seed, t_steps, store_snaps = CASES[name]
    mopt = model_opt(T)
    sd = make_state_dict(mopt, seed=0, gain=1.0)
    cond, noise = make_inputs(B, R, T, seed=seed)
    net = build_reference(mopt, sd)
    tabs = O.schedule_tables(mopt["beta_schedule"]["val"])
    inter = 1 | (T // 10)
    keep_k = sorted({T - 1 - t for t in t_steps} | {T - t for t in t_steps})   # k = steps done
    kept = {}
    snaps = [cond]
    t0 = time.time()
    with inject_noise(noise):
        img = torch.randn(cond.shape)
        if 0 in keep_k:
            kept[0] = img.clone()
        for k, i in enumerate(reversed(range(T)), start=1):
            img = net.p_sample(img, i, condition_x=cond)      # the reference's own step
            if k in keep_k:
                kept[k] = img.clone()
            if i % inter == 0:
                snaps.append(img.clone())
            if k % 50 == 0:
                print(f"  case {name}: step {k}/{T}  {time.time() - t0:.0f}s", flush=True)
    final = img
    # re-pin the oracle on the teacher-forced steps against the reference's own x_{t-1}
    d = 0.0
    for t in t_steps:
        xin = kept[T - 1 - t]
        z = noise[T - t] if t > 0 else torch.zeros_like(xin)
        with torch.no_grad():
            out = O.p_sample(sd, mopt, tabs, xin, t, cond, z)
        d = max(d, float((out - kept[T - t]).abs().max()))
    assert d <= TOL_STEP, f"oracle != reference on case {name}: {d}"
    stored_k = [k for k in keep_k if k > 0]          # x after 0 steps is noise[0]: regenerated, not stored
    extra = {"snapshots": torch.cat(snaps, 0).numpy()} if store_snaps else {}
    np.savez_compressed(
        os.path.join(OUT, fname), cond=cond.numpy(), t_steps=np.array(t_steps), keep_k=np.array(stored_k),
        xs=torch.stack([kept[k] for k in stored_k]).numpy(), final=final.numpy(),
        input_seed=seed, T=T, R=R, B=B, **extra,
        noise_sha256=hashlib.sha256(noise.numpy().tobytes()).hexdigest(),
        weight_seed=0, weight_gain=1.0, weight_sha256=state_dict_digest(sd))
    line = (f"{name} r{R} B{B} T{T} full reference chain ({time.time() - t0:.0f}s CPU): teacher-forced "
            f"t={t_steps} max|oracle-reference| = {d:.2e}; final image range "
            f"[{float(final.min()):.3f}, {float(final.max()):.3f}]")
    print(line, flush=True)
    with open(os.path.join(OUT, "README.md"), "a") as f:
        f.write("* " + line + " (`python -m oracle.make_golden_headline`)\n")


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    for c in (sys.argv[1:] or ["F", "E"]):
        run_case(c)
