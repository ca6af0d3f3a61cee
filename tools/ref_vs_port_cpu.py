"""Build container only (needs /root/reference): times the UNMODIFIED reference module and the oracle port on the same
CPU cores, same weights, same inputs - the evidence behind `cpu_baseline.kind = "port"` in bench.py (the reference is pure
Python and cannot travel to the GPU box; the port is what runs there).

    python tools/ref_vs_port_cpu.py [steps] > profiles/<round>_ref_vs_port_cpu.txt
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import make_golden as G
from oracle import sr3_oracle as O
from oracle.weights import make_inputs, make_state_dict


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    B, R, T = 2, 128, 600
    torch.set_num_threads(os.cpu_count())
    mopt = G.model_opt(T)
    sd = make_state_dict(mopt, seed=0, gain=1.0)
    cond, noise = make_inputs(B, R, steps + 1, seed=123)
    net = G.build_reference(mopt, sd)
    tabs = O.schedule_tables(mopt["beta_schedule"]["val"])
    sd_o = {k: v for k, v in sd.items()}
    res = {}
    with torch.no_grad():
        for name in ("reference", "port", "reference", "port"):      # interleaved, second pair is the one reported
            x = noise[0].clone()
            t0 = None
            for i in range(steps + 1):
                if i == 1:
                    t0 = time.perf_counter()      # first step = warm-up
                t = T - 1 - i
                if name == "reference":
                    with G.inject_noise([noise[1]]):
                        x = net.p_sample(x, t, condition_x=cond)
                else:
                    x = O.p_sample(sd_o, mopt, tabs, x, t, cond, noise[1])
            res[name] = ((time.perf_counter() - t0) / steps, x)
    dt_r, x_r = res["reference"]
    dt_p, x_p = res["port"]
    print(f"torch {torch.__version__}, {os.cpu_count()} cores, B={B}, R={R}, {steps} sampling steps after one warm-up step")
    print(f"reference module (model/sr, unmodified): {dt_r * 1e3:8.1f} ms per step")
    print(f"oracle port (oracle/sr3_oracle.py):      {dt_p * 1e3:8.1f} ms per step   ratio port / reference {dt_p / dt_r:.3f}")
    print(f"max |x_port - x_reference| after {steps + 1} steps: {float((x_r - x_p).abs().max()):.2e}")


if __name__ == "__main__":
    main()
