"""Per-launch breakdown of one sampling step (CUDA events between launches, see
b200sr3_profile_step). Writes a table to stdout and JSON to gpurun_out/step_profile.json.

    python tools/profile_step.py [B] [R] [T]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-super-resolution-face-reconstruction_b200"))


def main():
    import torch
    import b200sr3
    from b200sr3 import synthetic
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    R = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    T = int(sys.argv[3]) if len(sys.argv) > 3 else 600
    opt = {"phase": "val", "sr": {"model": b200sr3.configs.model_opt(T)}}
    net = b200sr3.define_G(opt)
    net.load_state_dict(synthetic.state_dict(net, 0, 1.0), strict=True)
    net = net.cuda().eval()
    net.set_new_noise_schedule(opt["sr"]["model"]["beta_schedule"]["val"], [torch.device("cuda")])
    net.profile_step(B, R)
    prof = net.profile_step(B, R)
    total = sum(p[1] for p in prof)
    rows = []
    for name, ms, fl, by, fx in prof:
        rate = f"{fl / ms / 1e9:8.1f} TF/s" if fl > 0 else (f"{by / ms / 1e6:8.1f} GB/s" if by > 0 else " " * 13)
        rows.append({"op": name, "ms": ms, "flops": fl, "bytes": by, "flops_executed": fx})
        print(f"{name:28s} {ms * 1e3:9.1f} us {100 * ms / total:5.1f}%  {rate}")
    kinds = {}
    for name, ms, fl, by, fx in prof:
        k = "conv" if fl > 0 else name.split(".")[-1]
        kinds[k] = kinds.get(k, 0.0) + ms
    print("---- total %.3f ms" % total)
    for k, v in sorted(kinds.items(), key=lambda kv: -kv[1]):
        print(f"{k:12s} {v:8.3f} ms {100 * v / total:5.1f}%")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "step_profile.json"), "w") as f:
        json.dump({"B": B, "R": R, "T": T, "total_ms": total, "by_kind_ms": kinds, "ops": rows}, f, indent=1)


if __name__ == "__main__":
    main()
