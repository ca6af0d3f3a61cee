"""Faces/s of every BASELINE.json config on one GPU (one warm-up chain + one timed chain each, Philox noise,
synthetic weights and inputs), followed by the SR -> MICA hand-off for config 5.

    python tools/config_sweep.py
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-super-resolution-face-reconstruction_b200"))
import torch
import b200sr3
from b200sr3 import synthetic, mica_handoff

GF = {32: 5.5629, 64: 22.2483, 128: 88.9896}
# name, faces per GPU used here
CONFIGS = [("sr_sr3_VGGF2_8_32_model2", 4), ("sr_sr3_VGGF2_16_64_model3", 64), ("sr_sr3_VGGF2_16_128_model3", 32),
           ("sr_sr3_VGGF2_8_128_model3", 32), ("sr_sr3_VGGF2_32_128_model2", 64)]


def main():
    for name, B in CONFIGS:
        opt = b200sr3.configs.named(name)
        mopt = opt["sr"]["model"]
        R, T = opt["r_resolution"], mopt["beta_schedule"]["val"]["n_timestep"]
        net = b200sr3.define_G(opt)
        net.load_state_dict(synthetic.state_dict(net, seed=0, gain=1.0), strict=True)
        net = net.to("cuda").eval()
        net.set_new_noise_schedule(mopt["beta_schedule"]["val"], [torch.device("cuda")])
        cond = synthetic.inputs(B, R, seed=123).cuda()
        net.super_resolution_batched(cond, seed=1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = net.super_resolution_batched(cond, seed=2)
        e1.record()
        torch.cuda.synchronize()
        s = e0.elapsed_time(e1) / 1e3
        line = {"config": name, "R": R, "T": T, "faces": B, "s_per_chain": s, "faces_per_s": B / s,
                "ms_per_sampling_step": s / T * 1e3, "tflops_reference_graph": B * T * GF[R] / s / 1e3,
                "finite": bool(torch.isfinite(out).all())}
        if name.endswith("32_128_model2"):
            mica_handoff.sr_to_mica(out)      # warm-up (first call loads the kernels)
            torch.cuda.synchronize()
            e0.record()
            h = mica_handoff.sr_to_mica(out)
            e1.record()
            torch.cuda.synchronize()
            line["mica_handoff_ms"] = e0.elapsed_time(e1)
            line["arcface_blob_shape"] = list(h["arcface"].shape)
        print(json.dumps(line), flush=True)
        del net
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
