"""Times the SR -> MICA hand-off kernels (config 5: 512 faces of 128x128) against the HBM roofline, with the
reference's per-image host path (tensor2img + cv2.resize + cv2.dnn.blobFromImages, incl. the D2H/H2D copies it
implies) timed beside it on a bounded sample.

    python tools/handoff_bench.py [B] [R]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-super-resolution-face-reconstruction_b200"))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from b200sr3 import mica_handoff as H


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    R = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    x = torch.rand(B, 3, R, R, device="cuda") * 2.2 - 1.1
    for _ in range(3):
        out = H.sr_to_mica(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 20
    e0.record()
    for _ in range(iters):
        out = H.sr_to_mica(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    # algorithmic bytes: fp32 SR in, u8 image out+in, u8 224 out, fp32 224 image out, fp32 blob out
    by = B * (3 * R * R * 4 + 2 * 3 * R * R + 224 * 224 * 3 + 224 * 224 * 3 * 4 + 112 * 112 * 3 * 4)
    peak = 6558.4
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    line = {"kernel": "tensor2img_kernel + mica_handoff_kernel", "B": B, "R": R, "ms": ms, "faces_per_s": B / ms * 1e3,
            "algorithmic_bytes": by, "achieved_GBs": by / ms / 1e6, "peak_GBs": peak, "frac": by / ms / 1e6 / peak}
    try:
        import cv2
        from oracle import mica_handoff_oracle as M
        n = min(B, 64)
        t0 = time.perf_counter()
        for b in range(n):
            sr = M.tensor2img(x[b:b + 1].cpu().numpy())[0]
            up = cv2.resize(sr, (224, 224))
            arc = cv2.dnn.blobFromImages([up], 1.0 / 127.5, (112, 112), (127.5, 127.5, 127.5), swapRB=True)[0]
            a = torch.tensor(arc).cuda()[None]
            im = torch.tensor((up / 255.).transpose(2, 0, 1)).cuda()[None]
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        line["cpu_reference_path"] = {"faces_per_s": n / dt, "sample": f"{n} faces, 1 host thread, per-image D2H + cv2 + H2D as the reference does"}
    except ImportError:
        pass
    print(json.dumps(line))


if __name__ == "__main__":
    main()
