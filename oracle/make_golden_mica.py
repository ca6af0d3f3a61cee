"""TEST INFRASTRUCTURE ONLY (oracle/): generate tests/golden/mica_handoff.npz from the REAL reference code and OpenCV.

Run in the build container only (needs /root/reference):

    python -m oracle.make_golden_mica

For every case it runs the reference's own core/metrics.tensor2img / tensor2tensor_img (imported from
/root/reference) and the OpenCV calls the reference makes (model/sr3d/model.py:127-131,372-376), asserts that
oracle.mica_handoff_oracle reproduces them bit for bit (uint8 / float32 blob) or to 1e-5 (the float model3 blob), and
stores inputs and the reference's outputs.  model/sr3d/model.py itself cannot be imported (pytorch3d, FLAME assets
absent - SURVEY.md 8c), so its two three-line methods are called here as the same cv2 / torch expressions they contain.
"""
import os
import sys

import numpy as np

from . import mica_handoff_oracle as M

REF = os.environ.get("B200SR3_REF", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CASES = [("r32", 2, 32, 1), ("r128", 1, 128, 2)]


def make_input(B, R, seed):
    """Images slightly outside [-1, 1] (the clamp matters), with exact .5 rounding ties planted."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(-1.15, 1.15, (B, 3, R, R)).astype(np.float32)
    ties = (np.arange(0, 255, 2, dtype=np.float32) + 0.5) / 255.0 * 2.0 - 1.0      # (t+1)/2*255 = k + 0.5
    x.reshape(-1)[: ties.size] = ties
    return x


def reference_outputs(x):
    import cv2
    import torch
    import torch.nn.functional as F
    sys.path.insert(0, REF)
    import core.metrics as Metrics
    sr, up, image, blob, tblob = [], [], [], [], []
    for b in range(x.shape[0]):
        t = torch.from_numpy(x[b:b + 1].copy())
        sr_img = Metrics.tensor2img(t)                                            # core/metrics.py:16-42
        sr_up = cv2.resize(sr_img, (224, 224))                                    # model/sr3d/model.py:374
        arc = cv2.dnn.blobFromImages([sr_up], 1.0 / 127.5, (112, 112), (127.5, 127.5, 127.5), swapRB=True)[0]
        img = (sr_up / 255.).transpose(2, 0, 1)                                   # :380-381 (float64)
        tt = Metrics.tensor2tensor_img(torch.from_numpy(x[b].copy())) * 255.0     # :477
        tb = (tt - 127.5) / 127.5                                                 # create_tensor_blob :105-124
        tb = F.interpolate(tb.unsqueeze(0), size=(112, 112), mode="bilinear", align_corners=False).squeeze(0)
        tb = tb[[2, 1, 0], :, :]
        sr.append(sr_img); up.append(sr_up); image.append(img.astype(np.float32)); blob.append(arc); tblob.append(tb.numpy())
    return (np.stack(sr), np.stack(up), np.stack(image), np.stack(blob), np.stack(tblob))


def main():
    out = {}
    for name, B, R, seed in CASES:
        x = make_input(B, R, seed)
        sr, up, image, blob, tblob = reference_outputs(x)
        o_sr, o_up, o_image, o_blob = M.handoff_u8(x)
        o_tblob = M.tensor_blob_f32(x)
        assert np.array_equal(o_sr, sr), name + ": tensor2img"
        assert np.array_equal(o_up, up), name + ": cv2.resize"
        assert np.array_equal(o_blob, blob), name + ": blobFromImages"
        assert np.array_equal(o_image, image), name + ": image"
        err = float(np.abs(o_tblob - tblob).max())
        assert err <= 1e-5, (name, err)   # float interpolation weights differ by an ulp of the source index (FMA contraction)
        print(f"{name}: oracle == reference/OpenCV (uint8 and blob bit-exact; model3 float blob max|d| {err:.1e})")
        out[name + "_x"] = x
        out[name + "_sr"] = sr
        out[name + "_up"] = up
        out[name + "_blob"] = blob
        out[name + "_tblob"] = tblob
    # more sizes, checked here but not stored
    import cv2
    rng = np.random.default_rng(7)
    for R in (8, 16, 64, 100, 224):
        for _ in range(5):
            im = rng.integers(0, 256, (R, R, 3), dtype=np.uint8)
            assert np.array_equal(cv2.resize(im, (224, 224)), M.resize_linear_u8(im)), R
    np.savez_compressed(os.path.join(OUT, "mica_handoff.npz"), **out)
    print("wrote", os.path.join(OUT, "mica_handoff.npz"))


if __name__ == "__main__":
    main()
