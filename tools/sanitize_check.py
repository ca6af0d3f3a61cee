"""A few small invocations that together launch every kernel variant of the library (one 8x16 tile / two 8x8 images /
five 4x4 images per tile, head, tail, stride 2, deep ring, affine+PReLU, attention, hand-off, encoder), written for
`compute-sanitizer --tool memcheck python tools/sanitize_check.py`. compute-sanitizer is closed on this GPU pool (the
call is refused), so the script only serves as a quick all-variants smoke run:

    python tools/sanitize_check.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-super-resolution-face-reconstruction_b200"))
import torch
import b200sr3
from b200sr3 import synthetic


def main():
    opt = {"phase": "val", "sr": {"model": b200sr3.configs.model_opt(4)}}
    net = b200sr3.define_G(opt)
    net.load_state_dict(synthetic.state_dict(net, 0, 1.0), strict=True)
    net = net.cuda().eval()
    net.set_new_noise_schedule(opt["sr"]["model"]["beta_schedule"]["val"], [torch.device("cuda")])
    for R, B in ((32, 3), (64, 5), (128, 2)):
        cond, noise = synthetic.inputs(B, R, 2, seed=R + B)
        cond, x, z = cond.cuda(), noise[0].cuda(), noise[1].cuda()
        y = net.p_sample(x, 2, condition_x=cond, noise=z)
        out = net.super_resolution_batched(cond, seed=9)
        torch.cuda.synchronize()
        assert bool(torch.isfinite(y).all()) and bool(torch.isfinite(out).all())
        print(f"sampler R={R} B={B}: one teacher-forced step and a T=4 Philox chain ran", flush=True)
        if R == 128:
            h = b200sr3.mica_handoff.sr_to_mica(out)
            t = b200sr3.mica_handoff.create_tensor_blob(out)
            torch.cuda.synchronize()
            print("hand-off ran", tuple(h["arcface"].shape), tuple(h["image"].shape), tuple(t.shape), flush=True)
    enc = b200sr3.MicaEncoder()
    enc.arcface.load_state_dict(synthetic.mica_state_dict(enc.arcface, 0), strict=True)
    enc.regressor.load_state_dict(synthetic.mica_state_dict(enc.regressor, 1), strict=True)
    identity, code = enc.cuda()(torch.rand(2, 3, 112, 112, device="cuda") * 2 - 1)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(code).all()) and bool(torch.isfinite(identity).all())
    print("encoder B=2 ran", tuple(identity.shape), tuple(code.shape), flush=True)
    print("DONE")


if __name__ == "__main__":
    main()
