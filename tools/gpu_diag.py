"""GPU bring-up diagnostics: every case runs in its own subprocess under a timeout, so one trap
or hang cannot poison the CUDA context of the others.

    python tools/gpu_diag.py all            # everything, summary on stdout + gpurun_out/diag.json
    python tools/gpu_diag.py conv B Cin H W Cout k stride up res
    python tools/gpu_diag.py unet R B
"""
import ctypes as C
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-super-resolution-face-reconstruction_b200"))
sys.path.insert(0, ROOT)


def conv_case(B, Cin, H, W, Cout, k, stride, up, res, iters=0):
    import torch
    import torch.nn.functional as F
    from b200sr3 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(B * 1000 + Cin + Cout + H)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    b = torch.randn(Cout, generator=g)
    Ho = (H * (2 if up else 1)) // stride
    Wo = (W * (2 if up else 1)) // stride
    r = torch.randn(B, Cout, Ho, Wo, generator=g) if res else None
    xb, wb = x.bfloat16().float(), w.bfloat16().float()
    xin = F.interpolate(xb, scale_factor=2, mode="nearest") if up else xb
    ref = F.conv2d(xin.double(), wb.double(), b.double(), stride=stride, padding=k // 2)
    if res:
        ref = ref + r.bfloat16().double()
    xd, wd, bd = x.cuda(), w.cuda(), b.cuda()
    rd = r.cuda() if res else None
    y = torch.empty(B, Cout, Ho, Wo, device="cuda")
    ms = C.c_float(0)
    P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p()
    rc = lib.b200sr3_conv2d(0, P(xd), P(wd), P(bd), P(rd), B, Cin, H, W, Cout, k, stride, int(up), P(y), iters,
                            C.byref(ms), C.c_void_p(0))
    if rc != 0:
        return {"ok": False, "error": lib.b200sr3_last_error().decode()}
    torch.cuda.synchronize()
    got = y.cpu().double()
    err = (got - ref).abs().max().item()
    # bf16 output rounding bounds the error: |ref| * 2^-9 plus accumulation noise
    tol = float(ref.abs().max()) * 2 ** -8 + 1e-3
    flops = 2.0 * B * Ho * Wo * Cout * Cin * k * k
    out = {"ok": err <= tol, "max_err": err, "tol": tol, "ref_absmax": float(ref.abs().max())}
    if iters:
        out["ms"] = ms.value
        out["tflops"] = flops / (ms.value * 1e-3) / 1e12 if ms.value > 0 else None
    return out


def unet_case(R, B):
    import numpy as np
    import torch
    import b200sr3
    from oracle import sr3_oracle as O
    from oracle.weights import make_inputs, make_state_dict
    opt = b200sr3.configs.named("sr_sr3_VGGF2_8_32_model2")
    mopt = opt["sr"]["model"]
    sd = make_state_dict(mopt, seed=0, gain=1.0)
    net = b200sr3.define_G(opt)
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    cond, noise = make_inputs(B, R, 2, seed=9)
    x = noise[0]
    nl = 0.6
    taps = {}
    with torch.no_grad():
        eps_ref = O.unet_forward(sd, mopt, torch.cat([cond, x], 1), torch.full((B, 1), nl), taps)
    t0 = time.time()
    eps = net.unet_eps(cond.cuda(), x.cuda(), nl).cpu()
    res = {"eps_max_err": float((eps - eps_ref).abs().max()), "eps_ref_std": float(eps_ref.std()),
           "eps_has_nan": bool(torch.isnan(eps).any()), "first_call_s": time.time() - t0, "layers": {}}
    eng = net._engine()
    for name, ref in taps.items():
        if name == "final_conv":
            continue
        buf = torch.empty(ref.shape, device="cuda")
        c, h, w = C.c_int(), C.c_int(), C.c_int()
        rc = eng.lib.b200sr3_layer_output(eng.handle, name.encode(), C.c_void_p(buf.data_ptr()), C.byref(c), C.byref(h),
                                          C.byref(w), C.c_void_p(0))
        if rc != 0:
            res["layers"][name] = "error: " + eng.lib.b200sr3_last_error().decode()
            continue
        d = (buf.cpu() - ref).abs().max().item()
        res["layers"][name] = [round(d, 5), round(float(ref.abs().max()), 3)]
    res["ok"] = res["eps_max_err"] < 0.05 * max(1.0, res["eps_ref_std"]) and not res["eps_has_nan"]
    return res


CONV_CASES = [
    # B Cin  H   W  Cout k s up res
    (1, 64, 16, 16, 64, 3, 1, 0, 0),     # smallest: one M tile per 8 rows, K = 9 blocks
    (1, 64, 16, 16, 64, 1, 1, 0, 0),     # 1x1
    (2, 128, 32, 32, 128, 3, 1, 0, 1),   # BLOCK_N 128 path + residual
    (1, 256, 16, 16, 256, 3, 1, 0, 0),   # BLOCK_N by heuristic
    (4, 512, 8, 8, 512, 3, 1, 0, 0),     # two images per tile
    (2, 64, 32, 32, 64, 3, 2, 0, 0),     # stride 2 (parity maps)
    (2, 128, 8, 8, 128, 3, 1, 1, 0),     # upsample + conv
    (1, 64, 128, 128, 64, 3, 1, 0, 0),   # one image row per tile
    (3, 512, 2, 2, 512, 3, 1, 0, 0),     # tiny spatial, ragged batch
    (2, 192, 16, 16, 64, 3, 1, 0, 0),    # Cin multiple of 64, not a power of two
    (2, 512, 8, 8, 1536, 1, 1, 0, 0),    # qkv projection
]


def run_sub(args, timeout):
    t0 = time.time()
    try:
        p = subprocess.run([sys.executable, os.path.abspath(__file__)] + [str(a) for a in args], capture_output=True,
                           text=True, timeout=timeout)
        line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
        if line:
            r = json.loads(line[-1][7:])
        else:
            r = {"ok": False, "error": "no result", "rc": p.returncode, "stderr": p.stderr[-1500:], "stdout": p.stdout[-500:]}
    except subprocess.TimeoutExpired:
        r = {"ok": False, "error": f"timeout after {timeout}s"}
    r["wall_s"] = round(time.time() - t0, 1)
    return r


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "all"
    if mode == "conv":
        a = [int(v) for v in sys.argv[2:]]
        print("RESULT " + json.dumps(conv_case(*a)))
    elif mode == "unet":
        print("RESULT " + json.dumps(unet_case(int(sys.argv[2]), int(sys.argv[3]))))
    else:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        results = {"conv": [], "unet": []}
        for c in CONV_CASES:
            r = run_sub(["conv"] + list(c), 120)
            results["conv"].append({"case": c, **r})
            print("conv", c, json.dumps(r), flush=True)
        for bn in (64, 128, 256):          # every tile shape on one mid-size layer, with timing
            os.environ["B200SR3_BLOCK_N"] = str(bn)
            c = (8, 256, 32, 32, 256, 3, 1, 0, 0, 20)
            r = run_sub(["conv"] + list(c), 120)
            results["conv"].append({"case": c, "block_n": bn, **r})
            print("conv", c, "BLOCK_N", bn, json.dumps(r), flush=True)
        os.environ.pop("B200SR3_BLOCK_N", None)
        for R, B in ((32, 2), (64, 1)):
            r = run_sub(["unet", R, B], 300)
            results["unet"].append({"R": R, "B": B, **r})
            print("unet", R, B, json.dumps(r), flush=True)
        with open(os.path.join(ROOT, "gpurun_out", "diag.json"), "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
