"""Power and SM clock per conv layer class: each class runs back to back for ~2 s (through b200sr3_conv_block's timing
loop) while nvidia-smi is sampled every 100 ms. Writes a markdown table to stdout.

    python tools/power_by_class.py [B] [class filter]

With the timing build (B200SR3_LIB=.../libb200sr3_timing.so) B200SR3_CONV_ABLATE=<mask> removes one role's work
(results then wrong): 1 no global stores, 2 no transform, 4 no weight loads, 8 no halo loads, 16 no TMEM loads,
32 no statistics math - under the power cap the time that disappears is that role's share of the energy.
"""
import ctypes as C
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-super-resolution-face-reconstruction_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from halo_bench import SHAPES

CLASSES = ["c1 64->64 @128", "c1 128+64->64 @128", "c2 64->64+res128 @128", "c1 256+128->128 @64", "c2 128->128+res384 @64",
           "c1 512+256->256 @32", "c1 512+512->512 @16", "c1 512->512 @8", "up 128->128 @64->128"]


def sample_smi(path, stop):
    q = "clocks.sm,power.draw,clocks_event_reasons.sw_power_cap"
    with open(path, "w") as f:
        p = subprocess.Popen(["nvidia-smi", "-i", "0", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"], stdout=f)
        stop.wait()
        p.terminate()


def main():
    import threading
    import torch
    from b200sr3 import _lib
    lib = _lib.load()
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p()
    print("| layer class (B=%d) | us per launch (2 s loop) | TFLOP/s (reference graph) | SM clock MHz (median) | power W (median) | sw_power_cap |" % B)
    print("|---|---|---|---|---|---|")
    for name, c0, c1, cr0, cr1, h, cout, up in SHAPES:
        if name not in CLASSES or (len(sys.argv) > 2 and sys.argv[2] not in name):
            continue
        rn = lambda *s: torch.randn(*s, device="cuda")
        x0 = rn(B, c0, h, h)
        x1 = rn(B, c1, h, h) if c1 else None
        r0 = rn(B, cr0, h, h) if cr0 else None
        r1 = rn(B, cr1, h, h) if cr1 else None
        use_gn = not up
        gamma = torch.ones(c0 + c1, device="cuda") if use_gn else None
        beta = torch.zeros(c0 + c1, device="cuda") if use_gn else None
        w = rn(cout, c0 + c1, 3, 3) / (9 * (c0 + c1)) ** 0.5
        wres = rn(cout, cr0 + cr1, 1, 1) if cr0 + cr1 else None
        b = rn(cout)
        ho = h * (2 if up else 1)
        y = torch.empty(B, cout, ho, ho, device="cuda")
        ms = C.c_float(0)
        args = lambda iters: (0, P(x0), c0, P(x1), c1, P(gamma), P(beta), 32, 1, P(w), P(b), P(r0), cr0, P(r1), cr1, P(wres), B, h,
                              h, cout, up, P(y), C.c_void_p(), iters, C.byref(ms), C.c_void_p(0))
        _lib.check(lib.b200sr3_conv_block(*args(20)))
        iters = max(50, int(2000.0 / max(ms.value, 1e-3)))          # ~2 s
        fd, path = tempfile.mkstemp(suffix=".csv")
        os.close(fd)
        stop = threading.Event()
        th = threading.Thread(target=sample_smi, args=(path, stop))
        th.start()
        time.sleep(0.3)
        _lib.check(lib.b200sr3_conv_block(*args(iters)))
        stop.set()
        th.join()
        clk, pw, cap = [], [], 0
        for line in open(path):
            f = [v.strip() for v in line.split(",")]
            if len(f) >= 3:
                try:
                    clk.append(float(f[0])); pw.append(float(f[1])); cap += f[2].lower().startswith("active")
                except ValueError:
                    pass
        os.unlink(path)
        # drop the samples before / after the loop: keep the upper half by power
        idx = sorted(range(len(pw)), key=lambda i: pw[i])[len(pw) // 2:]
        med = lambda v: sorted(v)[len(v) // 2] if v else float("nan")
        flops = 2.0 * B * ho * ho * cout * (9 * (c0 + c1) + cr0 + cr1)
        print(f"| {name} | {ms.value * 1e3:.1f} | {flops / ms.value / 1e9:.0f} | {med([clk[i] for i in idx]):.0f} | "
              f"{med([pw[i] for i in idx]):.0f} | {cap}/{len(pw)} samples |", flush=True)
        del x0, x1, r0, r1, w, y
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
