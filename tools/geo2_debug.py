"""Debug aid: run one small halo-conv case under each B200SR3_CONV_ABLATE mask in a fresh process."""
import os
import subprocess
import sys

CODE = r'''
import sys, os, ctypes as C
sys.path.insert(0, "3d-super-resolution-face-reconstruction_b200")
import torch
from b200sr3 import _lib
lib = _lib.load()
B, Cin, H, Cout, gn, up = [int(v) for v in sys.argv[1:7]]
x = torch.randn(B, Cin, H, H, device="cuda")
w = torch.randn(Cout, Cin, 3, 3, device="cuda") / (9 * Cin) ** 0.5
b = torch.randn(Cout, device="cuda")
gamma = torch.ones(Cin, device="cuda") if gn else None
beta = torch.zeros(Cin, device="cuda") if gn else None
s = 2 if up else 1
y = torch.empty(B, Cout, H * s, H * s, device="cuda")
st = torch.empty(B, Cout, 2, device="cuda")
P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p()
ms = C.c_float(0)
rc = lib.b200sr3_conv_block(0, P(x), Cin, None, 0, P(gamma), P(beta), 32, 1, P(w), P(b), None, 0, None, 0, None, B, H, H, Cout, up,
                            P(y), P(st), 0, C.byref(ms), C.c_void_p(0))
print("rc", rc, lib.b200sr3_last_error().decode()[:200] if rc else "ok", float(y.abs().max()) if rc == 0 else "")
'''
case = sys.argv[1:7] if len(sys.argv) >= 7 else ["1", "64", "4", "64", "0", "0"]
for mask in (0, 1, 8, 4, 2, 16, 32, 1 | 8, 1 | 32, 1 | 8 | 4 | 16 | 32):
    env = dict(os.environ, B200SR3_CONV_ABLATE=str(mask))
    r = subprocess.run([sys.executable, "-c", CODE] + case, env=env, capture_output=True, text=True, timeout=120)
    print("ablate", mask, "->", (r.stdout.strip().splitlines() or ["<no output>"])[-1], r.stderr.strip().splitlines()[-1:] if r.returncode else "")
