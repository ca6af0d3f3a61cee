"""Layer-by-layer parity report of the MICA identity encoder against the CPU oracle (development aid).

    python tools/arcface_check.py [B]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-super-resolution-face-reconstruction_b200"))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
import b200sr3
from oracle import arcface_oracle as A


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    arc_sd, map_sd = A.make_arcface_state_dict(0), A.make_mapping_state_dict(0)
    enc = b200sr3.MicaEncoder()
    enc.arcface.load_state_dict(arc_sd, strict=True)
    enc.regressor.load_state_dict(map_sd, strict=True)
    enc = enc.cuda()
    blob = A.make_blob(B, seed=0)
    taps = {}
    with torch.no_grad():
        emb_ref = A.arcface_forward(arc_sd, blob, taps)
        id_ref = F.normalize(emb_ref)
        sh_ref = A.mapping_forward(map_sd, id_ref)
    out = enc.encode(blob.cuda(), want=("embedding", "identity", "shape_code"))
    for name, t in taps.items():
        got = enc.layer_output(name, tuple(t.shape), "cuda").cpu()
        rms = float(t.pow(2).mean().sqrt())
        err = float((got - t).pow(2).mean().sqrt())
        if name in ("stem",) or name.endswith(".0") or name.endswith(".2") or err > 0.03 * rms:
            print(f"{name:12s} rms {rms:8.4f}  err/rms {err / rms:.4f}  max {float((got - t).abs().max()):.4f}")
    for k, ref in (("embedding", emb_ref), ("identity", id_ref), ("shape_code", sh_ref)):
        got = out[k].cpu()
        rms = float(ref.pow(2).mean().sqrt())
        cos = float(F.cosine_similarity(got, ref).min())
        print(f"{k:12s} rms {rms:.4f} err/rms {float((got - ref).pow(2).mean().sqrt()) / rms:.4f} "
              f"max|err| {float((got - ref).abs().max()):.4e} min cosine {cos:.6f}")
    prof, total, conv = enc.profile(32)
    ms = sum(p[1] for p in prof)
    fl = sum(p[2] for p in prof)
    print(f"B=32 eager pass {ms:.3f} ms over {len(prof)} launches ({conv} tcgen05 convs); listed FLOPs {fl / 1e9:.1f} G")
    for n, m, f in sorted(prof, key=lambda p: -p[1])[:8]:
        print(f"   {n:20s} {m * 1e3:8.1f} us")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    x = A.make_blob(64, seed=1).cuda()
    enc.encode(x)
    e0.record()
    for _ in range(5):
        enc.encode(x)
    e1.record()
    torch.cuda.synchronize()
    print(f"B=64 graph: {e0.elapsed_time(e1) / 5:.3f} ms per batch -> {64 / (e0.elapsed_time(e1) / 5e3):.0f} faces/s")


if __name__ == "__main__":
    main()
