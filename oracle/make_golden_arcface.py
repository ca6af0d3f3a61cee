"""TEST INFRASTRUCTURE ONLY (oracle/): goldens of the MICA identity encoder from the REAL reference.

Run in the build container only (needs /root/reference):

    python -m oracle.make_golden_arcface

Builds the unmodified reference Arcface (model/mica/arcface.py:165-200) and MappingNetwork
(model/mica/generator.py:31-60, z_dim 512 / hidden 300 / mapping_layers 3 / n_shape 300 as
config/default/config.py:133 and model/sr3d/model.py:68-75 construct them), loads oracle.arcface_oracle's seeded
weights with strict=True (proving the key / shape contract), runs encode_mica's F.normalize(arcface(blob))
(model/sr3d/model.py:167) and the regressor on a seeded blob, checks the oracle restatement against them and stores
the reference's outputs. Weights (261 MB) and the blob are regenerated from their seeds, pinned by sha256.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

from . import arcface_oracle as A
from .make_golden import OUT, REF


def main():
    torch.set_num_threads(os.cpu_count())
    sys.path.insert(0, REF)
    from model.mica.arcface import Arcface              # noqa: the reference's own modules
    from model.mica.generator import MappingNetwork     # noqa
    arc_sd = A.make_arcface_state_dict(seed=0)
    map_sd = A.make_mapping_state_dict(seed=0)
    arc = Arcface().eval()
    res = arc.load_state_dict(arc_sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    reg = MappingNetwork(512, 300, 300, 3).eval()
    res = reg.load_state_dict(map_sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    blob = A.make_blob(2, seed=0)
    with torch.no_grad():
        emb = arc(blob)                                 # arcface.py:179-200
        ident = F.normalize(emb)                        # model/sr3d/model.py:167
        shape = reg(ident)                              # generator.py:86-88
        taps = {}
        emb_o = A.arcface_forward(arc_sd, blob, taps)
        ident_o, shape_o = A.mica_encode(arc_sd, map_sd, blob)
    d_emb = float((emb - emb_o).abs().max())
    d_id = float((ident - ident_o).abs().max())
    d_sh = float((shape - shape_o).abs().max())
    assert d_emb <= 1e-4 * float(emb.abs().max()) and d_id <= 1e-5 and d_sh <= 1e-5, (d_emb, d_id, d_sh)
    rms = {k: float(v.pow(2).mean().sqrt()) for k, v in taps.items()}
    np.savez_compressed(os.path.join(OUT, "arcface_b2.npz"), embedding=emb.numpy(), identity=ident.numpy(),
                        shape_code=shape.numpy(), blob_seed=0, arcface_weight_seed=0, mapping_weight_seed=0,
                        arcface_sha256=A.digest(arc_sd), mapping_sha256=A.digest(map_sd),
                        blob_sha256=__import__("hashlib").sha256(blob.numpy().tobytes()).hexdigest(),
                        tap_names=np.array(list(rms)), tap_rms=np.array(list(rms.values())))
    line = (f"I ArcFace iResNet-100 + MappingNetwork, B=2 blob 112x112: max|oracle-reference| embedding {d_emb:.2e} "
            f"(|emb|max {float(emb.abs().max()):.2f}), identity {d_id:.2e}, shape code {d_sh:.2e}; "
            f"residual-stream rms stem {rms['stem']:.2f} -> layer4.2 {rms['layer4.2']:.2f}")
    print(line)
    with open(os.path.join(OUT, "README.md"), "a") as f:
        f.write("* " + line + " (`python -m oracle.make_golden_arcface`)\n")


if __name__ == "__main__":
    main()
