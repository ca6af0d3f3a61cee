"""Parity of the MICA identity encoder on the device (ArcFace iResNet-100 + F.normalize +
MappingNetwork, csrc/arcface.cu through the C ABI) with the reference modules.

Golden: tests/golden/arcface_b2.npz, produced by the UNMODIFIED reference (model/mica/arcface.py, model/mica/generator.py;
oracle/make_golden_arcface.py) on seeded weights and a seeded blob; plus the CPU oracle layer by layer.

Stated tolerances (bf16 operands and bf16 activation storage through 100 convolutions, fp32 accumulation):
  residual stream after every block: rms error <= 2.5 % of the tensor's rms (measured 0.2 % -> 1.3 % from stem to layer4)
  512-d embedding: rms error <= 2.5 % of its rms; identity (the normalised embedding): cosine >= 0.9995 per face
  shape code: rms error <= 1 % of its rms (the regressor runs in fp32)
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import b200sr3
from oracle import arcface_oracle as A

pytestmark = pytest.mark.gpu

TOL_BLOCK, TOL_EMB, MIN_COS, TOL_SHAPE = 0.025, 0.025, 0.9995, 0.01


@pytest.fixture(scope="module")
def weights():
    return A.make_arcface_state_dict(0), A.make_mapping_state_dict(0)


@pytest.fixture(scope="module")
def encoder(weights):
    enc = b200sr3.MicaEncoder()
    enc.arcface.load_state_dict(weights[0], strict=True)        # the reference's key / shape contract
    enc.regressor.load_state_dict(weights[1], strict=True)
    return enc.cuda().eval()


def _rel(got, ref):
    return float((got - ref).pow(2).mean().sqrt()) / float(ref.pow(2).mean().sqrt())


def test_encoder_matches_the_reference_golden(golden_dir, weights, encoder):
    g = np.load(os.path.join(golden_dir, "arcface_b2.npz"))
    assert A.digest(weights[0]) == str(g["arcface_sha256"]) and A.digest(weights[1]) == str(g["mapping_sha256"])
    blob = A.make_blob(2, seed=int(g["blob_seed"]))
    out = encoder.encode(blob.cuda(), want=("embedding", "identity", "shape_code"))
    emb, ident, shape = (torch.from_numpy(g[k]) for k in ("embedding", "identity", "shape_code"))
    assert _rel(out["embedding"].cpu(), emb) <= TOL_EMB
    assert float(F.cosine_similarity(out["identity"].cpu(), ident).min()) >= MIN_COS
    assert float((out["identity"].norm(dim=1) - 1).abs().max()) < 1e-5          # F.normalize
    assert _rel(out["shape_code"].cpu(), shape) <= TOL_SHAPE
    # the module surface of the reference: Arcface()(images) -> embedding, MicaEncoder()(images) -> (identity, shape)
    assert torch.equal(encoder.arcface(blob.cuda()), out["embedding"])
    ident2, shape2 = encoder(blob.cuda())
    assert torch.equal(ident2, out["identity"]) and torch.equal(shape2, out["shape_code"])


def test_every_block_against_the_oracle(weights, encoder):
    """stem + the 49 IBasicBlocks: stride-1 blocks (identity K segment), the four stride-2 blocks (input-parity views +
    1x1 stride-2 downsample segment), partial tiles at 56 / 28 / 14 / 7 px."""
    blob = A.make_blob(3, seed=5)
    taps = {}
    with torch.no_grad():
        A.arcface_forward(weights[0], blob, taps)
    encoder.encode(blob.cuda())
    assert len(taps) == 50
    worst = 0.0
    for name, t in taps.items():
        got = encoder.layer_output(name, tuple(t.shape), "cuda").cpu()
        worst = max(worst, _rel(got, t))
        assert _rel(got, t) <= TOL_BLOCK, name
    print(f"worst block error {worst:.4f} of rms")


def test_batch_invariance_and_sizes(encoder):
    """A face's codes do not depend on its batch (no cross-sample op: BatchNorm runs on running statistics), including
    batches that leave the last 8-image group of the fc kernel and the last tile wave ragged."""
    blob = A.make_blob(11, seed=9).cuda()
    full = encoder.encode(blob, want=("embedding", "identity", "shape_code"))
    for sl in (slice(0, 1), slice(1, 4), slice(4, 11)):
        part = encoder.encode(blob[sl].contiguous(), want=("embedding", "identity", "shape_code"))
        for k in full:
            assert torch.equal(part[k], full[k][sl]), k
    assert all(torch.isfinite(v).all() for v in full.values())
    with pytest.raises(ValueError):
        encoder.encode(torch.zeros(1, 3, 64, 64, device="cuda"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        encoder.encode(torch.zeros(1, 3, 112, 112))


def test_sr_output_to_shape_code_end_to_end(encoder):
    """Config 5's tail: SR image -> tensor2img -> resize 224 -> ArcFace blob (bit-exact kernels, handoff.cu) -> identity
    and shape code, all on the device, against the same chain on the CPU (OpenCV + oracle)."""
    import cv2
    from b200sr3 import mica_handoff
    g = torch.Generator().manual_seed(3)
    sr = (torch.rand(2, 3, 128, 128, generator=g) * 2 - 1)
    h = mica_handoff.sr_to_mica(sr.cuda())
    ident, shape = encoder(h["arcface"])
    blobs = []
    for i in range(2):
        img = ((sr[i].clamp(-1, 1) + 1) / 2 * 255).numpy().round().astype(np.uint8).transpose(1, 2, 0)   # core/metrics.py:16-42
        up = cv2.resize(img, (224, 224))
        blobs.append(cv2.dnn.blobFromImages([up], 1.0 / 127.5, (112, 112), (127.5, 127.5, 127.5), swapRB=True)[0])
    blob = torch.from_numpy(np.stack(blobs))
    assert torch.equal(h["arcface"].cpu(), blob)
    w = A.make_arcface_state_dict(0), A.make_mapping_state_dict(0)
    with torch.no_grad():
        ident_ref, shape_ref = A.mica_encode(w[0], w[1], blob)
    assert float(F.cosine_similarity(ident.cpu(), ident_ref).min()) >= MIN_COS
    assert _rel(shape.cpu(), shape_ref) <= TOL_SHAPE
