"""The SR -> MICA hand-off oracle (oracle/mica_handoff_oracle.py) against the golden outputs of the reference's
core/metrics.tensor2img and of OpenCV (tests/golden/mica_handoff.npz, made by oracle/make_golden_mica.py), and
against the installed cv2 directly when it imports. No GPU."""
import os

import numpy as np
import pytest

from oracle import mica_handoff_oracle as M


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "mica_handoff.npz"))


@pytest.mark.parametrize("case", ["r32", "r128"])
def test_oracle_matches_reference_goldens(gold, case):
    x = gold[case + "_x"]
    sr, up, image, blob = M.handoff_u8(x)
    assert np.array_equal(sr, gold[case + "_sr"])           # tensor2img, incl. the planted .5 ties and the clamp
    assert np.array_equal(up, gold[case + "_up"])           # cv2.resize, bit-exact
    assert np.array_equal(blob, gold[case + "_blob"])       # cv2.dnn.blobFromImages, bit-exact float32
    assert np.array_equal(image, (gold[case + "_up"].astype(np.float64) / 255.0).astype(np.float32).transpose(0, 3, 1, 2))
    assert float(np.abs(M.tensor_blob_f32(x) - gold[case + "_tblob"]).max()) <= 1e-5


def test_oracle_matches_installed_opencv():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    for R in (8, 16, 32, 64, 100, 128, 224, 300):
        im = rng.integers(0, 256, (R, R, 3), dtype=np.uint8)
        up = cv2.resize(im, (224, 224))
        assert np.array_equal(M.resize_linear_u8(im), up), R
        blob = cv2.dnn.blobFromImages([up], 1.0 / 127.5, (112, 112), (127.5, 127.5, 127.5), swapRB=True)[0]
        assert np.array_equal(M.arcface_blob(up), blob), R


def test_edge_cases():
    # constant images stay constant through both resizes; extremes saturate exactly
    for v, u in ((-3.0, 0), (-1.0, 0), (0.0, 128), (1.0, 255), (7.0, 255)):      # (0+1)/2*255 = 127.5 -> 128 (even)
        x = np.full((1, 3, 16, 16), v, np.float32)
        sr, up, image, blob = M.handoff_u8(x)
        assert (sr == u).all() and (up == u).all()
        assert np.allclose(blob, (u - 127.5) / 127.5, atol=1e-7)
    # channel swap: a pure-red image lands in blob channel 2
    x = np.full((1, 3, 8, 8), -1.0, np.float32)
    x[:, 0] = 1.0
    blob = M.handoff_u8(x)[3]
    assert (blob[0, 2] == 1.0).all() and (blob[0, 0] == -1.0).all() and (blob[0, 1] == -1.0).all()
