"""GPU parity of the SR -> MICA hand-off kernels (csrc/handoff.cu, through the C ABI) against the oracle, the golden
outputs of OpenCV / the reference, and the installed cv2: bit-exact for the uint8 images and the ArcFace blob."""
import os

import numpy as np
import pytest
import torch

from oracle import mica_handoff_oracle as M

pytestmark = pytest.mark.gpu


def _run(x):
    from b200sr3 import mica_handoff as H
    out = H.sr_to_mica(torch.from_numpy(x).cuda())
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}


@pytest.mark.parametrize("case", ["r32", "r128"])
def test_goldens_bit_exact(golden_dir, case):
    g = np.load(os.path.join(golden_dir, "mica_handoff.npz"))
    x = g[case + "_x"]
    got = _run(x)
    assert np.array_equal(got["sr_img"], g[case + "_sr"])
    assert np.array_equal(got["up"], g[case + "_up"])
    assert np.array_equal(got["arcface"], g[case + "_blob"])
    assert np.array_equal(got["image"], (g[case + "_up"].astype(np.float64) / 255.0).astype(np.float32).transpose(0, 3, 1, 2))
    from b200sr3 import mica_handoff as H
    tb = H.create_tensor_blob(torch.from_numpy(x).cuda()).cpu().numpy()
    assert float(np.abs(tb - g[case + "_tblob"]).max()) <= 1e-5            # float path: stated tolerance


@pytest.mark.parametrize("R,B", [(8, 3), (16, 5), (32, 4), (64, 7), (100, 2), (128, 9), (224, 2), (300, 1)])
def test_matches_oracle_every_size(R, B):
    rng = np.random.default_rng(R * 31 + B)
    x = rng.uniform(-1.2, 1.2, (B, 3, R, R)).astype(np.float32)
    got = _run(x)
    sr, up, image, blob = M.handoff_u8(x)
    assert np.array_equal(got["sr_img"], sr)
    assert np.array_equal(got["up"], up)
    assert np.array_equal(got["image"], image)
    assert np.array_equal(got["arcface"], blob)


def test_matches_installed_opencv_at_baseline_size():
    """Config 5 shape (32 -> 128 output feeding MICA): 64 faces, against cv2 itself."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    x = rng.uniform(-1.1, 1.1, (64, 3, 128, 128)).astype(np.float32)
    got = _run(x)
    for b in range(0, 64, 7):
        up = cv2.resize(got["sr_img"][b], (224, 224))
        assert np.array_equal(got["up"][b], up)
        blob = cv2.dnn.blobFromImages([up], 1.0 / 127.5, (112, 112), (127.5, 127.5, 127.5), swapRB=True)[0]
        assert np.array_equal(got["arcface"][b], blob)


def test_size_independent_properties():
    from b200sr3 import mica_handoff as H
    # constant image -> constant outputs; saturation; idempotence of tensor2img on its own output
    for v, u in ((-3.0, 0), (0.0, 128), (1.0, 255)):
        out = H.sr_to_mica(torch.full((2, 3, 128, 128), v, device="cuda"))
        assert bool((out["sr_img"] == u).all()) and bool((out["up"] == u).all())
    x = torch.rand(4, 3, 64, 64, device="cuda") * 2 - 1
    a = H.tensor2img(x)
    back = a.permute(0, 3, 1, 2).float() / 255.0 * 2 - 1
    assert torch.equal(H.tensor2img(back), a)
    # batch invariance: an image's outputs do not depend on its neighbours
    full = H.sr_to_mica(x)
    one = H.sr_to_mica(x[2:3])
    for k in ("sr_img", "up", "image", "arcface"):
        assert torch.equal(full[k][2:3], one[k])
    # create_arcface_embeddings on the up image equals the fused blob
    assert torch.equal(H.create_arcface_embeddings(full["up"]), full["arcface"])


def test_rejects_cpu_tensors_and_bad_shapes():
    from b200sr3 import mica_handoff as H, _lib
    with pytest.raises(ValueError):
        H.tensor2img(torch.zeros(1, 3, 8, 8))
    with pytest.raises(_lib.B200Error):
        H.sr_to_mica(torch.zeros(1, 3, 448, 448, device="cuda"))        # OpenCV's INTER_AREA shortcut: not implemented
