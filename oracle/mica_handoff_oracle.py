"""TEST INFRASTRUCTURE ONLY (oracle/): CPU restatement of the SR -> MICA hand-off.

Not product code. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may import this.

What it restates (reference = zouiner/3d-super-resolution-Face-reconstruction):
  * core/metrics.py:16-42            tensor2img: clamp to [-1,1] -> [0,1] -> *255 -> round half to even -> uint8 HWC
  * model/sr3d/model.py:372-376      sr_up_img = cv2.resize(sr_img, (224, 224))           (model2 path, also :484)
  * model/sr3d/model.py:127-131      create_arcface_embeddings = cv2.dnn.blobFromImages(img, 1/127.5, (112,112),
                                     mean 127.5, swapRB=True)
  * model/sr3d/model.py:380-382      image = sr_up_img / 255.  (HWC -> CHW)
  * core/metrics.py:44-50 + model/sr3d/model.py:105-124,477-481   model3 path: tensor2tensor_img(x)*255 ->
                                     create_tensor_blob: (v-127.5)/127.5 -> F.interpolate(bilinear, 112) -> swap R/B

The arithmetic of the model2 path lives in a third-party dependency that is not under /root/reference: OpenCV
(requirements.txt pins opencv-python 4.9.0.80; this image has 4.13.0). Its published algorithm, restated here in
integer numpy:
  * cv::resize INTER_LINEAR on CV_8U (imgproc/src/resize.cpp): per axis fx = float((d+0.5)*scale-0.5),
    s = floor(fx), fx -= s; along x the border taps are clamped (s<0 -> s=0,fx=0; s>=w-1 -> s=w-1,fx=0), along y
    only the ROW index is clipped and the weights are kept; coefficients are round-half-even of (1-fx, fx) * 2048 as
    int16; horizontal pass in int32, vertical pass ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2.
  * a 2x down-scale with INTER_LINEAR is executed as INTER_AREA: (a+b+c+d+2)>>2 (resize.cpp, "is_area_fast").
  * blobFromImages (dnn/src/dnn_utils.cpp): resize on the uint8 image, convert to float32, subtract the mean, multiply
    by the scale factor (two float32 roundings), then swap channels 0 and 2 into NCHW.

Parity pin: the reference has no tests for this path. The restatement is pinned against OpenCV itself and against
the reference's own core/metrics.tensor2img, both run in the build container by oracle/make_golden_mica.py (bit-exact
on every case, sizes 8..224); their outputs are committed as tests/golden/mica_handoff.npz and re-checked by
tests/test_mica_handoff_cpu.py wherever the suite runs (and against the installed cv2 directly when it imports).
"""
import numpy as np

UP = 224        # cv2.resize target (model/sr3d/model.py:374)
BLOB = 112      # ArcFace input (model/sr3d/model.py:129)


def tensor2img(x):
    """core/metrics.py:16-42 for a [B,3,H,W] float32 batch -> uint8 [B,H,W,3] (RGB), image by image."""
    x = np.asarray(x, dtype=np.float32)
    t = np.clip(x, np.float32(-1), np.float32(1))
    t = (t - np.float32(-1)) / np.float32(2)                 # (tensor - min) / (max - min), float32
    img = np.round(t * np.float32(255.0))                    # numpy round = half to even
    return np.transpose(img, (0, 2, 3, 1)).astype(np.uint8)


def _linear_coefs(dn, sn, clamp):
    scale = 1.0 / (float(dn) / float(sn))                    # cv::resize: inv_scale = dsize/ssize; scale = 1/inv_scale
    d = np.arange(dn, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp:
        lo = s < 0
        f[lo] = 0
        s[lo] = 0
        hi = s >= sn - 1
        f[hi] = 0
        s[hi] = sn - 1
    a0 = np.rint((np.float32(1) - f) * np.float32(2048)).astype(np.int64)
    a1 = np.rint(f * np.float32(2048)).astype(np.int64)
    return np.clip(s, 0, sn - 1), np.clip(s + 1, 0, sn - 1), a0, a1


def resize_linear_u8(img, dw=UP, dh=UP):
    """cv2.resize(img, (dw, dh)) for one uint8 HWC image, INTER_LINEAR (not a 2x down-scale)."""
    sh, sw, _ = img.shape
    xs, xs1, xa0, xa1 = _linear_coefs(dw, sw, True)
    ys, ys1, ya0, ya1 = _linear_coefs(dh, sh, False)
    s = img.astype(np.int64)
    h = s[:, xs, :] * xa0[None, :, None] + s[:, xs1, :] * xa1[None, :, None]
    r0, r1 = h[ys], h[ys1]
    out = (((ya0[:, None, None] * (r0 >> 4)) >> 16) + ((ya1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return out.astype(np.uint8)


def area2x2_u8(img):
    s = img.astype(np.int64)
    return ((s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2).astype(np.uint8)


def arcface_blob(up):
    """cv2.dnn.blobFromImages([up], 1/127.5, (112,112), (127.5,)*3, swapRB=True)[0] for a 224x224x3 uint8 image."""
    assert up.shape == (UP, UP, 3)
    small = area2x2_u8(up).astype(np.float32)
    v = (small - np.float32(127.5)) * np.float32(1.0 / 127.5)
    return np.ascontiguousarray(v[:, :, ::-1].transpose(2, 0, 1))


def handoff_u8(x):
    """model2 path for a batch: returns (sr_img u8 [B,R,R,3], up u8 [B,224,224,3], image f32 [B,3,224,224],
    arcface f32 [B,3,112,112]). `image` is sr_up_img/255 (float64 in the reference) rounded to float32."""
    sr = tensor2img(x)
    up = np.stack([resize_linear_u8(im) for im in sr])
    image = (up.astype(np.float64) / 255.0).astype(np.float32).transpose(0, 3, 1, 2)
    blob = np.stack([arcface_blob(u) for u in up])
    return sr, up, np.ascontiguousarray(image), blob


def tensor_blob_f32(x):
    """model3 path (create_tensor_blob on tensor2tensor_img(x)*255): float32 bilinear (align_corners=False) resize of
    (v-127.5)/127.5 to 112x112, channels swapped. Written out with explicit float32 steps in torch's order
    (aten upsample_bilinear2d: source index = max((d+0.5)*scale-0.5, 0), lambda in float32)."""
    x = np.asarray(x, dtype=np.float32)
    B, C, H, W = x.shape
    t = np.clip(x, np.float32(-1), np.float32(1))
    t = (t - np.float32(-1)) / np.float32(2)
    v = (t * np.float32(255.0) - np.float32(127.5)) / np.float32(127.5)

    def idx(dn, sn):
        scale = np.float32(sn / dn)
        d = np.arange(dn, dtype=np.float32)
        src = np.maximum((d + np.float32(0.5)) * scale - np.float32(0.5), np.float32(0))
        i0 = np.minimum(src.astype(np.int64), sn - 1)
        i1 = np.minimum(i0 + 1, sn - 1)
        l1 = (src - i0.astype(np.float32)).astype(np.float32)
        return i0, i1, np.float32(1) - l1, l1

    y0, y1, hy0, hy1 = idx(BLOB, H)
    x0, x1, hx0, hx1 = idx(BLOB, W)
    top = hx0[None, None, None, :] * v[:, :, y0][:, :, :, x0] + hx1[None, None, None, :] * v[:, :, y0][:, :, :, x1]
    bot = hx0[None, None, None, :] * v[:, :, y1][:, :, :, x0] + hx1[None, None, None, :] * v[:, :, y1][:, :, :, x1]
    out = hy0[None, None, :, None] * top + hy1[None, None, :, None] * bot
    return np.ascontiguousarray(out[:, ::-1].astype(np.float32))
