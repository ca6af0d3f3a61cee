"""SR -> MICA hand-off on the device: the host mirror of the reference's per-image CPU round trip.

Reference call sites (model/sr3d/model.py, identical code in test_val :366-382 and the training loop :470-486):

    sr_img = Metrics.tensor2img(visuals['SR'])                 # core/metrics.py:16-42 -> uint8 HWC on the host
    sr_up_img = cv2.resize(sr_img, (224, 224))                 # :374
    temp_arcface = self.create_arcface_embeddings(sr_up_img)   # :127-131 cv2.dnn.blobFromImages(112, swapRB)
    temp_arcface = torch.tensor(temp_arcface).cuda()[None]
    sr_up_img = torch.tensor((sr_up_img / 255.).transpose(2, 0, 1)).cuda()[None]
    ... self.encode_mica(sr_up_img, temp_arcface)

Here the same tensors are produced by two kernels of libb200sr3 without leaving the GPU, for the whole batch at once,
bit-identical to OpenCV's fixed-point arithmetic (tests/test_gpu_mica_handoff.py). Names follow the reference:
`tensor2img`, `create_arcface_embeddings`, `create_tensor_blob`. There is no CPU fallback: CPU tensors are an error.
"""
import ctypes as C

import torch

from . import _lib

UP = 224
BLOB = 112


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p()


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _batched(t, name):
    if t.dim() == 3:
        t = t.unsqueeze(0)
    if t.dim() != 4:
        raise TypeError(f"{name}: expected a 3-D or 4-D tensor, got {t.dim()}-D")
    if t.device.type != "cuda":
        raise ValueError(f"{name}: expects a CUDA tensor (b200sr3 has no CPU fallback)")
    return t


@torch.no_grad()
def tensor2img(tensor):
    """core/metrics.py:16-42 for a batch: fp32 [B,3|1,H,W] (or [C,H,W]) in any range -> uint8 [B,H,W,C] on the same
    device. Unlike the reference a 4-D input is NOT tiled into a make_grid mosaic: every image is converted on its
    own, which is what the hand-off needs (the reference only ever passes single images on this path)."""
    x = _batched(tensor, "tensor2img").float().contiguous()
    B, Cn, H, W = x.shape
    img = torch.empty((B, H, W, Cn), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().b200sr3_tensor2img(_ptr(x), B, Cn, H, W, _ptr(img), _stream()))
    return img


@torch.no_grad()
def sr_to_mica(sr, want_up=True, want_image=True, want_arcface=True):
    """model/sr3d/model.py:366-382 for a batch of SR outputs: sr fp32 [B,3,R,R] in [-1,1] ->
    dict(sr_img uint8 [B,R,R,3], up uint8 [B,224,224,3], image fp32 [B,3,224,224], arcface fp32 [B,3,112,112])."""
    img = tensor2img(sr)
    B, R, W, Cn = img.shape
    if R != W or Cn != 3:
        raise ValueError("sr_to_mica: expects square RGB images")
    dev = img.device
    up = torch.empty((B, UP, UP, 3), dtype=torch.uint8, device=dev) if want_up else None
    image = torch.empty((B, 3, UP, UP), dtype=torch.float32, device=dev) if want_image else None
    blob = torch.empty((B, 3, BLOB, BLOB), dtype=torch.float32, device=dev) if want_arcface else None
    lib = _lib.load()
    with torch.cuda.device(dev):
        for lo in range(0, B, 65535):
            hi = min(B, lo + 65535)
            _lib.check(lib.b200sr3_mica_handoff(_ptr(img[lo:hi]), hi - lo, R, _ptr(up[lo:hi] if want_up else None),
                                                _ptr(image[lo:hi] if want_image else None),
                                                _ptr(blob[lo:hi] if want_arcface else None), _stream()))
    return {"sr_img": img, "up": up, "image": image, "arcface": blob}


@torch.no_grad()
def create_arcface_embeddings(images_u8):
    """model/sr3d/model.py:127-131 for uint8 [B,224,224,3] (or [224,224,3]) CUDA images -> fp32 [B,3,112,112]."""
    x = images_u8 if images_u8.dim() == 4 else images_u8.unsqueeze(0)
    if x.dtype != torch.uint8 or x.shape[1:] != (UP, UP, 3) or x.device.type != "cuda":
        raise ValueError("create_arcface_embeddings: expects uint8 CUDA [B,224,224,3]")
    x = x.contiguous()
    blob = torch.empty((x.shape[0], 3, BLOB, BLOB), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().b200sr3_mica_handoff(_ptr(x), x.shape[0], UP, None, None, _ptr(blob), _stream()))
    return blob


@torch.no_grad()
def create_tensor_blob(sr):
    """The model3 variant, model/sr3d/model.py:477-481: create_tensor_blob (:105-124) of tensor2tensor_img(sr) * 255
    (core/metrics.py:44-50). sr fp32 [B,3,R,R] (or [3,R,R]) -> fp32 [B,3,112,112]."""
    x = _batched(sr, "create_tensor_blob").float().contiguous()
    B, Cn, H, W = x.shape
    if Cn != 3 or H != W:
        raise ValueError("create_tensor_blob: expects square RGB images")
    blob = torch.empty((B, 3, BLOB, BLOB), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        for lo in range(0, B, 65535):
            hi = min(B, lo + 65535)
            _lib.check(_lib.load().b200sr3_tensor_blob(_ptr(x[lo:hi]), hi - lo, H, _ptr(blob[lo:hi]), _stream()))
    return blob
