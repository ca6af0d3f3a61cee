"""Drop-in for the MICA identity encoder: `Arcface` (model/mica/arcface.py:165-200),
`MappingNetwork` (model/mica/generator.py:31-60) and the `encode_mica` step that joins them
(model/sr3d/model.py:164-170: `F.normalize(self.arcface(arcface_imgs))`, then `self.regressor(arcface)`,
generator.py:86-88). Same constructor arguments, parameter / buffer names and shapes as the reference modules, so their
checkpoints load with strict=True; the forward pass runs in libb200sr3 (csrc/arcface.cu) on a B200. There is no
PyTorch or CPU fallback: a CPU call raises. Only inference (eval mode) is implemented: BatchNorm uses its running
statistics and dropout is the identity, as in the reference's `with torch.no_grad()` evaluation path.
"""
import ctypes as C
import weakref

import torch
from torch import nn

from . import _lib

LAYERS = (3, 13, 30, 3)          # iResNet-100, arcface.py:167
PLANES = (64, 128, 256, 512)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p()


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _no_fallback(what):
    raise RuntimeError(f"b200sr3: {what} runs on a CUDA (sm_100a) device only; there is no CPU fallback "
                       "(use the reference module on CPU)")


class _Holder(nn.Module):
    """Parameter container: the forward pass of this subtree happens in libb200sr3."""

    def forward(self, *a, **k):
        _no_fallback(type(self).__name__)


def _block(inplanes, planes, stride, downsample):
    """The tensors of one IBasicBlock (arcface.py:44-53) under the reference's attribute names."""
    m = _Holder()
    m.bn1 = nn.BatchNorm2d(inplanes, eps=1e-5)
    m.conv1 = nn.Conv2d(inplanes, planes, 3, stride=1, padding=1, bias=False)
    m.bn2 = nn.BatchNorm2d(planes, eps=1e-5)
    m.prelu = nn.PReLU(planes)
    m.conv2 = nn.Conv2d(planes, planes, 3, stride=stride, padding=1, bias=False)
    m.bn3 = nn.BatchNorm2d(planes, eps=1e-5)
    m.downsample = None
    if downsample:
        m.downsample = nn.Sequential(nn.Conv2d(inplanes, planes, 1, stride=stride, bias=False),
                                     nn.BatchNorm2d(planes, eps=1e-5))
    m.stride = stride
    return m


class Arcface(_Holder):
    """iResNet-100 with the reference's state_dict (925 entries, 65,156,160 parameters). `forward(images)` takes the
    ArcFace blob fp32 [B,3,112,112] on a CUDA device and returns the 512-d embedding (not normalised)."""

    def __init__(self, pretrained_path=None, **kwargs):
        super().__init__()
        if kwargs.get("fp16"):
            raise NotImplementedError("b200sr3: the fp16 autocast variant of the reference is not mirrored (bf16 operands, fp32 accumulation)")
        self.fp16 = False
        self.conv1 = nn.Conv2d(3, 64, 3, stride=1, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(64, eps=1e-5)
        self.prelu = nn.PReLU(64)
        inplanes = 64
        for li, (n, planes) in enumerate(zip(LAYERS, PLANES), start=1):
            blocks = []
            for bi in range(n):
                blocks.append(_block(inplanes, planes, 2 if bi == 0 else 1, bi == 0))
                inplanes = planes
            setattr(self, f"layer{li}", nn.Sequential(*blocks))
        self.bn2 = nn.BatchNorm2d(512, eps=1e-5)
        self.dropout = nn.Dropout(p=kwargs.get("dropout", 0), inplace=True)
        self.fc = nn.Linear(512 * 49, kwargs.get("num_features", 512))
        self.features = nn.BatchNorm1d(kwargs.get("num_features", 512), eps=1e-5)
        nn.init.constant_(self.features.weight, 1.0)
        self.features.weight.requires_grad = False
        if pretrained_path is not None:
            import os
            if os.path.exists(pretrained_path):
                self.load_state_dict(torch.load(pretrained_path, map_location="cpu", weights_only=True))
        for p in self.parameters():          # inference only
            p.requires_grad = False
        self._owner = None                   # MicaEncoder that runs this module

    def forward(self, images):
        enc = self._owner() if self._owner is not None else None
        if enc is None:
            enc = MicaEncoder(arcface=self)
            object.__setattr__(self, "_keepalive", enc)
        return enc.encode(images, want=("embedding",))["embedding"]

    forward_arcface = forward


class MappingNetwork(_Holder):
    """generator.py:31-60 (hidden <= 5, i.e. no skip connections: the reference uses mapping_layers = 3)."""

    def __init__(self, z_dim, map_hidden_dim, map_output_dim, hidden=2):
        super().__init__()
        if hidden > 5:
            raise NotImplementedError("b200sr3: mapping networks with skip connections (hidden > 5) are not mirrored")
        self.skips = []
        self.network = nn.ModuleList([nn.Linear(z_dim, map_hidden_dim)] +
                                     [nn.Linear(map_hidden_dim, map_hidden_dim) for _ in range(hidden)])
        self.output = nn.Linear(map_hidden_dim, map_output_dim)
        for lin in self.network:
            nn.init.kaiming_normal_(lin.weight, a=0.2, mode="fan_in", nonlinearity="leaky_relu")
        with torch.no_grad():
            self.output.weight *= 0.25
        self.dims = (z_dim, map_hidden_dim, map_output_dim, hidden)


class MicaEncoder(nn.Module):
    """encode_mica + regressor for a batch: arcface blob [B,3,112,112] -> identity code [B,512] and shape code
    [B,n_shape]. Holds the two reference-shaped modules (`arcface`, `regressor`) and one engine per device."""

    def __init__(self, arcface=None, regressor=None, z_dim=512, map_hidden_dim=300, n_shape=300, mapping_layers=3):
        super().__init__()
        self.arcface = arcface if arcface is not None else Arcface()
        self.regressor = regressor if regressor is not None else MappingNetwork(z_dim, map_hidden_dim, n_shape, mapping_layers)
        object.__setattr__(self.arcface, "_owner", weakref.ref(self))      # (a plain attribute: no module cycle)
        self._engines = {}
        self._weights_serial = 0

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_engines"] = {}
        return state

    def invalidate_weights(self):
        """Call after editing parameters through `.data` (see GaussianDiffusion.invalidate_weights)."""
        self._weights_serial += 1

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.invalidate_weights()
        return out

    def _tensors(self):
        for prefix, mod in (("arcface.", self.arcface), ("regressor.", self.regressor)):
            for k, v in mod.state_dict().items():
                if not k.endswith("num_batches_tracked"):
                    yield prefix + k, v

    def _engine(self, dev):
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        eng = self._engines.get(idx)
        lib = _lib.load()
        if eng is None:
            z, hdim, out, hidden = self.regressor.dims
            handle = C.c_void_p()
            _lib.check(lib.b200sr3_mica_create(idx, z, hdim, hidden, out, C.byref(handle)))
            eng = self._engines[idx] = {"handle": handle, "version": None}
        tensors = list(self._tensors())
        ptrs = vers = 0
        for _, t in tensors:
            ptrs ^= t.data_ptr()
            vers += t._version
        version = (self._weights_serial, ptrs, vers)
        if eng["version"] != version:
            with torch.cuda.device(idx):
                staged = [(k, t.detach().to(device=f"cuda:{idx}", dtype=torch.float32).contiguous()) for k, t in tensors]
                torch.cuda.current_stream().synchronize()
                for k, t in staged:
                    shape = (C.c_int64 * max(t.dim(), 1))(*(t.shape if t.dim() else (1,)))
                    _lib.check(lib.b200sr3_mica_load_tensor(eng["handle"], k.encode(), _ptr(t), shape, max(t.dim(), 1)))
                del staged
                _lib.check(lib.b200sr3_mica_finalize_weights(eng["handle"], _stream()))
            eng["version"] = version
        return eng

    def __del__(self):
        try:
            lib = _lib.load()
            for eng in self._engines.values():
                lib.b200sr3_mica_destroy(eng["handle"])
            self._engines = {}
        except Exception:
            pass

    @torch.no_grad()
    def encode(self, arcface_imgs, want=("identity", "shape_code")):
        if arcface_imgs.device.type != "cuda":
            _no_fallback("MicaEncoder.encode")
        x = arcface_imgs.detach().to(dtype=torch.float32).contiguous()
        if x.dim() != 4 or tuple(x.shape[1:]) != (3, 112, 112):
            raise ValueError("b200sr3: the ArcFace blob is [B,3,112,112]")
        dev = x.device
        eng = self._engine(dev)
        B = x.shape[0]
        n_shape = self.regressor.dims[2]
        outs = {"embedding": (B, 512), "identity": (B, 512), "shape_code": (B, n_shape)}
        bufs = {k: (torch.empty(s, dtype=torch.float32, device=dev) if k in want else None) for k, s in outs.items()}
        with torch.cuda.device(dev):
            _lib.check(_lib.load().b200sr3_mica_encode(eng["handle"], _ptr(x), B, _ptr(bufs["embedding"]),
                                                       _ptr(bufs["identity"]), _ptr(bufs["shape_code"]), _stream()))
        return {k: v for k, v in bufs.items() if v is not None}

    def forward(self, arcface_imgs):
        """(identity, shape_code): codedict['arcface'] of model/sr3d/model.py:167 and the regressor's output."""
        out = self.encode(arcface_imgs)
        return out["identity"], out["shape_code"]

    def layer_output(self, name, shape, device):
        eng = self._engine(torch.device(device))
        buf = torch.empty(shape, dtype=torch.float32, device=device)
        c, h, w = C.c_int(), C.c_int(), C.c_int()
        _lib.check(_lib.load().b200sr3_mica_layer_output(eng["handle"], name.encode(), _ptr(buf), C.byref(c), C.byref(h),
                                                         C.byref(w), _stream()))
        assert (c.value, h.value, w.value) == tuple(shape[1:]), (name, c.value, h.value, w.value)
        return buf

    def profile(self, B, device="cuda"):
        eng = self._engine(torch.device(device))
        n_max = 512
        ms = (C.c_float * n_max)()
        fl = (C.c_double * n_max)()
        names = C.create_string_buffer(48 * n_max)
        n, total, conv = C.c_int(), C.c_int64(), C.c_int64()
        with torch.cuda.device(device):
            _lib.check(_lib.load().b200sr3_mica_profile(eng["handle"], B, n_max, ms, fl, names, len(names), C.byref(n),
                                                        C.byref(total), C.byref(conv), _stream()))
        nm = names.value.decode().split("\n")
        return [(nm[i], ms[i], fl[i]) for i in range(n.value)], total.value, conv.value
