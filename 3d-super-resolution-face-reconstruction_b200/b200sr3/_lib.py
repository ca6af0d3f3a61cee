"""ctypes binding of libb200sr3.so (include/b200sr3.h). No torch types cross this boundary:
only raw pointers, sizes and a stream handle."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200SR3_LIB selects another build of the same library (same-box A/B runs of two kernel variants)
LIB_PATH = os.environ.get("B200SR3_LIB") or os.path.join(_HERE, "libb200sr3.so")
MAX_LEVELS = 8
ABI_VERSION = 2          # B200SR3_ABI_VERSION of include/b200sr3.h
NOISE_INJECTED = 1
NOISE_PHILOX = 2


class Config(C.Structure):
    _fields_ = [
        ("in_channel", C.c_int32), ("out_channel", C.c_int32), ("inner_channel", C.c_int32),
        ("norm_groups", C.c_int32), ("res_blocks", C.c_int32), ("n_mults", C.c_int32),
        ("channel_mults", C.c_int32 * MAX_LEVELS), ("n_attn_res", C.c_int32),
        ("attn_res", C.c_int32 * MAX_LEVELS), ("image_size", C.c_int32), ("conditional", C.c_int32),
    ]


class B200Error(RuntimeError):
    """Raised when a libb200sr3 call returns non-zero; carries b200sr3_last_error()."""


_P = C.c_void_p
_SIGNATURES = {
    "b200sr3_last_error": (C.c_char_p, []),
    "b200sr3_abi_version": (C.c_int, []),
    "b200sr3_create": (C.c_int, [C.POINTER(Config), C.c_int, C.POINTER(_P)]),
    "b200sr3_destroy": (C.c_int, [_P]),
    "b200sr3_num_tensors": (C.c_int, [_P]),
    "b200sr3_tensor_info": (C.c_int, [_P, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int64), C.POINTER(C.c_int)]),
    "b200sr3_load_tensor": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), C.c_int]),
    "b200sr3_finalize_weights": (C.c_int, [_P, _P]),
    "b200sr3_set_schedule": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, _P, _P, _P]),
    "b200sr3_unet_forward": (C.c_int, [_P, _P, _P, C.c_float, C.c_int, C.c_int, _P, _P]),
    "b200sr3_step": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "b200sr3_sample": (C.c_int, [_P, _P, C.c_int, _P, C.c_uint64, C.c_int64, C.c_int, C.c_int, _P, _P, _P]),
    "b200sr3_philox_normal": (C.c_int, [_P, C.c_uint64, C.c_int, C.c_int64, C.c_int, C.c_int, _P, _P]),
    "b200sr3_num_snapshots": (C.c_int, [_P]),
    "b200sr3_sample_host": (C.c_int, [_P, _P, C.c_uint64, C.c_int64, C.c_int, C.c_int, _P, _P]),
    "b200sr3_layer_output": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), _P]),
    "b200sr3_last_launch_count": (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "b200sr3_profile_step": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_double),
                                       C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_char_p, C.c_int,
                                       C.POINTER(C.c_int), _P]),
    "b200sr3_conv2d": (C.c_int, [C.c_int, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_int, _P, C.c_int, C.POINTER(C.c_float), _P]),
    "b200sr3_conv_block": (C.c_int, [C.c_int, _P, C.c_int, _P, C.c_int, _P, _P, C.c_int, C.c_int, _P, _P, _P, C.c_int,
                                     _P, C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, C.c_int,
                                     C.POINTER(C.c_float), _P]),
    "b200sr3_tensor2img": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "b200sr3_mica_handoff": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P, _P]),
    "b200sr3_tensor_blob": (C.c_int, [_P, C.c_int, C.c_int, _P, _P]),
    "b200sr3_mica_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "b200sr3_mica_destroy": (C.c_int, [_P]),
    "b200sr3_mica_num_tensors": (C.c_int, [_P]),
    "b200sr3_mica_tensor_info": (C.c_int, [_P, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int64), C.POINTER(C.c_int)]),
    "b200sr3_mica_load_tensor": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), C.c_int]),
    "b200sr3_mica_finalize_weights": (C.c_int, [_P, _P]),
    "b200sr3_mica_encode": (C.c_int, [_P, _P, C.c_int, _P, _P, _P, _P]),
    "b200sr3_mica_layer_output": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), _P]),
    "b200sr3_mica_profile": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_double), C.c_char_p,
                                       C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_int64), _P]),
}
EXPORTS = tuple(_SIGNATURES)

_lib = None


def load():
    """Load the shared library (once). There is no fallback: a missing library is an error."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with "
            "`python 3d-super-resolution-face-reconstruction_b200/build.py` "
            "(or __graft_entry__.build()). b200sr3 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.b200sr3_abi_version() != ABI_VERSION:
        raise ImportError("libb200sr3.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise B200Error(load().b200sr3_last_error().decode("utf-8", "replace"))


def make_config(unet_opt, diffusion_opt):
    """opt['sr']['model']['unet'] / ['diffusion'] -> b200sr3_config (model/sr/networks.py:89-108)."""
    cfg = Config()
    cfg.in_channel = int(unet_opt["in_channel"])
    out_ch = unet_opt.get("out_channel")
    cfg.out_channel = int(out_ch if out_ch is not None else unet_opt["in_channel"])
    cfg.inner_channel = int(unet_opt["inner_channel"])
    cfg.norm_groups = int(unet_opt.get("norm_groups") or 32)
    cfg.res_blocks = int(unet_opt["res_blocks"])
    mults = list(unet_opt["channel_multiplier"])
    if len(mults) > MAX_LEVELS:
        raise ValueError("channel_multiplier has more than %d levels" % MAX_LEVELS)
    cfg.n_mults = len(mults)
    for i, m in enumerate(mults):
        cfg.channel_mults[i] = int(m)
    attn = unet_opt["attn_res"]
    attn = list(attn) if isinstance(attn, (list, tuple)) else [attn]
    cfg.n_attn_res = len(attn)
    for i, a in enumerate(attn):
        cfg.attn_res[i] = int(a)
    cfg.image_size = int(diffusion_opt["image_size"])
    cfg.conditional = 1 if diffusion_opt.get("conditional", True) else 0
    return cfg
