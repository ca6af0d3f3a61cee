"""The C-ABI library builds, loads without a GPU/driver and exports every declared symbol."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "b200sr3.h")).read()
    return sorted(set(re.findall(r"B200SR3_API\s+[\w\s\*]+?\b(b200sr3_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    from b200sr3 import _lib
    assert _declared() == sorted(_lib.EXPORTS)


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    for name in _declared():
        assert hasattr(lib, name), name
    header = open(os.path.join(ROOT, "include", "b200sr3.h")).read()
    declared_version = int(re.search(r"#define\s+B200SR3_ABI_VERSION\s+(\d+)", header).group(1))
    from b200sr3 import _lib
    assert lib.b200sr3_abi_version() == declared_version == _lib.ABI_VERSION


def test_library_has_blackwell_code(built_lib):
    import subprocess
    sass = subprocess.run(["cuobjdump", "-sass", built_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):       # tcgen05.mma, TMA load, tcgen05.ld
        assert mnemonic in sass, mnemonic


def test_create_fails_loudly_without_gpu(built_lib):
    import torch
    if torch.cuda.is_available():
        return
    from b200sr3 import _lib
    lib = _lib.load()
    import b200sr3
    mopt = b200sr3.configs.model_opt(10)
    cfg = _lib.make_config(mopt["unet"], mopt["diffusion"])
    h = ctypes.c_void_p()
    rc = lib.b200sr3_create(ctypes.byref(cfg), 0, ctypes.byref(h))
    assert rc != 0 and not h
    assert b"no CUDA device" in lib.b200sr3_last_error() or b"CPU fallback" in lib.b200sr3_last_error()
