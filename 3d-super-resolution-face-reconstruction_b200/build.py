"""Build libb200sr3.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python 3d-super-resolution-face-reconstruction_b200/build.py [--force] [--verbose]

The library links cudart statically and resolves libcuda at run time, so it loads on a host
without a driver (symbol checks) and fails loudly, not silently, when asked to compute there.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "b200sr3", "libb200sr3.so")
SOURCES = ["abi.cu", "engine.cu", "conv_umma.cu", "conv_halo.cu", "kernels.cu", "handoff.cu", "arcface.cu"]
HEADERS = ["common.cuh", "conv_umma.cuh", "conv_halo.cuh", "kernels.cuh", "engine.cuh", "arcface.cuh", os.path.join("..", "..", "include", "b200sr3.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
         "-cudart", "static"]


def _digest(flags):
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(flags).encode())
    return h.hexdigest()


def build(force=False, verbose=False, timing=False, extra_flags=(), out=None):
    """timing=True builds libb200sr3_timing.so with the per-role cycle counters compiled in (-DB200SR3_ROLE_TIMING=1);
    select it with B200SR3_LIB=<path> for tools/halo_bench.py runs under B200SR3_CONV_TIMING=1. `extra_flags` / `out`
    build a variant library beside the default one (same-box A/B runs through B200SR3_LIB)."""
    flags = list(FLAGS) + list(extra_flags)
    target = out or OUT
    if timing:
        target = target.replace("libb200sr3.so", "libb200sr3_timing.so")
        flags.append("-DB200SR3_ROLE_TIMING=1")
    stamp = target + ".stamp"
    dig = _digest(flags)
    if not force and os.path.exists(target) and os.path.exists(stamp) and open(stamp).read() == dig:
        return target
    tag = os.path.basename(target).replace("libb200sr3", "").replace(".so", "")
    objdir = os.path.join(HERE, "build" + tag)
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [NVCC] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out_text, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {src}\n{out_text}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    link = [NVCC, "-shared", "-cudart", "static", "-o", target] + objs + ["-Xlinker", "--exclude-libs,ALL"]
    subprocess.run(link, check=True)
    with open(stamp, "w") as f:
        f.write(dig)
    return target


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, timing="--timing" in sys.argv))
