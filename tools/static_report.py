"""Static evidence from the built library, no GPU needed: per kernel the register / stack / static shared-memory
figures (`cuobjdump --dump-resource-usage`) and the count of the SASS mnemonics that prove the Blackwell paths
(`cuobjdump -sass`: UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor load / store, LDTM = tcgen05.ld, UTCBAR =
tcgen05.commit, HMMA = mma.sync, MUFU.TANH, FFMA2 / FADD2 = packed fp32, RED...64 = the int64 statistics adds).

    python tools/static_report.py [path/to/libb200sr3.so] > profiles/<round>_static_report.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "3d-super-resolution-face-reconstruction_b200", "b200sr3", "libb200sr3.so")
MNEMONICS = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "HMMA", "MUFU.TANH", "FFMA2", "FADD2", "RED64", "SPILL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def short(name):
    name = re.sub(r"\b(b200sr3::|void )", "", name)
    name = re.sub(r"\((.*)\)$", "", name)
    return name


def main():
    so = sys.argv[1] if len(sys.argv) > 1 else SO
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", so], capture_output=True, text=True).stdout
    usage, arch, archs = {}, None, collections.Counter()
    lines = res.split("\n")
    for i, l in enumerate(lines):
        m = re.match(r"arch = (\S+)", l)
        if m:
            arch = m.group(1)
            archs[arch] += 1
        m = re.match(r"\s*Function (\S+):", l)
        if m and arch == "sm_100a":
            kv = dict(re.findall(r"(REG|STACK|SHARED|LOCAL):(\d+)", lines[i + 1]))
            usage[m.group(1)] = {k: int(v) for k, v in kv.items()}
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    counts, cur, ninstr = collections.defaultdict(collections.Counter), None, collections.Counter()
    for l in sass.split("\n"):
        m = re.match(r"\s*Function : (\S+)", l)
        if m:
            cur = m.group(1)
            continue
        if cur is None or "/*" not in l:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if not m:
            continue
        op = m.group(1)
        ninstr[cur] += 1
        for key in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "HMMA", "MUFU.TANH", "FFMA2", "FADD2"):
            if op.startswith(key):
                counts[cur][key] += 1
        if op.startswith("RED") and ".64" in op:
            counts[cur]["RED64"] += 1
        if op.startswith(("STL", "LDL")):
            counts[cur]["SPILL"] += 1
    dm = demangle(sorted(usage))
    print("# Static report of `%s`\n" % os.path.relpath(so, ROOT))
    print("Fatbin ELF images: " + ", ".join(f"{n} × {a}" for a, n in sorted(archs.items())) +
          " (the sm_52 image is the empty device-link stub `nvcc -shared` adds; it holds no function).\n")
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    print("Totals over all kernels: " + ", ".join(f"{k} ×{tot[k]}" for k in MNEMONICS if k != "SPILL") +
          f"; local-memory instructions {tot['SPILL']} - in the conv kernels these are the argument block of the out-of-line "
          "barrier-timeout diagnostic (`STL` right before the `vprintf` call, after the slow path's `RET`), not register "
          "spills in a hot loop.\n")
    print("| kernel | regs | stack B | static smem B | SASS instr | " + " | ".join(MNEMONICS[:-1]) + " | LDL/STL |")
    print("|---|---|---|---|---|" + "---|" * len(MNEMONICS))
    for name in sorted(usage, key=lambda n: short(dm[n])):
        u, c = usage[name], counts[name]
        print(f"| `{short(dm[name])}` | {u.get('REG')} | {u.get('STACK')} | {u.get('SHARED')} | {ninstr[name]} | " +
              " | ".join(str(c[k]) if c[k] else "" for k in MNEMONICS) + " |")


if __name__ == "__main__":
    main()
