#!/usr/bin/env python3
"""SASS evidence per kernel of gomel_b200/libgomelcuda.so (no GPU needed):
    python profiles/sass_summary.py > profiles/r02_sass_summary.md
opcode counts from `cuobjdump -sass`, registers / shared memory from `cuobjdump -res-usage`."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gomel_b200", "libgomelcuda.so")
KEY = ["FADD2", "FFMA2", "FMUL2", "FADD", "FFMA", "FMUL", "DADD", "DFMA", "DMUL", "MUFU", "LDS", "STS", "LDG", "STG", "RED", "REDG", "ATOMG",
       "SHFL", "UBLKCP", "SYNCS", "CCTL", "BAR", "UTMALDG", "UTCHMMA", "HMMA", "LDTM", "MOV"]


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
    except OSError:
        return n


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            usage[cur] = line.strip()
            cur = None
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
        if cur is not None and m:
            cur[m.group(1).split(".")[0]] += 1
    print("# SASS summary of gomel_b200/libgomelcuda.so (round 2, final build)\n")
    print(f"`cuobjdump -sass` / `-res-usage`; target architectures in the fatbin: {', '.join(arch)}.  Static instruction counts "
          "(one CTA's code, loops not multiplied).\n")
    print("Blackwell-specific evidence: packed FP32 (`FADD2` / `FFMA2` / `FMUL2`, sm_100 only) carries the float32 FFT butterflies; the "
          "magnitude rows of `k_gl_iter` arrive by TMA bulk copy (`UBLKCP` + `SYNCS` mbarrier); the float64 kernel accumulates "
          "its overlap-add with fire-and-forget `REDG.E.ADD.F64` and pulls the next pair's lines with `CCTL.E.PF2` (L2 prefetch).  "
          "There are no tensor-core instructions (`UTC*MMA`, `HMMA`): the path is FFTs and banded sums, not a GEMM "
          "(profiles/r02_mel_tf32_compare.md shows the dense tensor-core projection losing on time and accuracy).\n")
    hot = [k for k in kernels if re.search(r"k_gl_iter|k_stft_fwd|k_istft_phase|k_mags_from_mel", k)]
    cols = [c for c in KEY if any(kernels[k][c] for k in hot)]
    print("| kernel | total | " + " | ".join(cols) + " | resources |")
    print("|---|---|" + "---|" * (len(cols) + 1))
    for k in hot:
        c = kernels[k]
        name = demangle(k)
        name = re.sub(r"\(.*", "", name).replace("gomel::", "").replace("void ", "")
        print(f"| `{name}` | {sum(c.values())} | " + " | ".join(str(c[x]) for x in cols) + f" | {usage.get(k, '')} |")


if __name__ == "__main__":
    sys.exit(main())
