#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of libgamc_b200.so (cuobjdump -sass): the instructions that prove which hardware paths the
kernels use -- DMMA (FP64 tensor cores), UTCIMMA / UTCBAR / LDTM (tcgen05.mma kind::i8, tcgen05.commit, tcgen05.ld: the integer
route of the metric), UTMALDG (TMA bulk tensor loads), SYNCS (mbarrier), LDGSTS (cp.async); no legacy HMMA / IMMA.  Usage: tools/sass_summary.py > profiles/rNN_sass_summary.md"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gamcsampler.jl_b200", "libgamc_b200.so")
WATCH = ["DMMA", "DFMA", "UTCIMMA", "UTCBAR", "LDTM", "UTCATOMSWS", "UTMALDG", "SYNCS", "LDGSTS", "UTCHMMA", "HMMA", "IMMA", "REDG", "ATOMG", "MUFU", "SHFL", "BAR", "LDG", "STG", "LDS", "STS"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"\(.*", "", o.replace("(anonymous namespace)::", "").replace("void ", "")) for o in out]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    archs = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
    kernels = OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            kernels[cur][op] += 1
            kernels[cur]["_total"] += 1
    names = demangle(list(kernels))
    print("# SASS summary of libgamc_b200.so (`cuobjdump -sass`, this build)\n")
    print(f"Architectures in the fat binary: {', '.join(archs)}.  {len(kernels)} kernels.\n")
    tot = Counter()
    for c in kernels.values():
        tot.update(c)
    print("Totals: " + ", ".join(f"{w} {tot[w]}" for w in WATCH) + f"; instructions {tot['_total']}.\n")
    print("tcgen05 (`UTCIMMA`, `UTCBAR` = tcgen05.commit, `LDTM` = tcgen05.ld, `UTCATOMSWS` = TMEM allocation) appears in `k_metric_i8` only: tcgen05 has no "
          "FP64 kind, so the FP64 tensor path is `DMMA.8x8x4` and the metric reaches the 5th-generation tensor cores through exact int8 digit products.\n")
    cols = ["DMMA", "DFMA", "UTCIMMA", "UTCBAR", "LDTM", "UTMALDG", "SYNCS", "LDGSTS", "REDG", "MUFU", "SHFL", "BAR"]
    print("| kernel | instructions | " + " | ".join(cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    for (mangled, c), name in zip(kernels.items(), names):
        print(f"| `{name[:90]}` | {c['_total']} | " + " | ".join(str(c[w]) for w in cols) + " |")


if __name__ == "__main__":
    sys.exit(main())
