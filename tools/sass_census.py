#!/usr/bin/env python
"""tools/sass_census.py — what the compiled kernels of libct2b200.so are made of: per kernel family, the number of
instantiations, the largest instantiation (SASS instructions) and how many tcgen05 / TMA / tensor-memory / mbarrier /
mma.sync / cp.async / cluster-barrier instructions the instantiations contain (cuobjdump -sass; the mnemonics that prove
tcgen05 and TMA are listed in B200_PROFILING.md).  Runs without a GPU.

    python tools/sass_census.py > profiles/r02_sass_census.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ctranslate2_b200", "libct2b200.so")
COLS = [("tcgen05.mma (UTC*MMA)", r"\bUTC[A-Z0-9]*MMA"), ("TMA load (UTMALDG)", r"\bUTMALDG"), ("TMA prefetch (UTMAPF)", r"\bUTMAPF"),
        ("bulk copy (UBLKCP)", r"\bUBLKCP"), ("tcgen05.ld/st (LDTM/STTM)", r"\b(?:LDTM|STTM)"), ("mbarrier (SYNCS)", r"\bSYNCS"),
        ("mma.sync (HMMA/IMMA)", r"\b(?:HMMA|IMMA)"), ("cp.async (LDGSTS)", r"\bLDGSTS"), ("cluster barrier (UCGABAR)", r"\bUCGABAR")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    parts = re.split(r"\n\s*Function : ", sass)[1:]
    names = [p.split("\n", 1)[0].strip() for p in parts]
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    agg = collections.OrderedDict()
    for body, d in zip(parts, dem):
        m = re.search(r"(\w+_kernel)", d)
        key = m.group(1) if m else d[:48]
        a = agg.setdefault(key, [0, 0] + [0] * len(COLS))
        a[0] += 1
        a[1] = max(a[1], len(re.findall(r"^\s+/\*[0-9a-f]+\*/\s+\S", body, re.M)))
        for i, (_, pat) in enumerate(COLS):
            a[2 + i] += len(re.findall(pat, body))
    print("# SASS census of `ctranslate2_b200/libct2b200.so` (sm_100a), produced by `tools/sass_census.py`\n")
    print("Counts are summed over the template instantiations of a kernel family; 0 = the family does not use that unit.\n")
    print("| kernel | instantiations | largest (SASS instr.) | " + " | ".join(c for c, _ in COLS) + " |")
    print("|---|---|---|" + "---|" * len(COLS))
    for k, a in agg.items():
        print("| `%s` | %d | %d | %s |" % (k, a[0], a[1], " | ".join(str(x) for x in a[2:])))


if __name__ == "__main__":
    sys.exit(main())
