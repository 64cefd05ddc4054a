#!/usr/bin/env python
"""SASS opcode evidence for the tensor-core / TMA paths: per kernel of libmaven_sm100.so, the count of the mnemonics that prove
tcgen05 (UTC*MMA), TMEM loads (LDTM), TMA (UTMALDG / UTMASTG), warp-level MMA (HMMA) and cp.async (LDGSTS).
Usage: python scripts/sass_histogram.py > profiles/<tag>_sass_histogram.md   (runs cuobjdump here, no GPU needed)"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "multimodal-supernovae_b200", "libmaven_sm100.so")
WANT = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "LDGSTS", "SYNCS", "MUFU.EX2", "FFMA", "LDS", "STS", "LDG", "STG", "BAR"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = name.replace("(anonymous namespace)::", "").replace("mvn::", "").replace("void ", "")
            name = re.sub(r"\(.*", "", name)
            cur = per.setdefault(name, collections.Counter())
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            cur["_total"] += 1
            for w in WANT:
                if op.startswith(w):
                    cur[w] += 1
    cols = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "HMMA", "MUFU.EX2", "FFMA", "LDS", "STS", "LDG", "STG", "BAR", "_total"]
    print("# SASS opcode histogram of libmaven_sm100.so (sm_100a), per kernel\n")
    print("`cuobjdump -sass multimodal-supernovae_b200/libmaven_sm100.so`, counted by scripts/sass_histogram.py.  UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld,")
    print("UTMALDG / UTMASTG = cp.async.bulk.tensor (TMA load / store), HMMA = mma.sync (warp-level tensor path).\n")
    print("| kernel | " + " | ".join(c.strip("_") for c in cols) + " |")
    print("|---|" + "---|" * len(cols))
    for name, c in per.items():
        if c["_total"] < 40:
            continue
        print(f"| `{name[:70]}` | " + " | ".join(str(c[k]) for k in cols) + " |")


if __name__ == "__main__":
    main()
