#!/usr/bin/env python
"""Per-kernel SASS opcode summary of libbsm_b200.so (cuobjdump -sass): what proves the Blackwell-native paths
(UBLKCP = cp.async.bulk, UTMALDG = cp.async.bulk.tensor, SYNCS = mbarrier, DMMA = FP64 tensor cores, LDGSTS = cp.async).
    python profiles/tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
LIB = ROOT / "blocksparsematrices.jl_b200" / "libbsm_b200.so"
KEYS = ["UBLKCP", "UTMALDG", "SYNCS", "DMMA", "LDGSTS", "LDS", "STS", "LDG", "STG", "DFMA", "FFMA", "SHFL", "BAR", "ATOM", "MEMBAR", "LD", "ST"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    fn, counts, total = None, {}, {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            fn = re.sub(r"\(.*", "", fn)
            counts[fn] = collections.Counter()
            total[fn] = 0
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and fn:
            counts[fn][m.group(1)] += 1
            total[fn] += 1
    print(f"{'kernel':92s} {'instr':>6s} " + " ".join(f"{k:>7s}" for k in KEYS))
    for fn in sorted(counts):
        if not fn.startswith(("void bsm::", "bsm::", "void (anonymous", "(anonymous")) and "bsm" not in fn:
            continue
        print(f"{fn[:92]:92s} {total[fn]:6d} " + " ".join(f"{counts[fn].get(k, 0):7d}" for k in KEYS))


if __name__ == "__main__":
    main()
