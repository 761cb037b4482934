#!/usr/bin/env python
"""Prints the SASS instructions with the most warp-stall samples from `ncu --page source --csv` output."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for k, r in enumerate(rows[2:]):
    try:
        data.append((int(r[ci["# Samples"]]), k, r))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
for s, k, r in sorted(data, key=lambda t: -t[0])[:n]:
    top = sorted(((int(r[ci[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
    print(f"{s:6d} {100 * s / tot:5.1f}%  #{k:4d} {r[ci['Source']].strip()[:70]:70s} {top}")
