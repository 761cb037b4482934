#!/usr/bin/env python
"""Collects the bench / probe JSON lines of a round from gpurun_out/ into profiles/<round>_bench_lines.jsonl
(every line tagged with `_run` = the file it came from).   python profiles/tools/collect_lines.py r02"""
import glob
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
out = []
for f in sorted(glob.glob(str(ROOT / "gpurun_out" / f"{tag}*.json"))):
    if f.endswith(".last_call.json"):
        continue
    for ln in open(f).read().splitlines():
        ln = ln.strip()
        if not ln.startswith("{"):
            continue
        try:
            d = json.loads(ln)
        except ValueError:
            continue
        d["_run"] = Path(f).stem
        out.append(d)
with open(ROOT / "profiles" / f"{tag}_bench_lines.jsonl", "w") as fh:
    for d in out:
        fh.write(json.dumps(d) + "\n")
print(len(out), "lines")
