#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page source --csv` (several kernels concatenated) into, per kernel: the warp-state sample
histogram and the N SASS lines with the most samples.   usage: python profiles/ncu_source_stalls.py source.csv [N]"""
import csv
import sys
from collections import Counter

path = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 10
kern, hdr, rows, idx = None, None, [], 0


def flush():
    global rows, idx
    if kern is None or not rows:
        return
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = Counter()
    for r in rows:
        for i in stall_cols:
            try:
                tot[hdr[i][6:]] += int(r[i])
            except (ValueError, IndexError):
                pass
    n = sum(tot.values())
    print(f"== {kern[:110]} #{idx}")
    print(f"   warp-state samples {n}: " + ", ".join(f"{k} {v}" for k, v in tot.most_common(9)))
    si = hdr.index("# Samples")
    ei = hdr.index("Instructions Executed")
    def samples(r):
        try:
            return int(r[si])
        except (ValueError, IndexError):
            return 0
    for r in sorted(rows, key=samples, reverse=True)[:topn]:
        st = sorted(((int(r[i]) if r[i].isdigit() else 0, hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
        print(f"   {samples(r):7d} samples {r[ei]:>9s} exec  {r[1].strip()[:84]:84s} {[(b, a) for a, b in st]}")
    rows = []
    idx += 1


for r in csv.reader(open(path, newline="")):
    if not r:
        continue
    if r[0] == "Kernel Name":
        flush()
        kern = r[1]
        hdr = None
    elif r[0] == "Address":
        hdr = r
    elif hdr is not None:
        rows.append(r)
flush()
