#!/usr/bin/env python
"""Per-kernel share of a step from an ncu launch list (`--metrics gpu__time_duration.sum --csv`):
usage: python profiles/launch_share.py LAUNCHES.csv [> OUT.txt]
Groups launches by kernel name (template arguments kept, parameter list dropped); prints count, total us and share.
ncu serialises launches and runs them cold-cache, so only the SHARES are comparable with bench.py's in-step table."""
import csv
import re
import sys
from collections import defaultdict


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) >= 15 and r[0].isdigit()]
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        name = re.sub(r"\(.*$", "", r[4]).replace("void ", "").replace("cnx::", "")
        name = re.sub(r"CUtensorMap_st", "TM", name)
        ns = float(r[14]) * ({"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r[13], 1.0))
        # ncu prints kernel names without their outermost namespace (cnx:: / at::): library kernels are recognised by theirs
        ours = re.match(r"^(native::|cuda::|cub::|at::|nvjet|nccl|cutlass|cublas|cudnn|sm\d+_)", name) is None
        a = agg[("cnx " if ours else "lib ") + name[:100]]
        a[0] += 1
        a[1] += ns / 1e3
    tot = sum(v[1] for v in agg.values())
    print(f"{len(rows)} launches, {tot / 1e3:.3f} ms total (serialised, cold cache)")
    ours = sum(v[1] for k, v in agg.items() if k.startswith("cnx "))
    print(f"libcnx kernels: {ours / tot:.1%} of the time; library (ATen/NCCL) kernels: {1 - ours / tot:.1%}")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{us / tot:7.2%} {us:10.1f} us {n:5d}x  {k}")


if __name__ == "__main__":
    main()
