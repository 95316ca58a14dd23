#!/usr/bin/env python
"""Per-kernel SASS evidence for libcnx.so: counts of the Blackwell-only instructions in every kernel
(`cuobjdump -sass`): UTCHMMA/UTCQMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st (TMEM), UTMALDG/UTMASTG = TMA tensor
load/store (cp.async.bulk.tensor), UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, FFMA2 = packed fp32x2 FMA,
plus legacy HMMA (mma.sync) which must be absent from the hot GEMMs.
usage: python profiles/sass_summary.py [lib] > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "imageclassification_b200", "lib", "libcnx.so")
OPS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "FFMA2", "FFMA", "HMMA", "LDG", "STG", "LDS", "STS"]

out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = {}
kern, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        continue
    if kern is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        counts[kern]["_n"] += 1
        for o in OPS:
            if op == o or (o not in ("FFMA", "LDG", "STG", "LDS", "STS") and op.startswith(o)):
                counts[kern][o] += 1
                break
names = list(counts)
dem = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines() if names else []
if len(dem) != len(names):
    dem = names
arch = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout
print(f"# SASS summary of {os.path.relpath(lib, ROOT)}  ({os.path.getsize(lib) / 1e6:.1f} MB; ELF: "
      f"{', '.join(sorted(set(re.findall(r'sm_\w+', arch))))}; {len(names)} kernels)")
print("# columns: instructions | " + " ".join(OPS) + " | kernel")
for k, d in zip(names, dem):
    c = counts[k]
    for o in OPS:
        total[o] += c[o]
    short = re.sub(r"\((?:int|bool|unsigned int|long|cnx::\w+(?:::\w+)*)\)", "", d)      # (int)256 -> 256 in template arguments
    short = re.sub(r"\(.*$", "", short)
    short = re.sub(r"^void ", "", short)
    print(f"{c['_n']:7d} | " + " ".join(f"{c[o]:5d}" for o in OPS) + f" | {short[:150]}")
print("# totals: " + ", ".join(f"{o}={total[o]}" for o in OPS))
tc = [d for k, d in zip(names, dem) if counts[k]["UTCHMMA"] + counts[k]["UTCQMMA"] > 0]
print(f"# kernels issuing tcgen05.mma (UTCHMMA): {len(tc)}; kernels with TMA loads (UTMALDG): "
      f"{sum(1 for k in names if counts[k]['UTMALDG'] > 0)}; kernels with legacy HMMA (mma.sync): "
      f"{sum(1 for k in names if counts[k]['HMMA'] > 0)}")
