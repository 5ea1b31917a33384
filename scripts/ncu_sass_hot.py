#!/usr/bin/env python3
"""Per-SASS-instruction executed counts from `ncu --page source --csv`, grouped into regions.
usage: ncu -i rep --page source --csv > src.csv; scripts/ncu_sass_hot.py src.csv [warp_steps]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, iex, iwf, ism = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("L1 Wavefronts Shared"), hdr.index("# Samples")
ws = float(sys.argv[2]) if len(sys.argv) > 2 else None
tot = 0
body = []
for r in rows[2:]:
    if len(r) <= iex: continue
    ex = int(r[iex] or 0); tot += ex
    body.append((r[isrc].strip(), ex, int(r[iwf] or 0), int(r[ism] or 0)))
print("total warp-instructions", tot, "per warp-step", tot / ws if ws else "")
# histogram by opcode weighted by exec count
from collections import Counter
c = Counter(); w = Counter(); smp = Counter()
for src, ex, wf, sm in body:
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    c[op] += ex; w[op] += wf; smp[op] += sm
for op, ex in c.most_common(40):
    print(f"{op:28s} {ex:14d} {ex / ws if ws else 0:8.2f}/step  wavefronts {w[op] / ws if ws else 0:6.2f}/step  samples {smp[op]}")
if len(sys.argv) > 3:
    for i, (src, ex, wf, sm) in enumerate(body):
        print(f"{i:5d} {ex:12d} {wf:12d} {sm:6d}  {src}")
