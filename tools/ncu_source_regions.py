"""Aggregate an `ncu --page source --csv` dump (SASS view) into code regions: executed warp-instructions,
stall samples and the dominant stall reasons per region.  Regions are given as hex offsets from the kernel start.

    ncu -i rep.ncu-rep --page source --csv --kernel-name regex:NAME > src.csv
    python tools/ncu_source_regions.py src.csv 0x3e60 0x5c80 0x7d40 0x99e0 0xc890
"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
body = []
for r in rows[hdr_i + 1:]:
    if not r or not r[0].startswith("0x"):
        break          # the dump may hold further views (other kernels / source lines) after the SASS table
    if len(r) == len(hdr):
        body.append(r)
base = int(body[0][0], 16)
cuts = [int(x, 16) for x in sys.argv[2:]] + [1 << 62]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
regions = [dict(inst=0, samples=0, n=0, st=Counter(), ops=Counter()) for _ in range(len(cuts) + 1)]
for r in body:
    off = int(r[0], 16) - base
    k = next(i for i, c in enumerate(cuts) if off < c)
    R = regions[k]
    ex = int(r[col["Instructions Executed"]] or 0)
    R["inst"] += ex
    R["n"] += 1
    R["samples"] += int(r[col["# Samples"]] or 0)
    for s in stalls:
        R["st"][s] += int(r[col[s]] or 0)
    op = r[col["Source"]].split()
    op = (op[1] if op and op[0].startswith("@") else op[0] if op else "?").split(".")[0]
    R["ops"][op] += ex
tot = sum(R["inst"] for R in regions) or 1
tots = sum(R["samples"] for R in regions) or 1
lo = 0
for R, hi in zip(regions, cuts):
    if R["n"]:
        top = ", ".join(f"{k[6:]} {v}" for k, v in R["st"].most_common(4))
        ops = ", ".join(f"{k} {100 * v / max(R['inst'], 1):.0f}%" for k, v in R["ops"].most_common(6))
        print(f"[{lo:#7x}, {min(hi, 1 << 24):#9x})  static {R['n']:5d}  executed {100 * R['inst'] / tot:5.1f} %  "
              f"samples {100 * R['samples'] / tots:5.1f} %\n      stalls: {top}\n      ops: {ops}")
    lo = hi
print(f"total warp-instructions executed {tot}")
