"""Opcode mix (executed warp-instructions and stall samples per SASS opcode) of one kernel in an .ncu-rep."""
import collections
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
cnt, smp = collections.Counter(), collections.Counter()
for r in rows[hi + 1:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) < len(hdr) or r[0] == "Address":
        continue
    toks = [o for o in r[idx["Source"]].strip().split() if not o.startswith("@")]
    op = toks[0].split(".")[0]
    cnt[op] += int(r[idx["Instructions Executed"]] or 0)
    smp[op] += int(r[idx["# Samples"]] or 0)
tot, ts = sum(cnt.values()), sum(smp.values())
print(f"{kern}: {tot} warp-instructions")
for op, n in cnt.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 30):
    print(f"{op:10s} {100 * n / tot:6.2f}% instr  {100 * smp[op] / max(ts, 1):6.2f}% samples")
