"""Top stall-sampled SASS instructions of one kernel from an .ncu-rep (needs --import-source on / -lineinfo).

    python tools/ncu_hot.py gpurun_out/prof.ncu-rep <kernel regex> [rows]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[0] == "Address" or r[0] == "Kernel Name":
        if r and r[0] == "Kernel Name":
            break            # first kernel instance only
        continue
    data.append(r)
tot = sum(int(r[idx["# Samples"]] or 0) for r in data)
execd = sum(int(r[idx["Instructions Executed"]] or 0) for r in data)
print(f"{kern}: {len(data)} SASS instructions, {execd} warp-instructions executed, {tot} stall samples")
order = sorted(range(len(data)), key=lambda i: -int(data[i][idx["# Samples"]] or 0))[:n]
for i in sorted(order):
    r = data[i]
    print(f"{i:5d} {int(r[idx['# Samples']] or 0):6d} {100 * int(r[idx['# Samples']] or 0) / max(tot, 1):5.1f}% "
          f"{r[idx['Instructions Executed']]:>9s}  {r[idx['Source']].strip()[:120]}")
