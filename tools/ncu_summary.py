"""Summarise an .ncu-rep (full set) or an ncu launch-list CSV into a small text table for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep  > profiles/rNN_ncu_xxx.txt
    python tools/ncu_summary.py --launches gpurun_out/launches.csv > profiles/rNN_launches_xxx.txt
"""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active%"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma_pipe%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
]


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {path}: ncu --set full --clock-control none (cold-cache, serialised launches)")
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0]
        print(f"\n{name}")
        for k, short in KEYS:
            if k in idx:
                print(f"  {short:16s} {r[idx[k]]} {units[idx[k]]}")


def launches(path):
    rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 5]
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    tot = defaultdict(lambda: [0.0, 0])
    for r in rows[1:]:
        try:
            v = float(r[idx["Metric Value"]].replace(",", ""))
        except (ValueError, KeyError):
            continue
        unit = r[idx["Metric Unit"]]
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        name = r[idx["Kernel Name"]].split("(")[0][:90]
        tot[name][0] += v * scale
        tot[name][1] += 1
    total = sum(v[0] for v in tot.values())
    own = sum(v[0] for k, v in tot.items() if "lmnet" in k or k.startswith(("dw_", "na2d_", "void dw_", "void na2d_", "bnact", "void bnact", "void ln_", "ln_")))
    print(f"# {path}: ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)")
    print(f"# total {total:.2f} ms over {sum(v[1] for v in tot.values())} launches; lmnet_b200 kernels {own:.2f} ms = {100 * own / max(total, 1e-9):.1f} %")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0])[:40]:
        print(f"{v[0]:10.3f} ms {100 * v[0] / total:6.2f} % {v[1]:6d}  {k}")


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[1])
