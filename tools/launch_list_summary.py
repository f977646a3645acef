"""Summarise an ncu launch-list CSV (gpu__time_duration + dram bytes per launch) into profiles/:
    python tools/launch_list_summary.py gpurun_out/launches_r1.csv profiles/r01_launches_bench.txt profiles/r01_traffic.json
"""
import collections
import csv
import json
import sys

src, out_txt, out_json = sys.argv[1:4]
rows = [r for r in csv.reader(open(src, errors="ignore")) if len(r) > 8]
idx = {h: i for i, h in enumerate(rows[0])}
t = collections.defaultdict(lambda: {"ms": 0.0, "n": 0, "rd": 0.0, "wr": 0.0})
for r in rows[1:]:
    name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
    m, u = r[idx["Metric Name"]], r[idx["Metric Unit"]]
    try:
        v = float(r[idx["Metric Value"]].replace(",", ""))
    except ValueError:
        continue
    if m == "gpu__time_duration.sum":
        t[name]["ms"] += v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
        t[name]["n"] += 1
    else:
        t[name]["rd" if "read" in m else "wr"] += v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
total = sum(v["ms"] for v in t.values())
own = {k: v for k, v in t.items() if k.startswith("lmnet::")}
own_ms = sum(v["ms"] for v in own.values())
lines = ["# ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none",
         "#   LMNET_NCU_RANGE=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-graph",
         "#   (the profiled range is exactly the timed step; cold-cache, serialised launches: compare SHARES, not absolutes)",
         f"# {sum(v['n'] for v in t.values())} launches, {total:.1f} ms total; lmnet_b200 kernels {own_ms:.1f} ms = {100 * own_ms / total:.1f} % of device time",
         f"{'ms':>9} {'share%':>7} {'launches':>8} {'DRAM MB/launch':>15}  kernel"]
for k, v in sorted(t.items(), key=lambda kv: -kv[1]["ms"])[:60]:
    lines.append(f"{v['ms']:9.3f} {100 * v['ms'] / total:7.2f} {v['n']:8d} {(v['rd'] + v['wr']) / max(v['n'], 1) / 1e6:15.2f}  {k[:120]}")
open(out_txt, "w").write("\n".join(lines) + "\n")
merged = collections.defaultdict(lambda: {"bytes": 0.0, "n": 0, "ms": 0.0})
for k, v in own.items():
    base = k.replace("lmnet::", "").split("<")[0]
    for suffix in ("_tma_kernel", "_mma_kernel", "_fast_kernel", "_cl_kernel", "_kernel"):
        if base.endswith(suffix):
            base = base[:-len(suffix)]
            break
    if base.endswith("_cl"):
        base = base[:-3]
    base = {"conv3x3_fwd": "conv3x3", "upsample2x_cl_fwd": "upsample2x_fwd", "upsample2x_cl_bwd": "upsample2x_bwd"}.get(base, base)
    base = {"bnact_stats": "bn_stats", "bnact_apply": "bn_apply", "bnact_bwd_reduce": "bn_bwd_reduce", "bnact_bwd_apply": "bn_bwd_apply",
            "bnact_fin_fwd": "bn_fin_fwd", "bnact_fin_bwd": "bn_fin_bwd", "drpb_reduce": "na2d_drpb_reduce"}.get(base, base)
    merged[base]["bytes"] += v["rd"] + v["wr"]
    merged[base]["n"] += v["n"]
    merged[base]["ms"] += v["ms"]
import hashlib, glob, os
_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_h = hashlib.sha256()
for _f in sorted(glob.glob(os.path.join(_root, "lm-net_b200", "csrc", "*.cu")) + glob.glob(os.path.join(_root, "lm-net_b200", "csrc", "*.cuh"))):
    _h.update(os.path.basename(_f).encode())
    _h.update(open(_f, "rb").read())
_out = {"_csrc_sha256": _h.hexdigest()[:16]}       # bench.py refuses this file once the kernels change
_out.update({k: {"dram_bytes_per_launch": v["bytes"] / v["n"], "launches_per_step": v["n"], "ncu_ms_per_step": v["ms"],
               "share_of_step": v["ms"] / total} for k, v in sorted(merged.items())})
json.dump(_out, open(out_json, "w"), indent=1)
print("\n".join(lines[:34]))
