"""Pretty-print a bench.py JSON line (stdin or file): headline numbers, roofline and per-kernel table."""
import json
import sys

src = open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin
for ln in src:
    ln = ln.strip()
    if not ln.startswith("{"):
        continue
    d = json.loads(ln)
    print(f"value {d.get('value')} {d.get('unit')}  ms/step {d.get('ms_per_step')}  e2e {d.get('e2e', {}).get('value')}  "
          f"launches {d.get('gpu_launches')}  clocks {d.get('clocks')}")
    print("roofline:", d.get("roofline"))
    print("cpu_baseline:", d.get("cpu_baseline"))
    for k, v in (d.get("kernels") or {}).items():
        gbs = v["achieved_GBs"]
        print(f"  {k:22s} {v['ms_per_step']:8.3f} ms/step  {v['launches_per_step']:3d} launches  "
              f"{'' if gbs is None else f'{gbs:8.1f} GB/s algorithmic'}")
