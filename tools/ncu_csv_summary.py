"""Summarise an `ncu --page raw --csv` export (made on the GPU box, so that the heavy .ncu-rep never travels).
    python tools/ncu_csv_summary.py gpurun_out/xxx_raw.csv > profiles/rNN_ncu_xxx.txt"""
import csv
import sys

KEYS = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
        ("launch__occupancy_limit_registers", "occ_lim_regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active%"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__t_sector_hit_rate.pct", "l2_hit%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma_pipe%"),
        ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "hmma%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
        ("smsp__inst_executed.sum", "warp_insts")]
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
print(f"# {sys.argv[1]}: ncu --set full --clock-control none (cold-cache, serialised launches)")
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    print("\n" + r[idx["Kernel Name"]][:150])
    for k, short in KEYS:
        if k in idx:
            print(f"  {short:16s} {r[idx[k]]} {units[idx[k]]}")
    stalls = []
    for h, i in idx.items():
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            if v > 0:
                stalls.append((v, h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
    stalls.sort(reverse=True)
    print("  stalls/issue     " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:7]))
