#!/bin/bash
# Round-end evidence run (one gpurun call, one GPU): bench (own arm + reference arm + the other BASELINE configs), ncu launch
# list of exactly the timed bench step (-> profiles/rNN_launches_bench.txt + source-stamped rNN_traffic.json), ncu --set full
# of the hot kernels exported to CSV ON THE BOX (the .ncu-rep files are ~9 MB per kernel and never travel), step profile.
R=${1:-r02}
mkdir -p gpurun_out
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/${R}_bench_final.json 2> gpurun_out/${R}_bench_final.err
python tools/summarize_bench.py gpurun_out/${R}_bench_final.json | head -30
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_ref.json 2> gpurun_out/${R}_bench_ref.err
head -c 300 gpurun_out/${R}_bench_ref.json; echo
timeout 300 python bench.py --workload cfg1 --steps 3 --warmup 1 > gpurun_out/${R}_bench_cfg1.json 2> /dev/null
timeout 300 python bench.py --workload na2d --steps 10 --warmup 3 > gpurun_out/${R}_bench_na2d.json 2> /dev/null
timeout 300 python bench.py --workload infer1024 --steps 10 --warmup 3 > gpurun_out/${R}_bench_infer1024.json 2> /dev/null
LMNET_NCU_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none --csv --log-file /tmp/launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-graph > gpurun_out/${R}_launches.log 2>&1
python tools/launch_list_summary.py /tmp/launches.csv gpurun_out/${R}_launches_bench.txt gpurun_out/${R}_traffic.json | head -5
for spec in "reparam 1 dw_|pixel_gemm|wgrad|bnact|se_gate 28" "reparam 3 dw_bwd_dx|dw_apply|dw_stats|dw_bwd_reduce 4" "conv 1 conv3x3 4" "natt 1 na2d_stream|ln_|pixel_gemm 12" "na 1 na2d_stream 2"; do
  set -- $spec
  timeout 600 ncu --set full --clock-control none -k regex:"$3" -c $4 -o /tmp/ncu_$1_l$2 python tools/run_block.py --unit $1 --level $2 --iters 1 > gpurun_out/${R}_ncu_$1_l$2.log 2>&1
  ncu -i /tmp/ncu_$1_l$2.ncu-rep --page raw --csv > gpurun_out/${R}_ncu_$1_l$2_raw.csv 2>/dev/null
  rm -f /tmp/ncu_$1_l$2.ncu-rep
done
timeout 300 python tools/profile_step.py --rows 130 --out gpurun_out/${R}_step_profile_final.txt | head -4 | cut -c1-200
timeout 200 python tools/bench_dw.py --iters 5 > gpurun_out/${R}_bench_dw.txt 2>&1
timeout 200 python tools/profile_ops.py --out gpurun_out/${R}_glue_ops.txt > /dev/null 2>&1
timeout 100 python tools/e2e_probe.py > gpurun_out/${R}_e2e_probe.txt 2>&1
ls -la gpurun_out | head -40
