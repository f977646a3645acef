#!/bin/bash
# Round-end evidence run (one gpurun call): bench (own arm + reference arm), ncu launch list of the bench step,
# ncu --set full of the NA streaming kernels, na2d micro-benchmark + high-res inference, torch-profiler step profile.
mkdir -p gpurun_out
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python tools/summarize_bench.py gpurun_out/bench_final.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
head -c 400 gpurun_out/bench_ref.json; echo
LMNET_NCU_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none --csv --log-file gpurun_out/launches_final.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-graph > gpurun_out/launches_final.log 2>&1
wc -l gpurun_out/launches_final.csv
for lvl in 1 4; do
  timeout 200 ncu --set full --import-source on --clock-control none -k regex:na2d_stream -c 2 -o gpurun_out/ncu_na_stream_final_l$lvl \
    python tools/run_block.py --unit na --level $lvl --iters 1 > gpurun_out/ncu_na_final_l$lvl.log 2>&1
done
timeout 300 python tools/bench_na2d.py --out gpurun_out/na2d_microbench_final.txt | tail -16
timeout 300 python tools/profile_step.py --out gpurun_out/step_profile_final.txt | head -30
