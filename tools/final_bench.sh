#!/bin/bash
# Last run of a round (one gpurun call, one GPU): the bench line at default flags, the reference arm, and the ncu launch list of
# exactly the timed step with its source stamp (bench.py refuses a traffic file whose stamp does not match the tree), so
# that the committed numbers belong to the committed code.  The heavy ncu --set full captures are tools/final_evidence.sh.
R=${1:-r02}
mkdir -p gpurun_out
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/${R}_bench_final.json 2> gpurun_out/${R}_bench_final.err
python tools/summarize_bench.py gpurun_out/${R}_bench_final.json | head -12
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_ref.json 2> gpurun_out/${R}_bench_ref.err
head -c 300 gpurun_out/${R}_bench_ref.json; echo
LMNET_NCU_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none --csv --log-file /tmp/launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile --no-graph > gpurun_out/${R}_launches.log 2>&1
python tools/launch_list_summary.py /tmp/launches.csv gpurun_out/${R}_launches_bench.txt gpurun_out/${R}_traffic.json | head -5
timeout 300 python tools/profile_step.py --rows 130 --out gpurun_out/${R}_step_profile_final.txt | head -3 | cut -c1-200
timeout 200 python tools/bench_dw.py --iters 5 > gpurun_out/${R}_bench_dw.txt 2>&1; tail -1 gpurun_out/${R}_bench_dw.txt
