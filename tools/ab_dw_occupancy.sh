#!/bin/bash
# A/B of the depthwise tensor-core kernels' occupancy knobs: __launch_bounds__ min-blocks (variant libraries
# built with -DLMNET_DW_MINBLOCKS=n) x persistent CTA target (LMNET_DW_CTAS).  Prints the per-kernel table.
mkdir -p gpurun_out
for cfg in "default 592" "mb5 740" "mb5 592" "mb6 888" "default 888"; do
  set -- $cfg
  lib=""; [ "$1" != default ] && lib="$PWD/lm-net_b200/lmnet_b200/liblmnet_b200_$1.so"
  echo "=== lib=$1 ctas=$2"
  LMNET_B200_LIB=$lib LMNET_DW_CTAS=$2 timeout 300 python bench.py --steps 15 --warmup 5 --no-cpu-baseline 2>/dev/null \
    | tee gpurun_out/ab_$1_$2.json | python tools/summarize_bench.py | grep -E "value|dw_"
done
