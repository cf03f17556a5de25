#!/bin/bash
# full ncu captures of one kernel at given launch-skip counts
# usage (under gpurun): bash tools/gpu_ncu2.sh <tag> <workload> <kernel-regex> <skip> [<skip> ...]
TAG=$1; WL=$2; K=$3; shift 3
mkdir -p gpurun_out
for S in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -f -o gpurun_out/${TAG}_${WL}_${K}_s${S} \
      python bench.py --workload $WL --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_${K}_s${S}.log 2>&1
  echo "ncu $K skip $S exit $?"
done
