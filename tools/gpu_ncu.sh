#!/bin/bash
# ncu evidence: launch list of one bench step + full captures of the top kernels.
# usage (under gpurun): bash tools/gpu_ncu.sh <tag> <workload> [kernel-regex ...]
TAG=${1:-r01}; WL=${2:-taxol}; shift 2
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${TAG}_launches_${WL}.csv \
    python bench.py --workload $WL --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch_${WL}.log 2>&1
for K in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/${TAG}_${WL}_${K} \
      python bench.py --workload $WL --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_${K}.log 2>&1
  echo "ncu $K exit $?"
done
