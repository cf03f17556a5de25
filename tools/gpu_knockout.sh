#!/bin/bash
# timing experiments with the GXB_KNOCKOUT diagnostic builds (results are garbage by construction)
# usage (under gpurun): bash tools/gpu_knockout.sh <tag> <workload> [<workload> ...]
TAG=$1; shift
for WL in "$@"; do
  for V in ko1 ko2 ko6 ko7; do
    GAUXC_B200_LIB=$PWD/gauxc_b200/libgauxc_b200_$V.so bash tools/gpu_bench_only.sh ${TAG}_$V $WL 3 2>&1 | tail -1
  done
done
