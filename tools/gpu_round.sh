#!/bin/bash
# One GPU session: parity tests, bench lines, ncu launch list + full capture of the top kernels.
# usage (under gpurun): bash tools/gpu_round.sh <tag> [workload]
TAG=${1:-r01}
WL=${2:-taxol}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
nproc >> gpurun_out/${TAG}_gpu.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -3 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --workload benzene --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_benzene.json 2> gpurun_out/${TAG}_bench_benzene.err
cat gpurun_out/${TAG}_bench_benzene.json
timeout 900 python bench.py --workload $WL --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_${WL}.json 2> gpurun_out/${TAG}_bench_${WL}.err
cat gpurun_out/${TAG}_bench_${WL}.json; tail -3 gpurun_out/${TAG}_bench_${WL}.err
