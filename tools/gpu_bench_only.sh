#!/bin/bash
# usage: bash tools/gpu_bench_only.sh <tag> <workload> [steps]   (env passes through)
TAG=$1; WL=$2; ST=${3:-5}
mkdir -p gpurun_out
timeout 900 python bench.py --workload $WL --steps $ST --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_${WL}.json 2> gpurun_out/${TAG}_bench_${WL}.err
tail -2 gpurun_out/${TAG}_bench_${WL}.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench_${WL}.json').read().strip().splitlines()[-1])
    r=d['roofline']
    print('${TAG} ${WL}', 'ms/step %.2f'%d['ms_per_step'], 'e2e %.2f'%d['e2e']['ms_per_step'], {k:round(v,2) for k,v in r['kernel_ms_per_step'].items()}, 'frac %.3f'%r['frac'], 'exc', d['exc'], 'nel', d['n_el'], 'ssf %.1f'%d['ssf_weights_ms'])
except Exception as e:
    print('${WL} bench failed', e)
PY
