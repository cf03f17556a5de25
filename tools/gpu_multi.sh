#!/bin/bash
# multi-GPU bench lines: bash tools/gpu_multi.sh <tag> <workload> <N> [<N> ...]   (under gpurun --gpus >= max N)
TAG=$1; WL=$2; shift 2
mkdir -p gpurun_out
for N in "$@"; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
      bench.py --gpus $N --workload $WL --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_${WL}_${N}gpu.json 2> gpurun_out/${TAG}_bench_${WL}_${N}gpu.err
  echo "N=$N rc=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/${TAG}_bench_${WL}_${N}gpu.json').read().strip().splitlines() if l.startswith('{')][-1])
    r=d['roofline']
    print('${WL} N=$N', 'ms/step %.2f'%d['ms_per_step'], 'e2e %.2f'%d['e2e']['ms_per_step'], 'Mpts/s %.1f'%(d['value']/1e6), {k:round(v,2) for k,v in r['kernel_ms_per_step'].items()}, 'exc', d['exc'], 'nel', d['n_el'], 'ssf %.1f'%d['ssf_weights_ms'])
except Exception as e:
    print('${WL} N=$N bench failed', e)
PY
  tail -2 gpurun_out/${TAG}_bench_${WL}_${N}gpu.err
done
