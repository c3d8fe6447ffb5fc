#!/bin/bash
# Usage: bash scripts/gpu_multi.sh <ngpus> [tag]   (under gpurun --gpus N): sharded-driver parity on N ranks, then
# the weak-scaling bench (with the strong-scaling sharded_c4 sub-record) at 1 and N GPUs.
N=${1:-2}; TAG=${2:-multi}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/dist_parity_main.py 2>&1 | grep -v "^\[W\|Warning\|warn" | tail -8
for g in 1 $N; do
  if [ $g = 1 ]; then
    python bench.py --gpus 1 --steps 300 --warmup 10 --no-cpu-baseline > gpurun_out/${TAG}_scale_${g}.json 2> gpurun_out/${TAG}_scale_${g}.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $g --steps 300 --warmup 10 > gpurun_out/${TAG}_scale_${g}.json 2> gpurun_out/${TAG}_scale_${g}.err
  fi
  python - <<PY
import json
try:
  j = json.loads(open('gpurun_out/${TAG}_scale_${g}.json').read().strip().splitlines()[-1])
  print('gpus ${g}: panos/s %.0f ms/step %.4f e2e %.0f compact e2e %.0f' % (j['value'], j['ms_per_step'], j['e2e']['value'], j['e2e']['compact_out']['value']))
  if 'sharded_c4' in j['extra']:
    print('   sharded c4:', {k: (round(v['ms_per_step'], 3) if 'ms_per_step' in v else v) for k, v in j['extra']['sharded_c4'].items() if isinstance(v, dict)})
  print('   c4 one GPU ms', j['extra'].get('c4', {}).get('ms_per_step'))
except Exception as e:
  print('failed', e, open('gpurun_out/${TAG}_scale_${g}.err').read()[-2500:])
PY
done
