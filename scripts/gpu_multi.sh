#!/bin/bash
# Usage: bash scripts/gpu_multi.sh <ngpus>   (under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/dist_parity_main.py 2>&1 | grep -v "^\[W\|Warning\|warn" | tail -8
for g in 1 $N; do
  if [ $g = 1 ]; then
    python bench.py --gpus 1 --steps 500 --warmup 20 --no-cpu-baseline > gpurun_out/scale_${g}.json 2> gpurun_out/scale_${g}.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $g --steps 500 --warmup 20 > gpurun_out/scale_${g}.json 2> gpurun_out/scale_${g}.err
  fi
  python - <<PY
import json
try:
  j = json.loads(open('gpurun_out/scale_${g}.json').read().strip().splitlines()[-1])
  print('gpus ${g}: panos/s %.0f ms/step %.4f e2e %.0f' % (j['value'], j['ms_per_step'], j['e2e']['value']))
except Exception as e:
  print('failed', e, open('gpurun_out/scale_${g}.err').read()[-2000:])
PY
done
