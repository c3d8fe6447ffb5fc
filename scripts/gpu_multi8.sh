#!/bin/bash
# 8-GPU run: sharded-driver parity, weak scaling of c2 at 1/2/4/8, strong scaling of the c4 sweep
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_parity_main.py 2>&1 | grep "DIST PARITY"
for g in 1 2 4 8; do
  if [ $g = 1 ]; then
    python bench.py --gpus 1 --steps 1000 --warmup 20 --no-cpu-baseline > gpurun_out/scale_${g}.json 2> gpurun_out/scale_${g}.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 2951$g bench.py --gpus $g --steps 1000 --warmup 20 > gpurun_out/scale_${g}.json 2> gpurun_out/scale_${g}.err
  fi
  python - <<PY
import json
try:
  j = json.loads(open('gpurun_out/scale_${g}.json').read().strip().splitlines()[-1])
  print('weak c2 gpus ${g}: panos/s %.0f ms/step %.4f e2e %.0f' % (j['value'], j['ms_per_step'], j['e2e']['value']))
except Exception as e:
  print('failed', e, open('gpurun_out/scale_${g}.err').read()[-1500:])
PY
done
for g in 1 2 4 8; do
  if [ $g = 1 ]; then
    python bench.py --gpus 1 --config c4 --sharded --steps 50 --warmup 5 > gpurun_out/shard_${g}.json 2> gpurun_out/shard_${g}.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 2952$g bench.py --gpus $g --config c4 --sharded --steps 50 --warmup 5 > gpurun_out/shard_${g}.json 2> gpurun_out/shard_${g}.err
  fi
  python - <<PY
import json
try:
  j = json.loads(open('gpurun_out/shard_${g}.json').read().strip().splitlines()[-1])
  print('strong c4 (64 poses, all-gather) gpus ${g}: panos/s %.0f ms/step %.4f' % (j['value'], j['ms_per_step']))
except Exception as e:
  print('failed', e, open('gpurun_out/shard_${g}.err').read()[-1500:])
PY
done
