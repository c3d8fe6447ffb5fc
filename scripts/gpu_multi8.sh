#!/bin/bash
# Usage: bash scripts/gpu_multi8.sh [tag]   (under gpurun --gpus 8): sharded-driver parity on 8 ranks, the weak-scaling
# bench with the strong-scaling sharded_c4 sub-record, and config c5 at its stated size (batch 64 = 8 per GPU).
TAG=${1:-m8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_parity_main.py 2>&1 | grep -v "^\[W\|Warning\|warn" | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 300 --warmup 10 > gpurun_out/${TAG}_scale_8.json 2> gpurun_out/${TAG}_scale_8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --config c5 --n-override 8 --steps 10 --warmup 3 --e2e-steps 2 --no-extras > gpurun_out/${TAG}_c5_8.json 2> gpurun_out/${TAG}_c5_8.err
python - <<PY
import json
for f in ('gpurun_out/${TAG}_scale_8.json', 'gpurun_out/${TAG}_c5_8.json'):
  try:
    j = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, 'panos/s %.0f ms/step %.4f e2e %.0f compact e2e %.0f' % (j['value'], j['ms_per_step'], j['e2e']['value'], j['e2e']['compact_out']['value']))
    if 'sharded_c4' in j.get('extra', {}):
      print('   sharded c4:', {k: round(v['ms_per_step'], 3) for k, v in j['extra']['sharded_c4'].items() if isinstance(v, dict)})
  except Exception as e:
    print('failed', f, e, open(f.replace('.json', '.err')).read()[-1500:])
PY
