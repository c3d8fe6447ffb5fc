#!/bin/bash
# Usage: bash scripts/gpu_multi8b.sh [tag]   (under gpurun --gpus 8): sharded-driver parity on 8 ranks (NCCL and
# multicast wires), the strong-scaling sharded c4 record alone, then the full weak-scaling bench line.
TAG=${1:-m8b}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "compact" 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_parity_main.py 2>&1 | grep -v "^\[W\|Warning\|warn\|^\*\*\*\|OMP_NUM\|^$" | tail -6
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 scripts/sharded_probe.py 2>&1 | grep -v "^\[W\|Warning\|warn\|^\*\*\*\|OMP_NUM\|^$" | tail -4 | tee gpurun_out/${TAG}_sharded.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 300 --warmup 10 > gpurun_out/${TAG}_scale_8.json 2> gpurun_out/${TAG}_scale_8.err
python - <<PY
import json
for f in ('gpurun_out/${TAG}_scale_8.json',):
  try:
    j = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, 'panos/s %.0f ms/step %.4f e2e %.0f (blocking %.0f) compact e2e %.0f (blocking %.0f)' % (j['value'], j['ms_per_step'], j['e2e']['value'], j['e2e']['blocking_call']['value'], j['e2e']['compact_out']['value'], j['e2e']['compact_out']['blocking_call']['value']))
    print('   ceiling', j['e2e']['host_link_ceiling_gbs'], j['e2e']['numa'])
    if 'sharded_c4' in j.get('extra', {}):
      print('   sharded c4:', {k: (round(v['ms_per_step'], 3) if 'ms_per_step' in v else v) for k, v in j['extra']['sharded_c4'].items() if isinstance(v, dict)})
  except Exception as e:
    print('failed', f, e, open(f.replace('.json', '.err')).read()[-1500:])
PY
