#!/bin/bash
# Quick check after a kernel change: GPU parity tests + c2 bench (room, rand, key64) + c3; summaries to stdout.
TAG=${1:-q}
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import json, sys
try:
  j = json.load(open(sys.argv[1]))
  print(sys.argv[1].split('/')[-1], 'us/step %.2f' % (j['ms_per_step'] * 1e3), [(k['name'], round(k['ms'] * 1e3, 1)) for k in j['kernels']], j['clocks']['sm_mhz'])
except Exception as e:
  print(sys.argv[1], 'parse failed', e)
  print(open(sys.argv[1].replace('.json', '.err')).read()[-1500:])
PY
}
python -m pytest tests -m gpu -q 2>&1 | tail -${2:-6}
python bench.py --steps 1000 --warmup 20 --no-cpu-baseline --e2e-steps 1 --no-extras > gpurun_out/${TAG}_room.json 2> gpurun_out/${TAG}_room.err; summ gpurun_out/${TAG}_room.json
python bench.py --steps 1000 --warmup 20 --no-cpu-baseline --e2e-steps 1 --no-extras --dist rand > gpurun_out/${TAG}_rand.json 2> gpurun_out/${TAG}_rand.err; summ gpurun_out/${TAG}_rand.json
python bench.py --steps 500 --warmup 20 --no-cpu-baseline --e2e-steps 1 --no-extras --key64 > gpurun_out/${TAG}_k64.json 2> gpurun_out/${TAG}_k64.err; summ gpurun_out/${TAG}_k64.json
python bench.py --config c3 --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 1 --no-extras > gpurun_out/${TAG}_c3.json 2> gpurun_out/${TAG}_c3.err; summ gpurun_out/${TAG}_c3.json
