#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for k in "" "--key64"; do
  for d in room rand; do
    python bench.py --steps 500 --warmup 20 --dist $d --no-cpu-baseline --e2e-steps 2 $k > gpurun_out/key_${d}${k}.json 2> gpurun_out/key_${d}${k}.err
    python - <<PY
import json
try:
  j = json.load(open('gpurun_out/key_${d}${k}.json'))
  print('key "${k}" ${d}: panos/s %.0f ms/step %.4f' % (j['value'], j['ms_per_step']), [(x['name'][:11], round(x['ms'] * 1e3, 1)) for x in j['kernels']])
except Exception as e:
  print('failed', e, open('gpurun_out/key_${d}${k}.err').read()[-1500:])
PY
  done
done
