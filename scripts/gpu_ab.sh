#!/bin/bash
# A/B of projection modes on the c2 bench + the fast-path tests
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "certified" 2>&1 | tail -15
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for m in 0 1; do
  for d in room rand; do
    python bench.py --steps 500 --warmup 20 --dist $d --no-cpu-baseline --e2e-steps 2 --proj-mode $m > gpurun_out/ab_${m}_${d}.json 2> gpurun_out/ab_${m}_${d}.err
    python - <<PY
import json
try:
  j = json.load(open('gpurun_out/ab_${m}_${d}.json'))
  print('proj ${m} ${d}: panos/s %.0f ms/step %.4f' % (j['value'], j['ms_per_step']), [(k['name'][:11], round(k['ms'] * 1e3, 1)) for k in j['kernels']])
except Exception as e:
  print('failed', e, open('gpurun_out/ab_${m}_${d}.err').read()[-1500:])
PY
  done
done
