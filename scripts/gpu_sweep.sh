#!/bin/bash
# chunk-size sweep of the c2 bench (device-resident value only)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for mb in 16 32 64 112 256; do
  for d in room rand; do
    python bench.py --steps 500 --warmup 20 --dist $d --no-cpu-baseline --e2e-steps 2 --chunk-mb $mb > gpurun_out/sweep_${mb}_${d}.json 2> gpurun_out/sweep_${mb}_${d}.err
    python - <<PY
import json
try:
  j = json.load(open('gpurun_out/sweep_${mb}_${d}.json'))
  print('chunk ${mb}MB ${d}: panos/s %.0f ms/step %.4f launches %d' % (j['value'], j['ms_per_step'], j['gpu_launches']), [(k['name'][:11], round(k['ms'] * 1e3, 1)) for k in j['kernels']])
except Exception as e:
  print('failed', e, open('gpurun_out/sweep_${mb}_${d}.err').read()[-1500:])
PY
  done
done
