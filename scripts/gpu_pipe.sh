#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for pc in 0 2 3 4 8; do
  for d in room rand; do
    python bench.py --steps 500 --warmup 20 --dist $d --no-cpu-baseline --e2e-steps 2 --pipeline $pc > gpurun_out/pipe_${pc}_${d}.json 2> gpurun_out/pipe_${pc}_${d}.err
    python - <<PY
import json
try:
  j = json.load(open('gpurun_out/pipe_${pc}_${d}.json'))
  print('pipeline ${pc} ${d}: panos/s %.0f ms/step %.4f launches/step %.1f step_frac %.3f' % (j['value'], j['ms_per_step'], j['gpu_launches'] / j['steps'], j['roofline_step']['frac_of_timed_step']))
except Exception as e:
  print('failed', e, open('gpurun_out/pipe_${pc}_${d}.err').read()[-1500:])
PY
  done
done
