#!/bin/bash
# Other BASELINE.json configs at (near) full size + sanitizer passes.
mkdir -p gpurun_out
for c in c1 c3 c4; do
  python bench.py --config $c --steps 60 --warmup 5 --no-cpu-baseline --e2e-steps 2 > gpurun_out/shape_${c}.json 2> gpurun_out/shape_${c}.err
  python - <<PY
import json
try:
  j = json.load(open('gpurun_out/shape_${c}.json'))
  print('${c}: panos/s %.0f ms/step %.4f Mpts/s %.0f step_frac %.3f e2e %.0f' % (j['value'], j['ms_per_step'], j['mpoints_per_s'], j['roofline_step']['frac_of_timed_step'], j['e2e']['value']), [(k['name'][:11], round(k['ms'] * 1e3, 1)) for k in j['kernels']])
except Exception as e:
  print('${c} failed', e, open('gpurun_out/shape_${c}.err').read()[-1500:])
PY
done
python bench.py --config c5 --n-override 4 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/shape_c5n4.json 2> gpurun_out/shape_c5n4.err
python - <<PY
import json
try:
  j = json.load(open('gpurun_out/shape_c5n4.json'))
  print('c5(N=4): panos/s %.1f ms/step %.3f Mpts/s %.0f step_frac %.3f' % (j['value'], j['ms_per_step'], j['mpoints_per_s'], j['roofline_step']['frac_of_timed_step']), [(k['name'][:11], round(k['ms'] * 1e3, 1)) for k in j['kernels']])
except Exception as e:
  print('c5 failed', e, open('gpurun_out/shape_c5n4.err').read()[-1500:])
PY
compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests -m gpu -q -x -k "tiny or chunked or single_frame or pose_sweep or project_feats or empty" 2>&1 | tail -6
compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests -m gpu -q -x -k "tiny or trajectory" 2>&1 | tail -6
