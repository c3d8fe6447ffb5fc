#!/bin/bash
# Usage (on the GPU box, via gpurun): bash scripts/gpu_check.sh <tag> [tests|notests] [ncu|noncu]
# Runs the GPU parity tests, smoke(), the bench (both depth distributions), and optionally the
# ncu launch list + one --set full capture; everything lands in gpurun_out/.
TAG=${1:-run}; TESTS=${2:-tests}; NCU=${3:-ncu}
mkdir -p gpurun_out
if [ "$TESTS" = tests ]; then
  python -m pytest tests -m gpu -x -q 2>&1 | tail -25
  python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
fi
python bench.py --steps 500 --warmup 20 > gpurun_out/bench_${TAG}_room.json 2> gpurun_out/bench_${TAG}_room.err
python - <<PY
import json
for d in ('room',):
  try:
    j = json.load(open('gpurun_out/bench_${TAG}_%s.json' % d))
    print(d, 'panos/s %.0f  ms/step %.4f  Mpts/s %.0f  e2e %.0f panos/s  step_frac %.3f' % (j['value'], j['ms_per_step'], j['mpoints_per_s'], j['e2e']['value'], j['roofline_step']['frac_of_timed_step']))
    print('  kernels', [(k['name'], round(k['ms'] * 1e3, 1)) for k in j['kernels']], 'clocks', j['clocks'])
    print('  cpu', j.get('cpu_baseline', {}).get('value'))
  except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench_${TAG}_%s.err' % d).read()[-2000:])
PY
python bench.py --steps 500 --warmup 20 --dist rand --no-cpu-baseline > gpurun_out/bench_${TAG}_rand.json 2> gpurun_out/bench_${TAG}_rand.err
python - <<PY
import json
try:
  j = json.load(open('gpurun_out/bench_${TAG}_rand.json'))
  print('rand panos/s %.0f  ms/step %.4f' % (j['value'], j['ms_per_step']), [(k['name'], round(k['ms'] * 1e3, 1)) for k in j['kernels']])
except Exception as e:
  print('bench parse failed', e); print(open('gpurun_out/bench_${TAG}_rand.err').read()[-2000:])
PY
if [ "$NCU" = ncu ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 12 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launch_${TAG}.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:"splat_depth|splat_feat|resolve" -s 12 -c 3 -o gpurun_out/prof_${TAG} python bench.py --steps 12 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full_${TAG}.log 2>&1
fi
ls gpurun_out | head -50
