#!/bin/bash
# Final checks of a round on one GPU: full GPU test suite, smoke, randomised parity sweep, sanitizers on a subset,
# the default bench line, and config c5 at its stated batch (64) without the host-buffer leg.
TAG=${1:-final}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python tests/tools/gpu_fuzz.py 60 11 2>&1 | tail -2
SUB="fused_single_frame or ragged or trajectory or compact or rollout or planted or project_cloud_rgb"
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SUB and not 2048 and not 512" 2>&1 | grep -E "ERROR SUMMARY|passed|failed" | tail -2
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_single_frame or ragged or compact" 2>&1 | grep -E "RACECHECK SUMMARY|passed|failed" | tail -2
(time python bench.py --steps 300 --warmup 10) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -4 gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
python bench.py --config c5 --steps 5 --warmup 3 --e2e-steps 0 --no-extras --no-cpu-baseline > gpurun_out/${TAG}_c5_n64.json 2> gpurun_out/${TAG}_c5_n64.err
python - <<PY
import json
for f in ('gpurun_out/${TAG}_bench.json', 'gpurun_out/${TAG}_bench_reference.json', 'gpurun_out/${TAG}_c5_n64.json'):
  try:
    j = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, 'value %.1f ms/step %.4f' % (j['value'], j['ms_per_step']), 'e2e', j.get('e2e', {}).get('value'))
  except Exception as e:
    print('failed', f, e, open(f.replace('.json', '.err')).read()[-1200:])
PY
