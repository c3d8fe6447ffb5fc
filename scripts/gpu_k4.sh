#!/bin/bash
# A/B of the resolve (K4) variants: SE3DS_K4 = 0 (one quad per thread), 1 (streaming grid, prefetch),
# 2 (1 + RGB stores staged through shared memory); SE3DS_K4_BLOCKS = resident blocks per SM of the grid.
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import json, sys
try:
  j = json.load(open(sys.argv[1]))
  print(sys.argv[1].split('/')[-1], 'us/step %.2f' % (j['ms_per_step'] * 1e3), [(k['name'], round(k['ms'] * 1e3, 1)) for k in j['kernels']], j['clocks']['sm_mhz'])
except Exception as e:
  print(sys.argv[1], 'parse failed', e)
PY
}
SE3DS_K4=3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
SE3DS_K4=2 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for cfg in "0 8" "3 8" "2 12" "2 16" "2 24" "3 8" "0 8"; do
  set -- $cfg
  SE3DS_K4=$1 SE3DS_K4_BLOCKS=$2 python bench.py --steps 1000 --warmup 20 --no-cpu-baseline --e2e-steps 1 > gpurun_out/k4_$1_$2.json 2> gpurun_out/k4_$1_$2.err
  summ gpurun_out/k4_$1_$2.json
done
for cfg in "0 8" "3 8" "2 16"; do
  set -- $cfg
  SE3DS_K4=$1 SE3DS_K4_BLOCKS=$2 python bench.py --config c3 --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 1 > gpurun_out/k4c3_$1_$2.json 2> gpurun_out/k4c3_$1_$2.err
  summ gpurun_out/k4c3_$1_$2.json
  SE3DS_K4=$1 SE3DS_K4_BLOCKS=$2 python bench.py --key64 --steps 500 --warmup 20 --no-cpu-baseline --e2e-steps 1 > gpurun_out/k4k64_$1_$2.json 2> gpurun_out/k4k64_$1_$2.err
  summ gpurun_out/k4k64_$1_$2.json
done
