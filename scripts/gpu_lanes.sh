#!/bin/bash
# A/B of the concurrent chunk lanes (se3ds_ws_lanes) and of bench --streams.
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import json, sys
try:
  j = json.load(open(sys.argv[1]))
  print(sys.argv[1].split('/')[-1], 'us/step %.2f' % (j['ms_per_step'] * 1e3), 'launches', j['gpu_launches'], j['clocks']['sm_mhz'])
except Exception as e:
  print(sys.argv[1], 'parse failed', e)
PY
}
B="python bench.py --no-cpu-baseline --e2e-steps 1"
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
$B --steps 1000 --warmup 20 --lanes 1 > gpurun_out/l1.json 2> gpurun_out/l1.err; summ gpurun_out/l1.json
$B --steps 1000 --warmup 20 --lanes 2 > gpurun_out/l2s.json 2> gpurun_out/l2s.err; summ gpurun_out/l2s.json
$B --steps 1000 --warmup 20 --lanes 3 > gpurun_out/l3s.json 2> gpurun_out/l3s.err; summ gpurun_out/l3s.json
$B --steps 1000 --warmup 20 --lanes 1 --streams 2 > gpurun_out/s2l1.json 2> gpurun_out/s2l1.err; summ gpurun_out/s2l1.json
$B --steps 1000 --warmup 20 --lanes 2 --streams 2 > gpurun_out/s2l2.json 2> gpurun_out/s2l2.err; summ gpurun_out/s2l2.json
$B --steps 1000 --warmup 20 --lanes 1 --streams 3 > gpurun_out/s3l1.json 2> gpurun_out/s3l1.err; summ gpurun_out/s3l1.json
$B --steps 1000 --warmup 20 --lanes 1 --streams 2 --dist rand > gpurun_out/s2l1r.json 2> gpurun_out/s2l1r.err; summ gpurun_out/s2l1r.json
for l in 1 2; do
  $B --config c3 --steps 100 --warmup 5 --lanes $l > gpurun_out/c3_l$l.json 2> gpurun_out/c3_l$l.err; summ gpurun_out/c3_l$l.json
  $B --config c4 --steps 200 --warmup 5 --lanes $l > gpurun_out/c4_l$l.json 2> gpurun_out/c4_l$l.err; summ gpurun_out/c4_l$l.json
done
$B --config c3 --steps 100 --warmup 5 --lanes 1 --streams 2 > gpurun_out/c3_s2.json 2> gpurun_out/c3_s2.err; summ gpurun_out/c3_s2.json
