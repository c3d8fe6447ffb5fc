#!/bin/bash
# A/B of prebuilt library variants (se3ds_b200/lib/variant_*.so copied over the library in turn).
mkdir -p gpurun_out
cp se3ds_b200/lib/libse3ds_geom.so /tmp/orig.so
for v in "$@"; do
  cp se3ds_b200/lib/variant_$v.so se3ds_b200/lib/libse3ds_geom.so
  touch se3ds_b200/lib/libse3ds_geom.so
  for cfg in "c2 1000" "c3 100"; do
    set -- $cfg
    python bench.py --config $1 --steps $2 --warmup 10 --no-cpu-baseline --e2e-steps 1 > gpurun_out/var_${v}_$1.json 2>gpurun_out/var_${v}_$1.err
    python - $v $1 <<'PY'
import json, sys
try:
  j = json.load(open('gpurun_out/var_%s_%s.json' % (sys.argv[1], sys.argv[2])))
  print(sys.argv[1], sys.argv[2], 'us/step %.2f' % (j['ms_per_step'] * 1e3), [(k['name'][:11], round(k['ms'] * 1e3, 1)) for k in j['kernels']])
except Exception as e:
  print('failed', sys.argv[1:], e)
PY
  done
done
cp /tmp/orig.so se3ds_b200/lib/libse3ds_geom.so
