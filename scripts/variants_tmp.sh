b() { python bench.py --steps 1000 --warmup 20 --no-cpu-baseline --e2e-steps 1 "$@" 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('us/step %.2f'%(j['ms_per_step']*1e3), [(k['name'][:11], round(k['ms']*1e3,1)) for k in j['kernels']])"; }
echo "default (flip, late wait)"; b; b
echo "flip, wait first"; SE3DS_WAITFIRST=1 b; SE3DS_WAITFIRST=1 b
echo "no flip, wait first"; SE3DS_NOFLIP=1 b; SE3DS_NOFLIP=1 b
echo "no flip, wait first, no pdl"; SE3DS_NOFLIP=1 b --no-pdl
