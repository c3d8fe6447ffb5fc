#!/bin/bash
# Last run of a round on a tight GPU budget: parity first (stop if it fails), then the ncu evidence and the bench
# line of the same build.
TAG=${1:-last}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -2 | tee gpurun_out/${TAG}_pytest.txt
grep -q "failed\|error" gpurun_out/${TAG}_pytest.txt && exit 1
python tests/tools/gpu_fuzz.py 40 13 2>&1 | tail -1 | tee gpurun_out/${TAG}_fuzz.txt
grep -q "FUZZ OK" gpurun_out/${TAG}_fuzz.txt || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 12 --warmup 3 --no-cpu-baseline --e2e-steps 0 --no-extras > gpurun_out/ncu_launch_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"splat_depth|splat_feat|resolve" -s 12 -c 3 -o gpurun_out/prof_${TAG} python bench.py --steps 12 --warmup 3 --no-cpu-baseline --e2e-steps 0 --no-extras > gpurun_out/ncu_full_${TAG}.log 2>&1
ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,lts__t_sector_hit_rate.pct -k regex:"splat_depth|splat_feat|resolve" -s 12 -c 6 --csv --log-file gpurun_out/warm_${TAG}.csv python bench.py --steps 12 --warmup 3 --no-cpu-baseline --e2e-steps 0 --no-extras > gpurun_out/ncu_warm_${TAG}.log 2>&1
python bench.py --steps 300 --warmup 10 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -2 gpurun_out/${TAG}_bench.err
python -c "
import json; j=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1]); print(j['value'], j['ms_per_step'], [round(k['ms']*1e3,2) for k in j['kernels']], j['roofline']['frac'], j['roofline']['dominant_kernel'], j['e2e']['value'], {k: v.get('ms_per_step') for k, v in j['extra'].items()})"
