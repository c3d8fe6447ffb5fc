#!/bin/bash
# ncu evidence of the final build (launch list, --set full cold, warm-cache traffic) + the default bench line again.
TAG=${1:-r02f}
mkdir -p gpurun_out
bash scripts/gpu_ncu.sh $TAG
ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,lts__t_sector_hit_rate.pct -k regex:"splat_depth|splat_feat|resolve" -s 12 -c 6 --csv --log-file gpurun_out/warm_${TAG}.csv python bench.py --steps 12 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-extras > gpurun_out/ncu_warm_${TAG}.log 2>&1
python bench.py --steps 300 --warmup 10 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -2 gpurun_out/${TAG}_bench.err
python -c "
import json; j=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1]); print(j['value'], j['ms_per_step'], j['roofline']['frac'], j['roofline']['dominant_kernel'], j['e2e']['value'])"
tail -8 gpurun_out/warm_${TAG}.csv | cut -c1-400
