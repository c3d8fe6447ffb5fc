#!/bin/bash
# ncu launch list + one --set full capture of the three fused kernels (c2), outputs in gpurun_out/.
TAG=${1:-n}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 12 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launch_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"splat_depth|splat_feat|resolve" -s 12 -c 3 -o gpurun_out/prof_${TAG} python bench.py --steps 12 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out | grep ${TAG}
