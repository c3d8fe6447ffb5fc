set -x
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -40
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
python bench.py --steps 300 --warmup 20 > gpurun_out/bench_room.json 2> gpurun_out/bench_room.err; tail -c 4000 gpurun_out/bench_room.json; tail -5 gpurun_out/bench_room.err
python bench.py --steps 300 --warmup 20 --dist rand --no-cpu-baseline > gpurun_out/bench_rand.json 2> gpurun_out/bench_rand.err; tail -c 2500 gpurun_out/bench_rand.json; tail -5 gpurun_out/bench_rand.err
python bench.py --steps 300 --warmup 20 --no-graph --no-cpu-baseline > gpurun_out/bench_room_nograph.json 2>&1; tail -c 1500 gpurun_out/bench_room_nograph.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 12 --warmup 3 --no-graph --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"splat_depth|splat_feat|resolve" -s 12 -c 3 -o gpurun_out/prof_r01 python bench.py --steps 12 --warmup 3 --no-graph --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
