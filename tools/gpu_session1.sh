#!/bin/bash
# GPU session: parity tests, bench, ncu launch lists and one full capture of the fused trajectory kernel.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 100 --warmup 5 > gpurun_out/bench_s2.json 2> gpurun_out/bench_s2.err; echo "bench rc=$?"
python tools/prof_run.py --workload C_full --iters 2 > gpurun_out/plain_full.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_full.csv \
    python tools/prof_run.py --workload C_full --iters 2 > gpurun_out/ncu_full.log 2>&1
python tools/prof_run.py --workload C_default --iters 2 --warm 150 > gpurun_out/plain_restricted.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1360 -c 40 --csv --log-file gpurun_out/launches_restricted.csv \
    python tools/prof_run.py --workload C_default --iters 2 --warm 150 > gpurun_out/ncu_restricted.log 2>&1
python tools/prof_run.py --workload C_full --iters 2 > gpurun_out/plain_full2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_traj -s 1 -c 1 -f -o gpurun_out/prof_traj_v9 \
    python tools/prof_run.py --workload C_full --iters 2 > gpurun_out/ncu_full_set.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_bench.log 2>&1
ls -la gpurun_out
