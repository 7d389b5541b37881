#!/bin/bash
# Evidence session for the final round-1 code (v11: carried predictor, pipelined download): bench line, launch lists
# of both regimes and of bench.py itself, one `ncu --set full` capture of the trajectory kernel, timeline.
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_v11.json 2> gpurun_out/bench_v11.err; echo "bench rc=$?"
python tools/prof_run.py --workload C_full --iters 2 > gpurun_out/plain_full.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_full_v11.csv \
    python tools/prof_run.py --workload C_full --iters 2 > gpurun_out/ncu_full.log 2>&1
python tools/prof_run.py --workload C_default --iters 3 --warm 150 > gpurun_out/plain_restricted.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 30 --csv --log-file gpurun_out/launches_restricted_v11.csv \
    python tools/prof_run.py --workload C_default --iters 3 --warm 150 > gpurun_out/ncu_restricted.log 2>&1
python tools/prof_run.py --workload C_full --iters 2 > gpurun_out/plain_full2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_traj -s 1 -c 1 -f -o gpurun_out/prof_traj_v11 \
    python tools/prof_run.py --workload C_full --iters 2 > gpurun_out/ncu_full_set.log 2>&1
python tools/ncu_top.py gpurun_out/prof_traj_v11.ncu-rep 30 > gpurun_out/ncu_full_k_traj_v11.txt 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_v11.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_bench.log 2>&1
python tools/timeline.py > gpurun_out/timeline_v11.jsonl 2>&1
head -14 gpurun_out/ncu_full_k_traj_v11.txt
tail -c 600 gpurun_out/bench_v11.json
ls -la gpurun_out | head -40
