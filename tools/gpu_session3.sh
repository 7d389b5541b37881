#!/bin/bash
# Evidence session for the final round-1 code (v12: carried predictor, pipelined download, multi-block select): bench line, launch lists
# of both regimes and of bench.py itself, one `ncu --set full` capture of the trajectory kernel, timeline.
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_v12.json 2> gpurun_out/bench_v12.err; echo "bench rc=$?"
python tools/prof_run.py --workload C_full --iters 2 > gpurun_out/plain_full.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_full_v12.csv \
    python tools/prof_run.py --workload C_full --iters 2 > gpurun_out/ncu_full.log 2>&1
python tools/prof_run.py --workload C_default --iters 3 --warm 150 > gpurun_out/plain_restricted.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 905 -c 30 --csv --log-file gpurun_out/launches_restricted_v12.csv \
    python tools/prof_run.py --workload C_default --iters 3 --warm 150 > gpurun_out/ncu_restricted.log 2>&1
python tools/prof_run.py --workload C_full --iters 2 > gpurun_out/plain_full2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_traj -s 1 -c 1 -f -o gpurun_out/prof_traj_v12 \
    python tools/prof_run.py --workload C_full --iters 2 > gpurun_out/ncu_full_set.log 2>&1
python tools/ncu_top.py gpurun_out/prof_traj_v12.ncu-rep 30 > gpurun_out/ncu_full_k_traj_v12.txt 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_v12.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_bench.log 2>&1
python tools/timeline.py > gpurun_out/timeline_v12.jsonl 2>&1
head -14 gpurun_out/ncu_full_k_traj_v12.txt
tail -c 600 gpurun_out/bench_v12.json
ls -la gpurun_out | head -40
