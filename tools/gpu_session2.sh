#!/bin/bash
# GPU session (after the pipelined download / chunk cache change): full bench line, launch list of the default-cut
# regime (sampling iterations only), per-iteration timeline.
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_v10.json 2> gpurun_out/bench_v10.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_v10.json
python tools/prof_run.py --workload C_default --iters 3 --warm 150 > gpurun_out/plain_restricted.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1195 -c 40 --csv --log-file gpurun_out/launches_restricted_v10.csv \
    python tools/prof_run.py --workload C_default --iters 3 --warm 150 > gpurun_out/ncu_restricted.log 2>&1
tail -3 gpurun_out/launches_restricted_v10.csv
python tools/timeline.py > gpurun_out/timeline_v10.jsonl 2>&1
cut -c1-400 gpurun_out/timeline_v10.jsonl
