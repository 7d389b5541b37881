#!/bin/bash
# round-2 evidence session (one B200): the bench line, its launch list, ncu --set full of the dominant kernels,
# launch lists of the small workloads. Every ncu command follows a plain run of the same command that exited 0.
mkdir -p gpurun_out
B="python bench.py --gpus 1 --steps 20 --warmup 5"
$B > gpurun_out/r02_bench_default_n1.json 2> gpurun_out/r02_bench_default_n1.err; echo "bench rc=$?"
S="python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-secondary --no-extra-blocks"
$S > gpurun_out/plain_S.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_bench_py.csv $S > gpurun_out/ncu_S.log 2>&1; echo "launch list rc=$?"
$S > gpurun_out/plain_S2.log 2>&1 && ncu --set full --clock-control none --import-source on -k 'regex:^k_traj$' -s 141 -c 1 -f -o gpurun_out/r02_k_traj_full $S > gpurun_out/ncu_S2.log 2>&1; echo "k_traj full rc=$?"
# 150 warm-up iterations come first: 6 launches per iteration on config A (n < 300: no row-partitioned kernel), 7 on B and C
for W in A B_default C_default; do
  P="python tools/prof_run.py --workload $W --iters 30 --warm 150"
  SKIP=1085; [ $W = A ] && SKIP=930
  $P > gpurun_out/plain_$W.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c 140 --csv --log-file gpurun_out/r02_launches_$W.csv $P > gpurun_out/ncu_$W.log 2>&1; echo "$W rc=$?"
done
for NT in 200000 25000; do
  CMD="python bench.py --workload E --n-total $NT --steps 2 --warmup 3 --no-cpu-baseline"
  $CMD > gpurun_out/plain_E$NT.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_(lv|grad)_sweep' -s 30 -c 2 -f -o gpurun_out/r02_tall_final_$NT $CMD > gpurun_out/ncu_E$NT.log 2>&1; echo "tall $NT rc=$?"
done
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_*.csv
