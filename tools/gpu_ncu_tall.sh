#!/bin/bash
# ncu --set full of the two tall-path sweeps at 200000 rows per GPU
mkdir -p gpurun_out
NT=${1:-200000}
CMD="python bench.py --workload E --n-total $NT --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_E$NT.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_(lv|grad)_sweep' -s 30 -c 2 -f -o gpurun_out/r02_tall_v2_$NT $CMD > gpurun_out/ncu_E$NT.log 2>&1
echo "NT=$NT rc=$?"; tail -2 gpurun_out/ncu_E$NT.log
