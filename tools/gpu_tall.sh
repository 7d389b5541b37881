#!/bin/bash
# tall path: parity, then config E at N = 1 (200000 rows) and the 25000-row shard of the N = 8 run on one GPU
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tall.py tests/test_gpu_edges.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for NT in 200000 25000; do
  timeout 300 python bench.py --workload E --n-total $NT --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/E_n1_$NT.json 2> gpurun_out/E_n1_$NT.err; echo "NT=$NT rc=$?"
  python - gpurun_out/E_n1_$NT.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['two_sweep_bound']['frac_iteration'], d['roofline']['sweep_share_of_step'], d['clocks'])
PY
done
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:'k_(lv|grad)_sweep' -s 30 -c 4 --csv --log-file gpurun_out/tall_launches.csv python bench.py --workload E --n-total 200000 --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
grep -E "k_lv|k_grad" gpurun_out/tall_launches.csv | awk -F'","' '{print $5, $(NF-2), $NF}' | head -8
