#!/bin/bash
# N-GPU evidence: replicas of the default workload, the tall row-sharded workload at full size, config D chains.
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/bench_C_n1_box$N.json 2> gpurun_out/bench_C_n1_box$N.err; echo "C n1 rc=$?"
timeout 600 $TR bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/bench_C_n$N.json 2> gpurun_out/bench_C_n$N.err; echo "C n$N rc=$?"
timeout 900 $TR bench.py --gpus $N --workload E --steps 5 --warmup 3 > gpurun_out/bench_E_n$N.json 2> gpurun_out/bench_E_n$N.err; echo "E n$N rc=$?"
timeout 600 $TR bench.py --gpus $N --workload D --steps 200 --chains-per-gpu 4 --no-cpu-baseline > gpurun_out/bench_D_n$N.json 2> gpurun_out/bench_D_n$N.err; echo "D n$N rc=$?"
for f in gpurun_out/bench_C_n1_box$N.json gpurun_out/bench_C_n$N.json gpurun_out/bench_E_n$N.json gpurun_out/bench_D_n$N.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print(sys.argv[1], "n_gpus", d["n_gpus"], "value", round(d["value"],2), "ms/step", round(d["ms_per_step"],3), "frac", r.get("frac"), d["config"].get("rows_per_rank"), d.get("one_after_the_other"))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
for f in gpurun_out/bench_*_n$N.err; do echo $f; grep -v "OMP_NUM\|^\*\*\*" $f | tail -n 3; done
