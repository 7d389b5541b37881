#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi.log
tail -4 gpurun_out/pytest_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
timeout 900 $TR bench.py --gpus $N --workload E --n-total 50000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_E50k_n$N.json 2> gpurun_out/bench_E50k_n$N.err; echo "E rc=$?"
timeout 600 python bench.py --workload E --n-total 50000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_E50k_n1.json 2> gpurun_out/bench_E50k_n1.err; echo "E n1 rc=$?"
for f in gpurun_out/bench_E50k_n1.json gpurun_out/bench_E50k_n$N.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
print(sys.argv[1], d["n_gpus"], round(d["value"],2), "it/s", round(r["achieved"]), "GB/s", round(r["frac"],3), d["config"]["rows_per_rank"])
PY
done
