#!/bin/bash
# 2-GPU session: row-shard parity tests, replica bench, tall workload at reduced and full size.
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_host_cpp.py tests/test_gpu_parity.py -k 'multi or host or sharded or eval' -m gpu -x -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi.log
tail -15 gpurun_out/pytest_multi.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/bench_replicas_n1_samebox.json 2> gpurun_out/bench_replicas_n1_samebox.err; echo "n1 rc=$?"
tail -c 900 gpurun_out/bench_replicas_n1_samebox.json | head -c 500; echo
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/bench_replicas_n$N.json 2> gpurun_out/bench_replicas_n$N.err; echo "replicas rc=$?"
tail -c 600 gpurun_out/bench_replicas_n$N.json
timeout 600 python bench.py --workload E --n-total 40000 --steps 3 --warmup 3 > gpurun_out/bench_E40k_n1.json 2> gpurun_out/bench_E40k_n1.err; echo "E40k n1 rc=$?"
timeout 600 $TR bench.py --gpus $N --workload E --n-total 40000 --steps 3 --warmup 3 > gpurun_out/bench_E40k_n$N.json 2> gpurun_out/bench_E40k_n$N.err; echo "E40k n$N rc=$?"
timeout 900 $TR bench.py --gpus $N --workload E --steps 3 --warmup 3 > gpurun_out/bench_E_n$N.json 2> gpurun_out/bench_E_n$N.err; echo "E full n$N rc=$?"
for f in gpurun_out/bench_E*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["n_gpus"], d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"].get("sweep_share_of_step"), d["config"]["rows_per_rank"])
except Exception as e:
    print("ERR", e)
PY
done
for f in gpurun_out/bench_E*.err; do echo $f; tail -n 3 $f; done
