#!/bin/bash
# N = 2: the row-shard parity tests, then the driver's N = 2 command
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
( time timeout 800 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err ) 2>&1 | grep real; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_n$N.err
python - $N <<'PY'
import json,sys
N=sys.argv[1]
d=json.loads(open(f'gpurun_out/bench_n{N}.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step','config','run','clocks','gpu_launches')})
print('parity', d.get('sharded_parity'))
e=d['E_row_sharded']; print('E', e['value'], e['ms_per_step'], e['run'], e['clocks'], e['roofline']['frac'], e['roofline']['sweep_share_of_step'], e['roofline']['two_sweep_bound'])
c=d['D_chains']; print('D', c['value'], c['config'], c['clocks'], c['one_after_the_other'])
PY
