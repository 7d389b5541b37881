#!/bin/bash
# the driver's N = 1 commands: both arms back to back
mkdir -p gpurun_out
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err ) 2>&1 | grep real; echo "ours rc=$?"
tail -c 400 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','config','run','clocks','gpu_launches')})
print('e2e', d['e2e']['value'], 'roofline', {k:d['roofline'][k] for k in ('achieved','frac','frac_iteration')})
print('default_cut', {k:d['default_cut'].get(k) for k in ('value','mean_nuvar','hmc_reject','ratio_to_cpu')})
for k,v in d['other_configs'].items(): print(k, v.get('value'), v.get('ratio_to_cpu'), v.get('hmc_reject'))
e=d['E_row_sharded']; print('E', e['value'], e['ms_per_step'], e['run'], e['clocks'], e['roofline']['frac'], e['roofline']['sweep_share_of_step'], e['roofline']['two_sweep_bound'], e.get('cpu_baseline'))
c=d['D_chains']; print('D', c['value'], c['run'], c['clocks'], c['one_after_the_other'], c.get('cpu_baseline'))
print('cpu', d['cpu_baseline'])
PY
