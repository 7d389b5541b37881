#!/bin/bash
# the driver's N = 1 commands: our arm (and optionally the reference arm)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_posterior.py tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -3
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err ) 2>&1 | grep real; echo "ours rc=$?"
tail -c 400 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','run','clocks','gpu_launches')})
print('e2e', d['e2e']['value'], 'roofline', {k:d['roofline'][k] for k in ('achieved','frac','frac_iteration')})
print('default_cut', {k:d['default_cut'].get(k) for k in ('value','mean_nuvar','hmc_reject','ratio_to_cpu')})
for k,v in d['other_configs'].items(): print(k, {kk:v.get(kk) for kk in ('value','ratio_to_cpu','hmc_reject','kernel_ms','fp64_tflops','fp64_peak_measured_tflops','frac_of_dmma_peak','call_seconds_with_host_X_ts','probs_row_sums_ok')})
e=d['E_row_sharded']; print('E', e['value'], e['ms_per_step'], e['clocks'], e['roofline']['frac'], e['roofline']['sweep_share_of_step'], e['roofline']['two_sweep_bound']['frac_iteration'], e.get('cpu_baseline',{}).get('value'))
c=d['D_chains']; print('D', c['value'], c['clocks'], c['one_after_the_other'], c.get('cpu_baseline',{}).get('value'))
print('cpu', d['cpu_baseline'])
PY
