#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_generators.py tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -8
timeout 300 python bench.py --workload E --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/E_n1_gen.json 2> gpurun_out/E_n1_gen.err; echo "E rc=$?"; tail -3 gpurun_out/E_n1_gen.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/E_n1_gen.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['run'], d['roofline']['frac'], d['clocks'])
PY
