#!/bin/bash
# What the driver runs at round end, in one go: GPU tests, smoke(), the reference arm, our arm.
mkdir -p gpurun_out
python -m pytest tests/ -x -q -m gpu > gpurun_out/rehearsal_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 2 gpurun_out/rehearsal_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/rehearsal_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/rehearsal_smoke.log
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > gpurun_out/rehearsal_ref.json 2> gpurun_out/rehearsal_ref.err ) 2>&1 | grep real; echo "ref rc=$?"
( time python bench.py --gpus 1 --steps 100 --warmup 5 > gpurun_out/rehearsal_ours.json 2> gpurun_out/rehearsal_ours.err ) 2>&1 | grep real; echo "ours rc=$?"
wc -l gpurun_out/rehearsal_ref.json gpurun_out/rehearsal_ours.json
python - <<'PY'
import json
r=json.loads(open("gpurun_out/rehearsal_ref.json").read().strip().splitlines()[-1])
o=json.loads(open("gpurun_out/rehearsal_ours.json").read().strip().splitlines()[-1])
print("ref", r["value"], r["steps"], r.get("all_cores",{}).get("value"))
print("ours", o["value"], "e2e", o["e2e"]["value"], "frac", o["roofline"]["frac"], "traffic", o["roofline"]["traffic"], "launches", o["gpu_launches"], o["clocks"])
print("ratio e2e/ref", o["e2e"]["value"]/r["value"])
PY
