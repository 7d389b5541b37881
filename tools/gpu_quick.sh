#!/bin/bash
# parity tests + bench (no ncu)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 100 --warmup 5 "$@" > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_quick.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], d["roofline"].get("fused_phase_ms_per_launch_block0"))
print("default_cut", d.get("default_cut",{}).get("value"), d.get("default_cut",{}).get("fused_phase_ms_per_launch_block0"))
PY
