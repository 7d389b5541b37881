#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for gw in 8 16; do
  echo "== HTLR_FUSED_GW=$gw"
  HTLR_FUSED_GW=$gw timeout 300 python tools/full_sweep.py 512 8192 20001
done
