#!/bin/bash
# the whole -m gpu suite with the measured margins of the tightened gates, then smoke()
mkdir -p gpurun_out; rm -f gpurun_out/margins.jsonl
HTLR_MARGINS_OUT=$PWD/gpurun_out/margins.jsonl timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
cat gpurun_out/margins.jsonl 2>/dev/null | head -20
