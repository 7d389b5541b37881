"""Same-box A/B of two builds of libhtlr_cuda.so on the bench workloads: HTLR_CUDA_LIB=<path> python tools/ab_quick.py
Prints it/s, the fused kernel's phase split (block 0) and the sweep GB/s for C_full, C_default, B_default, A."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

out = {"lib": os.environ.get("HTLR_CUDA_LIB", "in-tree")}
ALL = (("C_full", 60), ("C_default", 300), ("B_default", 300), ("B", 100), ("A", 300))
for wl, steps in [x for x in ALL if len(sys.argv) < 2 or x[0] in sys.argv[1:]]:
    r = bench.measure_workload(torch, wl, steps, 5, 0, 0, None, want_profile=True)
    pr = r["prof"]
    out[wl] = {"it_s": round(steps / (r["ms"] * 1e-3), 1), "us_per_it": round(1e3 * r["ms"] / steps, 1),
               "uvar": round(r["uvar"], 1), "rej": round(r["rej"], 3),
               "sweep_GBs": round(pr["sweep_bytes"] / 1e9 / max(pr["sweep_ms"] * 1e-3, 1e-12), 0),
               "phase_ms": [round(pr[k] / max(r["prof_steps"], 1), 4) for k in
                            ("fused_sweep_ms", "fused_barrier_ms", "fused_rowphase_ms")]}
print(json.dumps(out))
