"""Grid-mode fused trajectory in the full regime: per-step time and GB/s as the swept matrix grows through L2."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import htlr_b200
from tests.helpers import problem
n, C = 1000, 3
for u in [int(x) for x in (sys.argv[1:] or "2048 4096 8192 12000 16000 20001 30000".split())]:
    p = u - 1
    args = problem(n, p, C, seed=1, iters_rmc=60, iters_h=0, leap_L=50, hmc_sgmcut=0.0, init="zero")
    s = htlr_b200.Sampler(**args, seed=3, algo=2)
    s.run(5)
    s.profile_enable(True); s.profile_get(reset=True)
    s.run(10)
    pr = s.profile_get()
    s.profile_enable(False)
    st = torch.cuda.ExternalStream(s.stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.sync(); e0.record(st); s.run_async(30); e1.record(st); s.sync()
    us = e0.elapsed_time(e1) / 30 * 1e3
    k = pr["sweep_launches"] / 1
    print(json.dumps({"u": u, "MB": 8e-6 * n * u, "us_per_iter": us, "us_per_step_kernel": pr["sweep_ms"] * 1e3 / 10 / 51,
                      "GBs": pr["sweep_bytes"] / pr["sweep_ms"] / 1e6,
                      "phase_us_per_step": {q: pr[q] * 1e3 / 10 / 50 for q in ("fused_sweep_ms", "fused_barrier_ms", "fused_rowphase_ms")}}), flush=True)
    s.close()
