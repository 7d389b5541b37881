"""Cluster-mode trajectory kernel: per-step phase split (block 0's cycle counters) for a few small restricted sets."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import htlr_b200
from tests.helpers import problem
for n, C, u in ((62, 2, 208), (62, 2, 16), (200, 2, 256), (500, 4, 12), (500, 4, 64), (500, 2, 12), (1000, 3, 128), (1000, 3, 16)):
    args = problem(n, u - 1, C, seed=1, iters_rmc=60, iters_h=0, leap_L=50, hmc_sgmcut=0.0, init="zero")
    s = htlr_b200.Sampler(**args, seed=3, algo=3)
    s.run(5)
    s.profile_enable(True); s.profile_get(reset=True)
    s.run(20)
    pr = s.profile_get()
    print(json.dumps({"n": n, "K": C - 1, "u": u, "traj_us_per_launch": pr["sweep_ms"] * 1e3 / 20,
                      "per_step_us": {q: round(pr[q] * 1e3 / 20 / 50, 2) for q in ("fused_sweep_ms", "fused_barrier_ms", "fused_rowphase_ms")}}), flush=True)
    s.close()
