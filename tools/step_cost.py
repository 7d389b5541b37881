"""Per-leapfrog-step and fixed cost of one iteration for the trajectory kernels (ROWS_FIRST, FUSED_NO_SMALL, SMALL_FIRST) from runs with
L = 10, 50, 100 on a few small shapes: us_per_step = (t(L=100) - t(L=10)) / 90."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, htlr_b200
from tests.helpers import problem
for n, C, u in ((62, 2, 8), (500, 4, 8), (500, 2, 8), (1000, 3, 8)):
    for algo, name in ((7, "rows"), (5, "cluster"), (6, "small")):
        if name == "small" and n > 600: continue
        row = {"n": n, "K": C - 1, "u": u, "algo": name}
        for L in (10, 50, 100):
            args = problem(n, u - 1, C, seed=1, iters_rmc=400, iters_h=0, leap_L=L, hmc_sgmcut=0.0, init="zero")
            s = htlr_b200.Sampler(**args, seed=3, algo=algo)
            s.run(20)
            st = torch.cuda.ExternalStream(s.stream)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.sync(); e0.record(st); s.run_async(200); e1.record(st); s.sync()
            row[f"L{L}"] = round(e0.elapsed_time(e1) * 5.0, 1)
            s.close()
        row["us_per_step"] = round((row["L100"] - row["L10"]) / 90, 3)
        row["fixed_us"] = round(row["L10"] - 11 * row["us_per_step"], 1)
        print(json.dumps(row), flush=True)
