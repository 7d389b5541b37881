"""Grid-mode vs cluster-mode fused trajectory: time per MCMC iteration (L=50) over (n, u, K). cut = 0 and
p = u - 1 make every feature part of the restricted set, so u is exact."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import htlr_b200
from tests.helpers import problem

res = []
for n, C in ((62, 2), (200, 2), (500, 4), (1000, 3), (2000, 3)):
    for u in (16, 64, 128, 256, 512, 1024, 2048):
        p = u - 1
        args = problem(n, p, C, seed=1, iters_rmc=400, iters_h=0, leap_L=50, hmc_sgmcut=0.0, init="zero")
        row = {"n": n, "K": C - 1, "u": u, "MB": 8e-6 * n * u}
        for algo, name in ((2, "grid"), (3, "cluster")):
            s = htlr_b200.Sampler(**args, seed=3, algo=algo)
            s.run(20)
            st = torch.cuda.ExternalStream(s.stream)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.sync(); e0.record(st); s.run_async(100); e1.record(st); s.sync()
            row[name + "_us_per_iter"] = e0.elapsed_time(e1) * 10.0
            s.close()
        res.append(row)
        print(json.dumps(row), flush=True)
