"""Single-block kernel (AUTO, and SMALL_FIRST = whenever it fits) vs cluster/grid only (FUSED_NO_SMALL): time per iteration over
small (n, u, K)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, htlr_b200
from tests.helpers import problem
for n, C in ((62, 2), (200, 2), (500, 4), (1000, 3)):
    for u in (8, 16, 32, 48, 64, 96, 128, 256):
        if 8 * n * u > 200e3:
            continue
        args = problem(n, u - 1, C, seed=1, iters_rmc=400, iters_h=0, leap_L=50, hmc_sgmcut=0.0, init="zero")
        row = {"n": n, "K": C - 1, "u": u, "KB": 8e-3 * n * u}
        for algo, name in ((0, "auto"), (6, "small_first"), (5, "no_small")):
            s = htlr_b200.Sampler(**args, seed=3, algo=algo)
            s.run(20)
            st = torch.cuda.ExternalStream(s.stream)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.sync(); e0.record(st); s.run_async(200); e1.record(st); s.sync()
            row[name] = round(e0.elapsed_time(e1) * 5.0, 1)
            if algo == 0:
                row["small_ran"] = s.profile_get()["small_kernel_trajectories"] > 0
            s.close()
        print(json.dumps(row), flush=True)
