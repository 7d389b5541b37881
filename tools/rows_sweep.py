"""Row-partitioned cluster kernel (ROWS_FIRST = whenever it fits) vs the other trajectory kernels (FUSED_NO_SMALL: cluster / grid
mode of k_traj; SMALL_FIRST: single-block kernel): time per iteration (L = 50) over (n, u, K). cut = 0 and p = u - 1 make
every feature part of the restricted set, so u is exact."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, htlr_b200
from tests.helpers import problem
for n, C in [(n, C) for n in (300, 500, 1000, 2000) for C in (2, 3, 4, 5)]:
    for u in (4, 8, 12, 16, 24, 32, 64, 128):
        K = C - 1
        if u * K > 1024:
            continue
        args = problem(n, u - 1, C, seed=1, iters_rmc=400, iters_h=0, leap_L=50, hmc_sgmcut=0.0, init="zero")
        row = {"n": n, "K": K, "u": u, "uK": u * K}
        for algo, name in ((7, "rows"), (5, "cluster_or_grid"), (6, "small")):
            if name == "small" and n > 256:
                continue
            s = htlr_b200.Sampler(**args, seed=3, algo=algo)
            s.run(20)
            st = torch.cuda.ExternalStream(s.stream)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.sync(); e0.record(st); s.run_async(200); e1.record(st); s.sync()
            row[name] = round(e0.elapsed_time(e1) * 5.0, 1)
            s.close()
        print(json.dumps(row), flush=True)
