import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, bench, htlr_b200
args = bench.make_problem("C_full")
for step in (0.3, 0.1, 0.05, 0.03, 0.02):
    a = dict(args, leap_step=step, iters_h=0, iters_rmc=30)
    s = htlr_b200.Sampler(**a, seed=1)
    st = s.run(30)
    print(step, "rej", st["rej"].mean(), "loglike", st["loglike"][[0, 9, 29]], "finite", np.isfinite(st["loglike"]).all(), flush=True)
    s.close()
