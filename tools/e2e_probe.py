"""Where does the end-to-end time of htlr_fit_helper go? create / run / fetch / destroy, repeated."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, htlr_b200
args = bench.make_problem("C_full")
Xp = torch.from_numpy(np.ascontiguousarray(args["X"].T)).pin_memory()
args["X"] = Xp.numpy().T
for rep in range(4):
    t0 = time.perf_counter()
    s = htlr_b200.Sampler(**dict(args, iters_h=0, iters_rmc=100), seed=5)
    t1 = time.perf_counter()
    s.run(100)
    t2 = time.perf_counter()
    out = s.fetch()
    t3 = time.perf_counter()
    s.close()
    t4 = time.perf_counter()
    print("rep %d: create %.1f ms, run %.1f ms, fetch %.1f ms, destroy %.1f ms, total %.1f ms" %
          (rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3, (t4 - t0) * 1e3), flush=True)
