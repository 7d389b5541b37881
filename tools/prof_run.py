"""Minimal driver for ncu captures: a few sampling iterations of one workload, nothing else."""
import argparse, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, htlr_b200
ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="C_full"); ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--algo", type=int, default=0); ap.add_argument("--warm", type=int, default=0)
a = ap.parse_args()
args = bench.make_problem(a.workload)
s = htlr_b200.Sampler(**dict(args, iters_h=a.warm, iters_rmc=a.iters), seed=7, algo=a.algo, use_graph=False)
st = s.run(a.warm + a.iters)
print("uvar", st["uvar"][-3:], "rej", st["rej"][-3:], "launches", s.launch_count)
s.close()
