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
import ctypes
out=(ctypes.c_ulonglong*16)()
htlr_b200._lib.lib().htlr_cuda_debug_cycles(s.h, out)
nst=sum((5 if i<a.warm else 50) for i in range(a.warm+a.iters))
print("dbg ns/step blk0:", [round(out[q]/1.965/nst) for q in range(8)], "blk100:", [round(out[q]/1.965/nst) for q in range(8,16)])
s.close()
