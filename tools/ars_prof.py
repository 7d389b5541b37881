import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, bench, htlr_b200
from htlr_b200 import synth
args = bench.make_problem("C_default")
p, K = args["p"], args["K"]
deltas, sig = synth.init_state(p, K, init="random", seed=9)
deltas *= 0.05
a = dict(args, ptype="neg", deltas=deltas, sigmasbt=sig, iters_h=0, iters_rmc=6, leap_L=3)
s = htlr_b200.Sampler(**a, seed=3, use_graph=False)
s.run(6)
print("ok", s.launch_count)
s.close()
