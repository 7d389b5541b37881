"""ghs / neg priors: time of the per-feature variance update (warp-per-feature ARS, UpdateSigmasSgm
gibbs.cpp:351-388) inside a chain on the GPU, next to the CPU engine on the same inputs."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, htlr_b200
from htlr_b200 import synth
from oracle import oracle as O

args = bench.make_problem("C_default")
p, K = args["p"], args["K"]
deltas, sig = synth.init_state(p, K, init="random", seed=9)
deltas *= 0.05
for ptype in ("neg", "ghs"):
    a = dict(args, ptype=ptype, deltas=deltas, sigmasbt=sig, iters_h=0, iters_rmc=40, leap_L=10)
    s = htlr_b200.Sampler(**a, seed=3)
    s.run(5)
    s.profile_enable(True); s.profile_get(reset=True)
    s.run(20)
    pr = s.profile_get()
    st = s.state()
    s.close()
    v = st["var_deltas"][1:]
    g = np.random.default_rng(1)
    u = g.random((p, 64))
    t0 = time.perf_counter()
    O.ars_batch(ptype, v, K, 1.0, -10.0, u, max_nhull=1000)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    print(json.dumps({"ptype": ptype, "p": p, "K": K, "gpu_sigma_ms_per_iter": pr["sigma_ms"] / 20,
                      "gpu_total_ms_per_iter": pr["total_ms"] / 20, "cpu_ars_ms_per_iter (1 core)": cpu_ms,
                      "speedup": cpu_ms / (pr["sigma_ms"] / 20)}), flush=True)
