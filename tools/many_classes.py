"""K = 5..8 (6 to 9 classes): the fused trajectory kernel against the two-sweep path, time per iteration."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, htlr_b200
from tests.helpers import problem
for n, p, C, cut in ((500, 5000, 6, 0.05), (500, 5000, 6, 0.0), (200, 2000, 9, 0.05), (1000, 10000, 7, 0.0)):
    args = problem(n, p, C, seed=1, iters_rmc=400, iters_h=0, leap_L=50, hmc_sgmcut=cut, init="zero", leap_step=0.05)
    row = {"n": n, "p": p, "K": C - 1, "cut": cut}
    for algo, name in ((0, "fused_us"), (1, "two_sweep_us")):
        s = htlr_b200.Sampler(**args, seed=3, algo=algo)
        st0 = s.run(40)
        st = torch.cuda.ExternalStream(s.stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.sync(); e0.record(st); s.run_async(100); e1.record(st); s.sync()
        row[name] = round(e0.elapsed_time(e1) * 10.0, 1)
        row["u"] = float(st0["uvar"][-1])
        s.close()
    row["speedup"] = round(row["two_sweep_us"] / row["fused_us"], 2)
    print(json.dumps(row), flush=True)
