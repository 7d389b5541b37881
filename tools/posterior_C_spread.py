"""Between-chain spread on config C defaults: several GPU chains (device Philox, different keys) and a CPU oracle
chain on the Philox twin — rejection rate, mean log-likelihood, leading features."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, htlr_b200
from oracle import oracle as O
from tests.helpers import posterior_summary
args = dict(bench.make_problem("C_default"), iters_h=1000, iters_rmc=1000)
def summ(f, tag):
    f = dict(f); f["iter.rmc"] = 1000
    s = posterior_summary(f)
    print(json.dumps({"chain": tag, "rej": float(np.mean(np.ravel(f["mchmcrej"])[1:])), "nuvar": float(np.mean(np.ravel(f["mcuvar"])[1:])),
                      "loglike_last500": float(np.mean(np.ravel(f["mcloglike"])[-500:])),
                      "top8": (np.argsort(-s["sdb"])[:8] + 1).tolist(), "sdb_top8": np.sort(s["sdb"])[::-1][:8].round(3).tolist(),
                      "n_selected": int(len(s["nzero_idx"]))}), flush=True)
for seed in (11, 12, 13, 14):
    summ(htlr_b200.htlr_fit_helper(**args, seed=seed), f"gpu seed {seed}")
ora = O.OracleFit(**args); ora.set_seed(12); ora.initialize(); ora.run(2000)
summ(ora.fetch(), "cpu oracle, Philox twin seed 12")
