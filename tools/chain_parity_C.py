"""Whole-chain parity at config C's full size with injected variates: CUDA (every algorithm) vs the CPU oracle."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, htlr_b200
from oracle import oracle as O
from tests.helpers import relerr

ih, ir = int(os.environ.get("WARM", 30)), int(os.environ.get("ITER", 30))
args = dict(bench.make_problem("C_default"), iters_h=ih, iters_rmc=ir, keep_warmup_hist=True)
p, K = args["p"], args["K"]
T = ih + ir
g = np.random.default_rng(7)
st = dict(normals=g.standard_normal(T * (p + 1) * K), uniforms=g.random(T), gammas=g.gamma((1.0 + K) / 2, size=T * p))
t0 = time.perf_counter()
ora = O.OracleFit(**args); ora.initialize(); ora.set_inject(st["normals"], st["uniforms"], st["gammas"], None); ora.run(T)
b = ora.fetch(); so = ora.iter_stats()
print("oracle seconds", time.perf_counter() - t0, "uvar", so["uvar"][[0, ih - 1, ih, T - 1]], "rej mean", so["rej"].mean(), flush=True)
for algo, name in ((0, "auto"), (2, "grid"), (3, "cluster"), (1, "two-sweep")):
    s = htlr_b200.Sampler(**args, rng_mode=1, algo=algo)
    sg = s.run(T, **st)
    a = s.fetch(); s.close()
    same_acc = bool(np.array_equal(sg["rej"], so["rej"])); same_u = bool(np.array_equal(sg["uvar"], so["uvar"]))
    first_diff = int(np.argmax(sg["rej"] != so["rej"])) if not same_acc else -1
    print(json.dumps({"algo": name, "same_accepts": same_acc, "same_uvar": same_u, "first_accept_diff": first_diff,
                      "relerr_mcdeltas": relerr(a["mcdeltas"], b["mcdeltas"]), "relerr_sigmasbt": relerr(a["mcsigmasbt"], b["mcsigmasbt"]),
                      "relerr_loglike": relerr(np.ravel(a["mcloglike"]), b["mcloglike"]), "rej_gpu": float(sg["rej"].mean())}), flush=True)
