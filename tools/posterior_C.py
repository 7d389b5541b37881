"""Level-2 correctness on the named configuration: config C (gendata_FAM n=1000, p=20000, C=3) with htlr() defaults
(iter = 2000, warmup = 1000): the CUDA chain (device Philox) against the reference's own CPU sampler (verbatim
build, its own stream) — selected features and posterior sdb of the leading features. ~3 minutes of host CPU."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench, htlr_b200
from oracle import oracle as O
from tests.helpers import posterior_summary

iters_h = int(os.environ.get("WARM", 1000)); iters_rmc = int(os.environ.get("ITER", 1000))
args = dict(bench.make_problem("C_default"), iters_h=iters_h, iters_rmc=iters_rmc)
t0 = time.perf_counter()
fg = htlr_b200.htlr_fit_helper(**args, seed=11)
tg = time.perf_counter() - t0
t0 = time.perf_counter()
fr = O.ref_fit(**args)
tr = time.perf_counter() - t0
fr["iter.rmc"] = iters_rmc
sg, sr = posterior_summary(fg), posterior_summary(fr)
top = 12
og, orr = np.argsort(-sg["sdb"])[:top] + 1, np.argsort(-sr["sdb"])[:top] + 1
res = {"config": "C, htlr() defaults", "iters_h": iters_h, "iters_rmc": iters_rmc, "gpu_seconds": tg, "cpu_reference_seconds": tr,
       "nzero_idx_gpu": sg["nzero_idx"].tolist(), "nzero_idx_ref": sr["nzero_idx"].tolist(),
       "jaccard_nzero": len(set(sg["nzero_idx"]) & set(sr["nzero_idx"])) / max(1, len(set(sg["nzero_idx"]) | set(sr["nzero_idx"]))),
       "top_features_gpu": og.tolist(), "top_features_ref": orr.tolist(),
       "sdb_top_gpu": sg["sdb"][orr - 1].round(4).tolist(), "sdb_top_ref": sr["sdb"][orr - 1].round(4).tolist(),
       "mean_nuvar_gpu": float(np.mean(fg["mcuvar"][1:])), "mean_nuvar_ref": float(np.mean(fr["mcuvar"][1:])),
       "hmc_reject_gpu": float(np.mean(fg["mchmcrej"][1:])), "hmc_reject_ref": float(np.mean(fr["mchmcrej"][1:])),
       "mean_loglike_gpu": float(np.mean(np.ravel(fg["mcloglike"])[-iters_rmc // 2:])),
       "mean_loglike_ref": float(np.mean(np.ravel(fr["mcloglike"])[-iters_rmc // 2:]))}
print(json.dumps(res))
