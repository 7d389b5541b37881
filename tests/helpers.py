"""Shared problem builders for the parity tests (small shapes the CPU oracle finishes in seconds)."""
import numpy as np

from htlr_b200 import synth

DEFAULTS = dict(alpha=1.0, s=-10.0, eta=0.0, iters_rmc=6, iters_h=6, thin=1, leap_L=7, leap_L_h=3, leap_step=0.3,
                hmc_sgmcut=0.05, keep_warmup_hist=False, silence=1, legacy=False)


# Gates (north_star: 1e-10 relative for the deterministic pieces and for a full trajectory with injected variates).
# Whole chains with injected variates have identical accept histories on both sides, so they are held to the same
# 1e-10; measured on the GPU: 1e-13..1e-11 (profiles/r02_parity_margins.json).
TOL = 1e-10
CHAIN_TOL = 1e-10


def relerr(a, b):
    """Norm-wise relative error, the measure north_star's 1e-10 gate is stated in."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0))


def record_margin(case, **values):
    """HTLR_MARGINS_OUT=<file>: append the errors a test measured (how far inside its gate it is) as a JSON line."""
    import json
    import os
    out = os.environ.get("HTLR_MARGINS_OUT")
    if out:
        with open(out, "a") as f:
            f.write(json.dumps(dict(case=case, **{k: float(v) for k, v in values.items()})) + "\n")


def problem(n, p, C, seed=0, init="random", ptype="t", scale=0.5, **over):
    """A seeded MLR problem in htlr_fit_helper argument form."""
    d = synth.gendata_mlr(n, p, NC=C, seed=seed)
    y = d["y"].copy()
    y[:C] = np.arange(C)  # every class present
    X, ymat, ybase, K = synth.make_inputs(d["X"], y)
    g = np.random.default_rng(seed + 1000)
    if init == "random":
        deltas = np.asfortranarray(g.standard_normal((p + 1, K)) * scale)
    else:
        deltas = np.zeros((p + 1, K), order="F")
    v = synth.comp_vardeltas(deltas)[1:]
    sig = np.r_[2000.0, (np.exp(-10.0) + v) / 2 / g.gamma((1 + K) / 2, size=p)]
    args = dict(DEFAULTS)
    args.update(p=p, K=K, n=n, X=X, ymat=ymat, ybase=ybase, ptype=ptype, deltas=deltas, sigmasbt=sig)
    args.update(over)
    return args


def streams(args, seed=0, ars_width=0):
    """Injected variates sized for a whole run, reference consumption order."""
    g = np.random.default_rng(seed + 2000)
    T = (args["iters_h"] + args["iters_rmc"]) * args["thin"]
    p, K = args["p"], args["K"]
    out = dict(normals=g.standard_normal(T * (p + 1) * K), uniforms=g.random(T),
               gammas=g.gamma((args["alpha"] + K) / 2, size=T * p))
    if ars_width:
        out["ars_uniforms"] = g.random((T * (p + 1), ars_width))
    return out


def comp_lsl_np(lv):
    """comp_lsl (src/utils.cpp:59-70 via log_sum_exp): log-sum-exp over cbind(0, lv) along axis 1, max-shifted. Pinned
    to the reference's own output for the golden matrix (tests/golden/util_vectors.npz, htlr_ref_comp_lsl) by
    tests/test_oracle_pinned.py::test_posterior_restatements_are_pinned_to_reference_vectors."""
    full = np.concatenate([np.zeros_like(lv[:, :1]), lv], axis=1)
    m = full.max(axis=1, keepdims=True)
    return np.log(np.exp(full - m).sum(axis=1, keepdims=True)) + m


def comp_sdb_np(deltas):
    """comp_sdb (R/util.r:91-105): sqrt(comp_vardeltas(deltas) / C); comp_vardeltas pinned like comp_lsl_np."""
    C = deltas.shape[1] + 1
    return np.sqrt(synth.comp_vardeltas(deltas) / C)


def posterior_summary(fit, burn=None, thin=1, cut=0.1):
    """numpy restatement of the reference's posterior summaries for a fit dict (OutputR names):
    get_sample_indice / htlr_sdb / nzero_idx (R/mccoef.r:104-156), summary(method = mean) (R/mccoef.r:38-62),
    comp_sdb (R/util.r:91-105), comp_vardeltas (src/utils.cpp:51-56). Indices are R's 1-based ones."""
    mcd = np.asarray(fit["mcdeltas"])
    H = mcd.shape[2]
    iter_rmc = fit["mc.param"]["iter.rmc"] if "mc.param" in fit else fit["iter.rmc"]
    n_warm = H - iter_rmc
    n_burn = int(np.floor(iter_rmc * 0.2)) if burn is None else burn
    used = np.arange(n_warm + n_burn + 1, H + 1, thin)          # 1-based slice numbers, ignore.first = TRUE
    mean = mcd[:, :, used - 1].mean(axis=2)
    sdb = comp_sdb_np(mean[1:])
    return {"used": used, "mean_deltas": mean, "sdb": sdb, "nzero_idx": np.nonzero(sdb > cut * sdb.max())[0] + 1}


def predict_reference(fit, X_ts, burn=None, thin=1):
    """numpy restatement of htlr_predict (R/predict.R:28-92, rep.legacy = TRUE) for prepared X_ts (n_ts x (p+1))."""
    mcd = np.asarray(fit["mcdeltas"])
    H = mcd.shape[2]
    iter_rmc = fit["mc.param"]["iter.rmc"] if "mc.param" in fit else fit["iter.rmc"]
    n_burn = int(np.floor(iter_rmc * 0.1)) if burn is None else burn
    used = np.arange(H - iter_rmc + n_burn + 0, H + 1, thin)     # 1-based, ignore.first = FALSE
    used = used[used >= 1]
    lv = np.einsum("ij,jks->iks", X_ts, mcd[:, :, used - 1])     # n_ts x K x S
    lsl = comp_lsl_np(lv)                                        # n_ts x 1 x S
    probs = np.exp(lv - lsl).mean(axis=2)
    return np.c_[np.maximum(0.0, 1.0 - probs.sum(axis=1)), probs]
