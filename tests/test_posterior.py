"""Level-2 correctness (north_star): posterior coefficient means and the selected-feature set from the CUDA
chain agree with the reference CPU sampler within Monte-Carlo error on the same data; and the on-device
posterior summaries (htlr_cuda_summary) equal the reference's R formulas applied to the fetched trace."""
import numpy as np
import pytest

from htlr_b200 import synth
from tests.helpers import DEFAULTS, posterior_summary, relerr


def signal_problem(n=120, p=80, seed=3, n_signal=4):
    """Binary problem with a few strong features (so that feature selection has a right answer)."""
    g = np.random.default_rng(seed)
    X = g.standard_normal((n, p))
    beta = np.zeros(p)
    beta[:n_signal] = np.array([2.5, -2.0, 1.8, -1.5])[:n_signal]
    pr = 1 / (1 + np.exp(-(X @ beta)))
    y = (g.random(n) < pr).astype(int)
    y[:2] = [0, 1]
    Xa, ymat, ybase, K = synth.make_inputs(X, y)
    deltas, sig = synth.init_state(p, K, init="zero", seed=seed + 1)
    args = dict(DEFAULTS)
    args.update(p=p, K=K, n=n, X=Xa, ymat=ymat, ybase=ybase, ptype="t", deltas=deltas, sigmasbt=sig,
                iters_h=400, iters_rmc=1600, leap_L=20, leap_L_h=5, keep_warmup_hist=False)
    return args


def test_posterior_summary_restatement_shapes():
    """CPU: the numpy restatement follows get_sample_indice (R/mccoef.r:104-117) on a synthetic trace."""
    g = np.random.default_rng(0)
    fit = {"mcdeltas": g.standard_normal((6, 2, 11)), "iter.rmc": 10}
    s = posterior_summary(fit)
    assert list(s["used"]) == [4, 5, 6, 7, 8, 9, 10, 11]          # warm-up 1, burn floor(0.2*10) = 2, ignore first
    assert s["mean_deltas"].shape == (6, 2) and s["sdb"].shape == (5,)
    d = s["mean_deltas"][1:]
    assert np.allclose(s["sdb"], np.sqrt(((d ** 2).sum(1) - d.sum(1) ** 2 / 3) / 3))
    assert set(s["nzero_idx"]) == set(np.nonzero(s["sdb"] > 0.1 * s["sdb"].max())[0] + 1)


@pytest.mark.gpu
def test_device_summary_matches_reference_formulas():
    import htlr_b200
    from tests.helpers import problem
    args = problem(90, 70, 4, seed=12, iters_rmc=40, iters_h=10, keep_warmup_hist=True)
    s = htlr_b200.Sampler(**args, seed=8)
    s.run(50)
    fit = s.fetch()
    a, b = s.summary(), posterior_summary(fit)
    assert relerr(a["mean_deltas"], b["mean_deltas"]) < 1e-13
    assert relerr(a["sdb"], b["sdb"]) < 1e-12
    assert np.array_equal(a["nzero_idx"], b["nzero_idx"])
    # explicit range / thinning: slices 21, 24, ... (0-based)
    a2 = s.summary(first=21, last=50, thin=3)
    m2 = fit["mcdeltas"][:, :, 21:51:3].mean(axis=2)
    assert relerr(a2["mean_deltas"], m2) < 1e-13
    with pytest.raises(htlr_b200.HtlrCudaError, match="bad slice range"):
        s.summary(first=5, last=99)
    s.close()


def _batch_se(x, nb=8):
    """Monte-Carlo standard error of the mean of a (possibly autocorrelated) series, by batch means."""
    x = np.asarray(x)
    m = x.shape[-1] // nb
    bm = x[..., :m * nb].reshape(x.shape[:-1] + (nb, m)).mean(-1)
    return bm.std(-1, ddof=1) / np.sqrt(nb)


@pytest.mark.gpu
def test_posterior_agrees_with_cpu_reference_within_mc_error():
    import htlr_b200
    from oracle import oracle as O
    args = signal_problem()
    T = args["iters_h"] + args["iters_rmc"]
    gpu = htlr_b200.Sampler(**args, seed=101)
    gpu.run(T)
    fa = gpu.fetch()
    sa = gpu.summary()
    gpu.close()
    fits = []
    ora = O.OracleFit(**args)          # CPU restatement of the reference, its own (different) random stream
    ora.set_seed(202)
    ora.initialize()
    ora.run(T)
    fits.append(ora.fetch())
    if O.have_ref():                   # the reference's own sources compiled verbatim, when they travelled
        fits.append(O.ref_fit(**args))
    for f in fits:
        f["iter.rmc"] = args["iters_rmc"]
    used = posterior_summary(fa)["used"] - 1
    da = np.asarray(fa["mcdeltas"])[:, 0, used]
    for fb in fits:
        sb = posterior_summary(fb)
        db = np.asarray(fb["mcdeltas"])[:, 0, used]
        se = np.sqrt(_batch_se(da) ** 2 + _batch_se(db) ** 2)
        z = np.abs(da.mean(1) - db.mean(1)) / np.maximum(se, 1e-3)
        # 81 coefficients: allow the tail of the z distribution, forbid systematic disagreement
        assert np.median(z) < 1.5 and (z < 6).all(), (np.median(z), z.max())
        # the four planted features dominate both posteriors; selected sets agree on them
        top_a = set(np.argsort(-sa["sdb"])[:3] + 1)
        top_b = set(np.argsort(-sb["sdb"])[:3] + 1)
        assert top_a == top_b
        assert set(sa["nzero_idx"]) & {1, 2, 3, 4} == set(sb["nzero_idx"]) & {1, 2, 3, 4}
        jac = len(set(sa["nzero_idx"]) & set(sb["nzero_idx"])) / max(1, len(set(sa["nzero_idx"]) | set(sb["nzero_idx"])))
        assert jac >= 0.6, (sa["nzero_idx"], sb["nzero_idx"])
        # acceptance behaviour is comparable
        assert abs(np.mean(fa["mchmcrej"][1:]) - np.mean(fb["mchmcrej"][1:])) < 0.15


@pytest.mark.gpu
@pytest.mark.parametrize("n,p,C", [(90, 70, 4), (150, 333, 2), (60, 40, 3)])
def test_device_prediction_matches_reference_formulas(n, p, C):
    """htlr_cuda_predict against the numpy restatement of htlr_predict on the fetched trace (N2)."""
    import htlr_b200
    from tests.helpers import predict_reference, problem
    args = problem(n, p, C, seed=n + C, iters_rmc=30, iters_h=10, keep_warmup_hist=(C == 3))
    s = htlr_b200.Sampler(**args, seed=8)
    s.run(40)
    fit = s.fetch()
    g = np.random.default_rng(1)
    X_ts = np.c_[np.ones(77), g.standard_normal((77, p))]
    a, b = s.predict(X_ts), predict_reference(fit, X_ts)
    assert a.shape == (77, C)
    assert relerr(a, b) < 1e-11
    assert np.allclose(a.sum(1), 1.0, atol=1e-12)
    a2 = s.predict(X_ts[:5], first=3, last=29, thin=4)
    lv = np.einsum("ij,jks->iks", X_ts[:5], fit["mcdeltas"][:, :, 3:30:4])
    full = np.concatenate([np.zeros_like(lv[:, :1]), lv], axis=1)
    pr = np.exp(full - np.log(np.exp(full).sum(1, keepdims=True)))[:, 1:].mean(2)
    assert relerr(a2[:, 1:], pr) < 1e-11
    s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("n,p", [(62, 300), (63, 17), (2, 5), (1000, 40), (4097, 3)])
def test_device_std_matches_std_helper(n, p):
    """htlr_cuda_std against the restatement of std_helper (src/utils.cpp:36-48): exact medians, sample sd."""
    import htlr_b200
    g = np.random.default_rng(n + p)
    A = g.standard_normal((n, p)) * g.uniform(0.5, 5, p) + g.uniform(-3, 3, p)
    A[::7] = np.round(A[::7], 1)          # ties
    if n > 10:
        A[3, 0] = A[5, 0] = A[8, 0]       # repeated values around the middle
    out = htlr_b200.std(A)
    assert np.array_equal(out["median"], np.median(A, axis=0))
    assert relerr(out["sd"], A.std(axis=0, ddof=1)) < 1e-13
    assert relerr(out["X"], synth.std_median_sd(A)) < 1e-12
