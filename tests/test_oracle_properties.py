"""Analytic self-checks of the oracle that the reference never wrote, plus the two
reference properties that can be restated without R."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import problem, relerr, streams


def _fit(args, seed=1):
    f = O.OracleFit(**args)
    f.set_seed(seed)
    f.initialize()
    return f


def test_gradient_matches_finite_differences():
    args = problem(60, 25, 3, seed=5)
    f = _fit(args)
    g = np.random.default_rng(0)
    d = np.asfortranarray(g.standard_normal((26, 2)) * 0.3)
    iup = np.arange(26, dtype=np.int32)
    base = f.eval(iup, d)
    for (j, k) in [(0, 0), (3, 1), (25, 0), (11, 1)]:
        e = 1e-6
        dp, dm = d.copy(), d.copy()
        dp[j, k] += e
        dm[j, k] -= e
        fd = -(f.eval(iup, dp)["loglike"] - f.eval(iup, dm)["loglike"]) / (2 * e)
        assert abs(fd - base["grad"][j, k]) < 1e-6 * max(1.0, abs(fd))


def test_lv_is_fixed_plus_updated_part():
    """lv_fix + X[:,iup] d[iup] == X d: evaluate with complementary restricted sets."""
    args = problem(50, 40, 4, seed=6)
    f = _fit(args)
    g = np.random.default_rng(1)
    d = np.asfortranarray(g.standard_normal((41, 3)))
    allv = f.eval(np.arange(41, dtype=np.int32), d)["lv"]
    a = np.sort(g.choice(41, 13, replace=False)).astype(np.int32)
    b = np.setdiff1d(np.arange(41), a).astype(np.int32)
    assert relerr(f.eval(a, d)["lv"] + f.eval(b, d)["lv"], allv) < 1e-13


def test_pred_rows_sum_to_one_and_loglike_definition():
    args = problem(40, 30, 5, seed=7)
    f = _fit(args)
    d = np.asfortranarray(np.random.default_rng(2).standard_normal((31, 4)))
    r = f.eval(np.arange(31, dtype=np.int32), d)
    assert np.allclose(r["pred"].sum(1), 1.0, atol=1e-14)
    ll = np.log(r["pred"][np.arange(40), args["ybase"]]).sum()
    assert abs(ll - r["loglike"]) < 1e-10 * abs(ll)


def test_leapfrog_is_reversible():
    """Run L steps, flip the momenta, run L more: back at the start (up to rounding)."""
    args = problem(50, 30, 3, seed=8, hmc_sgmcut=0.0, scale=0.2)
    f = _fit(args)
    m0 = np.asfortranarray(np.random.default_rng(3).standard_normal((31, 2)))
    fwd = f.trajectory(m0, 12)
    args2 = dict(args, deltas=fwd["deltas"])
    f2 = _fit(args2)
    back = f2.trajectory(-fwd["momt"], 12)
    assert relerr(back["deltas"], args["deltas"]) < 1e-9
    assert relerr(-back["momt"], m0) < 1e-9


def test_energy_error_vanishes_with_step_size():
    args = problem(60, 20, 3, seed=9, hmc_sgmcut=0.0, scale=0.2)
    m0 = np.asfortranarray(np.random.default_rng(4).standard_normal((21, 2)))
    errs = []
    for step in (0.04, 0.02, 0.01):
        f = _fit(dict(args, leap_step=step))
        t = f.trajectory(m0, int(round(0.32 / step)))  # same integration time
        errs.append(abs(t["nenergy_new"] - t["nenergy_old"]))
    assert errs[1] < errs[0] * 0.5 and errs[2] < errs[1] * 0.5  # second-order integrator


def test_warmup_history_only_changes_recording():
    """tests/testthat/test-warmup.R:13-21."""
    args = problem(40, 30, 3, seed=10, iters_rmc=5, iters_h=4)
    outs = []
    for keep in (False, True):
        f = _fit(dict(args, keep_warmup_hist=keep), seed=9)
        f.run(9)
        outs.append(f.fetch())
    assert np.array_equal(outs[0]["mcdeltas"][:, :, 1:], outs[1]["mcdeltas"][:, :, 5:])
    assert np.array_equal(outs[0]["mcloglike"][1:], outs[1]["mcloglike"][5:])


def test_ghs_neg_zero_row_aborts_like_reference():
    """Quirk 6: var_deltas = 0 -> x0 = log(0) -> 'first tangent point doesn't have positive probability'."""
    args = problem(30, 20, 3, seed=11, ptype="ghs", init="zero")
    f = _fit(args)
    with pytest.raises(RuntimeError, match="first tangent point"):
        f.run(1)


def test_ybase_out_of_range_rejected():
    args = problem(30, 10, 3, seed=12)
    yb = args["ybase"].copy()
    yb[3] = 9
    with pytest.raises(RuntimeError, match="ybase out of range"):
        O.OracleFit(**dict(args, ybase=yb))


def test_t_prior_sigma_is_inverse_gamma():
    """spl_sgm_ig (utils.cpp:29-33): 1/sigma^2 * (alpha*w + v)/2 ~ Gamma((alpha+K)/2, 1)."""
    from scipy import stats
    args = problem(30, 3000, 2, seed=13, iters_rmc=1, iters_h=0, init="zero")
    f = _fit(args, seed=5)
    f.run(1)
    st = f.state()
    gam = (1.0 * np.exp(-10.0) + st["var_deltas"][1:]) / 2.0 / st["sigmasbt"][1:]
    assert stats.kstest(gam, "gamma", args=(1.0,)).pvalue > 1e-3


def test_ars_draws_follow_target_density():
    """ARS output distribution vs the normalised target on a grid (GHS and NEG)."""
    from scipy import stats
    g = np.random.default_rng(14)
    m = 4000
    for kind, K, v in (("ghs", 2, 0.3), ("neg", 3, 2.0)):
        u = g.random((m, 64))
        r = O.ars_batch(kind, np.full(m, v), K, 1.0, -1.0, u)
        xs = np.linspace(-40, 25, 200001)
        if kind == "neg":
            lf = -(K / 2 - 1) * xs - v / 2 / np.exp(xs) - 1.5 * np.log1p(2 * np.exp(xs + 1.0))
        else:
            lf = -(K - 1) / 2 * xs - v / 2 / np.exp(xs) - 1.0 * np.log1p(np.exp(xs + 1.0))
        pdf = np.exp(lf - lf.max())
        cdf = np.cumsum(pdf)
        cdf /= cdf[-1]
        assert stats.kstest(r["x"], lambda q: np.interp(q, xs, cdf)).pvalue > 1e-3


def test_keyed_rng_laws():
    from scipy import stats
    z = O.rng_dump(7, 0, 3, 4000, 2, "norm").ravel()
    u = O.rng_dump(7, 1, 3, 4000, 2, "unif").ravel()
    assert stats.kstest(z, "norm").pvalue > 1e-3 and stats.kstest(u, "uniform").pvalue > 1e-3
    for shape in (0.6, 1.0, 3.5):
        gm = O.rng_dump(7, 2, 3, 6000, 1, "gamma", shape).ravel()
        assert stats.kstest(gm, "gamma", args=(shape,)).pvalue > 1e-3
