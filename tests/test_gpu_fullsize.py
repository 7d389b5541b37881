"""Size-independent properties at BASELINE.json's full sizes (config C: n = 1000, p = 20000, C = 3), where the CPU
oracle would take minutes: leapfrog reversibility, agreement of the independent trajectory algorithms, and
linearity of the deterministic pieces."""
import numpy as np
import pytest

from htlr_b200 import synth
from tests.helpers import DEFAULTS, record_margin, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def config_c():
    d = synth.gendata_fam(1000, 20000, sd_g=0.5, stdx=True, seed=1003)
    X, ymat, ybase, K = synth.make_inputs(d["X"], d["y"])
    g = np.random.default_rng(5)
    deltas = np.asfortranarray(g.standard_normal((20001, K)) * 0.02)
    v = synth.comp_vardeltas(deltas)[1:]
    sig = np.r_[2000.0, (np.exp(-10.0) + v) / 2 / g.gamma((1 + K) / 2, size=20000)]
    args = dict(DEFAULTS)
    args.update(p=20000, K=K, n=1000, X=X, ymat=ymat, ybase=ybase, ptype="t", deltas=deltas, sigmasbt=sig,
                iters_rmc=3, iters_h=0, leap_L=50, leap_step=0.02)
    return args


@pytest.mark.parametrize("cut", [0.0, 0.05])
def test_trajectory_algorithms_agree_at_full_size(config_c, cut):
    """Two-sweep (reference structure) vs the fused kernels: same proposal to rounding, 50 leapfrog steps."""
    import htlr_b200
    args = dict(config_c, hmc_sgmcut=cut)
    g = np.random.default_rng(2)
    momt = np.asfortranarray(g.standard_normal((20001, args["K"])))
    outs = []
    for algo in (1, 2, 3):
        s = htlr_b200.Sampler(**args, algo=algo)
        outs.append(s.trajectory(momt, 50))
        s.close()
    assert outs[0]["nuvar"] == outs[1]["nuvar"] == outs[2]["nuvar"]
    assert outs[0]["nuvar"] == (20001 if cut == 0.0 else int((args["sigmasbt"] > cut * cut).sum()))
    for o in outs[1:]:
        assert relerr(o["deltas"], outs[0]["deltas"]) < 1e-10
        assert relerr(o["lv"], outs[0]["lv"]) < 1e-10
        assert abs(o["nenergy_new"] - outs[0]["nenergy_new"]) < 1e-10 * abs(outs[0]["nenergy_new"])
        assert abs(o["loglike"] - outs[0]["loglike"]) < 1e-10 * abs(outs[0]["loglike"])


def test_leapfrog_is_reversible_at_full_size(config_c):
    """Leapfrog is time-reversible: run L steps, flip the momenta, run L steps from the proposal: back at the start."""
    import htlr_b200
    args = dict(config_c, hmc_sgmcut=0.0)
    g = np.random.default_rng(3)
    momt = np.asfortranarray(g.standard_normal((20001, args["K"])))
    s = htlr_b200.Sampler(**args)
    fwd = s.trajectory(momt, 30)
    s.close()
    back_args = dict(args, deltas=fwd["deltas"])
    s = htlr_b200.Sampler(**back_args)
    back = s.trajectory(np.asfortranarray(-fwd["momt"]), 30)
    s.close()
    assert relerr(back["deltas"], args["deltas"]) < 1e-10
    assert relerr(-back["momt"], momt) < 1e-10
    # and the energy error of the trajectory is small for this step size (symplectic integrator)
    assert abs(fwd["nenergy_new"] - fwd["nenergy_old"]) < 0.5 * 50


def test_linear_predictor_and_gradient_are_linear_at_full_size(config_c):
    """lv(a d1 + b d2) = a lv(d1) + b lv(d2); the gradient at delta = 0 is X'(1/C - Y) exactly computable per column."""
    import htlr_b200
    args = dict(config_c, hmc_sgmcut=0.0)
    s = htlr_b200.Sampler(**args)
    g = np.random.default_rng(4)
    K = args["K"]
    iup = np.arange(20001, dtype=np.int32)
    d1 = np.asfortranarray(g.standard_normal((20001, K)) * 0.01)
    d2 = np.asfortranarray(g.standard_normal((20001, K)) * 0.01)
    e1, e2, e3 = s.eval(iup, d1), s.eval(iup, d2), s.eval(iup, np.asfortranarray(2.0 * d1 - 3.0 * d2))
    assert relerr(e3["lv"], 2.0 * e1["lv"] - 3.0 * e2["lv"]) < 1e-11
    z = s.eval(iup, np.zeros((20001, K), order="F"))
    assert np.allclose(z["pred"], 1.0 / (K + 1), atol=1e-15)
    assert abs(z["loglike"] + 1000 * np.log(K + 1)) < 1e-10
    resid = 1.0 / (K + 1) - args["ymat"]
    sub = g.choice(20001, 64, replace=False)
    assert relerr(z["grad"][sub], args["X"][:, sub].T @ resid) < 1e-11
    s.close()


@pytest.mark.parametrize("algo", [pytest.param(0, id="auto"), pytest.param(2, id="fused"), pytest.param(1, id="two-sweep")])
def test_full_size_chain_with_injected_variates_matches_oracle(algo):
    """Config C at full size with htlr() defaults, zero initial state, 20 warm-up + 20 sampling iterations with
    injected variates: identical accept decisions and restricted-set sizes, traces to 1e-10 (the oracle needs
    under a second here because the restricted set starts small)."""
    import htlr_b200
    from oracle import oracle as O
    d = synth.gendata_fam(1000, 20000, sd_g=0.5, stdx=True, seed=1003)
    X, ymat, ybase, K = synth.make_inputs(d["X"], d["y"])
    deltas, sig = synth.init_state(20000, K, init="zero", seed=1020)
    args = dict(DEFAULTS)
    args.update(p=20000, K=K, n=1000, X=X, ymat=ymat, ybase=ybase, ptype="t", deltas=deltas, sigmasbt=sig, iters_rmc=20,
                iters_h=20, leap_L=50, leap_L_h=5, keep_warmup_hist=True)
    g = np.random.default_rng(7)
    T = 40
    st = dict(normals=g.standard_normal(T * 20001 * K), uniforms=g.random(T), gammas=g.gamma((1.0 + K) / 2, size=T * 20000))
    ora = O.OracleFit(**args)
    ora.initialize()
    ora.set_inject(st["normals"], st["uniforms"], st["gammas"], None)
    ora.run(T)
    b, so = ora.fetch(), ora.iter_stats()
    s = htlr_b200.Sampler(**args, rng_mode=htlr_b200.fit.RNG_INJECT, algo=algo)
    sg = s.run(T, **st)
    a = s.fetch()
    s.close()
    assert np.array_equal(sg["uvar"], so["uvar"]) and np.array_equal(sg["rej"], so["rej"])
    assert so["uvar"].max() > 50     # the restricted set grows well past the intercept during the run
    record_margin(f"C_default_full_size_chain_algo{algo}", **{k: relerr(np.ravel(a[k]), np.ravel(b[k]))
                                                               for k in ("mcdeltas", "mcsigmasbt", "mcvardeltas", "mcloglike")})
    for k in ("mcdeltas", "mcsigmasbt", "mcvardeltas", "mcloglike"):
        assert relerr(np.ravel(a[k]), np.ravel(b[k])) < 1e-10, k


def _injected_chain_vs_oracle(args, T, algo, label):
    import htlr_b200
    from oracle import oracle as O
    p, K = args["p"], args["K"]
    g = np.random.default_rng(17)
    st = dict(normals=g.standard_normal(T * (p + 1) * K), uniforms=g.random(T),
              gammas=g.gamma((args["alpha"] + K) / 2, size=T * p))
    ora = O.OracleFit(**args)
    ora.initialize()
    ora.set_inject(st["normals"], st["uniforms"], st["gammas"], None)
    ora.run(T)
    b, so = ora.fetch(), ora.iter_stats()
    s = htlr_b200.Sampler(**args, rng_mode=htlr_b200.fit.RNG_INJECT, algo=algo)
    sg = s.run(T, **st)
    a = s.fetch()
    s.close()
    assert np.array_equal(sg["uvar"], so["uvar"]) and np.array_equal(sg["rej"], so["rej"])
    errs = {k: relerr(np.ravel(a[k]), np.ravel(b[k])) for k in ("mcdeltas", "mcsigmasbt", "mcvardeltas", "mcloglike")}
    record_margin(label, **errs)
    for k, e in errs.items():
        assert e < 1e-10, (k, e)
    return so


@pytest.mark.parametrize("algo", [pytest.param(0, id="auto"), pytest.param(1, id="two-sweep")])
def test_full_regime_full_size_chain_matches_oracle(algo):
    """Config C at full size in the FULL regime (cut = 0: all 20001 features move, every leapfrog step sweeps the whole
    160 MB of X - the regime the roofline is quoted on): 3 iterations of L = 50 with injected variates against the
    oracle (about 10 s of CPU), bench.py's step size."""
    d = synth.gendata_fam(1000, 20000, sd_g=0.5, stdx=True, seed=1003)
    X, ymat, ybase, K = synth.make_inputs(d["X"], d["y"])
    g = np.random.default_rng(5)
    deltas = np.asfortranarray(g.standard_normal((20001, K)) * 0.02)     # a moving chain: from zero every early proposal
    v = synth.comp_vardeltas(deltas)[1:]                                  # of the full regime is rejected
    sig = np.r_[2000.0, (np.exp(-10.0) + v) / 2 / g.gamma((1 + K) / 2, size=20000)]
    args = dict(DEFAULTS)
    args.update(p=20000, K=K, n=1000, X=X, ymat=ymat, ybase=ybase, ptype="t", deltas=deltas, sigmasbt=sig, iters_rmc=3,
                iters_h=0, leap_L=50, leap_step=0.02, hmc_sgmcut=0.0)
    so = _injected_chain_vs_oracle(args, 3, algo, f"C_full_regime_full_size_chain_algo{algo}")
    assert (so["uvar"] == 20001).all()
    assert so["rej"].sum() < 3       # at least one accepted proposal: the trace moved


@pytest.mark.parametrize("cut,T", [(0.05, 30), (0.0, 3)], ids=["default-cut", "full-regime"])
def test_config_b_shape_chain_matches_oracle(cut, T):
    """BASELINE configs[1]: gendata_MLR n = 500, p = 5000, C = 4 at full size."""
    d = synth.gendata_mlr(500, 5000, NC=4, seed=1002)
    y = d["y"].copy()
    y[:4] = np.arange(4)
    X, ymat, ybase, K = synth.make_inputs(d["X"], y)
    deltas, sig = synth.init_state(5000, K, init="zero", seed=1019)
    args = dict(DEFAULTS)
    args.update(p=5000, K=K, n=500, X=X, ymat=ymat, ybase=ybase, ptype="t", deltas=deltas, sigmasbt=sig, iters_rmc=T // 2,
                iters_h=T - T // 2, leap_L=50, leap_L_h=5, hmc_sgmcut=cut, keep_warmup_hist=True,
                leap_step=0.3 if cut > 0 else 0.05)
    _injected_chain_vs_oracle(args, T, 0, f"B_shape_cut{cut}")


def test_config_d_shape_chain_matches_oracle():
    """BASELINE configs[3]: one of the chains of gendata_MLR n = 200, p = 50000, C = 2 (K = 1) at full size, htlr()
    defaults, on the cluster-only algorithm the multi-chain workload uses."""
    d = synth.gendata_mlr(200, 50000, NC=2, seed=1004)
    y = d["y"].copy()
    y[:2] = np.arange(2)
    X, ymat, ybase, K = synth.make_inputs(d["X"], y)
    deltas, sig = synth.init_state(50000, K, init="zero", seed=1021)
    args = dict(DEFAULTS)
    args.update(p=50000, K=K, n=200, X=X, ymat=ymat, ybase=ybase, ptype="t", deltas=deltas, sigmasbt=sig, iters_rmc=15,
                iters_h=15, leap_L=50, leap_L_h=5, keep_warmup_hist=True)
    so = _injected_chain_vs_oracle(args, 30, 4, "D_shape_cluster_only")
    assert so["uvar"].max() > 100
