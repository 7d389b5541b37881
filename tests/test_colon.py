"""Config A: the reference's own data set (data/colon.rda, n = 62, p = 2000, binary) with htlr() defaults.
tests/golden/colon.npz holds the data and the chain the reference's sampler (its sources compiled verbatim,
oracle/_ref) produced from injected random streams; tests/golden/make_colon.py made it."""
import os

import numpy as np
import pytest

from htlr_b200 import synth
from tests.helpers import relerr

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "colon.npz")


def colon_case():
    d = np.load(GOLD)
    X, y = np.asarray(d["X"], dtype=np.float64), d["y"]
    Xa, ymat, ybase, K = synth.make_inputs(synth.std_median_sd(X), y)     # R/core.r:103-146
    p = X.shape[1]
    args = dict(p=p, K=K, n=62, X=Xa, ymat=ymat, ybase=ybase, ptype="t", alpha=1.0, s=-10.0, eta=0.0, iters_rmc=12,
                iters_h=12, thin=1, leap_L=50, leap_L_h=5, leap_step=0.3, hmc_sgmcut=0.05,
                deltas=np.zeros((p + 1, K), order="F"), sigmasbt=d["sigmasbt"], keep_warmup_hist=True, silence=1,
                legacy=False)
    g = np.random.default_rng(int(d["stream_seed"]))
    T = 24
    st = dict(normals=g.standard_normal(T * (p + 1) * K), uniforms=g.random(T), gammas=g.gamma((1.0 + K) / 2, size=T * p))
    return d, args, st


def check_against_golden(fit, d, tol):
    mcd = np.asarray(fit["mcdeltas"])
    nz = d["nz_rows"]
    assert relerr(mcd[nz], d["mcdeltas_nz"]) < tol
    rest = np.delete(np.arange(mcd.shape[0]), nz)
    assert np.abs(mcd[rest]).max() == 0.0                      # features never updated stay exactly zero
    assert np.array_equal(np.ravel(fit["mcuvar"]), d["mcuvar"])
    assert np.array_equal(np.ravel(fit["mchmcrej"]), d["mchmcrej"])
    assert relerr(np.ravel(fit["mcloglike"]), d["mcloglike"]) < tol
    assert relerr(np.asarray(fit["mcsigmasbt"])[:, -1], d["mcsigmasbt_last"]) < tol
    assert relerr(np.asarray(fit["mcvardeltas"])[:, -1], d["mcvardeltas_last"]) < tol


def test_colon_fixture_is_the_reference_dataset():
    d, args, _ = colon_case()
    assert d["X"].shape == (62, 2000) and int(d["y"].sum()) == 40
    Xs = args["X"][:, 1:]
    assert np.allclose(np.median(Xs, 0), 0, atol=1e-12) and np.allclose(Xs.std(0, ddof=1), 1)   # std_helper
    assert d["mcuvar"][1:].min() >= 1 and d["mchmcrej"].sum() > 0      # the chain exercises restriction and rejection


def test_oracle_reproduces_reference_chain_on_colon():
    """Pins the CPU restatement to the reference on the reference's own data (config A)."""
    from oracle import oracle as O
    d, args, st = colon_case()
    ora = O.OracleFit(**args)
    ora.initialize()
    ora.set_inject(st["normals"], st["uniforms"], st["gammas"], None)
    ora.run(24)
    check_against_golden(ora.fetch(), d, 1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("algo", [pytest.param(1, id="two-sweep"), pytest.param(2, id="fused"),
                                  pytest.param(3, id="cluster"), pytest.param(0, id="auto"), pytest.param(6, id="small-first")])
def test_cuda_reproduces_reference_chain_on_colon(algo):
    import htlr_b200
    d, args, st = colon_case()
    s = htlr_b200.Sampler(**args, rng_mode=htlr_b200.fit.RNG_INJECT, algo=algo)
    s.run(24, **st)
    check_against_golden(s.fetch(), d, 1e-10)
    s.close()
