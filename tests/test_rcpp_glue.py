"""the Rcpp glue (htlr_b200/host/rcpp_glue.cpp, the drop-in for the reference's src/main.cpp:7-25).
R and Rcpp do not exist in this image, so the glue is compiled against the repo's Rcpp/Armadillo stand-ins (the ones
the reference's own sources are compiled against for oracle/_ref) and driven through the same C entry point and the
same variate hooks as the reference build: one harness, two implementations of `htlr_fit_helper`."""
import ctypes
import os

import numpy as np
import pytest

from tests.helpers import problem, relerr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GLUE = os.path.join(ROOT, "oracle", "_glue", "libhtlr_glue_test.so")


def _streams(args, seed):
    g = np.random.default_rng(seed)
    T = args["iters_h"] + args["iters_rmc"]
    p, K = args["p"], args["K"]
    return dict(normals=g.standard_normal(T * (p + 1) * K), uniforms=g.random(T + 1),
                gammas=g.gamma((args["alpha"] + K) / 2, size=T * p))


def test_glue_builds_against_the_stand_ins_and_exports_the_entry_point():
    import subprocess
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "glue"], check=True)
    lib = ctypes.CDLL(GLUE)
    for sym in ("htlr_ref_fit", "htlr_ref_set_option", "htlr_ref_last_error", "htlr_ref_log"):
        assert hasattr(lib, sym), sym


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [dict(n=120, p=80, C=3, keep_warmup_hist=True), dict(n=62, p=300, C=2, keep_warmup_hist=False),
                                 dict(n=300, p=60, C=4, keep_warmup_hist=True, hmc_sgmcut=0.0, leap_step=0.1)],
                         ids=["C3-history", "K1", "K3-all-features"])
def test_r_order_mode_reproduces_the_reference_on_the_same_stream(cfg):
    """options(HTLR.cuda.rng = "R"): the glue draws from the host's generator in the reference's order, so the SAME
    sequential streams give the reference's chain: same names and shapes, same number of variates consumed, values to
    1e-10, identical accept history."""
    from oracle import oracle as O
    if not O.have_ref():
        pytest.skip("oracle/_ref not built")
    cfg = dict(cfg)
    args = problem(cfg.pop("n"), cfg.pop("p"), cfg.pop("C"), seed=31, iters_h=5, iters_rmc=7, leap_L=9, leap_L_h=3, **cfg)
    st = _streams(args, 5)
    ref = O.ref_fit(**args, **st)
    out = O.ref_fit(**args, **st, impl="glue", options={"HTLR.cuda.rng": "R"})
    assert np.array_equal(ref["consumed"], out["consumed"])       # normals, uniforms, gammas drawn from the host stream
    assert np.array_equal(ref["mcuvar"], out["mcuvar"]) and np.array_equal(ref["mchmcrej"], out["mchmcrej"])
    for k in ("mcdeltas", "mcsigmasbt", "mcvardeltas", "mclogw", "mcloglike", "DDNloglike"):
        assert ref[k].shape == out[k].shape, k
        assert relerr(out[k], ref[k]) < 1e-10, k


@pytest.mark.gpu
def test_default_mode_is_keyed_from_one_host_uniform():
    """Default RNG: the device stream, keyed by ONE unif_rand() draw inside the RNGScope - the same host stream gives the
    same chain (set.seed() reproducibility), another one a different chain; result layout as Fit::OutputR."""
    from oracle import oracle as O
    args = problem(90, 120, 3, seed=7, iters_h=4, iters_rmc=6, leap_L=8, leap_L_h=2)
    a = O.ref_fit(**args, uniforms=np.array([0.3, 0.9]), impl="glue")
    b = O.ref_fit(**args, uniforms=np.array([0.3, 0.1]), impl="glue")
    c = O.ref_fit(**args, uniforms=np.array([0.7, 0.9]), impl="glue")
    assert a["consumed"].tolist() == [0, 1, 0]
    assert np.array_equal(a["mcdeltas"], b["mcdeltas"]) and not np.array_equal(a["mcdeltas"], c["mcdeltas"])
    assert a["mcdeltas"].shape == (121, 2, 7) and a["mcsigmasbt"].shape == (121, 7)
    assert np.isfinite(a["mcloglike"]).all() and (a["mcuvar"][1:] >= 1).all()


@pytest.mark.gpu
def test_progress_lines_and_errors_match_the_reference():
    from oracle import oracle as O
    if not O.have_ref():
        pytest.skip("oracle/_ref not built")
    args = problem(70, 40, 3, seed=9, iters_h=2, iters_rmc=3, leap_L=5, leap_L_h=2, silence=0)
    st = _streams(args, 6)
    ref = O.ref_fit(**args, **st)
    out = O.ref_fit(**args, **st, impl="glue", options={"HTLR.cuda.rng": "R"})
    assert ref["log"].count("\n") == out["log"].count("\n") == 5
    assert out["log"] == ref["log"]                               # gibbs.cpp:119-121, same numbers to the printed digits
    with pytest.raises(RuntimeError, match="Unsupported prior type"):
        O.ref_fit(**dict(args, ptype=3), impl="glue")
    with pytest.raises(RuntimeError, match="host-order RNG mode"):
        O.ref_fit(**dict(args, ptype="ghs"), **st, impl="glue", options={"HTLR.cuda.rng": "R"})
