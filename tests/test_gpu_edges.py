"""Edge cases of the CUDA path against the oracle: smallest shapes, odd row counts, a restricted set that is only
the intercept, many classes (generic kernels), one leapfrog step, thinning, long trajectories, the largest row
count the fused kernel takes and the first one it does not."""
import numpy as np
import pytest

from tests.helpers import problem, relerr, streams

pytestmark = pytest.mark.gpu


def _chain(args, algo, T=None, tol=1e-10, **kw):
    import htlr_b200
    from oracle import oracle as O
    T = T or (args["iters_h"] + args["iters_rmc"])
    st = streams(args, seed=5)
    gpu = htlr_b200.Sampler(**args, rng_mode=htlr_b200.fit.RNG_INJECT, algo=algo, **kw)
    ora = O.OracleFit(**args)
    ora.initialize()
    ora.set_inject(st["normals"], st["uniforms"], st["gammas"], None)
    sg = gpu.run(T, **st)
    ora.run(T)
    so = ora.iter_stats()
    assert np.array_equal(sg["uvar"], so["uvar"]) and np.array_equal(sg["rej"], so["rej"])
    a, b = gpu.fetch(), ora.fetch()
    for k in ("mcdeltas", "mcsigmasbt", "mcvardeltas", "mcloglike", "mcuvar", "mchmcrej"):
        assert relerr(np.ravel(a[k]), np.ravel(b[k])) < tol, k
    gpu.close()
    return a


ALGOS = [pytest.param(1, id="two-sweep"), pytest.param(2, id="fused"), pytest.param(3, id="cluster"),
         pytest.param(0, id="auto"), pytest.param(4, id="cluster-only"), pytest.param(6, id="small-first"), pytest.param(7, id="rows-first")]


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("n,p,C", [(2, 1, 2), (3, 2, 3), (5, 1, 2), (31, 7, 5), (33, 3, 2), (63, 129, 4), (257, 5, 3)])
def test_tiny_and_odd_shapes(n, p, C, algo):
    args = problem(n, p, C, seed=n * 7 + p, iters_rmc=4, iters_h=3, leap_L=6, leap_L_h=2, hmc_sgmcut=0.0)
    _chain(args, algo)


@pytest.mark.parametrize("algo", ALGOS)
def test_only_the_intercept_is_updated(algo):
    """cut so large that no feature passes: iup = {intercept} (sigmab0 = 2000), u = 1 every iteration."""
    args = problem(120, 40, 3, seed=3, iters_rmc=5, iters_h=3, hmc_sgmcut=30.0)
    out = _chain(args, algo)
    assert (out["mcuvar"][1:] == 1).all()


@pytest.mark.parametrize("K1", [6, 9, 10, 17, 18, 34, 50])
def test_many_classes(K1):
    """The reference loops over any number of classes (gibbs.cpp:179-188). K = C - 1 up to 8: the fused kernel; 9..16:
    the two-sweep path with the classes in registers; beyond 16: the same kernels, 16 classes per pass over X (VERDICT
    r01 missing #3: K > 16 used to be refused)."""
    import htlr_b200
    args = problem(140, 30, K1, seed=K1, iters_rmc=3, iters_h=2, leap_L=4)
    _chain(args, 0)
    if K1 == 34:   # and the deterministic pieces on their own, with a ragged last class chunk
        from oracle import oracle as O
        gpu = htlr_b200.Sampler(**args)
        ora = O.OracleFit(**args)
        ora.initialize()
        g = np.random.default_rng(1)
        iup = np.sort(g.choice(31, size=11, replace=False)).astype(np.int32)
        d = np.asfortranarray(g.standard_normal((31, K1 - 1)) * 0.3)
        a, b = gpu.eval(iup, d), ora.eval(iup, d)
        for k in ("lv", "pred", "grad"):
            assert relerr(a[k], b[k]) < 1e-10, k
        X_ts = np.c_[np.ones(9), g.standard_normal((9, 30))]
        gpu.run(5)
        pr = gpu.predict(X_ts)
        from tests.helpers import predict_reference
        assert relerr(pr, predict_reference(gpu.fetch(), X_ts)) < 1e-11
        gpu.close()


@pytest.mark.parametrize("algo", ALGOS)
def test_one_leapfrog_step_and_thinning(algo):
    args = problem(90, 50, 3, seed=8, iters_rmc=4, iters_h=2, leap_L=1, leap_L_h=1, thin=3, keep_warmup_hist=True)
    _chain(args, algo)


@pytest.mark.parametrize("algo", ALGOS)
def test_long_trajectory(algo):
    args = problem(200, 30, 2, seed=5, iters_rmc=2, iters_h=1, leap_L=200, leap_L_h=3, leap_step=0.05, scale=0.2)
    _chain(args, algo, tol=1e-7)


@pytest.mark.parametrize("n", [2047, 2048, 2049, 3001])
def test_row_count_around_the_fused_limit(n):
    """n <= 2048 runs the fused kernel, above it AUTO falls back to the two-sweep path; both match the oracle."""
    import htlr_b200
    args = problem(n, 24, 3, seed=n, iters_rmc=2, iters_h=1, leap_L=5, hmc_sgmcut=0.0)
    _chain(args, 0)
    if n > 2048:
        with pytest.raises(htlr_b200.HtlrCudaError, match="fused trajectory not applicable"):
            htlr_b200.Sampler(**args, algo=2)


def test_restricted_set_larger_than_the_cluster_falls_back_or_errors():
    import htlr_b200
    args = problem(64, 12000, 2, seed=1, iters_rmc=2, iters_h=1, leap_L=3, hmc_sgmcut=0.0)   # u = 12001 > cluster capacity
    _chain(args, 3)                                  # FUSED_CLUSTER: grid-mode launch takes over
    s = htlr_b200.Sampler(**args, algo=4, seed=1)    # CLUSTER_ONLY: reported as an error
    with pytest.raises(htlr_b200.HtlrCudaError, match="exceeds the capacity"):
        s.run(1)
    s.close()


def test_concurrent_chains_on_one_gpu_match_their_solo_runs():
    """Config D's mode: several handles, separate streams, CLUSTER_ONLY; interleaving does not change a chain."""
    import htlr_b200
    args = problem(200, 300, 2, seed=2, iters_rmc=8, iters_h=4, leap_L=10)
    solo = []
    for c in range(3):
        s = htlr_b200.Sampler(**args, seed=100 + c, algo=4)
        s.run(12)
        solo.append(s.fetch())
        s.close()
    ss = [htlr_b200.Sampler(**args, seed=100 + c, algo=4) for c in range(3)]
    for it in range(12):
        for s in ss:
            s.run_async(1)
    for c, s in enumerate(ss):
        s.sync()
        out = s.fetch()
        for k in ("mcdeltas", "mcsigmasbt", "mcloglike", "mcuvar", "mchmcrej"):
            assert np.array_equal(out[k], solo[c][k]), k
        s.close()


def test_auto_uses_the_single_block_kernel_when_the_restricted_set_fits():
    """AUTO: a restricted set that fits in one SM's shared memory is run by k_traj_small; a large one is not."""
    import htlr_b200
    small = problem(100, 3, 3, seed=4, iters_rmc=6, iters_h=2, leap_L=8, hmc_sgmcut=0.0)       # 4 columns x 100 rows (limit: 4)
    big = problem(100, 400, 3, seed=4, iters_rmc=6, iters_h=2, leap_L=8, hmc_sgmcut=0.0)       # 401 columns
    for args, expect in ((small, 8), (big, 0)):
        s = htlr_b200.Sampler(**args, seed=2, algo=0)
        s.profile_get(reset=True)
        s.run(8)
        assert s.profile_get()["small_kernel_trajectories"] == expect
        s.close()
    s = htlr_b200.Sampler(**small, seed=2, algo=5)     # FUSED_NO_SMALL
    s.run(8)
    assert s.profile_get()["small_kernel_trajectories"] == 0
    s.close()


def test_auto_uses_the_row_partitioned_kernel_for_a_handful_of_columns_on_many_rows():
    """AUTO: a handful of columns (here 6 <= 12 at n = 1000, K = 2) goes to k_traj_rows (every block owns a slab of rows, only
    the gradient is exchanged); the same chain as the other algorithms to rounding, identical accept decisions."""
    import htlr_b200
    args = problem(1000, 5, 3, seed=11, iters_rmc=6, iters_h=2, leap_L=8, hmc_sgmcut=0.0)      # 6 columns x 1000 rows, K = 2
    outs = {}
    for algo in (0, 5):
        s = htlr_b200.Sampler(**args, seed=2, algo=algo)
        s.profile_get(reset=True)
        s.run(8)
        assert s.profile_get()["small_kernel_trajectories"] == (8 if algo == 0 else 0)
        outs[algo] = s.fetch()
        s.close()
    assert np.array_equal(outs[0]["mchmcrej"], outs[5]["mchmcrej"])
    assert relerr(outs[0]["mcdeltas"], outs[5]["mcdeltas"]) < 1e-10
    big = problem(1000, 60, 3, seed=11, iters_rmc=6, iters_h=2, leap_L=8, hmc_sgmcut=0.0)      # 61 columns: u K = 122
    s = htlr_b200.Sampler(**big, seed=2, algo=0)
    s.run(8)
    assert s.profile_get()["small_kernel_trajectories"] == 0
    s.close()


def test_run_past_the_end_is_refused():
    import htlr_b200
    args = problem(40, 10, 2, seed=1, iters_rmc=2, iters_h=1)
    s = htlr_b200.Sampler(**args, seed=1)
    s.run(3)
    with pytest.raises(htlr_b200.HtlrCudaError, match="past the configured number"):
        s.run(1)
    s.close()


def test_single_call_entry_point_with_tick_callback():
    """htlr_cuda_fit (the whole of htlr_fit_helper in one C call): same chain as create/run/fetch, the tick callback
    sees every unit of the pipelined run, and a non-zero return aborts with HTLR_E_STATE."""
    import ctypes
    import htlr_b200
    from htlr_b200 import _lib
    args = problem(70, 40, 3, seed=6, iters_rmc=300, iters_h=20, leap_L=4, leap_L_h=2)
    plain = htlr_b200.Sampler(**args, seed=9)     # blocking create / run / fetch
    plain.run(320)
    ref = plain.fetch()
    plain.close()
    L = _lib.lib()
    s = htlr_b200.Sampler(**dict(args, iters_rmc=1, iters_h=0), seed=9)   # only to borrow a filled htlr_cfg
    cfg = s.cfg
    s.close()
    cfg.iters_rmc, cfg.iters_h = args["iters_rmc"], args["iters_h"]
    nvar, K, H = 41, 2, 301
    X = np.asfortranarray(args["X"]); ym = np.asfortranarray(args["ymat"]); yb = np.ascontiguousarray(args["ybase"], dtype=np.int64)
    d0 = np.asfortranarray(args["deltas"]); s0 = np.ascontiguousarray(args["sigmasbt"])
    out = {k: np.zeros(sh, order="F") for k, sh in (("mcdeltas", (nvar, K, H)), ("mcsigmasbt", (nvar, H)),
                                                     ("mcvardeltas", (nvar, H)), ("mclogw", H), ("mcloglike", H),
                                                     ("mcuvar", H), ("mchmcrej", H), ("ddn", nvar))}
    seen = []

    def tick(user, done, stats, n_chunk):
        seen.append((done, n_chunk, stats[n_chunk - 1].uvar))
        return 0
    cb = _lib.TICK_FN(tick)
    err = ctypes.create_string_buffer(256)
    p = lambda a, t=_lib.c_dp: a.ctypes.data_as(t)
    rc = L.htlr_cuda_fit(ctypes.byref(cfg), p(X), 70, p(ym), p(yb, _lib.c_lp), p(d0), p(s0), None, cb, None,
                         p(out["mcdeltas"]), p(out["mcsigmasbt"]), p(out["mcvardeltas"]), p(out["mclogw"]),
                         p(out["mcloglike"]), p(out["mcuvar"]), p(out["mchmcrej"]), p(out["ddn"]), err, 256)
    assert rc == 0, err.value
    # units of 8 iterations: far inside the reference's poll interval of 256 (gibbs.cpp:124)
    assert [d for d, _, _ in seen] == list(range(8, 321, 8)) and all(c == 8 for _, c, _ in seen)
    for k in ("mcdeltas", "mcsigmasbt", "mcvardeltas"):
        assert np.array_equal(out[k], ref[k]), k
    for k in ("mclogw", "mcloglike", "mcuvar", "mchmcrej"):
        assert np.array_equal(out[k], ref[k].ravel()), k
    assert np.array_equal(out["ddn"], ref["mc.param"]["DDNloglike"].ravel())
    cb2 = _lib.TICK_FN(lambda user, done, stats, n_chunk: 1)
    rc = L.htlr_cuda_fit(ctypes.byref(cfg), p(X), 70, p(ym), p(yb, _lib.c_lp), p(d0), p(s0), None, cb2, None,
                         p(out["mcdeltas"]), p(out["mcsigmasbt"]), p(out["mcvardeltas"]), p(out["mclogw"]),
                         p(out["mcloglike"]), p(out["mcuvar"]), p(out["mchmcrej"]), p(out["ddn"]), err, 256)
    assert rc == 5 and b"interrupted" in err.value


def test_device_resident_inputs_give_the_same_chain():
    """htlr_cfg.x_on_device: X / ymat / ybase already in device memory (tall data generated on the GPU)."""
    import torch
    import htlr_b200
    args = problem(130, 60, 3, seed=8, iters_rmc=6, iters_h=3)
    host = htlr_b200.Sampler(**args, seed=4)
    host.run(9)
    a = host.fetch()
    host.close()
    n, nvar, K = 130, 61, 2
    ld = 144
    Xt = torch.zeros(nvar, ld, dtype=torch.float64, device="cuda")
    Xt[:, :n] = torch.from_numpy(np.ascontiguousarray(args["X"].T)).cuda()
    ym = torch.from_numpy(np.ascontiguousarray(np.asarray(args["ymat"]).T)).cuda().contiguous()
    yb = torch.from_numpy(np.ascontiguousarray(args["ybase"], dtype=np.int64)).cuda()
    dev = htlr_b200.Sampler(**dict(args, X=None, ymat=None, ybase=None), seed=4,
                            device_ptrs=(Xt.data_ptr(), ld, ym.data_ptr(), yb.data_ptr()))
    dev.run(9)
    b = dev.fetch()
    dev.close()
    for k in ("mcdeltas", "mcsigmasbt", "mcloglike", "mcuvar", "mchmcrej"):
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("p", [8191, 8192, 50000, 140000], ids=["one-block", "two-blocks", "13-blocks", "max-blocks"])
def test_restricted_set_selection_over_many_features(p):
    """WhichUpdate (gibbs.cpp:156-170) with p beyond one block's segment: k_select chains its block counts; the
    restricted set (ascending, same members) and the chain must equal the oracle's."""
    import htlr_b200
    from oracle import oracle as O
    args = problem(24, p, 2, seed=p % 97, iters_h=1, iters_rmc=2, leap_L=2, leap_L_h=2, hmc_sgmcut=0.05, init="zero")
    st = streams(args, seed=3)
    gpu = htlr_b200.Sampler(**args, rng_mode=htlr_b200.fit.RNG_INJECT)
    ora = O.OracleFit(**args)
    ora.initialize()
    ora.set_inject(st["normals"], st["uniforms"], st["gammas"], None)
    sg = gpu.run(3, **st)
    ora.run(3)
    so = ora.iter_stats()
    assert np.array_equal(sg["uvar"], so["uvar"]) and p / 400 < sg["uvar"][-1] < p / 20    # about 1 % of the features
    a, b = gpu.fetch(), ora.fetch()
    assert relerr(a["mcdeltas"], b["mcdeltas"]) < 1e-10 and relerr(a["mcsigmasbt"], b["mcsigmasbt"]) < 1e-10
    gpu.close()


@pytest.mark.parametrize("algo", [pytest.param(0, id="auto"), pytest.param(2, id="fused"), pytest.param(3, id="cluster"),
                                  pytest.param(6, id="small-first"), pytest.param(1, id="two-sweep")])
@pytest.mark.parametrize("shape", [(62, 400, 2, 0.05, 0.7), (300, 250, 4, 0.0, 0.1), (700, 150, 7, 0.05, 0.7)],
                         ids=["n62-K1", "n300-K3-full", "n700-K6"])
def test_linear_predictor_stays_consistent_with_the_coefficients(shape, algo):
    """The fused kernels never rebuild lv = X delta: they carry it from step to step and from iteration to iteration
    (as the reference's own lv_ drifts). After many accepted and rejected proposals the carried lv,
    the log-likelihood and the coefficients must still describe the same point."""
    import htlr_b200
    n, p, C, cut, step = shape
    args = problem(n, p, C, seed=n, iters_h=50, iters_rmc=350, leap_L=12, leap_L_h=4, hmc_sgmcut=cut, init="zero",
                   leap_step=step)   # step sizes with plenty of rejections
    s = htlr_b200.Sampler(**args, seed=5, algo=algo)
    st = s.run(400)
    state = s.state()
    s.close()
    assert 0.02 < st["rej"].mean() < 0.98                       # both branches exercised
    X, K = np.asarray(args["X"]), C - 1
    lv_ref = np.zeros((n, C))
    lv_ref[:, 1:] = X @ state["deltas"]
    assert relerr(state["lv"], lv_ref) < 1e-11
    m = lv_ref.max(axis=1, keepdims=True)
    lse = np.log(np.exp(lv_ref - m).sum(axis=1)) + m[:, 0]
    ll_ref = float((lv_ref[np.arange(n), np.asarray(args["ybase"], dtype=int)] - lse).sum())
    assert abs(state["loglike"] - ll_ref) < 1e-10 * abs(ll_ref)
