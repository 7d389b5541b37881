"""Tall X (columns too long for shared memory: the two-sweep kernels with several row tiles per column). Same parity
gates as the small shapes: UpdatePredProb / UpdateDNlogLike (src/gibbs.cpp:179-188, 239-249), Traject
(gibbs.cpp:390-419), whole chains with injected variates, device-resident inputs."""
import numpy as np
import pytest

from tests.helpers import problem, relerr, streams

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _pair(args, **gpu_kw):
    import htlr_b200
    from oracle import oracle as O
    gpu = htlr_b200.Sampler(**args, **gpu_kw)
    ora = O.OracleFit(**{k: v for k, v in args.items()})
    ora.initialize()
    return gpu, ora


@pytest.mark.parametrize("n,p,C", [(5000, 30, 3), (4608, 25, 4), (6001, 20, 2), (4097, 33, 6), (9000, 12, 17)],
                         ids=["ragged-last-tile", "whole-tiles", "odd-n-K1", "one-row-tile-K5", "K16"])
def test_eval_matches_oracle_on_tall_data(n, p, C):
    args = problem(n, p, C, seed=n + p)
    gpu, ora = _pair(args)
    g = np.random.default_rng(5)
    a0, b0 = gpu.fetch(), ora.fetch()
    assert relerr(a0["mc.param"]["DDNloglike"].ravel(), b0["DDNloglike"]) < TOL
    for u in (1, max(2, p // 3), p + 1):
        iup = np.sort(g.choice(p + 1, size=u, replace=False)).astype(np.int32)
        deltas = np.asfortranarray(g.standard_normal((p + 1, C - 1)) * 0.3)
        a, b = gpu.eval(iup, deltas), ora.eval(iup, deltas)
        assert relerr(a["lv"], b["lv"]) < TOL
        assert relerr(a["pred"], b["pred"]) < TOL
        assert abs(a["loglike"] - b["loglike"]) < TOL * abs(b["loglike"])
        assert relerr(a["grad"], b["grad"]) < TOL
    gpu.close()


def test_trajectory_matches_oracle_on_tall_data():
    n, p, C, L = 5000, 40, 3, 6
    args = problem(n, p, C, seed=77, hmc_sgmcut=0.0, scale=0.1)
    gpu, ora = _pair(args)
    momt = np.asfortranarray(np.random.default_rng(11).standard_normal((p + 1, C - 1)))
    a, b = gpu.trajectory(momt, L), ora.trajectory(momt, L)
    assert a["nuvar"] == b["nuvar"] == p + 1
    for k in ("deltas", "momt", "lv", "var_deltas"):
        assert relerr(a[k], b[k]) < TOL, k
    for k in ("loglike", "nenergy_old", "nenergy_new"):
        assert abs(a[k] - b[k]) < TOL * abs(b[k]), k
    gpu.close()


@pytest.mark.parametrize("use_graph", [False, True])
def test_chain_matches_oracle_on_tall_data(use_graph):
    import htlr_b200
    args = problem(4500, 30, 3, seed=19, iters_h=3, iters_rmc=6, leap_L=5, leap_L_h=2, keep_warmup_hist=True)
    st = streams(args, seed=4)
    gpu, ora = _pair(args, rng_mode=htlr_b200.fit.RNG_INJECT, use_graph=use_graph)
    ora.set_inject(st["normals"], st["uniforms"], st["gammas"], None)
    sg = gpu.run(9, **st)
    ora.run(9)
    so = ora.iter_stats()
    assert np.array_equal(sg["uvar"], so["uvar"]) and np.array_equal(sg["rej"], so["rej"])
    a, b = gpu.fetch(), ora.fetch()
    for k in ("mcdeltas", "mcsigmasbt", "mcvardeltas"):
        assert relerr(a[k], b[k]) < 1e-10, k
    assert relerr(a["mcloglike"].ravel(), b["mcloglike"]) < 1e-10
    gpu.close()


def test_device_resident_tall_inputs_give_the_same_chain():
    """x_on_device with several row tiles per column (the tall workload generates X on the GPU)."""
    import torch
    import htlr_b200
    n, p, C = 4700, 50, 3
    args = problem(n, p, C, seed=31, iters_h=2, iters_rmc=6, leap_L=4, leap_L_h=2)
    host = htlr_b200.Sampler(**args, seed=4)
    host.run(8)
    a = host.fetch()
    host.close()
    nvar, ld = p + 1, 4704
    Xt = torch.zeros(nvar, ld, dtype=torch.float64, device="cuda")
    Xt[:, :n] = torch.from_numpy(np.ascontiguousarray(args["X"].T)).cuda()
    ym = torch.from_numpy(np.ascontiguousarray(np.asarray(args["ymat"]).T)).cuda().contiguous()
    yb = torch.from_numpy(np.ascontiguousarray(args["ybase"], dtype=np.int64)).cuda()
    dev = htlr_b200.Sampler(**dict(args, X=None, ymat=None, ybase=None), seed=4,
                            device_ptrs=(Xt.data_ptr(), ld, ym.data_ptr(), yb.data_ptr()))
    dev.run(8)
    b = dev.fetch()
    dev.close()
    for k in ("mcdeltas", "mcsigmasbt", "mcloglike", "mcuvar", "mchmcrej"):
        assert np.array_equal(a[k], b[k]), k



@pytest.mark.parametrize("n,p,C", [(20500, 14, 2), (9100, 9, 9), (70001, 6, 5)], ids=["K1-five-tiles", "K8", "K4-35-tiles"])
def test_gradient_sweep_shapes(n, p, C):
    """Every register shape of the gradient sweep (rows per block 4096 / 2048 / 1024 / 512 for K <= 2 / 4 / 8 / 16) with
    several row tiles, ragged last tile, odd n; restricted sets shorter than one pass of columns."""
    args = problem(n, p, C, seed=n % 97)
    gpu, ora = _pair(args)
    g = np.random.default_rng(8)
    for u in (1, 3, p + 1):
        iup = np.sort(g.choice(p + 1, size=u, replace=False)).astype(np.int32)
        deltas = np.asfortranarray(g.standard_normal((p + 1, C - 1)) * 0.2)
        a, b = gpu.eval(iup, deltas), ora.eval(iup, deltas)
        assert relerr(a["lv"], b["lv"]) < TOL
        assert relerr(a["grad"], b["grad"]) < TOL
        assert abs(a["loglike"] - b["loglike"]) < TOL * abs(b["loglike"])
    gpu.close()


def test_pad_row_of_device_resident_x_is_neutralised():
    """with x_on_device and an odd n the kernels fetch whole row pairs; whatever sits in row n of the
    caller's buffer (here NaN) must not reach the gradient. create zeroes the pad row when ldx leaves room for it."""
    import torch
    import htlr_b200
    n, p, C = 301, 40, 3
    args = problem(n, p, C, seed=5, iters_h=2, iters_rmc=6, leap_L=6, leap_L_h=2, hmc_sgmcut=0.0)
    host = htlr_b200.Sampler(**args, seed=4)
    host.run(8)
    a = host.fetch()
    host.close()
    nvar, ld = p + 1, 304
    Xt = torch.full((nvar, ld), float("nan"), dtype=torch.float64, device="cuda")
    Xt[:, :n] = torch.from_numpy(np.ascontiguousarray(args["X"].T)).cuda()
    ym = torch.from_numpy(np.ascontiguousarray(np.asarray(args["ymat"]).T)).cuda().contiguous()
    yb = torch.from_numpy(np.ascontiguousarray(args["ybase"], dtype=np.int64)).cuda()
    for algo in (0, 1):
        dev = htlr_b200.Sampler(**dict(args, X=None, ymat=None, ybase=None), seed=4, algo=algo,
                                device_ptrs=(Xt.data_ptr(), ld, ym.data_ptr(), yb.data_ptr()))
        dev.run(8)
        b = dev.fetch()
        dev.close()
        assert np.isfinite(b["mcdeltas"]).all()
        if algo == 0:
            for k in ("mcdeltas", "mcsigmasbt", "mcloglike", "mcuvar", "mchmcrej"):
                assert np.array_equal(a[k], b[k]), k
    with pytest.raises(htlr_b200.fit.HtlrCudaError):
        htlr_b200.Sampler(**dict(args, X=None, ymat=None, ybase=None), seed=4,
                          device_ptrs=(Xt.data_ptr() + 8, ld, ym.data_ptr(), yb.data_ptr()))   # misaligned X
