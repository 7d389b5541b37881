"""the reference's data generators as device-side library entry points (htlr_cuda_gendata_mlr /
_fam: R/gendata.r:26-46, 105-113, src/utils.cpp:72-105). The device draws keyed Philox variates; the oracle holds an
independently written host twin of that generator, so the whole construction is restated in numpy from the same keyed
variates and compared element for element."""
import numpy as np
import pytest

from tests.helpers import relerr

pytestmark = pytest.mark.gpu

GAMMA_STEP = 0x6D6C7267616D
X_, BETA_, Y_, Z_, E_ = 8, 9, 11, 12, 13


def _buffers(torch, n, p, NC, ld=None):
    ld = ld or (n + 15) // 16 * 16
    X = torch.zeros(p + 1, ld, dtype=torch.float64, device="cuda")
    ymat = torch.zeros(NC - 1, n, dtype=torch.float64, device="cuda")
    ybase = torch.zeros(n, dtype=torch.int64, device="cuda")
    return X, ymat, ybase, ld


def test_gendata_mlr_matches_its_restatement_and_shards_consistently():
    import torch
    import htlr_b200
    from oracle import oracle as O
    n, p, NC, nu, w, seed = 300, 40, 4, 2.0, 1.5, 20241
    X, ymat, ybase, ld = _buffers(torch, n, p, NC)
    deltas = htlr_b200.gendata_mlr_device(n, p, NC, X.data_ptr(), ld, ymat.data_ptr(), ybase.data_ptr(), nu=nu, w=w,
                                          seed=seed, want_deltas=True)
    Xh = X.cpu().numpy()[:, :n].T
    # restatement from the same keyed variates
    gam = O.rng_dump(seed, 2, GAMMA_STEP, p, 1, "gamma", nu / 2)[:, 0]
    scale = np.r_[0.0, np.sqrt(w * (nu / 2) / gam)]
    betas = np.stack([O.rng_dump(seed, BETA_, c, p + 1, 1, "norm")[:, 0] for c in range(NC)], axis=1) * scale[:, None]
    Xr = np.stack([O.rng_dump(seed, X_, i, p + 1, 1, "norm")[:, 0] for i in range(n)])
    Xr[:, 0] = 1.0
    assert relerr(Xh, Xr) < 1e-13
    assert relerr(deltas, betas[:, 1:] - betas[:, [0]]) < 1e-12
    lv = Xr @ betas
    pr = np.exp(lv - lv.max(1, keepdims=True))
    u = np.array([O.rng_dump(seed, Y_, i, 1, 1, "unif")[0, 0] for i in range(n)]) * pr.sum(1)
    yr = np.minimum((u[:, None] > pr.cumsum(1)).sum(1), NC - 1)
    yh = ybase.cpu().numpy()
    assert (yh == yr).mean() > 0.99                      # a draw within rounding of a class boundary may flip
    ym = ymat.cpu().numpy().T
    assert np.array_equal(ym, (yh[:, None] == np.arange(1, NC)[None, :]).astype(float))
    # rows [150, 300) generated on their own are the same rows
    X2, ym2, yb2, ld2 = _buffers(torch, 150, p, NC)
    htlr_b200.gendata_mlr_device(150, p, NC, X2.data_ptr(), ld2, ym2.data_ptr(), yb2.data_ptr(), nu=nu, w=w, seed=seed,
                                 row0=150)
    assert np.array_equal(X2.cpu().numpy()[:, :150].T, Xh[150:])
    assert np.array_equal(yb2.cpu().numpy(), yh[150:])


def test_gendata_mlr_law():
    import torch
    import htlr_b200
    n, p, NC = 40000, 6, 3
    X, ymat, ybase, ld = _buffers(torch, n, p, NC)
    htlr_b200.gendata_mlr_device(n, p, NC, X.data_ptr(), ld, ymat.data_ptr(), ybase.data_ptr(), seed=5)
    Xh = X.cpu().numpy()[1:, :n]
    assert np.abs(Xh.mean(1)).max() < 5 / np.sqrt(n) and np.abs(Xh.var(1) - 1).max() < 0.05
    assert np.abs(np.corrcoef(Xh)[np.triu_indices(p, 1)]).max() < 0.03
    yh = ybase.cpu().numpy()
    assert set(np.unique(yh)) <= set(range(NC)) and len(np.unique(yh)) >= 2


@pytest.mark.parametrize("stdx", [False, True])
def test_gendata_fam_matches_its_restatement(stdx):
    import torch
    import htlr_b200
    from oracle import oracle as O
    n, p, NC, sd_g, seed = 240, 30, 3, 0.5, 77
    muj = np.zeros((p, NC))
    muj[:10] = np.array([[0, 1, 0], [0, 0, 0]] + [[0, 0, 1]] * 8, dtype=float) * 2
    block = np.array([[1, 0, 0], [2, 1, 0]] + [[0, 0, 1]] * 8, dtype=float)      # vignettes/simu.Rmd:49-84
    X, ymat, ybase, ld = _buffers(torch, n, p, NC)
    htlr_b200.gendata_fam_device(n, muj, block, X.data_ptr(), ld, ymat.data_ptr(), ybase.data_ptr(), sd_g=sd_g,
                                 stdx=stdx, seed=seed)
    Xh = X.cpu().numpy()[:, :n].T
    y = np.arange(n) % NC
    assert np.array_equal(ybase.cpu().numpy(), y) and np.all(Xh[:, 0] == 1.0)
    z = np.stack([O.rng_dump(seed, Z_, i, p, 1, "norm")[:, 0] for i in range(n)])
    e = np.stack([O.rng_dump(seed, E_, i, p, 1, "norm")[:, 0] for i in range(n)])
    A = np.eye(p)
    A[:10, :3] = block
    Xr = z @ A.T + muj[:, y].T + sd_g * e
    if stdx:
        Xr = (Xr - muj.mean(1)) / np.sqrt(np.diag(A @ A.T) + sd_g ** 2 + muj.var(1))
    assert relerr(Xh[:, 1:], Xr) < 1e-12
    if stdx:    # and the generated data are standardised in law
        assert np.abs(Xh[:, 11:].std(0) - 1).mean() < 0.08


def test_generated_data_feeds_the_sampler_without_touching_the_host():
    import torch
    import htlr_b200
    from htlr_b200 import synth
    n, p, NC = 4000, 60, 3
    X, ymat, ybase, ld = _buffers(torch, n, p, NC)
    htlr_b200.gendata_mlr_device(n, p, NC, X.data_ptr(), ld, ymat.data_ptr(), ybase.data_ptr(), seed=9)
    ybase[:NC] = torch.arange(NC, device="cuda")
    ymat[:] = (ybase[None, :] == torch.arange(1, NC, device="cuda")[:, None]).to(torch.float64)
    deltas, sig = synth.init_state(p, NC - 1, init="zero", seed=3)
    s = htlr_b200.Sampler(p, NC - 1, n, None, None, None, "t", 1.0, -10.0, 0.0, 6, 4, 1, 10, 3, 0.3, 0.05, deltas, sig,
                          seed=1, device_ptrs=(X.data_ptr(), ld, ymat.data_ptr(), ybase.data_ptr()))
    st = s.run(10)
    assert np.isfinite(st["loglike"]).all() and (st["uvar"] >= 1).all()
    assert abs(st["loglike"][0] + n * np.log(NC)) < 0.05 * n      # starts at the uniform predictive
    s.close()
