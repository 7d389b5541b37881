"""Synthetic inputs for the tests and the benchmark, restating the LAW of the reference's generators
(not product code: the reference's R generators stay as they are).

  gendata_mlr  — R/gendata.r:26-46   (X ~ N(0,1); t-distributed coefficients; y ~ softmax([1,X] betas))
  gendata_fam  — R/gendata.r:105-113 + src/utils.cpp:72-105 with the vignette's factor pattern
                 (vignettes/simu.Rmd:49-84): 10 informative/correlated features, the rest noise; the
                 p x p loading matrix A = I_p + a 10 x 3 block is applied sparsely, never materialised.
  init_state   — R/core.r:185-193: zero / "random" coefficients and
                 sigmasbt = c(sigmab0, (alpha*e^s + v_j)/2 / Gamma((alpha+K)/2))   (utils.cpp:29-33)
  make_inputs  — R/core.r:103-110,144: ybase / one-hot ymat / intercept column

Random numbers come from NumPy's Philox generator, not R's stream (R is not available).
"""
import numpy as np


def _rng(seed):
    return np.random.Generator(np.random.Philox(seed))


def comp_vardeltas(deltas):
    """src/utils.cpp:51-56."""
    d = np.asarray(deltas, dtype=np.float64)
    return (d * d).sum(1) - d.sum(1) ** 2 / (d.shape[1] + 1)


def gendata_mlr(n, p, NC=3, nu=2.0, w=1.0, seed=0):
    g = _rng(seed)
    X = g.standard_normal((n, p))
    sigmasbt = 1.0 / g.gamma(nu / 2, 2.0 / nu, size=p) * w
    betas = g.standard_normal((p + 1, NC)) * np.r_[0.0, np.sqrt(sigmasbt)][:, None]
    deltas = betas[:, 1:] - betas[:, [0]]
    lv = np.c_[np.ones(n), X] @ betas
    pr = np.exp(lv - lv.max(1, keepdims=True))
    pr /= pr.sum(1, keepdims=True)
    cum = pr.cumsum(1)
    y = (g.random(n)[:, None] > cum).sum(1).clip(0, NC - 1)  # 0-based class
    return {"X": X, "y": y, "deltas": deltas}


_FAM_MEANS = np.array([[0, 1, 0], [0, 0, 0]] + [[0, 0, 1]] * 8, dtype=np.float64) * 2
_FAM_BLOCK = np.array([[1, 0, 0], [2, 1, 0]] + [[0, 0, 1]] * 8, dtype=np.float64)


def gendata_fam(n, p, sd_g=0.5, stdx=True, seed=0):
    """Three classes; features 1-10 informative / mutually correlated through 3 factors, rest noise."""
    assert p >= 10
    g = _rng(seed)
    C = 3
    y = np.arange(n) % C
    muj = np.zeros((p, C))
    muj[:10] = _FAM_MEANS
    # A = diag(1, p); A[0:10, 0:3] = block  ->  rows < 10 mix the first three factors (+ own factor for 3..9)
    z = g.standard_normal((n, p))          # factor scores, one row per case
    Az = z.copy()
    Az[:, :10] = z[:, :3] @ _FAM_BLOCK.T
    Az[:, 3:10] += z[:, 3:10]
    diagAAt = np.ones(p)
    diagAAt[:10] = (_FAM_BLOCK ** 2).sum(1)
    diagAAt[3:10] += 1.0
    X = Az + muj[:, y].T + sd_g * g.standard_normal((n, p))
    if stdx:
        mux = muj.mean(1)
        sdx = np.sqrt(diagAAt + sd_g ** 2 + muj.var(1))
        X = (X - mux) / sdx
    return {"X": X, "y": y}


def make_inputs(X, y):
    """ybase, one-hot ymat (classes 2..C) and the intercept column, R/core.r:103-110, 144."""
    y = np.asarray(y)
    ybase = (y - y.min()).astype(np.int64)
    C = int(ybase.max()) + 1
    K = C - 1
    ymat = np.zeros((X.shape[0], K), order="F")
    for k in range(K):
        ymat[:, k] = ybase == k + 1
    Xa = np.asfortranarray(np.c_[np.ones(X.shape[0]), X])
    return Xa, ymat, ybase, K


def init_state(p, K, init="zero", alpha=1.0, s=-10.0, sigmab0=2000.0, seed=0):
    g = _rng(seed)
    if init == "zero":
        deltas = np.zeros((p + 1, K), order="F")
    elif init == "random":
        deltas = np.asfortranarray(g.standard_normal((p + 1, K)) * 2)
    else:
        raise ValueError(init)
    v = comp_vardeltas(deltas)[1:]
    gam = g.gamma((alpha + K) / 2, size=p)
    sigmasbt = np.r_[sigmab0, 1.0 / gam * (alpha * np.exp(s) + v) / 2.0]
    return deltas, sigmasbt


def std_median_sd(X):
    """std_helper, src/utils.cpp:36-48: centre by the column median, scale by the sample sd."""
    X = np.asarray(X, dtype=np.float64)
    return (X - np.median(X, 0)) / X.std(0, ddof=1)
