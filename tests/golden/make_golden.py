"""Regenerates tests/golden/*.npz from the REFERENCE ITSELF: the reference's own gibbs.cpp / sampler.cpp /
utils.cpp / ars.cpp / main.cpp compiled verbatim against the API shims (oracle/_ref, needs /root/reference at
build time; `make -C oracle ref`). The fixtures let the oracle restatement be pinned on machines where the
reference tree is absent (the GPU box, a fresh clone).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tests.helpers import problem  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (n, p, C, overrides)
    "t_basic": (40, 30, 3, dict(ptype="t")),
    "t_thin_warm_logw": (30, 40, 2, dict(ptype="t", thin=2, keep_warmup_hist=True, eta=0.5)),
    "t_reject": (45, 35, 3, dict(ptype="t", leap_step=0.9, keep_warmup_hist=True)),
    "t_full": (35, 25, 4, dict(ptype="t", hmc_sgmcut=0.0)),
    "ghs": (40, 30, 3, dict(ptype="ghs")),
    "neg": (36, 28, 2, dict(ptype="neg", thin=2)),
}


def case_inputs(name):
    n, p, C, over = CASES[name]
    args = problem(n, p, C, seed=sum(map(ord, name)), **over)
    g = np.random.default_rng(len(name))
    T = (args["iters_h"] + args["iters_rmc"]) * args["thin"]
    K = C - 1
    nrm = g.standard_normal(T * (p + 1) * K)
    uni = g.random(T * (1 + (p + 1) * 40))      # one shared sequential uniform stream (accept + ARS)
    gam = g.gamma((args["alpha"] + K) / 2, size=T * p)
    return args, nrm, uni, gam


def main():
    if not O.have_ref():
        raise SystemExit("oracle/_ref is not built; run `make -C oracle ref` where /root/reference exists")
    for name in CASES:
        args, nrm, uni, gam = case_inputs(name)
        ref = O.ref_fit(**args, normals=nrm, uniforms=uni, gammas=gam)
        keep = {k: ref[k] for k in ("mcdeltas", "mcsigmasbt", "mcvardeltas", "mclogw", "mcloglike", "mcuvar",
                                    "mchmcrej", "DDNloglike", "consumed")}
        np.savez_compressed(os.path.join(OUT, f"chain_{name}.npz"), **keep)
        print(name, "consumed", ref["consumed"], "rej", ref["mchmcrej"].mean())
    # ARS engine vectors from the verbatim ars.cpp
    g = np.random.default_rng(99)
    m = 400
    ars = {}
    for kind in ("ghs", "neg"):
        for K in (1, 2, 4):
            vard = np.exp(g.uniform(np.log(1e-12), np.log(50), m))
            u = g.random((m, 48))
            r = O.ars_batch(kind, vard, K, 1.0, -10.0, u, engine="ref")
            ars[f"{kind}_{K}_vard"] = vard
            ars[f"{kind}_{K}_u"] = u
            ars[f"{kind}_{K}_x"] = r["x"]
            ars[f"{kind}_{K}_nunif"] = r["n_unif"]
    np.savez_compressed(os.path.join(OUT, "ars_vectors.npz"), **ars)
    # the reference's exported helpers with restatable known answers (tests/testthat/test-util.r)
    import ctypes
    r = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libhtlr_ref.so"))
    dp = ctypes.POINTER(ctypes.c_double)
    A = np.asfortranarray(g.normal(100, 50, (80, 90)))
    vd, lsl = np.zeros(80), np.zeros(80)
    r.htlr_ref_comp_vardeltas(A.ctypes.data_as(dp), 80, 90, vd.ctypes.data_as(dp))
    r.htlr_ref_comp_lsl(A.ctypes.data_as(dp), 80, 90, lsl.ctypes.data_as(dp))
    r.htlr_ref_log_normcons.restype = ctypes.c_double
    lnc = r.htlr_ref_log_normcons(A.ctypes.data_as(dp), 80, 90)
    Xs, med, sd = np.zeros((80, 90), order="F"), np.zeros(90), np.zeros(90)
    r.htlr_ref_std(A.ctypes.data_as(dp), 80, 90, Xs.ctypes.data_as(dp), med.ctypes.data_as(dp), sd.ctypes.data_as(dp))
    np.savez_compressed(os.path.join(OUT, "util_vectors.npz"), A=A, comp_vardeltas=vd, comp_lsl=lsl,
                        log_normcons=lnc, std_X=Xs, std_median=med, std_sd=sd)
    print("written to", OUT)


if __name__ == "__main__":
    main()
