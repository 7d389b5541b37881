"""TEST INFRASTRUCTURE — ctypes bindings for the CPU oracle (oracle/liboracle.so) and, when built,
the reference's own sources compiled verbatim against API shims (oracle/_ref/*.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
Nothing in htlr_b200/ does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)
_lp = ctypes.POINTER(ctypes.c_int64)

PTYPES = {"t": 0, "ghs": 1, "neg": 2}


def _p(a, typ=_dp):
    return None if a is None else a.ctypes.data_as(typ)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def build(force=False):
    """Compile the oracle restatement (and oracle/_ref when /root/reference is present)."""
    lib = os.path.join(_HERE, "liboracle.so")
    if force or not os.path.exists(lib) or os.path.exists("/root/reference/src"):
        subprocess.run(["make", "-s", "-C", _HERE, "all"], check=True)
    return lib


class OracleCfg(ctypes.Structure):
    _fields_ = [("p", ctypes.c_int32), ("K", ctypes.c_int32), ("n", ctypes.c_int32), ("ptype", ctypes.c_int32),
                ("alpha", ctypes.c_double), ("s", ctypes.c_double), ("eta", ctypes.c_double),
                ("iters_rmc", ctypes.c_int32), ("iters_h", ctypes.c_int32), ("thin", ctypes.c_int32),
                ("leap_L", ctypes.c_int32), ("leap_L_h", ctypes.c_int32),
                ("leap_step", ctypes.c_double), ("hmc_sgmcut", ctypes.c_double),
                ("keep_warmup_hist", ctypes.c_int32), ("legacy", ctypes.c_int32),
                ("ars_max_nhull", ctypes.c_int32), ("pad_", ctypes.c_int32)]


class RefCfg(ctypes.Structure):
    _fields_ = [("p", ctypes.c_int32), ("K", ctypes.c_int32), ("n", ctypes.c_int32), ("ptype", ctypes.c_int32),
                ("alpha", ctypes.c_double), ("s", ctypes.c_double), ("eta", ctypes.c_double),
                ("iters_rmc", ctypes.c_int32), ("iters_h", ctypes.c_int32), ("thin", ctypes.c_int32),
                ("leap_L", ctypes.c_int32), ("leap_L_h", ctypes.c_int32),
                ("leap_step", ctypes.c_double), ("hmc_sgmcut", ctypes.c_double),
                ("keep_warmup_hist", ctypes.c_int32), ("legacy", ctypes.c_int32),
                ("silence", ctypes.c_int32), ("pad_", ctypes.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _lib = ctypes.CDLL(path)
        _lib.oracle_last_error.restype = ctypes.c_char_p
        _lib.oracle_set_seed.argtypes = [ctypes.c_void_p, ctypes.c_uint64]
        _lib.oracle_rng_dump.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int32,
                                         ctypes.c_int32, ctypes.c_int32, ctypes.c_double, _dp]
    return _lib


def _ref_lib(name):
    path = os.path.join(_HERE, "_ref", name)
    if not os.path.exists(path):
        return None
    return ctypes.CDLL(path)


def have_ref():
    return os.path.exists(os.path.join(_HERE, "_ref", "libhtlr_ref.so"))


def _check(rc):
    if rc != 0:
        raise RuntimeError(lib().oracle_last_error().decode())


class OracleFit:
    """The restated `Fit` sampler. Argument names follow htlr_fit_helper (src/main.cpp:7-14)."""

    def __init__(self, p, K, n, X, ymat, ybase, ptype, alpha, s, eta, iters_rmc, iters_h, thin, leap_L,
                 leap_L_h, leap_step, hmc_sgmcut, deltas, sigmasbt, keep_warmup_hist=False, silence=1,
                 legacy=False, ars_max_nhull=1000):
        L = lib()
        self.p, self.K, self.n, self.nvar, self.C = p, K, n, p + 1, K + 1
        self.iters_rmc, self.iters_h = iters_rmc, iters_h
        self._X = np.asfortranarray(X, dtype=np.float64)
        self._ymat = np.asfortranarray(np.reshape(ymat, (n, K), order="F"), dtype=np.float64)
        self._yb = np.ascontiguousarray(ybase, dtype=np.int64)
        d0 = np.asfortranarray(np.reshape(deltas, (p + 1, K), order="F"), dtype=np.float64)
        s0 = _f64(sigmasbt)
        assert self._X.shape == (n, p + 1)
        self.cfg = OracleCfg(p, K, n, PTYPES[ptype] if isinstance(ptype, str) else ptype, alpha, s, eta,
                             iters_rmc, iters_h, thin, leap_L, leap_L_h, leap_step, hmc_sgmcut,
                             int(keep_warmup_hist), int(legacy), ars_max_nhull, 0)
        self.h = ctypes.c_void_p()
        _check(L.oracle_create(ctypes.byref(self.cfg), _p(self._X), _p(self._ymat), _p(self._yb, _lp),
                               _p(d0), _p(s0), ctypes.byref(self.h)))
        self._keep = []
        self.hist_len = L.oracle_hist_len(self.h)

    def __del__(self):
        try:
            if self.h:
                lib().oracle_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_seed(self, seed):
        lib().oracle_set_seed(self.h, seed)

    def set_inject(self, normals=None, uniforms=None, gammas=None, ars_u=None):
        arrs = [None if a is None else _f64(a) for a in (normals, uniforms, gammas, ars_u)]
        self._keep = arrs
        nm, un, gm, au = arrs
        width = 0 if au is None else au.shape[-1]
        nblk = 0 if au is None else au.size // max(width, 1)
        lib().oracle_set_inject(self.h, _p(nm), ctypes.c_int64(0 if nm is None else nm.size), _p(un),
                                ctypes.c_int64(0 if un is None else un.size), _p(gm),
                                ctypes.c_int64(0 if gm is None else gm.size), _p(au),
                                ctypes.c_int64(width), ctypes.c_int64(nblk))

    def initialize(self):
        _check(lib().oracle_initialize(self.h))

    def run(self, n_iters):
        secs = ctypes.c_double()
        _check(lib().oracle_run(self.h, n_iters, ctypes.byref(secs)))
        return secs.value

    def fetch(self):
        H, nvar, K = self.hist_len, self.nvar, self.K
        out = {"mcdeltas": np.zeros((nvar, K, H), order="F"), "mcsigmasbt": np.zeros((nvar, H), order="F"),
               "mcvardeltas": np.zeros((nvar, H), order="F"), "mclogw": np.zeros(H), "mcloglike": np.zeros(H),
               "mcuvar": np.zeros(H), "mchmcrej": np.zeros(H), "DDNloglike": np.zeros(nvar)}
        lib().oracle_fetch(self.h, _p(out["mcdeltas"]), _p(out["mcsigmasbt"]), _p(out["mcvardeltas"]),
                           _p(out["mclogw"]), _p(out["mcloglike"]), _p(out["mcuvar"]), _p(out["mchmcrej"]),
                           _p(out["DDNloglike"]))
        return out

    def iter_stats(self):
        T = self.iters_h + self.iters_rmc
        o = {k: np.zeros(T) for k in ("loglike", "logw", "uvar", "rej")}
        lib().oracle_iter_stats(self.h, _p(o["loglike"]), _p(o["logw"]), _p(o["uvar"]), _p(o["rej"]))
        return o

    def state(self):
        n, C, nvar, K = self.n, self.C, self.nvar, self.K
        o = {"lv": np.zeros((n, C), order="F"), "pred": np.zeros((n, C), order="F"),
             "deltas": np.zeros((nvar, K), order="F"), "sigmasbt": np.zeros(nvar), "var_deltas": np.zeros(nvar),
             "momt": np.zeros((nvar, K), order="F"), "DNloglike": np.zeros((nvar, K), order="F"),
             "DNlogpost": np.zeros((nvar, K), order="F")}
        ll, lw, nu = ctypes.c_double(), ctypes.c_double(), ctypes.c_int32()
        iup = np.zeros(nvar, dtype=np.int32)
        lib().oracle_state(self.h, _p(o["lv"]), _p(o["pred"]), _p(o["deltas"]), _p(o["sigmasbt"]),
                           _p(o["var_deltas"]), _p(o["momt"]), _p(o["DNloglike"]), _p(o["DNlogpost"]),
                           ctypes.byref(ll), ctypes.byref(lw), ctypes.byref(nu), _p(iup, _ip))
        o.update(loglike=ll.value, logw=lw.value, iup=iup[:nu.value].copy())
        return o

    def eval(self, iup, deltas):
        """lv / pred / loglike / gradient for a given restricted set and coefficient matrix."""
        n, C, nvar, K = self.n, self.C, self.nvar, self.K
        iup = np.ascontiguousarray(iup, dtype=np.int32)
        d = np.asfortranarray(deltas, dtype=np.float64)
        u = iup.size
        lv, pred = np.zeros((n, C), order="F"), np.zeros((n, C), order="F")
        grad = np.zeros((u, K), order="F")
        ll = ctypes.c_double()
        _check(lib().oracle_eval(self.h, _p(iup, _ip), u, _p(d), _p(lv), _p(pred), ctypes.byref(ll), _p(grad)))
        return {"lv": lv, "pred": pred, "loglike": ll.value, "grad": grad}

    def trajectory(self, momt, L):
        n, C, nvar, K = self.n, self.C, self.nvar, self.K
        m = np.asfortranarray(momt, dtype=np.float64)
        o = {"deltas": np.zeros((nvar, K), order="F"), "momt": np.zeros((nvar, K), order="F"),
             "lv": np.zeros((n, C), order="F"), "var_deltas": np.zeros(nvar)}
        ll, e0, e1, nu = ctypes.c_double(), ctypes.c_double(), ctypes.c_double(), ctypes.c_int32()
        _check(lib().oracle_trajectory(self.h, _p(m), L, ctypes.c_double(0.0), _p(o["deltas"]), _p(o["momt"]),
                                       _p(o["lv"]), ctypes.byref(ll), _p(o["var_deltas"]), ctypes.byref(e0),
                                       ctypes.byref(e1), ctypes.byref(nu)))
        o.update(loglike=ll.value, nenergy_old=e0.value, nenergy_new=e1.value, nuvar=nu.value)
        return o

    def ars_effort(self):
        a, b, c, d = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int32()
        lib().oracle_ars_effort(self.h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), ctypes.byref(d))
        return {"evals": a.value, "unifs": b.value, "draws": c.value, "max_hulls": d.value}


def ars_batch(kind, vard, K, alpha, log_aw, u, max_nhull=1000, engine="port"):
    """One ARS draw of x = log sigma^2 per feature. engine: 'port' (restatement) or 'ref' (verbatim ars.cpp)."""
    vard, u = _f64(vard), _f64(u)
    m, width = vard.size, u.shape[1]
    x = np.zeros(m)
    nu, ne, nh = np.zeros(m, np.int32), np.zeros(m, np.int32), np.zeros(m, np.int32)
    k = PTYPES[kind] if isinstance(kind, str) else kind
    if engine == "port":
        _check(lib().oracle_ars_batch(k, m, _p(vard), ctypes.c_double(K), ctypes.c_double(alpha),
                                      ctypes.c_double(log_aw), _p(u), width, max_nhull, _p(x), _p(nu, _ip),
                                      _p(ne, _ip), _p(nh, _ip)))
    else:
        r = _ref_lib("libars_ref.so")
        r.ars_ref_last_error.restype = ctypes.c_char_p
        rc = r.ars_ref_batch(k, m, _p(vard), ctypes.c_double(K), ctypes.c_double(alpha), ctypes.c_double(log_aw),
                             _p(u), width, max_nhull, _p(x), _p(nu, _ip), _p(ne, _ip))
        if rc:
            raise RuntimeError(r.ars_ref_last_error().decode())
    return {"x": x, "n_unif": nu, "n_eval": ne, "n_hulls": nh}


def ars_logw(vard, K, nu, s, eta, x0, u, engine="port"):
    vard, u = _f64(vard), _f64(u)
    x, a, b = ctypes.c_double(), ctypes.c_int32(), ctypes.c_int32()
    args = (vard.size, _p(vard), ctypes.c_double(K), ctypes.c_double(nu), ctypes.c_double(s), ctypes.c_double(eta),
            ctypes.c_double(x0), _p(u), u.size, ctypes.byref(x), ctypes.byref(a), ctypes.byref(b))
    if engine == "port":
        _check(lib().oracle_ars_logw(*args))
    else:
        r = _ref_lib("libars_ref.so")
        r.ars_ref_last_error.restype = ctypes.c_char_p
        if r.ars_ref_logw(*args):
            raise RuntimeError(r.ars_ref_last_error().decode())
    return {"x": x.value, "n_unif": a.value, "n_eval": b.value}


def rng_dump(seed, purpose, step, n_a, n_b, kind, shape=1.0):
    out = np.zeros((n_a, n_b))
    lib().oracle_rng_dump(seed, purpose, step, n_a, n_b, {"unif": 0, "norm": 1, "gamma": 2}[kind], shape, _p(out))
    return out


def comp_vardeltas(deltas):
    d = np.asfortranarray(deltas, dtype=np.float64)
    out = np.zeros(d.shape[0])
    lib().oracle_comp_vardeltas(_p(d), d.shape[0], d.shape[1], _p(out))
    return out


def comp_lsl(A):
    a = np.asfortranarray(A, dtype=np.float64)
    out = np.zeros(a.shape[0])
    lib().oracle_comp_lsl(_p(a), a.shape[0], a.shape[1], _p(out))
    return out


def ref_fit(p, K, n, X, ymat, ybase, ptype, alpha, s, eta, iters_rmc, iters_h, thin, leap_L, leap_L_h,
            leap_step, hmc_sgmcut, deltas, sigmasbt, keep_warmup_hist=False, silence=1, legacy=False,
            normals=None, uniforms=None, gammas=None, impl="reference", options=None):
    """The reference's own htlr_fit_helper (main.cpp:7-25) compiled verbatim against the shims (impl="reference"),
    or - same C entry point, same variate hooks - the repo's Rcpp glue above the CUDA library (impl="glue",
    oracle/_glue/libhtlr_glue_test.so; `options` = R options for it, e.g. {"HTLR.cuda.rng": "R"})."""
    if impl == "glue":
        path = os.path.join(_HERE, "_glue", "libhtlr_glue_test.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/_glue/libhtlr_glue_test.so not built (make -C oracle glue)")
        r = ctypes.CDLL(path)
        r.htlr_ref_set_option.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
        for k in ("HTLR.cuda.rng", "HTLR.cuda.device"):
            r.htlr_ref_set_option(k.encode(), str((options or {}).get(k, "")).encode())
    else:
        r = _ref_lib("libhtlr_ref.so")
    if r is None:
        raise RuntimeError("oracle/_ref/libhtlr_ref.so not built (needs /root/reference at build time)")
    r.htlr_ref_last_error.restype = ctypes.c_char_p
    r.htlr_ref_log.restype = ctypes.c_char_p
    nvar = p + 1
    H = iters_rmc + 1 + (iters_h if keep_warmup_hist else 0)
    Xf = np.asfortranarray(X, dtype=np.float64)
    ym = np.asfortranarray(np.reshape(ymat, (n, K), order="F"), dtype=np.float64)
    yb = np.ascontiguousarray(ybase, dtype=np.int64)
    d0 = np.asfortranarray(np.reshape(deltas, (nvar, K), order="F"), dtype=np.float64)
    s0 = _f64(sigmasbt)
    cfg = RefCfg(p, K, n, PTYPES[ptype] if isinstance(ptype, str) else ptype, alpha, s, eta, iters_rmc, iters_h,
                 thin, leap_L, leap_L_h, leap_step, hmc_sgmcut, int(keep_warmup_hist), int(legacy), silence, 0)
    out = {"mcdeltas": np.zeros((nvar, K, H), order="F"), "mcsigmasbt": np.zeros((nvar, H), order="F"),
           "mcvardeltas": np.zeros((nvar, H), order="F"), "mclogw": np.zeros(H), "mcloglike": np.zeros(H),
           "mcuvar": np.zeros(H), "mchmcrej": np.zeros(H), "DDNloglike": np.zeros(nvar)}
    st = [None if a is None else _f64(a) for a in (normals, uniforms, gammas)]
    consumed = np.zeros(3, dtype=np.int64)
    secs = ctypes.c_double()
    rc = r.htlr_ref_fit(ctypes.byref(cfg), _p(Xf), _p(ym), _p(yb, _lp), _p(d0), _p(s0),
                        _p(st[0]), ctypes.c_int64(0 if st[0] is None else st[0].size),
                        _p(st[1]), ctypes.c_int64(0 if st[1] is None else st[1].size),
                        _p(st[2]), ctypes.c_int64(0 if st[2] is None else st[2].size),
                        _p(out["mcdeltas"]), _p(out["mcsigmasbt"]), _p(out["mcvardeltas"]), _p(out["mclogw"]),
                        _p(out["mcloglike"]), _p(out["mcuvar"]), _p(out["mchmcrej"]), _p(out["DDNloglike"]),
                        _p(consumed, _lp), ctypes.byref(secs))
    if rc:
        raise RuntimeError(r.htlr_ref_last_error().decode())
    out["consumed"] = consumed
    out["secs"] = secs.value
    out["log"] = r.htlr_ref_log().decode()
    return out
