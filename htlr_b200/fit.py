"""Host-side mirror of the reference's sampler entry point for the CUDA path.

`htlr_fit_helper(...)` has the 22 arguments of the reference's C++ function of the same name
(/root/reference/src/main.cpp:7-14, called from R/core.r:198-213) with the same meaning, and returns a
dict with the names and shapes of `Fit::OutputR` (src/gibbs.cpp:128-152): p, n, K, mc.param{...},
mcdeltas (nvar x K x H), mclogw, mcsigmasbt, mcvardeltas, mcloglike, mcuvar, mchmcrej.

`Sampler` exposes the same object as three phases (create / run in chunks / fetch), which is what the
C ABI offers an Rcpp host so that R stays interruptible.

Everything numeric happens in libhtlr_cuda.so on the GPU; this file only marshals arrays.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import HtlrCfg, HtlrInject, HtlrIterStats, HtlrProfile, c_dp, c_ip, c_lp

PTYPES = {"t": 0, "ghs": 1, "neg": 2}
RNG_DEVICE, RNG_INJECT = 0, 1


class HtlrCudaError(RuntimeError):
    """Raised for every non-zero status of the C ABI (the Rcpp glue would call Rcpp::stop)."""

    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


def _ptr(a, typ=c_dp):
    return None if a is None else a.ctypes.data_as(typ)


def _f64(a, order="C"):
    return np.require(a, dtype=np.float64, requirements=["F" if order == "F" else "C", "A"])


def _inject_struct(normals=None, uniforms=None, gammas=None, ars_uniforms=None):
    keep = [None if a is None else np.ascontiguousarray(a, dtype=np.float64)
            for a in (normals, uniforms, gammas, ars_uniforms)]
    inj = HtlrInject()
    inj.normals, inj.n_normals = _ptr(keep[0]), 0 if keep[0] is None else keep[0].size
    inj.uniforms, inj.n_uniforms = _ptr(keep[1]), 0 if keep[1] is None else keep[1].size
    inj.gammas, inj.n_gammas = _ptr(keep[2]), 0 if keep[2] is None else keep[2].size
    if keep[3] is not None:
        inj.ars_uniforms = _ptr(keep[3])
        inj.n_ars_blocks = keep[3].size // keep[3].shape[-1]
    return inj, keep


class Sampler:
    """One chain on one GPU (or one row shard of it). Mirrors `class Fit` (src/gibbs.h:91-100)."""

    def __init__(self, p, K, n, X, ymat, ybase, ptype, alpha, s, eta, iters_rmc, iters_h, thin, leap_L, leap_L_h,
                 leap_step, hmc_sgmcut, deltas, sigmasbt, keep_warmup_hist=False, silence=1, legacy=False, *,
                 device=0, rng_mode=RNG_DEVICE, seed=None, algo=0, use_graph=True, nccl_comm=None,
                 ars_inject_width=0, device_ptrs=None):
        """`device_ptrs=(X_ptr, ldx, ymat_ptr, ybase_ptr)`: the three n-sized inputs already live in device
        memory (column-major FP64 with even leading dimension ldx; ybase int64) and `X`, `ymat`, `ybase` are
        ignored — used for tall data generated on the GPU (htlr_cfg.x_on_device)."""
        L = _lib.lib()
        if seed is None:
            # the reference draws from R's global stream: two fits that do not fix a seed differ. Same here: a fresh
            # key per sampler unless one is given (kept in self.seed and in the fit's "mc.param")
            seed = int(np.random.SeedSequence().entropy) & ((1 << 64) - 1)
        self.seed = int(seed)
        if isinstance(ptype, str):
            if ptype not in PTYPES:
                raise HtlrCudaError(1, f"Unsupported prior type {ptype}")
            ptype = PTYPES[ptype]
        if device_ptrs is None:
            X = _f64(X, "F")
            if X.shape != (n, p + 1):
                raise HtlrCudaError(1, f"X must be n x (p+1) = {n} x {p + 1}, got {X.shape}")
            ymat = _f64(np.reshape(ymat, (n, K), order="F"), "F")
            ybase = np.ascontiguousarray(ybase, dtype=np.int64)
            if ybase.shape != (n,):
                raise HtlrCudaError(1, "ybase length mismatch")
            xp, ldx, ymp, ybp = _ptr(X), n, _ptr(ymat), _ptr(ybase, c_lp)
        else:
            xp, ldx, ymp, ybp = (ctypes.cast(device_ptrs[0], c_dp), int(device_ptrs[1]),
                                 ctypes.cast(device_ptrs[2], c_dp), ctypes.cast(device_ptrs[3], c_lp))
        deltas = _f64(np.reshape(deltas, (p + 1, K), order="F"), "F")
        sigmasbt = _f64(sigmasbt)
        if sigmasbt.shape != (p + 1,):
            raise HtlrCudaError(1, "sigmasbt length mismatch")
        self.p, self.K, self.n, self.nvar, self.C = p, K, n, p + 1, K + 1
        self.iters_rmc, self.iters_h, self.thin = iters_rmc, iters_h, thin
        self.leap_L, self.leap_L_h, self.leap_step = leap_L, leap_L_h, leap_step
        self.hmc_sgmcut = hmc_sgmcut
        cfg = HtlrCfg()
        cfg.struct_size, cfg.abi_version = ctypes.sizeof(HtlrCfg), _lib.ABI_VERSION
        cfg.p, cfg.K, cfg.n, cfg.ptype = p, K, n, ptype
        cfg.alpha, cfg.s, cfg.eta = alpha, s, eta
        cfg.iters_rmc, cfg.iters_h, cfg.thin, cfg.leap_L, cfg.leap_L_h = iters_rmc, iters_h, thin, leap_L, leap_L_h
        cfg.leap_step, cfg.hmc_sgmcut = leap_step, hmc_sgmcut
        cfg.keep_warmup_hist, cfg.silence, cfg.legacy = int(keep_warmup_hist), int(silence), int(legacy)
        cfg.device, cfg.rng_mode, cfg.algo, cfg.seed = device, rng_mode, algo, seed
        cfg.use_graph, cfg.x_on_device = int(use_graph), 0 if device_ptrs is None else 1
        cfg.nccl_comm = nccl_comm
        cfg.ars_inject_width = ars_inject_width
        self.cfg = cfg
        self.h = ctypes.c_void_p()
        rc = L.htlr_cuda_create(ctypes.byref(cfg), xp, ldx, ymp, ybp, _ptr(deltas), _ptr(sigmasbt),
                                ctypes.byref(self.h))
        if rc:
            raise HtlrCudaError(rc, L.htlr_cuda_last_error(None).decode())
        self.hist_len = L.htlr_cuda_hist_len(self.h)
        self.iters_done = 0

    # -- life cycle
    def close(self):
        if getattr(self, "h", None):
            _lib.lib().htlr_cuda_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise HtlrCudaError(rc, _lib.lib().htlr_cuda_last_error(self.h).decode())

    # -- sampling
    def run(self, n_iters, normals=None, uniforms=None, gammas=None, ars_uniforms=None):
        """Next n_iters passes of StartSampling's loop body; returns the per-iteration scalars."""
        st = (HtlrIterStats * max(n_iters, 1))()
        inj = None
        if self.cfg.rng_mode == RNG_INJECT:
            inj, keep = _inject_struct(normals, uniforms, gammas, ars_uniforms)
            inj = ctypes.byref(inj)
        self._check(_lib.lib().htlr_cuda_run(self.h, n_iters, inj, st))
        self.iters_done += n_iters
        return {"loglike": np.array([st[i].loglike for i in range(n_iters)]),
                "logw": np.array([st[i].logw for i in range(n_iters)]),
                "uvar": np.array([st[i].uvar for i in range(n_iters)]),
                "rej": np.array([st[i].rej for i in range(n_iters)])}

    def run_async(self, n_iters):
        self._check(_lib.lib().htlr_cuda_run_async(self.h, n_iters))
        self.iters_done += n_iters

    def sync(self):
        self._check(_lib.lib().htlr_cuda_sync(self.h))

    _TRACE = ("mcdeltas", "mcsigmasbt", "mcvardeltas", "mclogw", "mcloglike", "mcuvar", "mchmcrej")

    def trace_arrays(self):
        """Host arrays with the shapes of Fit::OutputR's trace members (gibbs.cpp:128-152)."""
        H, nvar, K = self.hist_len, self.nvar, self.K
        return {"mcdeltas": np.zeros((nvar, K, H), order="F"), "mcsigmasbt": np.zeros((nvar, H), order="F"),
                "mcvardeltas": np.zeros((nvar, H), order="F"), "mclogw": np.zeros((H, 1)),
                "mcloglike": np.zeros((H, 1)), "mcuvar": np.zeros((H, 1)), "mchmcrej": np.zeros((H, 1))}

    def collect(self, upto_iters, into, stats_from=None):
        """Pipelined download (htlr_cuda_collect): waits for the first `upto_iters` iterations only, copies the trace
        slices completed since the last collect into `into` (from trace_arrays(); the same dict every call) and
        returns the per-iteration scalars of iterations [stats_from, upto_iters)."""
        ns = 0 if stats_from is None else upto_iters - stats_from
        st = (HtlrIterStats * max(ns, 1))()
        self._check(_lib.lib().htlr_cuda_collect(self.h, upto_iters, 0 if stats_from is None else stats_from,
                                                 st if ns > 0 else None, *[_ptr(into[k]) for k in self._TRACE]))
        return {"loglike": np.array([st[i].loglike for i in range(ns)]),
                "logw": np.array([st[i].logw for i in range(ns)]),
                "uvar": np.array([st[i].uvar for i in range(ns)]),
                "rej": np.array([st[i].rej for i in range(ns)])}

    def fetch(self, collected=None):
        """Fit::OutputR (gibbs.cpp:128-152). `collected`: trace arrays already filled by collect() up to the last
        iteration — only the remaining members are downloaded then."""
        nvar = self.nvar
        ddn = np.zeros((1, nvar))
        if collected is not None:
            o = collected
            null = ctypes.cast(None, c_dp)
            self._check(_lib.lib().htlr_cuda_fetch(self.h, null, null, null, null, null, null, null, _ptr(ddn)))
        else:
            o = self.trace_arrays()
            self._check(_lib.lib().htlr_cuda_fetch(self.h, _ptr(o["mcdeltas"]), _ptr(o["mcsigmasbt"]),
                                                   _ptr(o["mcvardeltas"]), _ptr(o["mclogw"]), _ptr(o["mcloglike"]),
                                                   _ptr(o["mcuvar"]), _ptr(o["mchmcrej"]), _ptr(ddn)))
        sg = self.hmc_sgmcut
        out = {"p": self.p, "n": self.n, "K": self.K,
               "mc.param": {"iter.rmc": self.iters_rmc, "iter.warm": self.iters_h, "thin": self.thin,
                            "leap": self.leap_L, "leap.warm": self.leap_L_h, "leap.step": self.leap_step,
                            "sgmsq.cut": sg * sg if sg > 0 else sg, "DDNloglike": ddn},
               "cuda.seed": self.seed}
        out.update(o)
        return out

    def summary(self, first=-1, last=-1, thin=1, cut=0.1):
        """Posterior mean coefficients, htlr_sdb and nzero_idx (R/mccoef.r:38-62, 119-156) computed on the device
        from the resident trace. Default slice range = htlr_sdb's (warm-up and a further 20 % dropped).
        `nzero_idx` holds 1-based feature numbers like R's which()."""
        nvar, K = self.nvar, self.K
        mean = np.zeros((nvar, K), order="F")
        sdb = np.zeros(self.p)
        flags = np.zeros(self.p, dtype=np.int32)
        self._check(_lib.lib().htlr_cuda_summary(self.h, first, last, thin, float(cut), _ptr(mean), _ptr(sdb),
                                                 _ptr(flags, c_ip)))
        return {"mean_deltas": mean, "sdb": sdb, "nzero_idx": np.nonzero(flags)[0] + 1}

    def predict(self, X_ts, first=-1, last=-1, thin=1):
        """htlr_predict (R/predict.R:28-92) from the device-resident trace. `X_ts`: n_ts x (p+1), prepared like X
        (selected features, standardised, intercept column). Returns n_ts x C probabilities."""
        X_ts = _f64(np.atleast_2d(X_ts), "F")
        if X_ts.shape[1] != self.nvar:
            raise HtlrCudaError(1, f"X_ts must have p+1 = {self.nvar} columns")
        n_ts = X_ts.shape[0]
        probs = np.zeros((n_ts, self.C), order="F")
        self._check(_lib.lib().htlr_cuda_predict(self.h, _ptr(X_ts), n_ts, n_ts, first, last, thin, _ptr(probs)))
        return probs

    def predict_stats(self):
        """(kernel milliseconds, FP64 flops) of the last predict() call."""
        ms, fl = ctypes.c_double(), ctypes.c_double()
        bd = lambda x: ctypes.cast(ctypes.byref(x), c_dp)
        self._check(_lib.lib().htlr_cuda_predict_stats(self.h, bd(ms), bd(fl)))
        return ms.value, fl.value

    # -- deterministic pieces (parity hooks)
    def eval(self, iup, deltas):
        n, C, nvar, K = self.n, self.C, self.nvar, self.K
        iup = np.ascontiguousarray(iup, dtype=np.int32)
        d = _f64(deltas, "F")
        u = iup.size
        lv, pred, grad = np.zeros((n, C), order="F"), np.zeros((n, C), order="F"), np.zeros((u, K), order="F")
        ll = ctypes.c_double()
        self._check(_lib.lib().htlr_cuda_eval(self.h, _ptr(iup, c_ip), u, _ptr(d), _ptr(lv), _ptr(pred),
                                              ctypes.cast(ctypes.byref(ll), c_dp), _ptr(grad)))
        return {"lv": lv, "pred": pred, "loglike": ll.value, "grad": grad}

    def trajectory(self, momt, L):
        n, C, nvar, K = self.n, self.C, self.nvar, self.K
        m = _f64(momt, "F")
        o = {"deltas": np.zeros((nvar, K), order="F"), "momt": np.zeros((nvar, K), order="F"),
             "lv": np.zeros((n, C), order="F"), "var_deltas": np.zeros(nvar)}
        ll, e0, e1, nu = ctypes.c_double(), ctypes.c_double(), ctypes.c_double(), ctypes.c_int32()
        bd = lambda x: ctypes.cast(ctypes.byref(x), c_dp)
        self._check(_lib.lib().htlr_cuda_trajectory(self.h, _ptr(m), L, _ptr(o["deltas"]), _ptr(o["momt"]),
                                                    _ptr(o["lv"]), bd(ll), _ptr(o["var_deltas"]), bd(e0), bd(e1),
                                                    ctypes.cast(ctypes.byref(nu), c_ip)))
        o.update(loglike=ll.value, nenergy_old=e0.value, nenergy_new=e1.value, nuvar=nu.value)
        return o

    def state(self):
        n, C, nvar, K = self.n, self.C, self.nvar, self.K
        o = {"lv": np.zeros((n, C), order="F"), "deltas": np.zeros((nvar, K), order="F"),
             "sigmasbt": np.zeros(nvar), "var_deltas": np.zeros(nvar)}
        ll, lw = ctypes.c_double(), ctypes.c_double()
        bd = lambda x: ctypes.cast(ctypes.byref(x), c_dp)
        self._check(_lib.lib().htlr_cuda_get_state(self.h, _ptr(o["lv"]), _ptr(o["deltas"]), _ptr(o["sigmasbt"]),
                                                   _ptr(o["var_deltas"]), bd(ll), bd(lw)))
        o.update(loglike=ll.value, logw=lw.value)
        return o

    # -- measurement
    @property
    def stream(self):
        return _lib.lib().htlr_cuda_stream(self.h)

    def profile_enable(self, on=True):
        self._check(_lib.lib().htlr_cuda_profile_enable(self.h, int(on)))

    def profile_get(self, reset=False):
        pr = HtlrProfile()
        self._check(_lib.lib().htlr_cuda_profile_get(self.h, ctypes.byref(pr), int(reset)))
        out = {k: getattr(pr, k) for k in ("sweep_ms", "sweep_bytes", "sweep_launches", "total_ms",
                                           "kernel_launches", "sigma_ms", "fused_sweep_ms",
                                           "fused_barrier_ms", "fused_rowphase_ms", "small_kernel_trajectories",
                                           "allreduce_ms")}
        out["timeline_us"] = list(pr.timeline_us)
        return out

    @property
    def launch_count(self):
        return _lib.lib().htlr_cuda_launch_count(self.h)

    @property
    def ars_table_full(self):
        """ARS draws of this chain whose hull table (16 per feature, 32 for logw) filled up; see htlr_cuda.h."""
        return _lib.lib().htlr_cuda_ars_table_full(self.h)


def htlr_fit_helper(p, K, n, X, ymat, ybase, ptype, alpha, s, eta, iters_rmc, iters_h, thin, leap_L, leap_L_h,
                    leap_step, hmc_sgmcut, deltas, sigmasbt, keep_warmup_hist, silence, legacy, *, device=0,
                    seed=None, progress=None, **kw):
    """Drop-in for the reference's `htlr_fit_helper` (src/main.cpp:7-25): Fit ctor, StartSampling, OutputR.

    Pipelined like htlr_cuda_fit: units of 8 iterations, up to 64 in flight, each unit's trace slices downloaded
    while the following units run; `progress(done, stats)` (the reference polls for interrupts every 256
    iterations, gibbs.cpp:124) is called per unit; when `silence == 0` the reference's progress line
    (gibbs.cpp:119-121) is printed per iteration. Injected random streams are positional, so that mode runs in
    one blocking call.
    """
    streams = {k: kw.pop(k, None) for k in ("normals", "uniforms", "gammas", "ars_uniforms")}
    smp = Sampler(p, K, n, X, ymat, ybase, ptype, alpha, s, eta, iters_rmc, iters_h, thin, leap_L, leap_L_h,
                  leap_step, hmc_sgmcut, deltas, sigmasbt, keep_warmup_hist, silence, legacy, device=device,
                  seed=seed, **kw)
    try:
        total, done, enq = iters_h + iters_rmc, 0, 0
        inject = smp.cfg.rng_mode == RNG_INJECT
        want_stats = silence == 0 or progress is not None
        out = None if inject else smp.trace_arrays()
        unit, depth = 8, 64
        while done < total:
            if inject:
                upto = total
                st = smp.run(total, **streams)
            else:
                while enq < total and enq - done < depth:
                    nq = min(unit, total - enq)
                    smp.run_async(nq)
                    enq += nq
                upto = min(done + unit, total)
                st = smp.collect(upto, out, stats_from=done if want_stats else None)
            if silence == 0:
                for i in range(upto - done):
                    print("Iter%4d: deviance=%5.3f, logw=%6.2f, nuvar=%3.0f, hmcrej=%4.2f"
                          % (done + i - iters_h, -st["loglike"][i] / n, st["logw"][i], st["uvar"][i], st["rej"][i]))
            done = upto
            if progress is not None and progress(done, st):
                raise HtlrCudaError(5, "interrupted")
        return smp.fetch(collected=out)
    finally:
        smp.close()


def ars_sample(kind, var_deltas, K, alpha, log_aw, uniforms=None, seed=0, step=0, device=0):
    """Warp-per-feature ARS draw of x = log sigma^2_j for each feature (UpdateSigmasSgm, gibbs.cpp:351-372)."""
    L = _lib.lib()
    v = np.ascontiguousarray(var_deltas, dtype=np.float64)
    m = v.size
    u = None if uniforms is None else np.ascontiguousarray(uniforms, dtype=np.float64)
    width = 0 if u is None else u.shape[1]
    x = np.zeros(m)
    nu, ne, st = np.zeros(m, np.int32), np.zeros(m, np.int32), np.zeros(m, np.int32)
    k = PTYPES[kind] if isinstance(kind, str) else kind
    rc = L.htlr_cuda_ars(device, k, m, _ptr(v), float(K), float(alpha), float(log_aw), _ptr(u), width, seed, step,
                         _ptr(x), _ptr(nu, c_ip), _ptr(ne, c_ip), _ptr(st, c_ip))
    if rc:
        raise HtlrCudaError(rc, L.htlr_cuda_last_error(None).decode())
    return {"x": x, "n_unif": nu, "n_eval": ne, "status": st}


def gendata_mlr_device(n, p, NC, X_ptr, ldx, ymat_ptr, ybase_ptr, nu=2.0, w=1.0, seed=0, row0=0, device=0,
                       want_deltas=False):
    """gendata_MLR's law (R/gendata.r:26-46) for rows [row0, row0 + n) of a data set, generated on the GPU into the
    caller's device buffers (X n x (p+1) col-major with leading dimension ldx, ymat n x (NC-1), ybase int64)."""
    d = np.zeros((p + 1, NC - 1), order="F") if want_deltas else None
    rc = _lib.lib().htlr_cuda_gendata_mlr(device, n, p, NC, float(nu), float(w), seed, row0, X_ptr, ldx, ymat_ptr,
                                          ybase_ptr, _ptr(d))
    if rc:
        raise HtlrCudaError(rc, _lib.lib().htlr_cuda_last_error(None).decode())
    return d


def gendata_fam_device(n, muj, Ablock, X_ptr, ldx, ymat_ptr, ybase_ptr, sd_g=0.0, stdx=False, seed=0, row0=0, device=0):
    """gendata_FAM's law (R/gendata.r:105-113, src/utils.cpp:72-105) with A = I_p + a leading pb x kb block."""
    muj = _f64(muj, "F")
    p, NC = muj.shape
    Ab = _f64(np.atleast_2d(Ablock), "F")
    pb, kb = Ab.shape
    rc = _lib.lib().htlr_cuda_gendata_fam(device, n, p, NC, _ptr(muj), pb, kb, _ptr(Ab), float(sd_g), int(stdx), seed,
                                          row0, X_ptr, ldx, ymat_ptr, ybase_ptr)
    if rc:
        raise HtlrCudaError(rc, _lib.lib().htlr_cuda_last_error(None).decode())


def fp64_peak(device=0):
    """Measured FP64 peaks of the device in TFLOP/s: {"dfma": vector pipe, "dmma": tensor-core m8n8k4}."""
    a, b = ctypes.c_double(), ctypes.c_double()
    bd = lambda x: ctypes.cast(ctypes.byref(x), c_dp)
    rc = _lib.lib().htlr_cuda_fp64_peak(device, bd(a), bd(b))
    if rc:
        raise HtlrCudaError(rc, _lib.lib().htlr_cuda_last_error(None).decode())
    return {"dfma": a.value, "dmma": b.value}


def rng_dump(seed, purpose, step, n_a, n_b, kind, shape=1.0, device=0):
    L = _lib.lib()
    out = np.zeros((n_a, n_b))
    rc = L.htlr_cuda_rng_dump(device, seed, purpose, step, n_a, n_b, {"unif": 0, "norm": 1, "gamma": 2}[kind],
                              float(shape), _ptr(out))
    if rc:
        raise HtlrCudaError(rc, L.htlr_cuda_last_error(None).decode())
    return out


def std(A, device=0):
    """`std_helper` (src/utils.cpp:36-48) on the GPU: returns {"median", "sd", "X"} for an n x p matrix."""
    L = _lib.lib()
    A = _f64(A, "F")
    n, p = A.shape
    out = np.zeros((n, p), order="F")
    med, sd = np.zeros(p), np.zeros(p)
    rc = L.htlr_cuda_std(device, n, p, A.ctypes.data, n, 0, out.ctypes.data, n, _ptr(med), _ptr(sd))
    if rc:
        raise HtlrCudaError(rc, L.htlr_cuda_last_error(None).decode())
    return {"median": med, "sd": sd, "X": out}
