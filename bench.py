#!/usr/bin/env python
"""Benchmark of the HTLR sampler hot path on B200: MCMC iterations/sec + X-sweep HBM roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one sampling-phase MCMC iteration (gibbs.cpp:57-125 with L = leap = 50 leapfrog steps):
restricted-set selection, momenta, trajectory (X sweeps, softmax, gradient), Metropolis test, per-feature
variance update, trace recording. Workloads (BASELINE.json configs):

  C_full     gendata_FAM n=1000 p=20000 C=3, t prior df=1, htlr() defaults except cut=0: every feature is
             updated every iteration, so each leapfrog step sweeps all 160 MB of X (> 126 MB L2). This is
             the regime that exposes the HBM roofline north_star quotes; the default workload.
  C_default  same data, htlr() defaults (cut=0.05): restricted regime, latency-bound. Always measured too
             and reported in the "default_cut" block of the same line, with its own CPU baseline.
  B / D      gendata_MLR n=500 p=5000 C=4 / n=200 p=50000 C=2 (one chain).

N > 1 (torchrun): the headline `value` is independent chains of the default workload, one per GPU, no communication
(SURVEY 8e "replicas"), summed over ranks. The same line carries three more blocks, measured inside the same process
group after the replica measurement (each with its own clock record):

  sharded_parity  a small chain row-sharded over all N ranks (NCCL all-reduce of the gradient partials in every
                  leapfrog step) against the same chain on rank 0 alone: identical accept history and restricted-set
                  sizes, coefficients to 1e-10
  E_row_sharded   config E at full size (n = 200000, p = 10000, C = 5), rows sharded over the N ranks: the strong-
                  scaling point for this N (at N = 1 the single-GPU value, so the curve has an origin)
  D_chains        config D: 4 chains per GPU running concurrently, no communication

Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "C_full": dict(gen="fam", n=1000, p=20000, C=3, cut=0.0, seed=1003, leap_step=0.05,
                   desc="gendata_FAM n=1000 p=20000 C=3, t prior df=1, leap=50, cut=0 (all features updated: full X sweep "
                        "per leapfrog step), leap.stepsize=0.05 (the default 0.3 rejects every proposal when all 20001 "
                        "features move; same cost per iteration)"),
    "C_default": dict(gen="fam", n=1000, p=20000, C=3, cut=0.05, seed=1003,
                      desc="gendata_FAM n=1000 p=20000 C=3, t prior df=1, leap=50, cut=0.05 (htlr() defaults, restricted regime)"),
    "B": dict(gen="mlr", n=500, p=5000, C=4, cut=0.0, seed=1002, leap_step=0.05,
              desc="gendata_MLR n=500 p=5000 C=4, t prior df=1, leap=50, cut=0 (all features updated), leap.stepsize=0.05"),
    "B_default": dict(gen="mlr", n=500, p=5000, C=4, cut=0.05, seed=1002,
                      desc="gendata_MLR n=500 p=5000 C=4, t prior df=1, leap=50, htlr() defaults (cut=0.05)"),
    "D": dict(gen="mlr", n=200, p=50000, C=2, cut=0.05, seed=1004,
              desc="independent chains of gendata_MLR n=200 p=50000 C=2, t prior df=1, leap=50, htlr() defaults (cut=0.05), "
                   "several chains per GPU running concurrently, no communication"),
    "A": dict(gen="colon", n=62, p=2000, C=2, cut=0.05, seed=1001,
              desc="colon (n=62, p=2000, binary), t prior df=1, leap=50, htlr() defaults (cut=0.05)"),
    "tiny": dict(gen="mlr", n=120, p=400, C=3, cut=0.0, seed=5, desc="tiny smoke workload"),
    # config E: ONE chain over all ranks, X row-sharded, NCCL all-reduce of the u x K gradient per leapfrog step
    "E": dict(gen="mlr_device", n=200000, p=10000, C=5, cut=0.0, seed=1005, leap_step=0.02,
              desc="tall gendata_MLR n=200000 p=10000 C=5 generated on device, t prior df=1, leap=50, cut=0, "
                   "leap.stepsize=0.02, row-sharded over the ranks with an NCCL all-reduce of the gradient partials"),
}
E_CPU_ROWS = 2000   # rows of the CPU-reference sample of workload E (rate scaled by E_CPU_ROWS / n)
CHAIN_WARM_FULL = 60   # untimed iterations ahead of the warm-up steps in the full regime (the chain starts at zero)
HTLR_DEFAULTS = dict(ptype="t", alpha=1.0, s=-10.0, eta=0.0, thin=1, leap_L=50, leap_L_h=5, leap_step=0.3,
                     keep_warmup_hist=False, silence=1, legacy=False)


def make_problem(wl, chain_seed=0):
    from htlr_b200 import synth
    w = WORKLOADS[wl]
    if w["gen"] == "colon":
        # config A: the reference's own data set (data/colon.rda, parsed once into tests/golden/colon.npz by
        # tests/golden/make_colon.py), prepared as R/core.r:103-146 does
        d = np.load(os.path.join(ROOT, "tests", "golden", "colon.npz"))
        Xc, yc = np.asarray(d["X"], dtype=np.float64), d["y"]
        X, ymat, ybase, K = synth.make_inputs(synth.std_median_sd(Xc), yc)
        deltas, sig = synth.init_state(w["p"], K, init="zero", seed=w["seed"] + 17 + chain_seed)
        return dict(p=w["p"], K=K, n=w["n"], X=X, ymat=ymat, ybase=ybase, deltas=deltas, sigmasbt=sig,
                    hmc_sgmcut=w["cut"], **HTLR_DEFAULTS)
    if w["gen"] == "mlr_device":
        # CPU legs only: a host-generated row subsample of the same law (the reference's cost is linear in n)
        w = dict(w, n=E_CPU_ROWS)
        d = synth.gendata_mlr(w["n"], w["p"], NC=w["C"], seed=w["seed"])
        d["y"][:w["C"]] = np.arange(w["C"])
    elif w["gen"] == "fam":
        d = synth.gendata_fam(w["n"], w["p"], sd_g=0.5, stdx=True, seed=w["seed"])
    else:
        d = synth.gendata_mlr(w["n"], w["p"], NC=w["C"], seed=w["seed"])
        d["y"][:w["C"]] = np.arange(w["C"])
    X, ymat, ybase, K = synth.make_inputs(d["X"], d["y"])
    deltas, sig = synth.init_state(w["p"], K, init="zero", seed=w["seed"] + 17 + chain_seed)
    out = dict(p=w["p"], K=K, n=w["n"], X=X, ymat=ymat, ybase=ybase, deltas=deltas, sigmasbt=sig,
               hmc_sgmcut=w["cut"], **HTLR_DEFAULTS)
    if "leap_step" in w:
        out["leap_step"] = w["leap_step"]
    return out


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel_key):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full summary, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(kernel_key)
    except Exception:
        return None


def workload_config(wl):
    """The `config` object of a line: what is measured, nothing about how a run went - both arms print the same one."""
    w = WORKLOADS[wl]
    return {"workload": w["desc"], "regime": "full" if w["cut"] <= 0 else "restricted"}


# ---------------------------------------------------------------------------------------------------
def cpu_reference_rate(wl, max_iters, budget_s, warm_iters):
    """The reference's CPU sampler on this box's host cores, single thread (its hot loops are
    single-threaded): oracle/_ref (reference sources compiled verbatim against the API
    shims) when it was built, else the oracle port. Time-boxed sample of the same workload."""
    from oracle import oracle as O
    args = make_problem(wl)
    if O.have_ref():
        kind = "reference"
        def fit(iters_h, iters_rmc):
            a = dict(args, iters_h=iters_h, iters_rmc=iters_rmc)
            return O.ref_fit(**a)["secs"]
        base = fit(warm_iters, 1)                      # setup + warm-up + 1 sampling iteration
        m = int(max(1, min(max_iters, (budget_s - base) / max(base, 1e-3) - 1)))
        tm = fit(warm_iters, 1 + m) - base             # m more sampling iterations of the same chain
        rate = m / max(tm, 1e-9)
    else:
        kind = "port"
        f = O.OracleFit(**dict(args, iters_h=warm_iters, iters_rmc=max_iters))
        f.set_seed(1)
        f.initialize()
        if warm_iters:
            f.run(warm_iters)
        m, tm = 0, 0.0
        while m < max_iters and tm < budget_s:
            tm += f.run(1)
            m += 1
        rate = m / tm
    return rate, kind, m


def _ref_worker(wl, warm, m, q):
    """One independent reference chain in its own process: seconds for m sampling iterations."""
    from oracle import oracle as O
    args = make_problem(wl)
    base = O.ref_fit(**dict(args, iters_h=warm, iters_rmc=1))["secs"]
    tot = O.ref_fit(**dict(args, iters_h=warm, iters_rmc=1 + m))["secs"]
    q.put(max(tot - base, 1e-9))


def all_cores_reference(wl, warm, m, cores):
    """The reference's sampler is single-threaded, so the only way it can use every host core is one
    independent chain per core: aggregate iterations/sec of `cores` concurrent chains (they share memory bandwidth)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    ps = [ctx.Process(target=_ref_worker, args=(wl, warm, m, q)) for _ in range(cores)]
    for p in ps:
        p.start()
    secs = [q.get(timeout=1800) for _ in ps]
    for p in ps:
        p.join()
    return cores * m / max(secs)


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[a.workload]
    warm = 0 if w["cut"] <= 0 else 150
    rate, kind, m = cpu_reference_rate(a.workload, a.steps, budget_s=150.0, warm_iters=warm)
    cores = os.cpu_count()
    scale_note = ""
    if a.workload == "E":
        n_total = a.n_total or w["n"]
        rate *= E_CPU_ROWS / n_total
        scale_note = f"; timed on a {E_CPU_ROWS}-row sample of the same law and scaled linearly to n={n_total}"
    line = {"impl": "reference", "metric": "mcmc_iterations_per_sec", "value": rate, "unit": "it/s",
            "n_gpus": a.gpus, "steps": m, "warmup": a.warmup, "chain_warmup_iterations": warm, "ms_per_step": 1e3 / rate, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(a.workload),
            "cpu_baseline": {"value": rate, "unit": "it/s", "cores": 1, "kind": kind,
                             "sample": f"{m} sampling iterations (L=50) of the same workload, time-boxed to ~150 s; "
                                       f"single thread of {cores} host cores (the reference's hot loops are single-threaded)"
                                       + scale_note},
            "e2e": {"value": rate, "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    from oracle import oracle as O
    if cores and cores > 1 and O.have_ref() and a.workload != "E":
        mm = max(1, min(m, 3))
        agg = all_cores_reference(a.workload, warm, mm, cores)
        line["all_cores"] = {"value": agg, "unit": "it/s (sum over chains)", "cores": cores,
                             "what": f"{cores} independent reference chains, one per host core, {mm} sampling iterations "
                                     f"each, run concurrently: the only way the single-threaded reference can use every "
                                     f"core; `value` above is one chain, which is what one htlr() call runs"}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
def time_chain(torch, smp, steps, dist=None):
    """K iterations enqueued back to back, CUDA events on the library's own stream."""
    st = torch.cuda.ExternalStream(smp.stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    smp.sync()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record(st)
    smp.run_async(steps)
    e1.record(st)
    smp.sync()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    return e0.elapsed_time(e1)  # ms


ALGO = 0


def measure_workload(torch, wl, steps, warmup, device, chain_seed, dist, want_profile=True):
    import htlr_b200
    w = WORKLOADS[wl]
    args = make_problem(wl, chain_seed)
    warm_iters = CHAIN_WARM_FULL if w["cut"] <= 0 else 150   # let the chain (and the restricted set) settle first:
    total = warm_iters + warmup + steps                      # from a zero state the first proposals are all rejected
    prof_steps = min(steps, 20) if want_profile else 0
    smp = htlr_b200.Sampler(**dict(args, iters_h=warm_iters, iters_rmc=total + prof_steps), device=device,
                            seed=1000 + chain_seed, algo=ALGO)
    if warm_iters:
        smp.run(warm_iters)
    st_w = smp.run(max(warmup, 3))
    l0 = smp.launch_count
    ms = time_chain(torch, smp, steps, dist)
    launches = smp.launch_count - l0
    res = {"ms": ms, "launches": launches}
    if want_profile:
        smp.profile_enable(True)
        smp.profile_get(reset=True)
        smp.run(prof_steps)
        res["prof"] = smp.profile_get()
        res["prof_steps"] = prof_steps
        smp.profile_enable(False)
    out = smp.fetch()
    lo = max(warmup, 3) + 1   # trace slot of the first timed iteration (slot 0 is the initial state; warm-up
                              # iterations are not recorded)
    res["uvar"] = float(np.mean(out["mcuvar"][lo:lo + steps]))
    res["rej"] = float(np.mean(out["mchmcrej"][lo:lo + steps]))
    res["n"], res["p"], res["K"] = args["n"], args["p"], args["K"]
    res["chain_warm"] = warm_iters
    smp.close()
    return res


def measure_e2e(torch, wl, steps, device, chain_seed, dist):
    """Same metric through the public API with HOST buffers: one htlr_fit_helper call = H2D of X / ymat /
    initial state from pinned host memory, `steps` sampling iterations, D2H of the whole trace."""
    import htlr_b200
    args = make_problem(wl, chain_seed)
    Xp = torch.from_numpy(np.ascontiguousarray(args["X"].T)).pin_memory()   # column-major storage, pinned
    args["X"] = Xp.numpy().T
    w = WORKLOADS[wl]
    warm_iters = 0 if w["cut"] <= 0 else 150
    a = dict(args, iters_h=warm_iters, iters_rmc=steps)   # (a whole call: the chain starts from the zero state)
    times = []
    for rep in range(3):   # median of three whole calls (each: create + H2D, iterations, D2H of the trace, destroy)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = htlr_b200.htlr_fit_helper(**a, device=device, seed=2000 + chain_seed + rep, algo=ALGO)
        t1 = time.perf_counter()
        times.append(t1 - t0)
    t0, t1 = 0.0, sorted(times)[1]
    nvar, K, n = args["p"] + 1, args["K"], args["n"]
    h2d = 8 * (n * nvar + n * K + n + nvar * K + nvar)
    H = steps + 1
    d2h = 8 * (nvar * K * H + 2 * nvar * H + 4 * H + nvar)
    assert np.isfinite(out["mcloglike"]).all()
    measure_e2e.calls_s = [round(t, 5) for t in times]
    return (t1 - t0), h2d / steps, d2h / steps, warm_iters


def measure_chains(torch, steps, warmup, nch, dist, rank, world, local, cpu_baseline=True):
    """Config D: independent chains, `nch` of them per GPU running CONCURRENTLY (each handle has its own stream;
    CLUSTER_ONLY keeps every chain's trajectory on one 16-SM cluster so they overlap), no communication. Also times the
    same chains one after the other for the concurrency gain. Returns the result dict on rank 0, None elsewhere."""
    import htlr_b200
    w = WORKLOADS["D"]
    warm_iters = 150
    total = warm_iters + warmup + 2 * steps
    smps = []
    for c in range(nch):
        args = make_problem("D", chain_seed=rank * nch + c)
        smps.append(htlr_b200.Sampler(**dict(args, iters_h=warm_iters, iters_rmc=total - warm_iters), device=local,
                                      seed=5000 + rank * nch + c, algo=4))
    for s in smps:
        s.run(warm_iters + warmup)
    streams = [torch.cuda.ExternalStream(s.stream) for s in smps]
    ev = lambda: torch.cuda.Event(enable_timing=True)
    clocks = ClockSampler(local)   # over both timed phases (the concurrent one alone can be shorter than a sample period)
    clocks.start()
    # one after the other
    seq_ms = 0.0
    for s, st in zip(smps, streams):
        e0, e1 = ev(), ev()
        s.sync(); e0.record(st); s.run_async(steps); e1.record(st); s.sync()
        seq_ms += e0.elapsed_time(e1)
    # together
    for s in smps:
        s.sync()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = sum(s.launch_count for s in smps)
    e0 = ev(); e0.record(streams[0])
    ends = []
    for s, st in zip(smps, streams):
        s.run_async(steps)
        e1 = ev(); e1.record(st); ends.append(e1)
    for s in smps:
        s.sync()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    clk = clocks.stop()
    ms = max(e0.elapsed_time(e) for e in ends)
    launches = sum(s.launch_count for s in smps) - l0
    uvar, rej = [], []
    for s in smps:
        o = s.fetch()
        uvar.append(float(np.mean(o["mcuvar"][-steps:])))
        rej.append(float(np.mean(o["mchmcrej"][-steps:])))
        s.close()
    t = torch.tensor([ms, seq_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank != 0:
        return None
    ms_max, seq_max = float(t[0]), float(t[1])
    out = {"metric": "mcmc_iterations_per_sec", "value": world * nch * steps / (ms_max * 1e-3), "unit": "it/s",
           "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_max / steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": dict(workload_config("D"), chains_per_gpu=nch, chains_total=world * nch),
           "run": {"l2": "X[:, iup] of every chain is L2 / shared-memory resident (stated; latency-bound regime)",
                   "parallelism": f"{nch} chains per GPU on separate streams, {world} GPU(s), no communication",
                   "algorithm": "cluster-mode fused trajectory kernel (one 16-CTA cluster per chain)",
                   "mean_nuvar": float(np.mean(uvar)), "hmc_reject": float(np.mean(rej)),
                   "chain_warmup_iterations": warm_iters},
           "one_after_the_other": {"value": world * nch * steps / (seq_max * 1e-3), "unit": "it/s",
                                   "concurrency_gain": seq_max / ms_max},
           "gpu_launches": launches, "clocks": clk,
           "roofline": None, "roofline_note": "latency-bound regime (X[:, iup] of a chain is a few MB, L2 / shared-memory "
                                              "resident); the HBM roofline is reported on the default workload"}
    if cpu_baseline:
        rate, kind, m = cpu_reference_rate("D", 40, budget_s=12.0, warm_iters=150)
        out["cpu_baseline"] = {"value": rate, "unit": "it/s", "cores": 1, "kind": kind,
                               "sample": f"one chain, {m} sampling iterations after 150 warm-up iterations, 1 of "
                                         f"{os.cpu_count()} host cores"}
    return out


def make_tall_on_device(torch, n_total, p, C, rank, world, seed):
    """gendata_MLR's law (R/gendata.r:26-46) for this rank's row block, generated on the GPU by the library's own
    generator (htlr_cuda_gendata_mlr: keyed by the GLOBAL row, so the ranks' blocks are slices of ONE data set whatever
    the sharding; 16 GB of X at full size never exists on the host)."""
    import htlr_b200
    from htlr_b200 import shard
    lo, hi = shard.row_range(n_total, rank, world)
    nl = hi - lo
    ld = (nl + 15) // 16 * 16
    dev = torch.device("cuda", torch.cuda.current_device())
    Xt = torch.zeros(p + 1, ld, dtype=torch.float64, device=dev)     # row j = column j of X (column-major n x nvar)
    ymat = torch.zeros(C - 1, nl, dtype=torch.float64, device=dev)   # [k][i]
    ybase = torch.zeros(nl, dtype=torch.int64, device=dev)
    htlr_b200.gendata_mlr_device(nl, p, C, Xt.data_ptr(), ld, ymat.data_ptr(), ybase.data_ptr(), seed=seed, row0=lo,
                                 device=dev.index)
    # coefficients this heavy-tailed give near-degenerate class probabilities; that is the law, keep it
    if rank == 0:   # every class present (R/core.r:90-100 requires two cases per class; the first rows make sure)
        ybase[:2 * C] = torch.arange(2 * C, device=dev) % C
        ymat[:, :2 * C] = (ybase[None, :2 * C] == torch.arange(1, C, device=dev)[:, None]).to(torch.float64)
    return dict(Xt=Xt, ld=ld, ymat=ymat, ybase=ybase, n_local=nl, K=C - 1)


def _peer_access(torch):
    try:
        return [bool(torch.cuda.can_device_access_peer(0, d)) for d in range(1, min(8, torch.cuda.device_count()))]
    except Exception as e:   # never let a diagnostic break the measurement
        return [repr(e)[:80]]


def measure_tall(torch, steps, warmup, n_total, dist, rank, world, local, cpu_baseline=True):
    """Config E: ONE chain, rows sharded over the ranks, NCCL all-reduce of the gradient partials in every leapfrog
    step. Strong scaling: total work fixed as N grows. Returns the result dict on rank 0, None elsewhere."""
    import htlr_b200
    from htlr_b200 import shard, synth
    w = WORKLOADS["E"]
    n_total = n_total or w["n"]
    p, C = w["p"], w["C"]
    d = make_tall_on_device(torch, n_total, p, C, rank, world, w["seed"])
    K, nl = d["K"], d["n_local"]
    deltas, sig = synth.init_state(p, K, init="zero", seed=w["seed"] + 17)
    psteps = min(steps, 5)
    total = warmup + steps + psteps
    comm = None
    if world > 1:
        uid = shard.exchange_unique_id(dist, shard.nccl_unique_id)
        comm = shard.nccl_init(uid, rank, world, local)
    kw = dict(HTLR_DEFAULTS, leap_step=w.get("leap_step", HTLR_DEFAULTS["leap_step"]))
    smp = htlr_b200.Sampler(p, K, nl, None, None, None, hmc_sgmcut=w["cut"], deltas=deltas, sigmasbt=sig, iters_h=0,
                            iters_rmc=total, device=local, seed=1000, nccl_comm=comm,
                            device_ptrs=(d["Xt"].data_ptr(), d["ld"], d["ymat"].data_ptr(), d["ybase"].data_ptr()),
                            **kw)
    smp.run(warmup)
    clocks = ClockSampler(local)
    clocks.start()
    l0 = smp.launch_count
    ms = time_chain(torch, smp, steps, dist)
    launches = smp.launch_count - l0
    clk = clocks.stop()
    smp.profile_enable(True)
    smp.profile_get(reset=True)
    smp.run(psteps)
    pr = smp.profile_get()
    smp.profile_enable(False)
    out = smp.fetch()
    smp.close()
    if comm is not None:
        shard.nccl_destroy(comm)
    del d
    torch.cuda.empty_cache()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    same = None
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # every rank must hold the same trace bit for bit (replicated state, identical keys): checked here at the full
        # shard size on the recorded log-likelihoods, restricted-set sizes, accept history and the last coefficient slice
        flat = np.concatenate([np.ravel(out[k]) for k in ("mcloglike", "mcuvar", "mchmcrej")] +
                              [np.ravel(out["mcdeltas"][:, :, -1])])
        t0 = torch.from_numpy(flat.copy()).cuda()
        dist.broadcast(t0, src=0)
        ok = torch.tensor([1.0 if np.array_equal(t0.cpu().numpy(), flat) else 0.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        same = bool(float(ok[0]) == 1.0)
    ms_max = float(t[0])
    if rank != 0:
        return None
    peak, peak_src = measured_peak()
    ach = (pr["sweep_bytes"] / 1e9) / (pr["sweep_ms"] * 1e-3)
    u = float(np.mean(out["mcuvar"][warmup + 1:warmup + 1 + steps]))
    L = HTLR_DEFAULTS["leap_L"]
    # two X sweeps per leapfrog step are inherent when rows are sharded (the gradient of a column needs every rank's
    # rows before the column can move): 2 (L+1) sweeps of this rank's 8 n_local u bytes (+ the detach sweep)
    bound_bytes = 2.0 * (L + 1) * 8.0 * nl * u
    tt = ncu_traffic("tall_sweeps_200000_rows") if nl == 200000 else None
    tall_tr = 0.5 * (tt["k_lv_sweep<4>"] + tt["k_grad_sweep<4>"]) if tt else None
    tall_src = ("profiles/traffic.json (dram__bytes_read.sum per launch, mean of the two sweep kernels at 200000 rows per "
                "GPU, committed ncu capture; not re-measured in this run)") if tt else "not captured at this shard size"
    it_s = steps / (ms_max * 1e-3)
    line = {"metric": "mcmc_iterations_per_sec", "value": it_s, "unit": "it/s", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": ms_max / steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config("E"), n_total=n_total),
            "run": {"rows_per_rank": nl,
                    "l2": "X shard (%.1f GB) is larger than the 126 MB L2; no flush needed" % (8e-9 * nl * (p + 1)),
                    "parallelism": f"rows sharded over {world} rank(s), NCCL all-reduce (sum, f64) of the "
                                   f"{p + 1}x{K} gradient partials + log-likelihood per gradient evaluation"
                    if world > 1 else "1 GPU holds all rows",
                    "algorithm": "two-sweep (lv sweep + gradient sweep per leapfrog step)", "mean_nuvar": u,
                    "hmc_reject": float(np.mean(out["mchmcrej"][warmup + 1:warmup + 1 + steps])),
                    "all_ranks_bit_identical": same,
                    # whether rank 0's GPU can map its peers' memory: NCCL's all-reduce falls back to host shared memory
                    # when it cannot (43 us per call became 224 us on one 8-GPU lease of round 2)
                    "peer_access_from_gpu0": _peer_access(torch) if world > 1 else [],
                    "loglike_first_last": [float(np.ravel(out["mcloglike"])[1]), float(np.ravel(out["mcloglike"])[-1])]},
            "gpu_launches": launches, "clocks": clk,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": tall_tr, "traffic_source": tall_src,
                         "kernel": "k_lv_sweep + k_grad_sweep over this rank's row block",
                         "peak_source": peak_src, "launches_timed": pr["sweep_launches"],
                         "algorithmic_bytes_per_launch": pr["sweep_bytes"] / max(pr["sweep_launches"], 1),
                         "sweep_share_of_step": pr["sweep_ms"] / max(pr["total_ms"], 1e-9),
                         # NCCL all-reduce of the (p+1) x K gradient (+ log-likelihood), once per gradient evaluation:
                         # CUDA events around the call on the library's stream, same eager pass as the sweep times
                         "allreduce_share_of_step": pr["allreduce_ms"] / max(pr["total_ms"], 1e-9),
                         "allreduce_us_per_call": 1e3 * pr["allreduce_ms"] / max(psteps * (L + 1), 1) if world > 1 else None,
                         "two_sweep_bound": {"bytes_per_iteration_per_rank": bound_bytes,
                                             "it_s_at_measured_peak": peak * 1e9 / bound_bytes,
                                             "frac_iteration": it_s * bound_bytes / 1e9 / peak}}}
    if cpu_baseline:
        rate, kind, m = cpu_reference_rate("E", 3, budget_s=15.0, warm_iters=0)
        line["cpu_baseline"] = {"value": rate * E_CPU_ROWS / n_total, "unit": "it/s", "cores": 1, "kind": kind,
                                "sample": f"{m} sampling iterations on a {E_CPU_ROWS}-row sample of the same law, scaled "
                                          f"linearly to n={n_total} (the reference's cost is linear in n); 1 of "
                                          f"{os.cpu_count()} host cores"}
    return line


def sharded_parity(torch, dist, rank, world, local):
    """A small chain row-sharded over all ranks (device Philox, identical keys on every rank) against the same chain
    on rank 0 alone with the two-sweep algorithm: the accept history and the restricted-set sizes must be identical
    and the coefficient trace equal to 1e-10 (norm-relative); every rank must hold the same trace bit for bit."""
    import htlr_b200
    from htlr_b200 import synth
    n, p, C, iters_h, iters_rmc = 4000, 300, 3, 4, 8
    d = synth.gendata_mlr(n, p, NC=C, seed=4242)
    y = d["y"].copy()
    y[:C] = np.arange(C)
    X, ymat, ybase, K = synth.make_inputs(d["X"], y)
    g = np.random.default_rng(4243)
    deltas = np.asfortranarray(g.standard_normal((p + 1, K)) * 0.3)
    v = synth.comp_vardeltas(deltas)[1:]
    sig = np.r_[2000.0, (np.exp(-10.0) + v) / 2 / g.gamma((1 + K) / 2, size=p)]
    args = dict(HTLR_DEFAULTS, p=p, K=K, n=n, X=X, ymat=ymat, ybase=ybase, deltas=deltas, sigmasbt=sig,
                hmc_sgmcut=0.05, iters_h=iters_h, iters_rmc=iters_rmc, leap_L=12, leap_L_h=4, keep_warmup_hist=True)
    T = iters_h + iters_rmc
    s = htlr_b200.ShardedSampler(**args, dist=dist, device=local, seed=77)
    s.run(T)
    mine = s.fetch()
    s.close()
    # every rank's trace against rank 0's, bit for bit (replicated state evolves identically: no broadcast needed)
    flat = np.concatenate([np.ravel(mine[k]) for k in ("mcdeltas", "mcsigmasbt", "mcloglike", "mcuvar", "mchmcrej")])
    t0 = torch.from_numpy(flat.copy()).cuda()
    dist.broadcast(t0, src=0)
    same = torch.tensor([1.0 if np.array_equal(t0.cpu().numpy(), flat) else 0.0], device="cuda", dtype=torch.float64)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    if rank != 0:
        return None
    one = htlr_b200.Sampler(**args, device=local, seed=77, algo=1)
    one.run(T)
    ref = one.fetch()
    one.close()
    den = float(np.max(np.abs(ref["mcdeltas"])))
    rel = float(np.max(np.abs(mine["mcdeltas"] - ref["mcdeltas"])) / den)
    rel_ll = float(np.max(np.abs(mine["mcloglike"] - ref["mcloglike"])) / np.max(np.abs(ref["mcloglike"])))
    out = {"what": f"n={n} p={p} C={C}, {T} iterations (L=12), device Philox, rows sharded over {world} ranks with an NCCL "
                   f"all-reduce of the gradient partials per leapfrog step vs the same chain on rank 0 alone (two-sweep)",
           "mcuvar_identical": bool(np.array_equal(mine["mcuvar"], ref["mcuvar"])),
           "mchmcrej_identical": bool(np.array_equal(mine["mchmcrej"], ref["mchmcrej"])),
           "relerr_mcdeltas": rel, "relerr_mcloglike": rel_ll, "tolerance": 1e-10,
           "all_ranks_bit_identical": bool(float(same[0]) == 1.0),
           "accepted_proposals": int(T - float(np.sum(ref["mchmcrej"][1:]))),
           "mean_nuvar": float(np.mean(ref["mcuvar"][1:]))}
    out["ok"] = bool(out["mcuvar_identical"] and out["mchmcrej_identical"] and rel < 1e-10 and rel_ll < 1e-10
                     and out["all_ranks_bit_identical"])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C_full", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the default-cut (restricted) block")
    ap.add_argument("--no-extra-blocks", action="store_true",
                    help="skip the sharded_parity / E_row_sharded / D_chains blocks of the default workload's line")
    ap.add_argument("--algo", type=int, default=0, help="0 auto (fused when applicable), 1 two-sweep, 2 fused")
    ap.add_argument("--n-total", type=int, default=0, help="workload E only: total rows (default 200000)")
    ap.add_argument("--chains-per-gpu", type=int, default=4, help="workload D only")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)
    # ONE JSON line on stdout: native libraries (NCCL banner, ...) print to fd 1, so park it on stderr and
    # keep the real stdout for the result line
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    global ALGO
    ALGO = a.algo

    if a.impl == "reference":
        run_reference_arm(a)
        return

    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: htlr_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_mod

    if a.workload in ("E", "D"):
        if a.workload == "E":
            line = measure_tall(torch, a.steps, a.warmup, a.n_total, dist, rank, world, local, not a.no_cpu_baseline)
        else:
            line = measure_chains(torch, a.steps, a.warmup, a.chains_per_gpu, dist, rank, world, local,
                                  not a.no_cpu_baseline)
        if rank == 0:
            line["e2e"] = None
            line["e2e_note"] = "the host-buffer end-to-end number is measured on the default workload"
            print(json.dumps(line))
        if dist is not None:
            dist.destroy_process_group()
        return

    if world > 1:
        # the CPU baseline and the extra single-GPU blocks belong to the N = 1 line only
        a.no_cpu_baseline = True
        a.no_secondary = True

    w = WORKLOADS[a.workload]
    clocks = ClockSampler(local)
    clocks.start()
    res = measure_workload(torch, a.workload, a.steps, a.warmup, local, rank, dist)
    clk = clocks.stop()
    e2e_s, h2d, d2h, e2e_warm = measure_e2e(torch, a.workload, a.steps, local, rank, dist)

    ms = torch.tensor([res["ms"], e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max = float(ms[0]), float(ms[1])
    value = world * a.steps / (ms_max * 1e-3)
    e2e_value = world * a.steps / (e2e_ms_max * 1e-3)

    secondary = None
    if not a.no_secondary and a.workload == "C_full" and rank == 0:
        n2 = max(a.steps, 200)
        r2 = measure_workload(torch, "C_default", n2, a.warmup, local, rank, None, want_profile=True)
        secondary = {"workload": WORKLOADS["C_default"]["desc"], "value": n2 / (r2["ms"] * 1e-3),
                     "unit": "it/s", "ms_per_step": r2["ms"] / n2, "mean_nuvar": r2["uvar"],
                     "hmc_reject": r2["rej"], "gpu_launches": r2["launches"],
                     "fused_phase_ms_per_launch_block0": {k: r2["prof"][k] / max(r2["prof_steps"], 1) for k in
                                                          ("fused_sweep_ms", "fused_barrier_ms", "fused_rowphase_ms")},
                     "note": "restricted regime: X[:,iup] is L2-resident, bound by launch/sync latency not HBM"}
    others = None
    if not a.no_secondary and a.workload == "C_full" and rank == 0:
        # config B (BASELINE configs[1]) in both regimes, and one whole default run of config C
        # (htlr() defaults: iter = 2000, warmup = 1000) through the public entry point with host buffers
        others = {}
        for name, wl_name in (("A_colon_default_cut", "A"), ("B_default_cut", "B_default"), ("B_full", "B")):
            rb = measure_workload(torch, wl_name, 200, a.warmup, local, rank, None, want_profile=False)
            others[name] = {"workload": WORKLOADS[wl_name]["desc"], "value": 200 / (rb["ms"] * 1e-3), "unit": "it/s",
                            "ms_per_step": rb["ms"] / 200, "mean_nuvar": rb["uvar"], "hmc_reject": rb["rej"],
                            "gpu_launches": rb["launches"]}
        import htlr_b200
        args = make_problem("C_default")
        t0 = time.perf_counter()
        out = htlr_b200.htlr_fit_helper(**dict(args, iters_h=1000, iters_rmc=1000), device=local, seed=4242)
        dt = time.perf_counter() - t0
        others["C_whole_default_run"] = {"what": "htlr() defaults on config C through htlr_fit_helper with host buffers: "
                                                 "1000 warm-up iterations (leap.warm = 5) + 1000 sampling iterations "
                                                 "(leap = 50), H2D of X, D2H of the 1001-slice trace",
                                         "seconds": dt, "value": 2000 / dt, "unit": "it/s",
                                         "mean_nuvar_sampling": float(np.mean(out["mcuvar"][1:])),
                                         "hmc_reject_sampling": float(np.mean(out["mchmcrej"][1:]))}
        # N2: htlr_predict on the device from the resident trace of a whole default run - the one dense
        # contraction of the path, n_ts x (p+1) x K*S in FP64 on the DMMA path - against the measured FP64 peaks
        smp = htlr_b200.Sampler(**dict(args, iters_h=1000, iters_rmc=1000), device=local, seed=4243)
        smp.run_async(2000)
        smp.sync()
        Xts = np.asfortranarray(args["X"])
        t0 = time.perf_counter()
        probs = smp.predict(Xts)
        t_call = time.perf_counter() - t0
        kms = []
        for _ in range(3):
            smp.predict(Xts)
            kms.append(smp.predict_stats()[0])
        flops = smp.predict_stats()[1]
        smp.close()
        pk = htlr_b200.fp64_peak(local)
        km = sorted(kms)[1]
        others["predict_on_device"] = {
            "what": "htlr_predict (R/predict.R:28-92) of the %d training cases from the %d used samples of that run's "
                    "resident trace: cbind(1, X_ts) %%*%% mcdeltas[, , used] on the FP64 tensor-core path (DMMA m8n8k4), "
                    "log-sum-exp, mean over samples" % (Xts.shape[0], int(round(flops / (2.0 * Xts.shape[0] * Xts.shape[1] * args["K"])))),
            "kernel_ms": km, "fp64_tflops": flops / (km * 1e-3) / 1e12, "flops": flops,
            "fp64_peak_measured_tflops": pk, "frac_of_dmma_peak": flops / (km * 1e-3) / 1e12 / pk["dmma"],
            "call_seconds_with_host_X_ts": t_call, "h2d_bytes": 8 * Xts.size,
            "probs_row_sums_ok": bool(np.allclose(probs.sum(axis=1), 1.0, atol=1e-12))}
    if dist is not None:
        dist.barrier()

    # ---- the blocks for the paths that the replica measurement does not touch (every rank takes part)
    blocks = {}
    if a.workload == "C_full" and not a.no_extra_blocks:
        if world > 1:
            blocks["sharded_parity"] = sharded_parity(torch, dist, rank, world, local)
        e_steps = max(5, min(20, 3 * world))   # an iteration shrinks with the shard: keep the timed region near a second
        blocks["E_row_sharded"] = measure_tall(torch, e_steps, 3, a.n_total, dist, rank, world, local,
                                               cpu_baseline=(world == 1 and not a.no_cpu_baseline))
        blocks["D_chains"] = measure_chains(torch, max(a.steps, 300), a.warmup, a.chains_per_gpu, dist, rank, world, local,
                                            cpu_baseline=(world == 1 and not a.no_cpu_baseline))
        if dist is not None:
            dist.barrier()

    if rank == 0:
        peak, peak_src = measured_peak()
        pr = res["prof"]
        ach = (pr["sweep_bytes"] / 1e9) / (pr["sweep_ms"] * 1e-3) if pr["sweep_ms"] > 0 else None
        n, p, K = res["n"], res["p"], res["K"]
        u = res["uvar"]
        nvar = p + 1
        m_detach = u if u <= nvar // 2 else nvar - u
        L = HTLR_DEFAULTS["leap_L"]
        bytes_two = 8.0 * n * (m_detach + u + 2 * L * u)     # the reference's algorithm
        # what the fused kernels actually read: (L+1) sweeps of X[:, iup], no detach sweep (lv is carried from step to
        # step); SURVEY 8d's B_iter,min = 8 n [m + (L+1) u] still counts the detach sweep
        bytes_min = 8.0 * n * ((0 if a.algo != 1 else m_detach) + (L + 1) * u)
        it_s_1gpu = a.steps / (res["ms"] * 1e-3)
        tr = ncu_traffic("sweep")
        line = {
            "metric": "mcmc_iterations_per_sec", "value": value, "unit": "it/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_max / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(a.workload),
            "run": {"l2": "X (%.0f MB) is larger than the 126 MB L2; no flush needed" % (8e-6 * n * nvar)
                    if 8 * n * nvar > 126e6 else "X fits in L2 (stated; no flush: the sampler re-reads X by design)",
                    "parallelism": "1 chain per GPU, no communication" if world > 1 else "1 chain, 1 GPU",
                    "algorithm": "two-sweep (reference structure)" if a.algo == 1 else "fused one-sweep persistent trajectory kernel",
                    "mean_nuvar": u, "hmc_reject": res["rej"], "chain_warmup_iterations": res["chain_warm"]},
            "e2e": {"value": e2e_value, "unit": "it/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "what": "htlr_fit_helper() with pinned host buffers: H2D of X/ymat/state, %d sampling iterations "
                            "(after %d warm-up), D2H of the full trace (pipelined behind the sampling, units of 8 "
                            "iterations); median of 3 calls" % (a.steps, e2e_warm),
                    "calls_s": getattr(measure_e2e, "calls_s", None)},
            "gpu_launches": res["launches"],
            "clocks": clk,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                         "frac": (ach / peak) if ach else None,
                         # the same algorithmic bytes over the driver-visible time of a whole iteration (every kernel of
                         # it, graph replay): the kernel-level `frac` comes from a separate eager pass with events
                         # around the sweep launches
                         "frac_iteration": bytes_min * it_s_1gpu / 1e9 / peak,
                         "traffic": tr, "traffic_source": "profiles/traffic.json (ncu --set full capture of the same "
                                                          "kernel, committed; not re-measured in this run)" if tr else None,
                         "kernel": "k_lv_sweep + k_grad_sweep (X sweeps)" if a.algo == 1 else "k_traj (L+1 X sweeps per launch; no detach sweep: lv is carried across steps)", "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": pr["sweep_bytes"] / max(pr["sweep_launches"], 1),
                         "launches_timed": pr["sweep_launches"], "frac_of_nominal_8TBs": (ach / 8000.0) if ach else None,
                         "fused_phase_ms_per_launch_block0": {k: pr[k] / max(res["prof_steps"], 1) for k in
                                                              ("fused_sweep_ms", "fused_barrier_ms", "fused_rowphase_ms")},
                         "iteration_level": {"bytes_two_sweep": bytes_two, "bytes_min_one_sweep": bytes_min,
                                             "achieved_one_sweep_GBs": bytes_min * it_s_1gpu / 1e9,
                                             "achieved_two_sweep_GBs": bytes_two * it_s_1gpu / 1e9,
                                             "frac_two_sweep": bytes_two * it_s_1gpu / 1e9 / peak}},
        }
        if secondary:
            line["default_cut"] = secondary
        if others:
            line["other_configs"] = others
        for k, v in blocks.items():
            line[k] = v
        if not a.no_cpu_baseline:
            rate, kind, m = cpu_reference_rate(a.workload, 4, budget_s=25.0, warm_iters=0)
            line["cpu_baseline"] = {"value": rate, "unit": "it/s", "cores": 1, "kind": kind,
                                    "sample": f"{m} sampling iterations (L=50) of the same workload on 1 of "
                                              f"{os.cpu_count()} host cores (reference hot loops are single-threaded)"}
            if secondary:
                r2c, kind2, m2 = cpu_reference_rate("C_default", 400, budget_s=15.0, warm_iters=150)
                secondary["cpu_baseline"] = {"value": r2c, "unit": "it/s", "cores": 1, "kind": kind2,
                                             "sample": f"{m2} sampling iterations after 150 warm-up iterations"}
                secondary["ratio_to_cpu"] = secondary["value"] / r2c
            if others:
                for name, wl_name, warm in (("A_colon_default_cut", "A", 150), ("B_default_cut", "B_default", 150),
                                            ("B_full", "B", 0)):
                    rc_, kind_, m_ = cpu_reference_rate(wl_name, 200, budget_s=8.0, warm_iters=warm)
                    others[name]["cpu_baseline"] = {"value": rc_, "unit": "it/s", "cores": 1, "kind": kind_,
                                                    "sample": f"{m_} sampling iterations after {warm} warm-up iterations"}
                    others[name]["ratio_to_cpu"] = others[name]["value"] / rc_
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
