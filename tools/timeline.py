"""Where an iteration's time goes, kernel by kernel, in graph-replay mode (HTLR_TIMELINE=1: every kernel's first
thread stamps %globaltimer at its start; see htlr_profile.timeline_us). Also the breakdown of one end-to-end
htlr_fit_helper-style call (create + H2D, run, fetch, destroy) on the default workload."""
import sys, os, json, time
os.environ["HTLR_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, htlr_b200, bench

NAMES = ["tail(prev)+gap", "select", "begin", "detach", "softmax", "small", "cluster", "grid(traj)", "energy"]
# timeline_us[k] = time from the start of the previous kernel to the start of kernel k, i.e. the cost of the
# PREVIOUS kernel: index 0 (start of select) closes the previous iteration's tail kernel


def timeline(wl, iters=200, algo=0):
    args = bench.make_problem(wl)
    w = bench.WORKLOADS[wl]
    warm = 0 if w["cut"] <= 0 else 150
    s = htlr_b200.Sampler(**dict(args, iters_h=warm, iters_rmc=3 * iters + 50), seed=11, algo=algo)
    s.run(warm + 20)
    st = torch.cuda.ExternalStream(s.stream)

    def timed(k):
        s.profile_get(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.sync(); e0.record(st); s.run_async(k); e1.record(st); s.sync()
        return e0.elapsed_time(e1), s.profile_get()
    # the first interval of a run also holds the host-side gap since the previous run: difference of a run of
    # 2 * iters and a run of iters cancels it
    ms1, p1 = timed(iters)
    ms2, p2 = timed(2 * iters)
    tl = [(b - a) / iters for a, b in zip(p1["timeline_us"], p2["timeline_us"])]
    ends = tl[9:12]
    tl = tl[:9]
    row = {"workload": wl, "algo": algo, "us_per_iter": round(1e3 * ms2 / (2 * iters), 1),
           "small_share": p2["small_kernel_trajectories"] / (2 * iters),
           "stages_us": {NAMES[k]: round(tl[k], 2) for k in range(9)},
           # with the end marks the intervals split: kernel body (start -> end of its epilogue) and gap to the next start
           "tail_body": round(ends[0], 2), "energy_body": round(ends[1], 2), "begin_body": round(ends[2], 2),
           "sum": round(sum(tl) + sum(ends), 1)}
    s.close()
    print(json.dumps(row), flush=True)


def e2e_breakdown(wl="C_full", steps=100):
    args = bench.make_problem(wl)
    Xp = torch.from_numpy(np.ascontiguousarray(args["X"].T)).pin_memory()
    args["X"] = Xp.numpy().T
    a = dict(args, iters_h=0, iters_rmc=steps)
    for rep in range(6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = htlr_b200.htlr_fit_helper(**a, seed=5 + rep)
        t1 = time.perf_counter()
        print(json.dumps({"e2e": wl, "rep": rep, "pipelined_helper_ms": round(1e3 * (t1 - t0), 2),
                          "it_per_s": round(steps / (t1 - t0), 1)}), flush=True)
        del out
    for rep in range(6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        s = htlr_b200.Sampler(**a, seed=5 + rep)
        s.sync(); t1 = time.perf_counter()
        s.run(steps); t2 = time.perf_counter()
        out = s.fetch(); t3 = time.perf_counter()
        s.close(); t4 = time.perf_counter()
        print(json.dumps({"e2e": wl, "rep": rep, "create_ms": round(1e3 * (t1 - t0), 2), "run_ms": round(1e3 * (t2 - t1), 2),
                          "fetch_ms": round(1e3 * (t3 - t2), 2), "destroy_ms": round(1e3 * (t4 - t3), 2),
                          "total_ms": round(1e3 * (t4 - t0), 2)}), flush=True)
        del out


if __name__ == "__main__":
    if "--e2e-only" not in sys.argv:
        for wl in ("A", "B_default", "C_default", "B", "C_full"):
            timeline(wl)
    e2e_breakdown()
