"""Cluster-mode trajectory: time per iteration vs cluster size for small restricted sets (HTLR_CLUSTER_SIZE)."""
import sys, os, json, subprocess
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch, htlr_b200
    from tests.helpers import problem
    for n, C, u in ((500, 4, 12), (500, 4, 64), (62, 2, 64), (200, 2, 256), (1000, 3, 16), (1000, 3, 128)):
        args = problem(n, u - 1, C, seed=1, iters_rmc=400, iters_h=0, leap_L=50, hmc_sgmcut=0.0, init="zero")
        try:
            s = htlr_b200.Sampler(**args, seed=3, algo=4)
        except Exception as e:
            print(json.dumps({"cl": os.environ.get("HTLR_CLUSTER_SIZE"), "n": n, "K": C - 1, "u": u, "err": str(e)[:60]}), flush=True)
            continue
        s.run(20)
        st = torch.cuda.ExternalStream(s.stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.sync(); e0.record(st); s.run_async(200); e1.record(st); s.sync()
        print(json.dumps({"cl": os.environ.get("HTLR_CLUSTER_SIZE"), "n": n, "K": C - 1, "u": u, "us_per_iter": e0.elapsed_time(e1) * 5.0}), flush=True)
        s.close()
else:
    for cl in (1, 2, 4, 8, 16):
        subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, HTLR_CLUSTER_SIZE=str(cl)))
