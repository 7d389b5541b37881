"""Row-sharded X over 2 GPUs (NCCL all-reduce of the gradient partials inside every leapfrog step) against
the single-GPU chain and the CPU oracle. Needs a 2-GPU box: skipped elsewhere."""
import os
import socket

import numpy as np
import pytest

from tests.helpers import problem, relerr, streams

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, args, st, use_graph, q):
    import torch
    import torch.distributed as dist
    import htlr_b200
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        T = args["iters_h"] + args["iters_rmc"]
        kw = dict(rng_mode=htlr_b200.fit.RNG_INJECT) if st is not None else dict(seed=77)
        s = htlr_b200.ShardedSampler(**args, dist=dist, device=rank, use_graph=use_graph, **kw)
        stats = s.run(T, **(st or {}))
        out = s.fetch()
        s.close()
        q.put((rank, {k: np.array(v) for k, v in out.items() if k.startswith("mc") and k != "mc.param"},
               stats["uvar"], out["n"]))
    finally:
        dist.destroy_process_group()


def _run_sharded(args, st, use_graph=True, world=2):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, args, st, use_graph, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = sorted((q.get(timeout=300) for _ in procs), key=lambda x: x[0])
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    return res


def _need_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")


@pytest.mark.parametrize("use_graph", [True, False])
def test_row_sharded_chain_matches_oracle(use_graph):
    _need_two_gpus()
    from oracle import oracle as O
    args = problem(301, 120, 3, seed=5, iters_rmc=5, iters_h=4, keep_warmup_hist=True, leap_L=9)
    st = streams(args, seed=8)
    res = _run_sharded(args, st, use_graph)
    ora = O.OracleFit(**args)
    ora.initialize()
    ora.set_inject(st["normals"], st["uniforms"], st["gammas"], None)
    ora.run(9)
    b = ora.fetch()
    for rank, out, uvar, n_rep in res:
        assert n_rep == 301
        for k in ("mcdeltas", "mcsigmasbt", "mcvardeltas", "mclogw", "mcloglike", "mcuvar", "mchmcrej"):
            assert relerr(np.ravel(out[k]), np.ravel(b[k])) < 1e-10, (rank, k)
    # replicated state evolves bit-identically on both ranks (no broadcast needed)
    for k in res[0][1]:
        assert np.array_equal(res[0][1][k], res[1][1][k]), k


def test_row_sharded_device_rng_matches_single_gpu():
    _need_two_gpus()
    import htlr_b200
    args = problem(500, 200, 4, seed=9, iters_rmc=6, iters_h=6, leap_L=12)
    res = _run_sharded(args, None)
    one = htlr_b200.Sampler(**args, seed=77, algo=1)
    one.run(12)
    ref = one.fetch()
    one.close()
    for rank, out, uvar, n_rep in res:
        assert np.array_equal(out["mcuvar"], ref["mcuvar"])
        assert np.array_equal(out["mchmcrej"], ref["mchmcrej"])
        assert relerr(out["mcdeltas"], ref["mcdeltas"]) < 1e-10
        assert relerr(out["mcloglike"], ref["mcloglike"]) < 1e-10


def test_one_process_drives_chains_on_two_devices():
    """Chains / CV folds may also be spread over the GPUs from ONE process (a host thread per device): kernel
    attributes (opt-in shared memory, cluster size) are per device, the device-memory cache is keyed by device, and
    the same key gives the same chain on either GPU."""
    import threading
    import htlr_b200
    _need_two_gpus()
    cases = [problem(900, 700, 3, seed=3, iters_h=4, iters_rmc=12, leap_L=10, leap_L_h=3, hmc_sgmcut=0.0),   # grid mode
             problem(62, 300, 2, seed=4, iters_h=4, iters_rmc=12, leap_L=10, leap_L_h=3),                    # small / cluster
             problem(300, 40, 17, seed=5, iters_h=2, iters_rmc=6, leap_L=4, leap_L_h=2)]                     # K = 16: generic kernels
    out = {}

    def chain(dev, ci):
        s = htlr_b200.Sampler(**cases[ci], seed=100 + ci, device=dev)
        s.run(cases[ci]["iters_h"] + cases[ci]["iters_rmc"])
        out[(dev, ci)] = s.fetch()
        s.close()
    for ci in range(len(cases)):
        chain(0, ci)                       # device 0 first: configures every kernel there
    threads = [threading.Thread(target=lambda: [chain(1, ci) for ci in range(len(cases))]),
               threading.Thread(target=lambda: [chain(0, ci) for ci in reversed(range(len(cases)))])]
    first = {k: v for k, v in out.items()}
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for ci in range(len(cases)):
        for k in ("mcdeltas", "mcsigmasbt", "mcloglike", "mcuvar", "mchmcrej"):
            assert np.array_equal(out[(1, ci)][k], first[(0, ci)][k]), (ci, k)
            assert np.array_equal(out[(0, ci)][k], first[(0, ci)][k]), (ci, k)
