"""Host-side logic of the N > 1 paths on CPU: world_size-2 `gloo` process groups (no GPU, no compute calls).

Covers the row partition, the replica deal, the unique-id exchange and the throughput aggregation that
bench.py and ShardedSampler rely on. The data path itself (NCCL all-reduce inside the sampler) is covered
by tests/test_gpu_multi.py on a 2-GPU box."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from htlr_b200 import shard
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. the 128-byte id created on rank 0 reaches every rank unchanged
        uid = shard.exchange_unique_id(dist, lambda: bytes(range(128)))
        # 2. row shards tile the rows: gather every rank's block and compare with the full matrix
        g = np.random.default_rng(0)
        n, p, K = 11, 4, 2
        X = np.asfortranarray(g.standard_normal((n, p + 1)))
        ymat = np.asfortranarray((g.random((n, K)) > 0.5).astype(np.float64))
        ybase = g.integers(0, K + 1, n)
        Xr, ymr, ybr = shard.shard_rows(X, ymat, ybase, rank, world)
        assert Xr.flags.f_contiguous and ymr.flags.f_contiguous
        parts = [None] * world
        dist.all_gather_object(parts, (Xr, ymr, ybr))
        ok_rows = (np.array_equal(np.vstack([q_[0] for q_ in parts]), X)
                   and np.array_equal(np.vstack([q_[1] for q_ in parts]), ymat)
                   and np.array_equal(np.concatenate([q_[2] for q_ in parts]), ybase))
        # 3. a column-wise gradient partial summed over ranks equals the full gradient (what the NCCL
        #    all-reduce does on the device): X_r' R_r summed == X' R
        R = g.standard_normal((n, K))
        lo, hi = shard.row_range(n, rank, world)
        import torch
        part = torch.from_numpy(Xr.T @ R[lo:hi])
        dist.all_reduce(part)
        ok_grad = np.allclose(part.numpy(), X.T @ R, rtol=1e-13, atol=1e-13)
        # 4. replicas: whole-job rate = total iterations / slowest rank
        rate, tmax = shard.gather_rates(dist, iterations=100, seconds=1.0 + rank)
        chains = shard.chain_assignment(5, rank, world)
        q.put((rank, uid == bytes(range(128)), ok_rows, ok_grad, rate, tmax, chains))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_host_logic():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    for rank, uid_ok, ok_rows, ok_grad, rate, tmax, chains in res:
        assert uid_ok and ok_rows and ok_grad
        assert rate == pytest.approx(200 / 2.0) and tmax == 2.0
    assert res[0][6] == [0, 2, 4] and res[1][6] == [1, 3]


def test_row_range_properties():
    from htlr_b200.shard import chain_assignment, row_range
    for n in (1, 7, 8, 200000, 200003):
        for world in (1, 2, 3, 4, 8):
            if n < world:
                continue
            blocks = [row_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
    assert sorted(sum((chain_assignment(32, r, 8) for r in range(8)), [])) == list(range(32))
    with pytest.raises(ValueError):
        row_range(10, 3, 3)
