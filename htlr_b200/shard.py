"""Multi-GPU host plumbing: one process per GPU (torchrun), `torch.distributed` only for rendezvous.

Two ways the path partitions:

* replicas   — independent chains / CV folds, one sampler per rank, nothing exchanged on the data path.
               `chain_assignment` deals chain ids to ranks; `gather_rates` sums per-rank throughput.
* row shards — tall X: rank r owns rows `row_range(n, r, R)` of X / ymat / ybase. The library all-reduces the
               u x K gradient partials (+ local log-likelihood) over NCCL inside every leapfrog step; every
               O(nvar) quantity is replicated and evolves identically on all ranks (same Philox keys), so every
               rank ends with the same trace.

Nothing here computes: it only slices host arrays, exchanges the 128-byte NCCL id and builds samplers.
"""
import ctypes

import numpy as np

from . import _lib
from .fit import HtlrCudaError, Sampler


def row_range(n, rank, world):
    """Contiguous, balanced row block of rank `rank`: sizes differ by at most one, earlier ranks larger."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def chain_assignment(n_chains, rank, world):
    """Round-robin deal of chain / fold ids to ranks (config D: 32 chains over 8 GPUs -> 4 each)."""
    return list(range(rank, n_chains, world))


def shard_rows(X, ymat, ybase, rank, world):
    """This rank's rows of the three n-sized inputs (column-major copies, as the C ABI wants them)."""
    lo, hi = row_range(X.shape[0], rank, world)
    if hi - lo < 1:
        raise HtlrCudaError(1, f"rank {rank} of {world} would own no rows (n = {X.shape[0]})")
    return (np.asfortranarray(X[lo:hi]), np.asfortranarray(np.reshape(ymat, (X.shape[0], -1), order="F")[lo:hi]),
            np.ascontiguousarray(ybase[lo:hi]))


def exchange_unique_id(dist, make_id, src=0):
    """Rank `src` creates the 128-byte NCCL unique id with `make_id()`; everyone receives it through the
    process group `dist` (any backend: gloo on CPU, nccl on GPU). Returns bytes of length 128."""
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.zeros(128, dtype=torch.uint8, device=dev)
    if dist.get_rank() == src:
        raw = make_id()
        if len(raw) != 128:
            raise ValueError("NCCL unique id must be 128 bytes")
        t = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
    dist.broadcast(t, src=src)
    return bytes(t.cpu().numpy().tobytes())


def nccl_unique_id():
    buf = (ctypes.c_ubyte * 128)()
    rc = _lib.lib().htlr_cuda_nccl_unique_id(buf)
    if rc:
        raise HtlrCudaError(rc, _lib.lib().htlr_cuda_last_error(None).decode())
    return bytes(buf)


def nccl_init(uid, rank, world, device):
    comm = ctypes.c_void_p()
    buf = (ctypes.c_ubyte * 128).from_buffer_copy(uid)
    rc = _lib.lib().htlr_cuda_nccl_init(buf, rank, world, device, ctypes.byref(comm))
    if rc:
        raise HtlrCudaError(rc, _lib.lib().htlr_cuda_last_error(None).decode())
    return comm


def nccl_destroy(comm):
    _lib.lib().htlr_cuda_nccl_destroy(comm)


def gather_rates(dist, iterations, seconds):
    """Whole-job throughput of replicas: total iterations over all ranks / slowest rank's time."""
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    it = torch.tensor([float(iterations)], dtype=torch.float64, device=dev)
    tm = torch.tensor([float(seconds)], dtype=torch.float64, device=dev)
    dist.all_reduce(it, op=dist.ReduceOp.SUM)
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    return float(it) / float(tm), float(tm)


class ShardedSampler(Sampler):
    """One rank's handle of a row-sharded chain. Same constructor as `Sampler` with the FULL inputs; it keeps
    only this rank's rows. `dist` is an initialised torch.distributed module (nccl backend on GPUs)."""

    def __init__(self, p, K, n, X, ymat, ybase, *rest, dist, device=0, **kw):
        rank, world = dist.get_rank(), dist.get_world_size()
        Xr, ymr, ybr = shard_rows(X, ymat, ybase, rank, world)
        uid = exchange_unique_id(dist, nccl_unique_id)
        self.comm = nccl_init(uid, rank, world, device)
        self.n_total, self.rank, self.world = n, rank, world
        try:
            super().__init__(p, K, Xr.shape[0], Xr, ymr, ybr, *rest, device=device, nccl_comm=self.comm, **kw)
        except Exception:
            nccl_destroy(self.comm)
            self.comm = None
            raise

    def fetch(self):
        out = super().fetch()
        out["n"] = self.n_total   # OutputR reports the full n; self.n stays this rank's row count
        return out

    def close(self):
        super().close()
        if getattr(self, "comm", None):
            nccl_destroy(self.comm)
            self.comm = None
