"""Pipelined run + download (htlr_cuda_collect): the trace a host assembles unit by unit while the chain keeps running
is the trace of the blocking create / run / fetch sequence, bit for bit (Fit::OutputR, src/gibbs.cpp:128-152)."""
import numpy as np
import pytest

from tests.helpers import problem

pytestmark = pytest.mark.gpu
TRACE = ("mcdeltas", "mcsigmasbt", "mcvardeltas", "mclogw", "mcloglike", "mcuvar", "mchmcrej")


def blocking(args, seed, algo=0):
    import htlr_b200
    s = htlr_b200.Sampler(**args, seed=seed, algo=algo)
    st = s.run(args["iters_h"] + args["iters_rmc"])
    out = s.fetch()
    s.close()
    return out, st


@pytest.mark.parametrize("keep", [False, True], ids=["sampling-only", "keep-warmup"])
@pytest.mark.parametrize("iters", [(0, 5), (7, 30), (20, 131)], ids=["short", "medium", "odd-length"])
def test_pipelined_helper_equals_blocking_run(keep, iters):
    import htlr_b200
    args = problem(90, 300, 3, seed=3, iters_h=iters[0], iters_rmc=iters[1], leap_L=6, leap_L_h=3,
                   keep_warmup_hist=keep)
    ref, ref_st = blocking(args, seed=21)
    seen = []
    out = htlr_b200.htlr_fit_helper(**args, seed=21, progress=lambda done, st: seen.append((done, st)) or False)
    for k in TRACE:
        assert np.array_equal(out[k], ref[k]), k
    assert np.array_equal(out["mc.param"]["DDNloglike"], ref["mc.param"]["DDNloglike"])
    total = sum(iters)
    assert [d for d, _ in seen] == list(range(8, total, 8)) + [total]
    for k in ("loglike", "logw", "uvar", "rej"):
        assert np.array_equal(np.concatenate([st[k] for _, st in seen]), ref_st[k]), k


def test_collect_copies_only_finished_slices_and_leaves_later_work_running():
    import htlr_b200
    args = problem(120, 500, 3, seed=5, iters_h=4, iters_rmc=60, leap_L=8, leap_L_h=2)
    ref, _ = blocking(args, seed=8)
    s = htlr_b200.Sampler(**args, seed=8)
    out = s.trace_arrays()
    for _ in range(4):
        s.run_async(16)                 # 64 iterations enqueued, a completion mark after every 16
    s.collect(16, out)                  # iterations 0..15 = 4 warm-up + 12 sampling: slices 0..12
    assert np.array_equal(out["mcdeltas"][:, :, :13], ref["mcdeltas"][:, :, :13])
    assert not out["mcdeltas"][:, :, 13:].any() and not out["mcsigmasbt"][:, 13:].any()
    s.collect(20, out)                  # inside the second enqueue call: waits for its mark (32), copies up to 16
    assert np.array_equal(out["mcsigmasbt"][:, :17], ref["mcsigmasbt"][:, :17]) and not out["mcsigmasbt"][:, 17:].any()
    st = s.collect(64, out, stats_from=60)
    assert len(st["uvar"]) == 4 and np.array_equal(st["uvar"], ref["mcuvar"].ravel()[-4:])
    full = s.fetch(collected=out)
    for k in TRACE:
        assert np.array_equal(full[k], ref[k]), k
    with pytest.raises(htlr_b200.HtlrCudaError, match="collect past the enqueued"):
        s.collect(65, out)
    s.close()


def test_collect_without_marks_falls_back_to_a_full_sync():
    """More enqueue calls than the handle keeps marks for (256): early marks are recycled, collect still works."""
    import htlr_b200
    args = problem(40, 30, 2, seed=2, iters_h=0, iters_rmc=300, leap_L=2, leap_L_h=2)
    ref, _ = blocking(args, seed=4)
    s = htlr_b200.Sampler(**args, seed=4)
    out = s.trace_arrays()
    for _ in range(300):
        s.run_async(1)
    s.collect(10, out)
    assert np.array_equal(out["mcdeltas"][:, :, :11], ref["mcdeltas"][:, :, :11])
    s.collect(300, out)
    for k in TRACE:
        assert np.array_equal(out[k], ref[k]), k
    s.close()


@pytest.mark.parametrize("over", [dict(thin=3), dict(ptype="ghs", alpha=1.0), dict(ptype="neg", alpha=2.0, thin=2),
                                  dict(eta=0.5)], ids=["thin3", "ghs", "neg-thin2", "logw-ars"])
def test_pipelined_helper_with_thinning_and_the_other_priors(over):
    """The trace slice index does not depend on how the iterations were enqueued (units of 8 thin-groups, graphs of
    8 iterations only when thin = 1): thinning, ARS-based priors and the logw update give the blocking run's trace."""
    import htlr_b200
    args = problem(80, 120, 3, seed=12, iters_h=5, iters_rmc=21, leap_L=5, leap_L_h=2, keep_warmup_hist=True,
                   init="random", **over)
    ref, _ = blocking(args, seed=33)
    out = htlr_b200.htlr_fit_helper(**args, seed=33)
    for k in TRACE:
        assert np.array_equal(out[k], ref[k]), k


def test_unseeded_fits_differ_and_the_reported_seed_reproduces_them():
    """the reference draws from R's global stream, so two fits without set.seed() differ; a default seed
    of 0 silently made them identical. Now a fresh key per sampler, reported in the fit."""
    import htlr_b200
    args = problem(80, 60, 3, seed=2, iters_h=3, iters_rmc=5, leap_L=6, leap_L_h=2)
    a = htlr_b200.htlr_fit_helper(**args)
    b = htlr_b200.htlr_fit_helper(**args)
    assert a["cuda.seed"] != b["cuda.seed"] and not np.array_equal(a["mcdeltas"], b["mcdeltas"])
    c = htlr_b200.htlr_fit_helper(**args, seed=a["cuda.seed"])
    assert np.array_equal(a["mcdeltas"], c["mcdeltas"]) and np.array_equal(a["mcsigmasbt"], c["mcsigmasbt"])
