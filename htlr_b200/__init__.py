"""htlr_b200 — B200 (sm_100a) CUDA replacement for HTLR's restricted-Gibbs / HMC `Fit` sampler.

Public surface: `htlr_fit_helper` (the reference's entry point, same 22 arguments and result layout),
`Sampler` (create / run / fetch phases of the same object) and the parity hooks. The compute lives in
the in-tree C-ABI library `libhtlr_cuda.so` (include/htlr_cuda.h); importing this package without it,
or calling it without a B200, fails loudly — there is no CPU path.
"""
from .fit import (HtlrCudaError, Sampler, ars_sample, fp64_peak, gendata_fam_device, gendata_mlr_device,  # noqa: F401
                  htlr_fit_helper, rng_dump, std)
from . import _lib, shard  # noqa: F401
from .shard import ShardedSampler  # noqa: F401
