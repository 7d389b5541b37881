"""ctypes loader for the C-ABI CUDA library (htlr_b200/libhtlr_cuda.so, built in-tree for sm_100a).

Fails loudly when the library is missing: there is no CPU fallback anywhere in this package.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# HTLR_CUDA_LIB: another build of the same ABI (same-box A/B runs of two kernel versions: tools/ab/)
LIB_PATH = os.environ.get("HTLR_CUDA_LIB") or os.path.join(_HERE, "libhtlr_cuda.so")

c_dp = ctypes.POINTER(ctypes.c_double)
c_ip = ctypes.POINTER(ctypes.c_int32)
c_lp = ctypes.POINTER(ctypes.c_int64)

ABI_VERSION = 2


class HtlrCfg(ctypes.Structure):
    """Mirror of `htlr_cfg` in include/htlr_cuda.h."""
    _fields_ = [
        ("struct_size", ctypes.c_uint32), ("abi_version", ctypes.c_uint32),
        ("p", ctypes.c_int32), ("K", ctypes.c_int32), ("n", ctypes.c_int32), ("ptype", ctypes.c_int32),
        ("alpha", ctypes.c_double), ("s", ctypes.c_double), ("eta", ctypes.c_double),
        ("iters_rmc", ctypes.c_int32), ("iters_h", ctypes.c_int32), ("thin", ctypes.c_int32),
        ("leap_L", ctypes.c_int32), ("leap_L_h", ctypes.c_int32),
        ("leap_step", ctypes.c_double), ("hmc_sgmcut", ctypes.c_double),
        ("keep_warmup_hist", ctypes.c_int32), ("silence", ctypes.c_int32), ("legacy", ctypes.c_int32),
        ("device", ctypes.c_int32), ("rng_mode", ctypes.c_int32), ("algo", ctypes.c_int32),
        ("seed", ctypes.c_uint64),
        ("use_graph", ctypes.c_int32), ("x_on_device", ctypes.c_int32),
        ("nccl_comm", ctypes.c_void_p),
        ("ars_inject_width", ctypes.c_int32), ("reserved", ctypes.c_int32),
    ]


class HtlrInject(ctypes.Structure):
    _fields_ = [("normals", c_dp), ("n_normals", ctypes.c_int64),
                ("uniforms", c_dp), ("n_uniforms", ctypes.c_int64),
                ("gammas", c_dp), ("n_gammas", ctypes.c_int64),
                ("ars_uniforms", c_dp), ("n_ars_blocks", ctypes.c_int64)]


class HtlrIterStats(ctypes.Structure):
    _fields_ = [("loglike", ctypes.c_double), ("logw", ctypes.c_double), ("uvar", ctypes.c_double),
                ("rej", ctypes.c_double)]


class HtlrProfile(ctypes.Structure):
    _fields_ = [("sweep_ms", ctypes.c_double), ("sweep_bytes", ctypes.c_double),
                ("sweep_launches", ctypes.c_int64), ("total_ms", ctypes.c_double),
                ("kernel_launches", ctypes.c_int64), ("sigma_ms", ctypes.c_double),
                ("fused_sweep_ms", ctypes.c_double), ("fused_barrier_ms", ctypes.c_double),
                ("fused_rowphase_ms", ctypes.c_double), ("small_kernel_trajectories", ctypes.c_double),
                ("timeline_us", ctypes.c_double * 12), ("allreduce_ms", ctypes.c_double)]


TICK_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(HtlrIterStats),
                           ctypes.c_int32)

# every symbol include/htlr_cuda.h declares
EXPORTS = [
    "htlr_cuda_create", "htlr_cuda_destroy", "htlr_cuda_run", "htlr_cuda_fetch", "htlr_cuda_hist_len",
    "htlr_cuda_collect", "htlr_cuda_release_cache",
    "htlr_cuda_fit", "htlr_cuda_summary", "htlr_cuda_predict", "htlr_cuda_std", "htlr_cuda_eval", "htlr_cuda_trajectory", "htlr_cuda_get_state", "htlr_cuda_ars",
    "htlr_cuda_rng_dump", "htlr_cuda_stream", "htlr_cuda_profile_enable", "htlr_cuda_profile_get",
    "htlr_cuda_gendata_mlr", "htlr_cuda_gendata_fam", "htlr_cuda_launch_count", "htlr_cuda_ars_table_full", "htlr_cuda_predict_stats", "htlr_cuda_fp64_peak", "htlr_cuda_sync", "htlr_cuda_run_async", "htlr_cuda_nccl_unique_id",
    "htlr_cuda_nccl_init", "htlr_cuda_nccl_destroy", "htlr_cuda_last_error", "htlr_cuda_abi_version",
]

_lib = None


def build(verbose=False):
    """Compile the CUDA library in-tree (nvcc, sm_100a). Cross-compiles without a GPU."""
    subprocess.run(["make", "-s", "-j4", "-C", os.path.join(_HERE, "csrc")], check=True,
                   stdout=None if verbose else subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C htlr_b200/csrc`). htlr_b200 has no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    L.htlr_cuda_last_error.restype = ctypes.c_char_p
    L.htlr_cuda_last_error.argtypes = [ctypes.c_void_p]
    L.htlr_cuda_stream.restype = ctypes.c_void_p
    L.htlr_cuda_stream.argtypes = [ctypes.c_void_p]
    L.htlr_cuda_launch_count.restype = ctypes.c_int64
    L.htlr_cuda_launch_count.argtypes = [ctypes.c_void_p]
    L.htlr_cuda_gendata_mlr.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_double,
                                        ctypes.c_double, ctypes.c_uint64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                        ctypes.c_void_p, ctypes.c_void_p, c_dp]
    L.htlr_cuda_gendata_fam.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, c_dp,
                                        ctypes.c_int32, ctypes.c_int32, c_dp, ctypes.c_double, ctypes.c_int32,
                                        ctypes.c_uint64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                        ctypes.c_void_p, ctypes.c_void_p]
    L.htlr_cuda_predict_stats.argtypes = [ctypes.c_void_p, c_dp, c_dp]
    L.htlr_cuda_fp64_peak.argtypes = [ctypes.c_int32, c_dp, c_dp]
    L.htlr_cuda_ars_table_full.restype = ctypes.c_int64
    L.htlr_cuda_ars_table_full.argtypes = [ctypes.c_void_p]
    L.htlr_cuda_destroy.argtypes = [ctypes.c_void_p]
    L.htlr_cuda_destroy.restype = None
    L.htlr_cuda_release_cache.restype = None
    L.htlr_cuda_release_cache.argtypes = []
    L.htlr_cuda_hist_len.argtypes = [ctypes.c_void_p]
    L.htlr_cuda_create.argtypes = [ctypes.POINTER(HtlrCfg), c_dp, ctypes.c_int64, c_dp, c_lp, c_dp, c_dp,
                                   ctypes.POINTER(ctypes.c_void_p)]
    L.htlr_cuda_run.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(HtlrInject),
                                ctypes.POINTER(HtlrIterStats)]
    L.htlr_cuda_run_async.argtypes = [ctypes.c_void_p, ctypes.c_int32]
    L.htlr_cuda_sync.argtypes = [ctypes.c_void_p]
    L.htlr_cuda_fetch.argtypes = [ctypes.c_void_p] + [c_dp] * 8
    L.htlr_cuda_collect.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(HtlrIterStats)] + \
                                   [c_dp] * 7
    L.htlr_cuda_summary.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_double,
                                    c_dp, c_dp, c_ip]
    L.htlr_cuda_predict.argtypes = [ctypes.c_void_p, c_dp, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                    ctypes.c_int32, ctypes.c_int32, c_dp]
    L.htlr_cuda_std.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64,
                                ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64, c_dp, c_dp]
    L.htlr_cuda_get_state.argtypes = [ctypes.c_void_p] + [c_dp] * 6
    L.htlr_cuda_eval.argtypes = [ctypes.c_void_p, c_ip, ctypes.c_int32, c_dp, c_dp, c_dp, c_dp, c_dp]
    L.htlr_cuda_trajectory.argtypes = [ctypes.c_void_p, c_dp, ctypes.c_int32, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp,
                                       c_dp, c_ip]
    L.htlr_cuda_ars.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, c_dp, ctypes.c_double,
                                ctypes.c_double, ctypes.c_double, c_dp, ctypes.c_int32, ctypes.c_uint64,
                                ctypes.c_uint64, c_dp, c_ip, c_ip, c_ip]
    L.htlr_cuda_rng_dump.argtypes = [ctypes.c_int32, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint64,
                                     ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_double, c_dp]
    L.htlr_cuda_profile_enable.argtypes = [ctypes.c_void_p, ctypes.c_int32]
    L.htlr_cuda_profile_get.argtypes = [ctypes.c_void_p, ctypes.POINTER(HtlrProfile), ctypes.c_int32]
    L.htlr_cuda_fit.argtypes = [ctypes.POINTER(HtlrCfg), c_dp, ctypes.c_int64, c_dp, c_lp, c_dp, c_dp,
                                ctypes.POINTER(HtlrInject), TICK_FN, ctypes.c_void_p] + [c_dp] * 8 + \
                               [ctypes.c_char_p, ctypes.c_size_t]
    L.htlr_cuda_nccl_unique_id.argtypes = [ctypes.c_void_p]
    L.htlr_cuda_nccl_init.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                      ctypes.POINTER(ctypes.c_void_p)]
    L.htlr_cuda_nccl_destroy.argtypes = [ctypes.c_void_p]
    L.htlr_cuda_nccl_destroy.restype = None
    if L.htlr_cuda_abi_version() != ABI_VERSION:
        raise RuntimeError("libhtlr_cuda.so ABI version mismatch; rebuild")
    _lib = L
    return L
