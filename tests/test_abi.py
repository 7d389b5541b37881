"""The C-ABI library loads and exports every symbol include/htlr_cuda.h declares; without a GPU every
compute entry point fails loudly (there is no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "htlr_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(htlr_cuda_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import htlr_b200
    lib = htlr_b200._lib.lib()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(htlr_b200._lib.EXPORTS) == names
    assert lib.htlr_cuda_abi_version() == htlr_b200._lib.ABI_VERSION == 2


def test_cfg_struct_layout_matches_header():
    import htlr_b200
    # 2*u32, 4*i32, 3*f64, 5*i32 (+pad), 2*f64, 6*i32, u64, 2*i32, ptr, 2*i32
    assert ctypes.sizeof(htlr_b200._lib.HtlrCfg) == 144
    assert htlr_b200._lib.HtlrCfg.seed.offset == 112 and htlr_b200._lib.HtlrCfg.nccl_comm.offset == 128
    assert ctypes.sizeof(htlr_b200._lib.HtlrIterStats) == 32
    # htlr_profile: 10 eight-byte scalars + timeline_us[12] + allreduce_ms
    assert ctypes.sizeof(htlr_b200._lib.HtlrProfile) == 8 * 23
    assert htlr_b200._lib.HtlrProfile.timeline_us.offset == 80


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "htlr_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "liboracle" not in txt and "_ref/" not in txt, f


def test_fails_loudly_without_a_gpu():
    import htlr_b200
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(htlr_b200.HtlrCudaError, match="no CUDA device|no CPU fallback"):
        htlr_b200.Sampler(3, 1, 4, np.ones((4, 4)), np.zeros((4, 1)), np.zeros(4), "t", 1, -10, 0, 2, 2, 1, 3, 2,
                          0.3, 0.05, np.zeros((4, 1)), np.ones(4))
    with pytest.raises(htlr_b200.HtlrCudaError):
        htlr_b200.ars_sample("ghs", np.ones(4), 2, 1.0, -10.0)


def test_argument_validation_happens_before_any_device_work():
    import htlr_b200
    with pytest.raises(htlr_b200.HtlrCudaError, match="Unsupported prior"):
        htlr_b200.Sampler(3, 1, 4, np.ones((4, 4)), np.zeros((4, 1)), np.zeros(4), "laplace", 1, -10, 0, 2, 2, 1, 3,
                          2, 0.3, 0.05, np.zeros((4, 1)), np.ones(4))
    with pytest.raises(htlr_b200.HtlrCudaError, match="X must be"):
        htlr_b200.Sampler(3, 1, 4, np.ones((4, 3)), np.zeros((4, 1)), np.zeros(4), "t", 1, -10, 0, 2, 2, 1, 3, 2,
                          0.3, 0.05, np.zeros((4, 1)), np.ones(4))
