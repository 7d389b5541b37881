"""Properties of the BUILT library that the measured numbers depend on and a source change can silently lose (no GPU needed):
the trajectory kernel's L1 is all shared memory, so a spilled register is an L2 round trip — the instantiations the bench
workloads run (config C: K = 2 grid mode; config B full: K = 3 grid mode; config A: K = 1 cluster mode) must not use a stack
frame, and every instantiation must stay at 128 registers or fewer (16 warps per block)."""
import re
import shutil
import subprocess

import pytest

import htlr_b200._lib as L


def _res_usage():
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    out = subprocess.run(["cuobjdump", "-res-usage", L.LIB_PATH], capture_output=True, text=True, check=True).stdout
    use = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+)", out):
        use[m.group(1)] = (int(m.group(2)), int(m.group(3)))
    return use


def test_hot_trajectory_instantiations_have_no_stack_frame():
    use = _res_usage()
    traj = {k: v for k, v in use.items() if "6k_trajILi" in k}
    assert len(traj) >= 48                                   # K = 1..8 x group widths x grid / cluster mode
    assert max(r for r, _ in traj.values()) <= 128           # 512 threads per block: the register file is full at 128
    for inst in ("ILi2ELi2ELi8ELb0E", "ILi3ELi2ELi4ELb0E", "ILi1ELi1ELi1ELb1E", "ILi1ELi1ELi1ELb0E"):
        hits = [v for k, v in traj.items() if inst in k]
        assert hits and all(stack == 0 for _, stack in hits), (inst, hits)


def test_every_kernel_family_is_in_the_library():
    names = " ".join(_res_usage())
    for fam in ("k_traj", "k_traj_small", "k_traj_rows", "k_lv_sweep", "k_grad_sweep", "k_softmax", "k_select", "k_begin",
                "k_energy", "k_tail_t", "k_sigma_ars", "k_predict_gemm"):
        assert fam in names, fam
