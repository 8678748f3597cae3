"""The C-ABI library loads and exports every symbol include/ddpb.h declares; without a GPU every entry point that
needs the device fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    from ddp_pinocchio_b200 import capi
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__ as g
        g.build_cuda()
    return ctypes.CDLL(capi.LIB_PATH)


def test_every_declared_symbol_is_exported(pkg):
    hdr = open(os.path.join(ROOT, "include", "ddpb.h")).read()
    declared = set(re.findall(r"\b(ddpb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 35
    lib = _lib()
    for sym in sorted(declared):
        assert hasattr(lib, sym), sym
    from ddp_pinocchio_b200 import capi
    assert set(capi.EXPORTS) == declared


def test_model_validation_needs_no_gpu(pkg):
    from ddp_pinocchio_b200 import capi, models
    lib = _lib()
    lib.ddpb_last_error.restype = ctypes.c_char_p
    keep = []
    m = models.arm_6r()
    md = capi.model_desc_c(m, keep)
    h = ctypes.c_void_p()
    assert lib.ddpb_model_create(ctypes.byref(md), ctypes.byref(h)) == 0
    lib.ddpb_model_destroy(h)
    # a branching tree is rejected (serial chains only)
    m2 = models.arm_6r(); m2.parent = m2.parent.copy(); m2.parent[3] = 0
    md2 = capi.model_desc_c(m2, keep)
    assert lib.ddpb_model_create(ctypes.byref(md2), ctypes.byref(h)) == 2
    assert b"serial chains" in lib.ddpb_last_error()
    # pendulum_model_t's constructor asserts (pendulum_model.hpp:29-33)
    md3 = capi.model_desc_c(models.pendulum(mass=-1.0), keep)
    assert lib.ddpb_model_create(ctypes.byref(md3), ctypes.byref(h)) == 1


def test_fails_loudly_without_a_gpu(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from ddp_pinocchio_b200 import capi, synthetic
    with pytest.raises(capi.DdpbError, match="no CUDA device"):
        capi.Batch.from_workload(synthetic.workload("c2", B=2, T=4))


def test_model_functions_fail_loudly_without_a_gpu(pkg):
    """The Model concept's own functions run on the device too: argument errors are reported first, then the missing device."""
    import numpy as np
    import torch
    from ddp_pinocchio_b200 import capi, models
    mm = capi.Model(models.arm_6r())
    q = np.zeros((2, 6))
    with pytest.raises(capi.DdpbError, match="frame index"):
        mm.frame_coordinates(99, q)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.DdpbError, match="cudaSetDevice"):
        mm.dynamics_aba(q, q, q)


def test_backward_kernel_uses_tma_bulk_copies():
    """The backward pass stages slab records with TMA (cp.async.bulk -> SASS UBLKCP) completing on an mbarrier (SYNCS);
    checked on the built library's SASS (skipped where cuobjdump is not installed)."""
    import shutil
    import subprocess
    import pytest
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ddp-pinocchio_b200", "libddpb.so")
    # (kin_r: the all-revolute variant of the device code, csrc/variant.cuh)
    for fun in ("_ZN4ddpb5kin_r15backward_kernelILi6ELb0EEEvNS_8BatchDevE", "_ZN4ddpb5kin_r15backward_kernelILi7ELb1EEEvNS_8BatchDevE"):
        sass = subprocess.run([exe, "-sass", "-fun", fun, lib], capture_output=True, text=True).stdout
        assert "UBLKCP.S.G" in sass and "SYNCS.PHASECHK" in sass, fun
    # the lane-group kernel of the short chains streams its records with 16-byte cp.async (LDGSTS) and works in registers / shuffles
    sass = subprocess.run([exe, "-sass", "-fun", "_ZN4ddpb5kin_r18backward_lg_kernelILi2ELb0EEEvNS_8BatchDevE", lib], capture_output=True, text=True).stdout
    assert "LDGSTS" in sass and "SHFL.IDX" in sass and "DFMA" in sass
