"""The parity tests proper: CUDA kernels through the C ABI (libddpb.so) vs the CPU oracle."""
import numpy as np
import pytest

import parity_cases as pc

pytestmark = pytest.mark.gpu
CASES = pc.cases()
IDS = [c[0] for c in CASES]


def make(w):
    from ddp_pinocchio_b200 import capi
    return capi.Batch.from_workload(w)


@pytest.mark.parametrize("tag,w", CASES, ids=IDS)
def test_linearization(tag, w):
    pc.check_linearization(make, w)


@pytest.mark.parametrize("tag,w", CASES, ids=IDS)
def test_backward(tag, w):
    pc.check_backward(make, w)


@pytest.mark.parametrize("tag,w", CASES, ids=IDS)
def test_forward(tag, w):
    pc.check_forward(make, w)


@pytest.mark.parametrize("tag,w", [c for c in CASES if c[0] in ("c2_chain2r", "c3_arm6r_terminal", "c5_arm7r_fdvalues", "c1_pendulum_fdvalues")],
                         ids=lambda v: v if isinstance(v, str) else "")
def test_tensors_vs_quad(tag, w):
    pc.check_tensors_vs_quad(make, w)


@pytest.mark.parametrize("name,T,iters", [("c2", 40, 6), ("c3", 20, 5), ("c4", 12, 4), ("c5", 10, 3)])
def test_solve(name, T, iters):
    w = pc.S.workload(name, B=3, T=T)
    stats = pc.check_solve(make, w, iters)
    for s in stats:
        assert s["dev_vs_quad"] <= 50 * s["double_vs_quad"] + 1e-8, s
        assert s["dev_vs_double"] <= 50 * s["double_vs_quad"] + 1e-8, s
    assert sum(s["same_flow"] for s in stats) >= len(stats) - 1, stats


def test_gpu_matches_cpu_emulation_of_same_kernels():
    """The CUDA kernels and their CPU emulation run the same source; results agree to rounding (FMA contraction,
    sincos) on first-order quantities."""
    import hostemu_lib as he
    w = pc.S.workload("c3", B=4, T=6)
    g = w.generate()
    outs = []
    for mk in (make, he.EmuBatch.from_workload):
        bt = mk(w)
        bt.set_targets(g["targets"]); bt.make_trajectory(g["x0"], g["us"]); bt.compute_derivatives()
        outs.append(bt.get_derivatives(["fx", "fu", "f_val", "eq_x"]))
        bt.close()
    for f in outs[0]:
        assert pc.relerr(outs[0][f], outs[1][f]) < 1e-12
