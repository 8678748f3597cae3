"""Kernel arithmetic vs the oracle WITHOUT a GPU: the CTA bodies of csrc/*.cuh are compiled for the host and
run one thread at a time (tests/hostemu).  Same checks as tests/test_gpu_parity.py."""
import pytest

import parity_cases as pc
import hostemu_lib as he

CASES = pc.cases()
IDS = [c[0] for c in CASES]


def make(w):
    return he.EmuBatch.from_workload(w)


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


def test_2r_chain_vs_textbook_closed_form():
    pc.check_2r_vs_textbook(make, pc.S.workload("c2", B=3, T=8))


@pytest.mark.parametrize("name", ["c3", "c5"])
def test_arm_records_vs_lagrangian_mechanics(name):
    pc.check_vs_lagrangian(make, pc.S.workload(name, B=2, T=4))


@pytest.mark.parametrize("tag", ["scara", "prprr", "rp"])
def test_prismatic_records_vs_lagrangian_mechanics(tag):
    """Chains with sliding joints: f(x,u) and f_u = dt M^-1 of the records against Lagrange's equations (numpy kinematics only)."""
    pc.check_vs_lagrangian(make, pc.prismatic_workload(tag, 2, 4))


@pytest.mark.parametrize("tag,T,iters,floor", [("scara", 16, 4, 1e-6), ("scara_spatial", 12, 4, 1e-6), ("rp", 20, 5, 1e-8)])
def test_solve_prismatic(tag, T, iters, floor):
    """End to end on chains with a sliding joint.  The SCARA's light wrist turns about an axis 0.7 m from the world origin:
    in world-origin Pluecker coordinates its 0.003 kg m^2 is what is left of terms of 0.5 kg m^2, so the first-order records
    carry 1e-13 instead of 1e-15 (still 4 digits inside the 1e-9 tolerance) and the reference's forward-difference Hessians
    (eps = 2^-26) amplify that to 1e-7 in the trajectories: the floor of the end-to-end distance is 1e-6 there."""
    stats = pc.check_solve(make, pc.prismatic_workload(tag, 3, T), iters)
    for s in stats:
        assert s["dev_vs_quad"] <= 50 * s["double_vs_quad"] + floor, s
        assert s["dev_vs_double"] <= 50 * s["double_vs_quad"] + floor, s
    assert sum(s["same_flow"] for s in stats) >= len(stats) - 1, stats


@pytest.mark.parametrize("name,T,iters", [("c2", 40, 6), ("c3", 20, 5), ("c4", 12, 4), ("c5", 10, 3)])
def test_solve(name, T, iters):
    w = pc.S.workload(name, B=3, T=T)
    stats = pc.check_solve(make, w, iters)
    for s in stats:
        assert s["dev_vs_quad"] <= 50 * s["double_vs_quad"] + 1e-8, s
        assert s["dev_vs_double"] <= 50 * s["double_vs_quad"] + 1e-8, s
    assert sum(s["same_flow"] for s in stats) >= len(stats) - 1, stats


@pytest.mark.parametrize("name", ["c2", "c3"])
def test_slab_structure_is_enforced(name):
    pc.check_structure_is_enforced(make, pc.S.workload(name, B=2, T=5), RuntimeError)


@pytest.mark.parametrize("name", ["c2", "c3"])
def test_forward_many_halvings(name):
    seen = pc.check_forward_many_halvings(make, pc.S.workload(name, B=7, T=8))
    assert max(seen) > 4      # some problem needed more than four trials


@pytest.mark.parametrize("tag,m", pc.model_concept_models(), ids=[t for t, _ in pc.model_concept_models()])
def test_model_concept_functions(tag, m):
    """dynamics_aba / d_dynamics_aba / inverse_dynamics_rnea / frame_coordinates / d_frame_coordinates (csrc/model_eval.cuh)."""
    pc.check_model_concept(he.EmuModel, m)


@pytest.mark.parametrize("name", ["c2", "c3", "c4", "c5"])
def test_single_knot_batches(name):
    """Horizon 1: what the per-call forms of the facade's Dynamics / Problem concepts run on (ddpb.hpp detail::knot_derivatives)."""
    pc.check_linearization(make, pc.S.workload(name, B=3, T=1))
