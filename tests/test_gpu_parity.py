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


@pytest.mark.parametrize("name", ["c1", "c2", "c3", "c4", "c5"])
def test_against_committed_golden_vectors(name):
    """CUDA path vs tests/golden/*.npz (made by tests/golden/make_golden.py from the oracle)."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"{name}.npz"))
    w = pc.S.workload(name, B=2, T=int(z["xs0_0"].shape[0]) - 1)
    bt = make(w)
    bt.set_targets(z["targets"]); bt.make_trajectory(z["x0"], z["us0"]); bt.set_multipliers(None, z["mult_jac"])
    xs, _ = bt.get_trajectory()
    bt.compute_derivatives()
    D = bt.get_derivatives(["f_val", "fx", "fu", "eq_val", "eq_x", "eq_u"])
    for b in range(2):
        assert pc.relerr(xs[b], z[f"xs0_{b}"]) < pc.TOL_ROLLOUT
        for f in D:
            assert pc.relerr(D[f][b], z[f"{f}_{b}"]) < pc.TOL_FIRST, f
    bt.solve(int(z["iters"]), float(z["mu"]))
    xs, us = bt.get_trajectory()
    for b in range(2):
        # end-to-end: as close to the __float128 oracle as the double oracle is (DESIGN.md "Tolerances")
        e_ref = pc.relerr(z[f"xs_{b}"], z[f"xsq_{b}"])
        assert pc.relerr(xs[b], z[f"xsq_{b}"]) <= 50 * e_ref + 1e-8, (pc.relerr(xs[b], z[f"xsq_{b}"]), e_ref)
    bt.close()


def test_full_size_properties_c3():
    """BASELINE.json sizes (1024 x 100, 6R): size-independent properties instead of an oracle run --
    identical problems give identical results wherever they sit in the batch, the rollout reproduces x_{t+1} = f(x_t,u_t)
    of the derivative slabs bit for bit, top rows of the FD tensors are exactly zero, guarded status words are set."""
    w = pc.S.workload("c3")
    g = w.generate()
    for k in ("x0", "us", "targets", "mult_jac"):   # problems 0 and 777 become copies of problem 5
        g[k][0] = g[k][5]; g[k][777] = g[k][5]
    bt = make(w)
    bt.set_targets(g["targets"]); bt.make_trajectory(g["x0"], g["us"]); bt.set_multipliers(None, g["mult_jac"])
    xs, us = bt.get_trajectory()
    bt.compute_derivatives()
    D = bt.get_derivatives(["f_val", "fxx", "fuu", "fx"])
    n, nd = w.n, w.ndx
    assert np.array_equal(D["f_val"].reshape(w.B, w.T, w.nx), xs[:, 1:, :])          # f(x_t,u_t) == x_{t+1} of the rollout
    fxx = D["fxx"].reshape(w.B, w.T, nd, nd, nd)                                      # [.., j(right), i(left), k(out)]
    assert not fxx[..., :n].any() and fxx[..., n:].any()                              # rows k < n are exact zeros
    assert not D["fuu"].any()                                                         # same q => same M^-1 => exact zeros
    sc = bt.solve(3, w.mu)
    xs2, us2 = bt.get_trajectory()
    assert np.array_equal(xs2[0], xs2[5]) and np.array_equal(xs2[777], xs2[5]) and np.array_equal(us2[0], us2[5])
    assert np.isfinite(xs2[sc["status"] == 0]).all() and (sc["iterations"][sc["status"] == 0] == 3).all()
    assert (sc["status"] == 2).sum() <= w.B // 50
    bt.close()


@pytest.mark.parametrize("name,B,T", [("c2", 200, 30), ("c3", 160, 12)])
def test_solve_does_not_depend_on_subbatches(name, B, T):
    """ddpb_set_subbatches only changes how the batch is scheduled (streams); problems are independent, so every
    trajectory, scalar and counter is bit-identical for 1, 2 and 5 sub-batches (ragged last sub-batch included)."""
    w = pc.S.workload(name, B=B, T=T)
    g = w.generate()
    ref = None
    for nsub in (1, 2, 5):
        bt = make(w)
        bt.set_subbatches(nsub)
        bt.set_targets(g["targets"]); bt.make_trajectory(g["x0"], g["us"]); bt.set_multipliers(None, g["mult_jac"])
        sc = bt.solve(4, w.mu)
        xs, us = bt.get_trajectory()
        bt.close()
        cur = (xs, us, sc["mu"], sc["reg"], sc["rollouts"], sc["restarts"], sc["merit"])
        if ref is None:
            ref = cur
        else:
            for a, b in zip(ref, cur):
                assert np.array_equal(a, b, equal_nan=True)


def test_full_size_roundtrip_properties_c2():
    """BASELINE.json configs[1] at full size (4096 x 200, 2R chain): size-independent properties -- the derivative
    record's f(x_t,u_t) equals the rollout bit for bit, rows k < n of f_xx and all of f_uu are exact zeros, a solve keeps
    every non-diverged trajectory finite and counts exactly the requested iterations."""
    w = pc.S.workload("c2")
    g = w.generate()
    bt = make(w)
    bt.set_targets(g["targets"]); bt.make_trajectory(g["x0"], g["us"]); bt.set_multipliers(None, g["mult_jac"])
    xs, us = bt.get_trajectory()
    bt.compute_derivatives()
    D = bt.get_derivatives(["f_val", "fxx", "fuu"])
    n, nd = w.n, w.ndx
    assert np.array_equal(D["f_val"].reshape(w.B, w.T, w.nx), xs[:, 1:, :])
    fxx = D["fxx"].reshape(w.B, w.T, nd, nd, nd)
    assert not fxx[..., :n].any() and fxx[..., n:].any() and not D["fuu"].any()
    sc = bt.solve(2, w.mu)
    xs2, _ = bt.get_trajectory()
    ok = sc["status"] == 0
    assert np.isfinite(xs2[ok]).all() and (sc["iterations"][ok] == 2).all() and (sc["status"] == 2).sum() <= w.B // 50
    bt.close()


def test_repeated_solves_replay_identically():
    """ddpb_solve replays a captured CUDA graph; an odd iteration count leaves the trajectory buffers swapped, so the
    next solve runs with the roles exchanged.  Every repetition from the same start must give the same bits, and
    the phase-timed (direct, ungraphed) path must agree with the replayed one."""
    w = pc.S.workload("c3", B=96, T=10)
    g = w.generate()
    bt = make(w)
    outs = []
    for rep in range(5):
        if rep == 4:
            bt.enable_phase_timing(True)   # direct enqueue instead of the graph
        bt.set_targets(g["targets"]); bt.make_trajectory(g["x0"], g["us"]); bt.set_multipliers(None, g["mult_jac"])
        sc = bt.solve(3, w.mu)
        xs, us = bt.get_trajectory()
        outs.append((xs, us, sc["mu"], sc["rollouts"], sc["restarts"], sc["merit"]))
    bt.phase_times()
    bt.close()
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            assert np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("name", ["c2", "c3"])
def test_slab_structure_is_enforced(name):
    from ddp_pinocchio_b200 import capi
    pc.check_structure_is_enforced(make, pc.S.workload(name, B=2, T=5), capi.DdpbError)


@pytest.mark.parametrize("name,B", [("c2", 70), ("c3", 40), ("c5", 9)])
def test_forward_many_halvings(name, B):
    """All line-search trials run at once in two waves (8 or 16 trials, then the rest up to 34): the accepted step must be
    the one the reference's sequential halving finds, whichever wave resolves the problem."""
    seen = pc.check_forward_many_halvings(make, pc.S.workload(name, B=B, T=8))
    assert max(seen) > 16 and min(seen) == 1, seen      # both waves exercised, and the full step somewhere


def test_guarded_exit_while_other_kernels_run():
    """A problem whose Cholesky keeps failing at a knot t > 0 leaves the backward kernel through the guarded exit while
    the copy staged for knot t-1 is still in flight; the kernel must drain it before its shared memory is handed to
    another kernel.  Two batches on their own streams, driven from two host threads: batch A's backward pass hits the
    guard on every problem (doctored f_xx: the value function turns indefinite three knots into the pass) in a loop,
    batch B linearizes concurrently; B's records must equal the ones it produces alone, every time."""
    import threading
    import torch
    wa = pc.S.workload("c3", B=256, T=12)
    ga, bta, usa, orcs, rng = pc.setup(make, wa)
    bta.set_stream(torch.cuda.Stream().cuda_stream)
    bta.set_multipliers(rng.uniform(-1, 1, (wa.B, bta.ne_total)), ga["mult_jac"])
    bta.compute_derivatives()
    D = bta.get_derivatives()
    D["fxx"] = D["fxx"] - 1e12 * (np.abs(D["fxx"]) > 0)      # only state-dependent entries change: the structure is kept
    bta.set_derivatives(D)
    wb = pc.S.workload("c3", B=512, T=20)
    gb, btb, usb, _, _ = pc.setup(make, wb)
    btb.set_stream(torch.cuda.Stream().cuda_stream)
    btb.compute_derivatives()
    ref = btb.get_derivatives(["fx", "fxx", "f_val"])
    stop = threading.Event()
    seen_guard = []

    def hammer():
        while not stop.is_set():
            bta.set_scalar("status", 0.0)
            bta.backward_pass(1e4, 0.0)
            seen_guard.append(int((bta.get_scalars()["status"] == 2).sum()))

    th = threading.Thread(target=hammer)
    th.start()
    try:
        for _ in range(25):
            btb.compute_derivatives()
            cur = btb.get_derivatives(["fx", "fxx", "f_val"])
            for f in ref:
                assert np.array_equal(cur[f], ref[f]), f
    finally:
        stop.set(); th.join()
    assert seen_guard and max(seen_guard) > wa.B // 2, seen_guard[:5]     # the guarded exit was really taken
    bta.close(); btb.close()


@pytest.mark.parametrize("name,B,T,iters", [("c2", 33, 12, 3), ("c2", 1, 9, 3), ("c1", 5, 14, 4), ("c3", 9, 6, 2)])
def test_solve_ragged_batch_sizes(name, B, T, iters):
    """Batch sizes that do not fill a warp / a lane group / a block of 32 line-search problems: same checks as test_solve."""
    w = pc.S.workload(name, B=B, T=T)
    stats = pc.check_solve(make, w, iters)
    for s in stats:
        assert s["dev_vs_quad"] <= 50 * s["double_vs_quad"] + 1e-8, s
    assert sum(s["same_flow"] for s in stats) >= len(stats) - max(1, len(stats) // 10), stats


def make_model(m):
    from ddp_pinocchio_b200 import capi
    return capi.Model(m)


@pytest.mark.parametrize("tag,m", pc.model_concept_models(), ids=[t for t, _ in pc.model_concept_models()])
def test_model_concept_functions(tag, m):
    """ddpb_model_dynamics_aba / d_dynamics_aba / inverse_dynamics_rnea / frame_coordinates / d_frame_coordinates vs the oracle."""
    pc.check_model_concept(make_model, m)


def test_model_concept_functions_many_samples():
    """A batch that spans several CTAs with a ragged tail, against the single-sample calls."""
    from ddp_pinocchio_b200 import models
    mm = make_model(models.arm_6r())
    rng = np.random.RandomState(5)
    n = 64 * 3 + 17
    q, v, tau = rng.uniform(-2, 2, (n, 6)), rng.uniform(-1, 1, (n, 6)), rng.uniform(-5, 5, (n, 6))
    a = mm.dynamics_aba(q, v, tau)
    for s in (0, 63, 64, n - 1):
        assert np.array_equal(a[s], mm.dynamics_aba(q[s], v[s], tau[s])[0])
    assert pc.relerr(mm.inverse_dynamics_rnea(q, v, a), tau) < 1e-10


def test_model_concept_argument_checks():
    from ddp_pinocchio_b200 import capi, models
    mm = make_model(models.arm_6r())
    q = np.zeros((1, 6))
    with pytest.raises(capi.DdpbError, match="frame index"):
        mm.frame_coordinates(99, q)
    with pytest.raises(capi.DdpbError, match="Jacobian convention"):
        mm.d_frame_coordinates(0, q, jac_mode=7)
    pm = make_model(models.pendulum())
    with pytest.raises(capi.DdpbError, match="pendulum"):
        pm.inverse_dynamics_rnea(np.zeros((1, 1)), np.zeros((1, 1)), np.zeros((1, 1)))


@pytest.mark.parametrize("name", ["c2", "c3", "c4", "c5"])
def test_single_knot_batches(name):
    """Horizon 1: what the per-call forms of the facade's Dynamics / Problem concepts run on (ddpb.hpp detail::knot_derivatives)."""
    pc.check_linearization(make, pc.S.workload(name, B=3, T=1))
