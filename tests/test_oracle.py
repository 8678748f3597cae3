"""The oracle itself: identities of the rigid-body arithmetic (the part the reference delegates to Pinocchio and
that no reference test pins), __float128 finite differences, pendulum closed forms in mpmath, and the committed
golden vectors."""
import os

import mpmath as mp
import numpy as np
import pytest

import ddpb_loader

P = ddpb_loader.load()
import oracle_lib as ol  # noqa: E402
from parity_cases import planar_2r_closed_form, lagrangian_forward_dynamics  # noqa: E402

S = P.synthetic
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODELS = [P.models.chain_2r, P.models.arm_6r, P.models.arm_7r, lambda: P.models.chain_nr(5), lambda: P.models.chain_nr(1),
          P.models.scara_rrpr, lambda: P.models.chain_mixed(5, "PRPRR"), lambda: P.models.chain_mixed(2, "RP"),
          lambda: P.models.chain_mixed(1, "P")]
MODEL_IDS = ["2r", "6r", "7r", "5r_random", "1r_random", "scara_rrpr", "prprr_random", "rp_random", "p_random"]


@pytest.mark.parametrize("mk", MODELS, ids=MODEL_IDS)
def test_aba_satisfies_inverse_dynamics(mk):
    """M(q) aba(q,v,tau) + nle(q,v) = tau, with an independently written RNEA."""
    m = mk(); n = m.nq
    o = ol.Oracle(m, 4)
    rng = np.random.RandomState(0)
    for _ in range(5):
        q, v, tau = rng.uniform(-2, 2, n), rng.uniform(-1, 1, n), rng.uniform(-5, 5, n)
        a = o.aba(q, v, tau)
        assert np.abs(o.rnea(q, v, a) - tau).max() < 1e-12 * max(1.0, np.abs(tau).max())
        assert np.abs(a - o.aba(q, v, tau, prec=1)).max() < 1e-12 * max(1.0, np.abs(a).max())


@pytest.mark.parametrize("mk", MODELS, ids=MODEL_IDS)
def test_forward_dynamics_is_affine_in_the_torque(mk):
    """aba(q, v, tau + d) = aba(q, v, tau) + M^-1(q) d: the identity behind the torque-affine evaluation points of the
    value-based linearization (csrc/linearize.cuh), at the reference's step eps = 2^-13 and at a large step, in __float128
    (identity to rounding) and in double (the two sides differ by a few ulp of the acceleration: the noise a second ABA
    run carries into the stencil anyway)."""
    m = mk(); n = m.nq
    o = ol.Oracle(m, 4)
    rng = np.random.RandomState(7)
    for step in (2.0 ** -13, 3.0):
        q, v, tau = rng.uniform(-2, 2, n), rng.uniform(-1, 1, n), rng.uniform(-5, 5, n)
        d = np.zeros(n); d[rng.randint(n)] = step
        if n > 1:
            d[rng.randint(n)] += step
        for prec, tol in ((1, 1e-15), (0, 1e-12)):
            minv = np.asarray(o.d_aba(q, v, tau, prec=prec)[2]).reshape(n, n)
            a0, a1 = o.aba(q, v, tau, prec=prec), o.aba(q, v, tau + d, prec=prec)
            assert np.abs(a1 - (a0 + minv @ d)).max() <= tol * max(1.0, np.abs(a1).max()), (step, prec)


@pytest.mark.parametrize("mk", MODELS, ids=MODEL_IDS)
def test_aba_derivatives_vs_quad_central_differences(mk):
    m = mk(); n = m.nq
    o = ol.Oracle(m, 4)
    rng = np.random.RandomState(1)
    q, v, tau = rng.uniform(-2, 2, n), rng.uniform(-1, 1, n), rng.uniform(-5, 5, n)
    dq, dv, dt = o.d_aba(q, v, tau)
    h = 2e-5
    for J, which in ((dq, 0), (dv, 1), (dt, 2)):
        fd = np.zeros((n, n))
        for j in range(n):
            e = np.zeros(n); e[j] = h
            args = [q, v, tau]
            ap = list(args); ap[which] = args[which] + e
            am = list(args); am[which] = args[which] - e
            fd[:, j] = (o.aba(*ap, prec=1) - o.aba(*am, prec=1)) / (2 * h)
        assert np.abs(J - fd).max() <= 2e-8 * max(1.0, np.abs(J).max())
    assert np.abs(dt - dt.T).max() < 1e-12 * np.abs(dt).max()  # full symmetric M^-1 (pinocchio_model.ipp:391-399)


def test_aba_2r_chain_vs_textbook_closed_form():
    """External pin of the rigid-body arithmetic on the bench model (BASELINE.json configs[1]): the oracle's ABA and
    M^-1 against the published closed-form dynamics of the planar double pendulum."""
    m = P.models.chain_2r()
    assert np.allclose(m.axis, [[0, 1, 0], [0, 1, 0]]) and np.allclose(m.com[:, :2], 0) and np.allclose(m.placement_trans[:, :2], 0)
    o = ol.Oracle(m, 4)
    rng = np.random.RandomState(11)
    for _ in range(20):
        q, v, tau = rng.uniform(-3, 3, 2), rng.uniform(-2, 2, 2), rng.uniform(-5, 5, 2)
        a_ref, M = planar_2r_closed_form(m, q, v, tau)
        a = o.aba(q, v, tau)
        assert np.abs(a - a_ref).max() < 1e-12 * max(1.0, np.abs(a_ref).max())
        _, _, minv = o.d_aba(q, v, tau)
        assert np.abs(minv - np.linalg.inv(M)).max() < 1e-12 * np.abs(minv).max()


@pytest.mark.parametrize("mk", MODELS, ids=MODEL_IDS)
def test_aba_vs_lagrangian_mechanics(mk):
    """External check of ABA and M^-1 for every model (the arms have no closed form): Lagrange's equations with the
    mass matrix from geometric link Jacobians and Christoffel symbols by central differences."""
    m = mk(); n = m.nq
    o = ol.Oracle(m, 4)
    rng = np.random.RandomState(5)
    for _ in range(4):
        q, v, tau = rng.uniform(-2, 2, n), rng.uniform(-1, 1, n), rng.uniform(-5, 5, n)
        a_ref, M = lagrangian_forward_dynamics(m, q, v, tau)
        a = o.aba(q, v, tau)
        assert np.abs(a - a_ref).max() < 1e-9 * max(1.0, np.abs(a_ref).max()), np.abs(a - a_ref).max()  # measured: 4e-12
        _, _, minv = o.d_aba(q, v, tau)
        assert np.abs(minv @ M - np.eye(n)).max() < 1e-10 * np.linalg.cond(M)


@pytest.mark.parametrize("mk", [P.models.arm_6r, P.models.scara_rrpr, lambda: P.models.chain_mixed(5, "PRPRR")], ids=["6r", "scara_rrpr", "prprr_random"])
def test_frame_jacobian_conventions(mk):
    m = mk(); n = m.nq
    o = ol.Oracle(m, 4)
    q = np.random.RandomState(2).uniform(-2, 2, n)
    fr = m.n_frames - 1
    p, J_lwa = o.frame(fr, q, mode=1)
    _, J_world = o.frame(fr, q, mode=0)
    fd = np.zeros((3, n))
    for j in range(n):
        e = np.zeros(n); e[j] = 1e-6
        fd[:, j] = (o.frame(fr, q + e, 1, prec=1)[0] - o.frame(fr, q - e, 1, prec=1)[0]) / 2e-6
    assert np.abs(J_lwa - fd).max() < 1e-9                 # LOCAL_WORLD_ALIGNED = true dp/dq
    assert np.abs(p - P.models.frame_position(m, q, fr)).max() < 1e-13   # numpy FK of the input generator
    # WORLD convention (what the reference uses, Appendix B.1): velocity of the point at the world origin;
    # J_world(:,j) = J_lwa(:,j) - w_j x p, so the two agree only for a frame at the origin
    rev = m.joint_types == 0
    assert np.abs(J_world - J_lwa)[:, rev].max() > 1e-2
    assert np.abs(J_world - J_lwa)[:, ~rev].max() == 0.0 if (~rev).any() else True   # a prismatic column is its axis in both
    w = np.stack([np.cross(J_lwa[:, j], J_world[:, j] - J_lwa[:, j]) for j in range(n)])  # orthogonality sanity
    assert np.isfinite(w).all()


def test_pendulum_closed_forms_mpmath():
    """pendulum_model.hpp:105-130 + dynamics_t (problem.hpp:441-503), restated with mpmath at 40 digits."""
    mp.mp.dps = 40
    o = ol.Oracle(P.models.pendulum(1.0, 1.0), 4)
    rng = np.random.RandomState(3)
    for _ in range(10):
        q, v, u = rng.uniform(-3, 3), rng.uniform(-2, 2), rng.uniform(-4, 4)
        g = mp.mpf("9.81")
        a = -g / 1 * mp.sin(mp.mpf(q)) + mp.mpf(u) / 1
        f, fx, fu = o.first_order(np.array([q, v]), np.array([u]))
        dt = mp.mpf("0.01")
        assert abs(f[0] - float(mp.mpf(q) + dt * mp.mpf(v))) < 1e-15 * max(1, abs(f[0]))
        assert abs(f[1] - float(mp.mpf(v) + dt * a)) < 4e-16 * max(1, abs(f[1]))
        assert abs(fx[1, 0] - float(-g * mp.cos(mp.mpf(q)) * dt)) < 1e-16
        assert fx[0, 0] == 1.0 and fx[0, 1] == 0.01 and fx[1, 1] == 1.0 and fu[0, 0] == 0.0 and fu[1, 0] == 0.01


def test_fd_steps_are_the_reference_epsilons():
    # problem.hpp:107 sqrt(eps), :188 sqrt(sqrt(eps))
    assert np.sqrt(np.finfo(float).eps) == 2.0 ** -26 and np.sqrt(np.sqrt(np.finfo(float).eps)) == 2.0 ** -13


@pytest.mark.parametrize("name", ["c1", "c2", "c3", "c4", "c5"])
def test_oracle_reproduces_golden(name):
    """Regression pin of the oracle (tolerances rather than bits: libm's sin/cos may differ between CPUs)."""
    z = np.load(os.path.join(GOLD, f"{name}.npz"))
    w = S.workload(name, B=2, T=int(z["xs0_0"].shape[0]) - 1)
    g = w.generate()
    assert np.array_equal(g["x0"], z["x0"]) and np.array_equal(g["targets"], z["targets"]) and np.array_equal(g["mult_jac"], z["mult_jac"])
    for b in range(2):
        o = ol.from_workload(w, g["targets"][b])
        xs0 = o.rollout(g["x0"][b], g["us"][b])
        np.testing.assert_allclose(xs0, z[f"xs0_{b}"], rtol=1e-13, atol=1e-14)
        d = o.compute_derivatives(xs0, g["us"][b])
        for f in ("f_val", "fx", "fu", "eq_val", "eq_x", "eq_u"):
            np.testing.assert_allclose(d[f], z[f"{f}_{b}"], rtol=1e-12, atol=1e-13, err_msg=f)
        for f in ("fxx", "fux", "eq_xx"):
            np.testing.assert_allclose(d[f], z[f"{f}_{b}"], rtol=0, atol=1e-5, err_msg=f)
        r = o.solve(xs0, g["us"][b], g["mult_jac"][b], int(z["iters"]), float(z["mu"]))
        np.testing.assert_allclose(r["xs"], z[f"xs_{b}"], rtol=0, atol=1e-5 * max(1.0, np.abs(z[f"xs_{b}"]).max()))
        assert r["log"].shape == z[f"log_{b}"].shape


def test_flop_count_of_one_first_order_evaluation():
    """W_fo of BASELINE.md section 3, pinned with the counting scalar (oracle arithmetic, dense 6x6 algebra)."""
    for mk, lo, hi in ((P.models.arm_6r, 3e4, 8e4), (P.models.arm_7r, 4e4, 1e5), (P.models.chain_2r, 4e3, 2e4)):
        m = mk(); o = ol.Oracle(m, 4)
        x = np.linspace(-1, 1, 2 * m.nq); u = np.linspace(-2, 2, m.nq)
        assert lo < o.flops_first_order(x, u) < hi
