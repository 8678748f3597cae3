"""Parity checks shared by the CPU-emulated kernels (tests/test_emu_parity.py, no GPU) and the real CUDA
path through the C ABI (tests/test_gpu_parity.py, -m gpu).  `make(w)` builds a batch for a Workload."""
import numpy as np

import ddpb_loader

P = ddpb_loader.load()
import oracle_lib as ol  # noqa: E402

S = P.synthetic
FIRST_ORDER = ["lx", "lu", "lxx", "lux", "luu", "f_val", "fx", "fu", "eq_val", "eq_x", "eq_u"]
TENSORS = ["fxx", "fux", "fuu", "eq_xx", "eq_ux", "eq_uu"]
U = 2.0 ** -53
# Tolerances.  north_star: Jacobians and gains within 1e-9 relative, trajectories / cost within 1e-8 relative.
TOL_FIRST = 1e-10       # f, fx, fu, eq, eq_x, eq_u (relative to the block's max magnitude)
TOL_GAINS = 1e-9        # k, K given identical derivative slabs
TOL_ROLLOUT = 1e-10     # trajectories / merit given identical gains


def relerr(a, b):
    a = np.asarray(a, dtype=float).ravel(); b = np.asarray(b, dtype=float).ravel()
    if a.size == 0:
        return 0.0
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def cases():
    """(id, workload) pairs covering the five BASELINE.json configs at oracle-friendly sizes + edge cases."""
    out = []

    def add(tag, name, B, T, **kw):
        w = S.workload(name, B=B, T=T)
        for k, v in kw.items():
            setattr(w, k, v)
        out.append((tag, w))

    add("c1_pendulum_fdvalues", "c1", 2, 16)
    add("c1_pendulum_fdjac", "c1", 2, 16, hess_mode=S.HESS_FD_JACOBIANS)
    add("c2_chain2r", "c2", 5, 10)
    add("c2_chain2r_fdvalues", "c2", 3, 8, hess_mode=S.HESS_FD_VALUES)
    add("c3_arm6r_terminal", "c3", 3, 8)
    add("c3_arm6r_every_knot", "c3", 2, 6, eq_every_knot=True)
    add("c3_arm6r_unconstrained", "c3", 2, 5, eq_kind=S.EQ_NONE)
    add("c4_arm6r_spatial", "c4", 3, 8)
    add("c4_arm6r_spatial_lwa_jac", "c4", 2, 6, jac_mode=S.FJ_LOCAL_WORLD_ALIGNED)
    # value-based FD Hessians subtract eps*J: only meaningful with the true Jacobian (the reference's WORLD-convention
    # quirk would leave an O(1/eps) residue), so this combination uses LOCAL_WORLD_ALIGNED
    add("c4_arm6r_spatial_fdvalues", "c4", 2, 5, hess_mode=S.HESS_FD_VALUES, jac_mode=S.FJ_LOCAL_WORLD_ALIGNED)
    add("c5_arm7r_fdvalues", "c5", 2, 6)
    add("c5_arm7r_fdjac_every_knot", "c5", 2, 4, hess_mode=S.HESS_FD_JACOBIANS, eq_every_knot=True)
    w = S.Workload("chain3r_random", P.models.chain_nr(3), 3, 7)
    out.append(("chain3r_random", w))
    w = S.Workload("chain5r_random_spatial", P.models.chain_nr(5), 2, 6, eq_kind=S.EQ_SPATIAL, frame=5)
    out.append(("chain5r_random_spatial", w))
    w = S.Workload("chain1r", P.models.chain_nr(1), 2, 6)
    out.append(("chain1r", w))
    # prismatic joints (URDF type="prismatic"): every kernel family -- arms (4 and 5 joints, both Hessian modes, both
    # constraint kinds), the register-resident <= 2-joint kernels, a single sliding joint
    m = P.models.scara_rrpr()
    out.append(("scara_rrpr_spatial", S.Workload("scara_rrpr_spatial", m, 3, 8, eq_kind=S.EQ_SPATIAL, frame=m.n_frames - 1)))
    out.append(("scara_rrpr_fdvalues", S.Workload("scara_rrpr_fdvalues", P.models.scara_rrpr(), 2, 6, hess_mode=S.HESS_FD_VALUES)))
    out.append(("chain_prprr_random", S.Workload("chain_prprr_random", P.models.chain_mixed(5, "PRPRR"), 2, 6)))
    out.append(("chain_rp_random", S.Workload("chain_rp_random", P.models.chain_mixed(2, "RP"), 4, 9)))
    out.append(("chain_pr_random_fdvalues", S.Workload("chain_pr_random_fdvalues", P.models.chain_mixed(2, "PR"), 3, 7, hess_mode=S.HESS_FD_VALUES)))
    out.append(("chain_p_random", S.Workload("chain_p_random", P.models.chain_mixed(1, "P"), 2, 6)))
    return out


def prismatic_workload(tag, B, T):
    """Workloads on chains with prismatic joints (models.scara_rrpr, models.chain_mixed) for the external / end-to-end checks."""
    if tag == "scara":
        return S.Workload("scara_rrpr", P.models.scara_rrpr(), B, T)
    if tag == "scara_spatial":
        m = P.models.scara_rrpr()
        return S.Workload("scara_rrpr_spatial", m, B, T, eq_kind=S.EQ_SPATIAL, frame=m.n_frames - 1)
    if tag == "prprr":
        return S.Workload("chain_prprr_random", P.models.chain_mixed(5, "PRPRR"), B, T)
    if tag == "rp":
        return S.Workload("chain_rp_random", P.models.chain_mixed(2, "RP"), B, T)
    raise ValueError(tag)


def setup(make, w, seed=1, random_controls=True):
    g = w.generate()
    bt = make(w)
    bt.set_targets(g["targets"])
    rng = np.random.RandomState(seed)
    us = rng.uniform(-2, 2, (w.B, w.T, w.nu)) if random_controls else g["us"]
    bt.make_trajectory(g["x0"], us)
    orcs = [ol.from_workload(w, g["targets"][b]) for b in range(w.B)]
    return g, bt, us, orcs, rng


def check_linearization(make, w):
    """compute_derivatives vs the oracle on identical trajectories (problem.hpp:956-997)."""
    g, bt, us, orcs, _ = setup(make, w)
    xs, us_back = bt.get_trajectory()
    assert np.array_equal(us_back, us)
    bt.compute_derivatives()
    D = bt.get_derivatives()
    eps = 2.0 ** -26 if w.hess_mode == S.HESS_FD_JACOBIANS else 2.0 ** -13
    for b in range(w.B):
        xo = orcs[b].rollout(g["x0"][b], us[b])
        assert relerr(xs[b], xo) < TOL_ROLLOUT, "make_trajectory"
        d = orcs[b].compute_derivatives(xs[b], us[b])
        for f in FIRST_ORDER:
            assert D[f][b].shape == d[f].shape
            assert relerr(D[f][b], d[f]) < TOL_FIRST, (f, relerr(D[f][b], d[f]))
        # finite-difference tensors: both sides carry rounding noise ~ ulp(J)/eps (mode 1: ulp(f)/eps^2);
        # bound the absolute difference by that floor
        jscale = max(1.0, np.abs(d["fx"]).max(), np.abs(d["fu"]).max(), np.abs(d["f_val"]).max())
        floor = 4096 * U * jscale / (eps if w.hess_mode == S.HESS_FD_JACOBIANS else eps * eps)
        for f in TENSORS:
            if d[f].size:
                assert np.abs(D[f][b] - d[f]).max() <= floor, (f, np.abs(D[f][b] - d[f]).max(), floor)
    bt.close()


def planar_2r_closed_form(m, q, v, tau):
    """Textbook equations of motion of a planar two-link arm (e.g. Spong, Hutchinson, Vidyasagar, "Robot Modeling and
    Control", the planar elbow manipulator), angles measured from the hanging position:
        M11 = m1 lc1^2 + m2 (l1^2 + lc2^2 + 2 l1 lc2 cos q2) + I1 + I2,  M12 = m2 (lc2^2 + l1 lc2 cos q2) + I2,
        M22 = m2 lc2^2 + I2,  h = -m2 l1 lc2 sin q2,
        tau1 = M11 a1 + M12 a2 + h (2 v1 v2 + v2^2) + g1,   tau2 = M12 a1 + M22 a2 - h v1^2 + g2,
        g1 = (m1 lc1 + m2 l1) g sin q1 + m2 lc2 g sin(q1 + q2),  g2 = m2 lc2 g sin(q1 + q2).
    Written from the published formulas only -- no spatial algebra, nothing shared with the oracle or the kernels."""
    m1, m2 = m.mass
    l1 = -m.placement_trans[1][2]
    lc1, lc2 = -m.com[0][2], -m.com[1][2]
    I1, I2 = m.inertia[0][3], m.inertia[1][3]          # yy: about the joint axis direction, at the COM
    g = -m.gravity[2]
    c2, s2 = np.cos(q[1]), np.sin(q[1])
    M = np.array([[m1 * lc1**2 + m2 * (l1**2 + lc2**2 + 2 * l1 * lc2 * c2) + I1 + I2, m2 * (lc2**2 + l1 * lc2 * c2) + I2],
                  [m2 * (lc2**2 + l1 * lc2 * c2) + I2, m2 * lc2**2 + I2]])
    h = -m2 * l1 * lc2 * s2
    cor = np.array([h * (2 * v[0] * v[1] + v[1]**2), -h * v[0]**2])
    grav = np.array([(m1 * lc1 + m2 * l1) * g * np.sin(q[0]) + m2 * lc2 * g * np.sin(q[0] + q[1]), m2 * lc2 * g * np.sin(q[0] + q[1])])
    return np.linalg.solve(M, tau - cor - grav), M


def lagrangian_forward_dynamics(m, q, v, tau, h=1e-5):
    """Forward dynamics of a chain of revolute / prismatic joints from Lagrangian mechanics, written with numpy forward kinematics only:
        M(q) = sum_i m_i Jv_i^T Jv_i + Jw_i^T (R_i I_i R_i^T) Jw_i      (geometric link Jacobians at the COM),
        c_i  = sum_jk (dM_ij/dq_k - 1/2 dM_jk/dq_i) v_j v_k              (Christoffel symbols, dM/dq by central differences),
        g_j  = -sum_i m_i gravity . Jv_i[:, j],      qdd = M^-1 (tau - c - g).
    No spatial algebra, no recursion, nothing shared with the oracle or the kernels; accuracy is that of the
    central differences (~1e-9 relative)."""
    from ddp_pinocchio_b200.models import _rodrigues
    n = m.n_joints
    jt = m.joint_types

    def mass_and_gravity(q):
        R, p = np.eye(3), np.zeros(3)
        ax, org, M, g = [], [], np.zeros((n, n)), np.zeros(n)
        for i in range(n):
            p = p + R @ m.placement_trans[i]
            Rp = R @ m.placement_rot[i]
            ax.append(Rp @ m.axis[i]); org.append(p.copy())
            if jt[i]:    # prismatic: translation along the axis, no rotation
                R = Rp
                p = p + ax[i] * q[i]
            else:
                R = Rp @ _rodrigues(m.axis[i], q[i])
            c = p + R @ m.com[i]
            xx, xy, xz, yy, yz, zz = m.inertia[i]
            Iw = R @ np.array([[xx, xy, xz], [xy, yy, yz], [xz, yz, zz]]) @ R.T
            Jv, Jw = np.zeros((3, n)), np.zeros((3, n))
            for j in range(i + 1):
                if jt[j]:
                    Jv[:, j] = ax[j]
                else:
                    Jv[:, j] = np.cross(ax[j], c - org[j]); Jw[:, j] = ax[j]
            M += m.mass[i] * Jv.T @ Jv + Jw.T @ Iw @ Jw
            g -= m.mass[i] * (m.gravity @ Jv)
        return M, g

    M, g = mass_and_gravity(q)
    dM = np.zeros((n, n, n))
    for k in range(n):
        e = np.zeros(n); e[k] = h
        dM[:, :, k] = (mass_and_gravity(q + e)[0] - mass_and_gravity(q - e)[0]) / (2 * h)
    c = np.einsum("ijk,j,k->i", dM, v, v) - 0.5 * np.einsum("jki,j,k->i", dM, v, v)
    return np.linalg.solve(M, tau - c - g), M


def check_2r_vs_textbook(make, w):
    """f(x,u) and f_u = dt M^-1 of the 2R chain records against the published closed-form dynamics of the planar
    double pendulum (planar_2r_closed_form) -- an external pin that involves neither the oracle
    nor any spatial algebra."""
    g, bt, us, _, _ = setup(make, w)
    xs, _ = bt.get_trajectory()
    bt.compute_derivatives()
    D = bt.get_derivatives(["f_val", "fu"])
    m = w.model
    for b in range(w.B):
        for t in range(w.T):
            f = np.asarray(D["f_val"][b]).reshape(w.T, 4)[t]
            fu = np.asarray(D["fu"][b]).reshape(w.T, 2, 4)[t].T          # 4 x 2, stored column-major
            q, v = xs[b].reshape(w.T + 1, 4)[t, :2], xs[b].reshape(w.T + 1, 4)[t, 2:]
            a, M = planar_2r_closed_form(m, q, v, us[b].reshape(w.T, 2)[t])
            ref = np.concatenate([q + w.dt * v, v + w.dt * a])           # dynamics_t::eval_to, problem.hpp:441-461
            assert relerr(f, ref) < 1e-12, (b, t, relerr(f, ref))
            assert np.abs(fu[:2]).max() == 0.0 and relerr(fu[2:], w.dt * np.linalg.inv(M)) < 1e-12
    bt.close()


def check_vs_lagrangian(make, w, knots=3):
    """f(x,u) and f_u = dt M^-1 of the records of any chain against Lagrangian mechanics (lagrangian_forward_dynamics):
    the external check for the arms, which have no closed form.  Tolerance 1e-9 = the reference side's central
    differences (measured agreement ~4e-12)."""
    g, bt, us, _, _ = setup(make, w)
    xs, _ = bt.get_trajectory()
    bt.compute_derivatives()
    D = bt.get_derivatives(["f_val", "fu"])
    n, nx = w.nu, w.nx
    for b in range(w.B):
        for t in range(min(knots, w.T)):
            f = np.asarray(D["f_val"][b]).reshape(w.T, nx)[t]
            fu = np.asarray(D["fu"][b]).reshape(w.T, n, nx)[t].T
            x = xs[b].reshape(w.T + 1, nx)[t]
            a, M = lagrangian_forward_dynamics(w.model, x[:n], x[n:], us[b].reshape(w.T, n)[t])
            ref = np.concatenate([x[:n] + w.dt * x[n:], x[n:] + w.dt * a])
            assert relerr(f, ref) < 1e-9, (b, t, relerr(f, ref))
            assert np.abs(fu[n:] @ M - w.dt * np.eye(n)).max() < 1e-10 * np.linalg.cond(M) * w.dt
    bt.close()


def check_tensors_vs_quad(make, w):
    """The product's FD tensors are as close to the __float128 oracle as the double oracle's are."""
    g, bt, us, orcs, _ = setup(make, w)
    xs, _ = bt.get_trajectory()
    bt.compute_derivatives()
    D = bt.get_derivatives(TENSORS)
    for b in range(w.B):
        dd = orcs[b].compute_derivatives(xs[b], us[b])
        dq = orcs[b].compute_derivatives(xs[b], us[b], prec=1)
        for f in TENSORS:
            if dd[f].size:
                e_ref = np.abs(dd[f] - dq[f]).max()
                e_dev = np.abs(D[f][b] - dq[f]).max()
                assert e_dev <= 8 * e_ref + 1e-7, (f, e_dev, e_ref)
    bt.close()


def check_backward(make, w):
    """backward_pass on IDENTICAL derivative slabs (the oracle's) -> gains within 1e-9, same mu/reg/restarts."""
    g, bt, us, orcs, rng = setup(make, w)
    xs, _ = bt.get_trajectory()
    mval = rng.uniform(-1, 1, (w.B, bt.ne_total))
    mjac = g["mult_jac"]
    bt.set_multipliers(mval, mjac)
    ds = [orcs[b].compute_derivatives(xs[b], us[b]) for b in range(w.B)]
    bt.set_derivatives({f: np.stack([ds[b][f] for b in range(w.B)]) for f in ol.DERIV_FIELDS})
    D = bt.get_derivatives()
    for b in range(w.B):
        for f in ol.DERIV_FIELDS:
            if f not in ("lfx", "lfxx"):
                assert np.array_equal(D[f][b], ds[b][f]), f"slab round trip {f}"
    restarts_seen = 0
    for mu, reg in [(1e4, 0.0), (1e-3, 0.0), (10.0, 1.0)]:
        bt.set_scalar("status", 0.0)
        mu_e, reg_e = bt.backward_pass(mu, reg)
        _, k, K = bt.get_feedback()
        status = bt.get_scalars()["status"]
        for b in range(w.B):
            ko, Ko, muo, rego, rs = orcs[b].backward_pass(ds[b], xs[b], mval[b], mjac[b], reg, mu)
            restarts_seen += rs
            assert mu_e[b] == muo and reg_e[b] == rego, (mu_e[b], muo, reg_e[b], rego)
            if w.max_restarts and rs > w.max_restarts:
                assert status[b] == 2.0   # guarded mode: diverged, the feedback of the aborted pass is undefined
                continue
            assert status[b] == 0.0
            assert relerr(k[b], ko) < TOL_GAINS and relerr(K[b], Ko) < TOL_GAINS, (relerr(k[b], ko), relerr(K[b], Ko))
    bt.close()
    return restarts_seen


def check_forward(make, w):
    """forward_pass / cost_seq_aug / optimality on identical gains and multipliers."""
    g, bt, us, orcs, rng = setup(make, w)
    xs, _ = bt.get_trajectory()
    mval = rng.uniform(-1, 1, (w.B, bt.ne_total))
    bt.set_multipliers(mval, g["mult_jac"])
    ds = [orcs[b].compute_derivatives(xs[b], us[b]) for b in range(w.B)]
    bt.set_derivatives({f: np.stack([ds[b][f] for b in range(w.B)]) for f in ol.DERIV_FIELDS})
    mu_e, reg_e = bt.backward_pass(w.mu, 0.0)
    bt.set_scalar("mu", mu_e)
    bt.set_scalar("status", 0.0)
    step = bt.forward_pass()
    xn, un = bt.get_new_trajectory()
    _, k, K = bt.get_feedback()
    mo, mv, mj = bt.get_multipliers()
    co = bt.cost_seq_aug()
    oo, oc = bt.optimality()
    for b in range(w.B):
        xo, uo, so, ro = orcs[b].forward_pass(xs[b], us[b], mo[b], mv[b], mj[b], k[b], K[b], mu_e[b])
        assert step[b] == so, (step[b], so)
        assert relerr(xn[b], xo) < TOL_ROLLOUT and relerr(un[b], uo) < TOL_ROLLOUT
        cc = orcs[b].cost_seq_aug(xs[b], us[b], mo[b], mv[b], mj[b], mu_e[b])
        assert relerr(co[b], cc) < TOL_ROLLOUT
        a, c = orcs[b].optimality(ds[b], xs[b], mo[b], mv[b], mj[b], mu_e[b])
        assert abs(oo[b] - a) <= 1e-10 * max(a, 1e-300) and abs(oc[b] - c) <= 1e-12 * max(c, 1.0)
    bt.close()


def check_structure_is_enforced(make, w, exc):
    """The slab keeps the constants of the reference's cost / Euler scheme implicit: derivatives that do not have that
    structure (e.g. a dense random f_x) are refused by set_derivatives instead of being silently truncated."""
    g, bt, us, orcs, rng = setup(make, w)
    xs, _ = bt.get_trajectory()
    ds = [orcs[b].compute_derivatives(xs[b], us[b]) for b in range(w.B)]
    d = {f: np.stack([ds[b][f] for b in range(w.B)]) for f in ol.DERIV_FIELDS}
    bt.set_derivatives(d)                                  # the oracle's derivatives have the structure
    bad = dict(d); bad["fx"] = d["fx"] + 1e-3              # rows k < n of f_x are no longer [I, dt I]
    import pytest
    with pytest.raises(exc):
        bt.set_derivatives(bad)
    bad = dict(d); bad["luu"] = d["luu"] * 2.0             # l_uu != c I
    with pytest.raises(exc):
        bt.set_derivatives(bad)
    bt.close()


def check_forward_many_halvings(make, w, scale=3e3):
    """Line search with gains scaled so that the full step is rejected many times: the accepted step index (and the
    trajectory it produces) equals the oracle's sequential halving, including problems resolved by the second wave of
    concurrent trials and give-ups after 34 trials."""
    g, bt, us, orcs, rng = setup(make, w)
    xs, _ = bt.get_trajectory()
    mval = rng.uniform(-1, 1, (w.B, bt.ne_total))
    bt.set_multipliers(mval, g["mult_jac"])
    ds = [orcs[b].compute_derivatives(xs[b], us[b]) for b in range(w.B)]
    bt.set_derivatives({f: np.stack([ds[b][f] for b in range(w.B)]) for f in ol.DERIV_FIELDS})
    mu_e, _ = bt.backward_pass(w.mu, 0.0)
    _, k, K = bt.get_feedback()
    sc = np.ones(w.B); sc[::2] = scale; sc[1::3] = scale ** 2    # different problems need different numbers of halvings
    bt.set_feedback(k * sc[:, None, None], K)
    bt.set_scalar("mu", mu_e); bt.set_scalar("status", 0.0); bt.set_scalar("rollouts", 0.0)
    step = bt.forward_pass()
    xn, un = bt.get_new_trajectory()
    ro = bt.get_scalars()["rollouts"]
    mo, mv, mj = bt.get_multipliers()
    seen = set()
    for b in range(w.B):
        xo, uo, so, rollouts = orcs[b].forward_pass(xs[b], us[b], mo[b], mv[b], mj[b], k[b] * sc[b], K[b], mu_e[b])
        assert step[b] == so and int(ro[b]) == int(rollouts), (b, step[b], so, ro[b], rollouts)
        assert relerr(xn[b], xo) < TOL_ROLLOUT and relerr(un[b], uo) < TOL_ROLLOUT
        seen.add(int(rollouts))
    bt.close()
    return seen


def check_solve(make, w, iters):
    """End to end: same control flow as the double oracle and trajectories as close to the __float128 oracle
    as the double oracle's own (the FD Hessians make 1e-8 unreachable for ANY two double implementations,
    see DESIGN.md 'Parity')."""
    g = w.generate()
    bt = make(w)
    bt.set_targets(g["targets"])
    bt.make_trajectory(g["x0"], g["us"])
    bt.set_multipliers(None, g["mult_jac"])
    sc = bt.solve(iters, w.mu)
    xs, us = bt.get_trajectory()
    stats = []
    for b in range(w.B):
        o = ol.from_workload(w, g["targets"][b])
        xs0 = o.rollout(g["x0"][b], g["us"][b])
        rd = o.solve(xs0, g["us"][b], g["mult_jac"][b], iters, w.mu)
        rq = o.solve(xs0, g["us"][b], g["mult_jac"][b], iters, w.mu, prec=1)
        e_dev_d = relerr(xs[b], rd["xs"]); e_d_q = relerr(rd["xs"], rq["xs"]); e_dev_q = relerr(xs[b], rq["xs"])
        same_flow = (sc["mu"][b] == rd["log"][-1, 0] and sc["reg"][b] == rd["log"][-1, 1] and
                     int(sc["rollouts"][b]) == int(rd["log"][:, 7].sum()) and int(sc["restarts"][b]) == int(rd["log"][:, 8].sum()))
        stats.append(dict(b=b, dev_vs_double=e_dev_d, double_vs_quad=e_d_q, dev_vs_quad=e_dev_q, same_flow=same_flow))
        assert int(sc["iterations"][b]) == iters
    bt.close()
    return stats


# ---- the Model concept's own functions (capi.Model / hostemu_lib.EmuModel: dynamics_aba, d_dynamics_aba, inverse_dynamics_rnea,
# frame_coordinates, d_frame_coordinates for n samples) ---------------------------------------------------------------------
def model_concept_models():
    return [("pendulum", P.models.pendulum()), ("chain_2r", S.workload("c2", B=1, T=2).model), ("arm_6r", P.models.arm_6r()),
            ("arm_7r", S.workload("c5", B=1, T=2).model), ("scara_rrpr", P.models.scara_rrpr()), ("chain_prprr", P.models.chain_mixed(5, "PRPRR"))]


def check_model_concept(make_model, m, n=7, seed=3):
    """Every function against the oracle's (1e-10 relative; the oracle works in body coordinates with dense 6x6 algebra, the
    kernels in world-frame Pluecker coordinates), inverse dynamics of the forward dynamics gives the torques back, M^-1 is
    symmetric, the arms also against Lagrangian mechanics written with numpy kinematics only, and the frame Jacobian in the
    LOCAL_WORLD_ALIGNED convention against central differences of the frame position."""
    mm = make_model(m)
    nq = mm.nq
    rng = np.random.RandomState(seed)
    q = rng.uniform(-1.5, 1.5, (n, nq)); v = rng.uniform(-1, 1, (n, nq)); tau = rng.uniform(-3, 3, (n, nq))
    o = ol.Oracle(m, 2)
    a = mm.dynamics_aba(q, v, tau)
    dq, dv, dtau = mm.d_dynamics_aba(q, v, tau)
    worst = 0.0
    for s in range(n):
        worst = max(worst, relerr(a[s], o.aba(q[s], v[s], tau[s])))
        r = o.d_aba(q[s], v[s], tau[s])
        for mine, ref in zip((dq[s], dv[s], dtau[s]), r):
            ref = np.asarray(ref).reshape(nq, nq)
            worst = max(worst, float(np.abs(mine - ref).max() / max(1.0, np.abs(ref).max())))
        assert np.abs(dtau[s] - dtau[s].T).max() <= 1e-12 * np.abs(dtau[s]).max()
    assert worst < 1e-10, worst
    if m.kind == 0:
        return worst
    tau_back = mm.inverse_dynamics_rnea(q, v, a)
    assert relerr(tau_back, tau) < 1e-10
    for s in range(n):
        assert relerr(tau_back[s], o.rnea(q[s], v[s], a[s])) < 1e-10
    for s in range(2):
        a_l, M = lagrangian_forward_dynamics(m, q[s], v[s], tau[s])
        assert relerr(a[s], a_l) < 1e-7, (relerr(a[s], a_l))
        assert relerr(dtau[s], np.linalg.inv(M)) < 1e-9
    for f in range(m.n_frames):
        p = mm.frame_coordinates(f, q)
        for mode in (0, 1):
            J = mm.d_frame_coordinates(f, q, mode)
            for s in range(n):
                po, Jo = o.frame(f, q[s], mode)
                assert np.abs(p[s] - po).max() < 1e-12 * max(1.0, np.abs(po).max())
                assert np.abs(J[s] - np.asarray(Jo).reshape(3, nq)).max() < 1e-12 * max(1.0, np.abs(Jo).max())
        h = 1e-6
        J1 = mm.d_frame_coordinates(f, q[:1], 1)[0]
        for c in range(nq):
            e = np.zeros(nq); e[c] = h
            fd = (mm.frame_coordinates(f, q[:1] + e)[0] - mm.frame_coordinates(f, q[:1] - e)[0]) / (2 * h)
            assert np.abs(fd - J1[:, c]).max() < 1e-8
    return worst
