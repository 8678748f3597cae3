"""ctypes wrapper over oracle/liboracle.so (the CPU oracle; test infrastructure only)."""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)

DERIV_FIELDS = ["lfx", "lfxx", "lx", "lu", "lxx", "lux", "luu", "f_val", "fx", "fu", "fxx", "fux", "fuu",
                "eq_val", "eq_x", "eq_u", "eq_xx", "eq_ux", "eq_uu"]


class ModelDescC(C.Structure):
    _fields_ = [("kind", C.c_int), ("n_joints", C.c_int), ("parent", _ip), ("placement_rot", _dp), ("placement_trans", _dp),
                ("axis", _dp), ("mass", _dp), ("com", _dp), ("inertia", _dp), ("gravity", C.c_double * 3), ("n_frames", C.c_int),
                ("frame_parent", _ip), ("frame_rot", _dp), ("frame_trans", _dp), ("pend_mass", C.c_double), ("pend_length", C.c_double),
                ("joint_type", _ip)]


class ProblemDescC(C.Structure):
    _fields_ = [("T", C.c_int), ("dt", C.c_double), ("c", C.c_double), ("second_order_finite_diff", C.c_int), ("eq_kind", C.c_int),
                ("frame", C.c_int), ("jac_mode", C.c_int), ("ne_per_knot", _ip), ("targets", _dp), ("max_restarts", C.c_int)]


class DerivsC(C.Structure):
    _fields_ = [(f, _dp) for f in DERIV_FIELDS]


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build()
        _lib = C.CDLL(path)
        _lib.orc_create.restype = C.c_void_p
        _lib.orc_flops_first_order.restype = C.c_uint64
        _lib.orc_flops_eval.restype = C.c_uint64
        _lib.orc_solve_batch.restype = C.c_double
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def model_desc_c(m, keep):
    """Fill a C model descriptor from ddp_pinocchio_b200.models.ModelDesc; `keep` holds the arrays alive."""
    md = ModelDescC()
    md.kind = m.kind
    md.n_joints = m.n_joints
    md.pend_mass = m.pend_mass
    md.pend_length = m.pend_length
    for k in range(3):
        md.gravity[k] = float(m.gravity[k])
    if m.kind != 0:
        arrs = dict(parent=np.ascontiguousarray(m.parent, dtype=np.int32),
                    placement_rot=np.ascontiguousarray(m.placement_rot, dtype=np.float64),
                    placement_trans=np.ascontiguousarray(m.placement_trans, dtype=np.float64),
                    axis=np.ascontiguousarray(m.axis, dtype=np.float64), mass=np.ascontiguousarray(m.mass, dtype=np.float64),
                    com=np.ascontiguousarray(m.com, dtype=np.float64), inertia=np.ascontiguousarray(m.inertia, dtype=np.float64),
                    frame_parent=np.ascontiguousarray(m.frame_parent, dtype=np.int32),
                    frame_rot=np.ascontiguousarray(m.frame_rot, dtype=np.float64),
                    frame_trans=np.ascontiguousarray(m.frame_trans, dtype=np.float64),
                    joint_type=np.ascontiguousarray(m.joint_types, dtype=np.int32))
        keep.append(arrs)
        md.parent = _i(arrs["parent"]); md.placement_rot = _d(arrs["placement_rot"]); md.placement_trans = _d(arrs["placement_trans"])
        md.axis = _d(arrs["axis"]); md.mass = _d(arrs["mass"]); md.com = _d(arrs["com"]); md.inertia = _d(arrs["inertia"])
        md.n_frames = m.n_frames
        md.frame_parent = _i(arrs["frame_parent"]); md.frame_rot = _d(arrs["frame_rot"]); md.frame_trans = _d(arrs["frame_trans"])
        md.joint_type = _i(arrs["joint_type"])
    return md


class Oracle:
    """One single-problem instance of the reference path (model + horizon + constraint)."""

    def __init__(self, model, T, dt=0.01, c=1.0, hess_mode=0, eq_kind=0, frame=0, jac_mode=0, ne_per_knot=None, targets=None,
                 max_restarts=0):
        self._keep = []
        self.model = model
        md = model_desc_c(model, self._keep)
        ne = np.zeros(T, dtype=np.int32) if ne_per_knot is None else np.ascontiguousarray(ne_per_knot, dtype=np.int32)
        tg = np.zeros(max(int(ne.sum()), 1)) if targets is None else np.ascontiguousarray(targets, dtype=np.float64).ravel()
        self._keep += [ne, tg]
        pd = ProblemDescC(T, dt, c, 1 if hess_mode else 0, eq_kind, frame, jac_mode, _i(ne), _d(tg), max_restarts)
        self._md, self._pd = md, pd
        self.h = C.c_void_p(lib().orc_create(C.byref(md), C.byref(pd)))
        dims = np.zeros(6, dtype=np.int32)
        lib().orc_dims(self.h, _i(dims))
        self.nq, self.nx, self.ndx, self.nu, self.ne_total, self.T = [int(v) for v in dims]
        self.ne = ne
        self.eq_off = np.concatenate([[0], np.cumsum(ne)]).astype(int)

    def __del__(self):
        try:
            lib().orc_destroy(self.h)
        except Exception:
            pass

    # ---- sizes of the packed derivative sequences (ddp.hpp:84-105) ----
    def deriv_sizes(self):
        T, nd, nu, nx, ne = self.T, self.ndx, self.nu, self.nx, self.ne_total
        return dict(lfx=nd, lfxx=nd * nd, lx=T * nd, lu=T * nu, lxx=T * nd * nd, lux=T * nu * nd, luu=T * nu * nu, f_val=T * nx,
                    fx=T * nd * nd, fu=T * nd * nu, fxx=T * nd ** 3, fux=T * nd * nu * nd, fuu=T * nd * nu * nu, eq_val=ne,
                    eq_x=ne * nd, eq_u=ne * nu, eq_xx=ne * nd * nd, eq_ux=ne * nu * nd, eq_uu=ne * nu * nu)

    def _derivs_c(self, d):
        dc = DerivsC()
        for f in DERIV_FIELDS:
            setattr(dc, f, _d(d[f]))
        return dc

    def aba(self, q, v, tau, prec=0):
        a = np.zeros(self.nq)
        lib().orc_aba(self.h, prec, _d(np.ascontiguousarray(q, dtype=float)), _d(np.ascontiguousarray(v, dtype=float)),
                      _d(np.ascontiguousarray(tau, dtype=float)), _d(a))
        return a

    def rnea(self, q, v, a, prec=0):
        tau = np.zeros(self.nq)
        lib().orc_rnea(self.h, prec, _d(np.ascontiguousarray(q, dtype=float)), _d(np.ascontiguousarray(v, dtype=float)),
                       _d(np.ascontiguousarray(a, dtype=float)), _d(tau))
        return tau

    def d_aba(self, q, v, tau, prec=0):
        n = self.nq
        dq, dv, dt = np.zeros(n * n), np.zeros(n * n), np.zeros(n * n)
        lib().orc_d_aba(self.h, prec, _d(np.ascontiguousarray(q, dtype=float)), _d(np.ascontiguousarray(v, dtype=float)),
                        _d(np.ascontiguousarray(tau, dtype=float)), _d(dq), _d(dv), _d(dt))
        return [m.reshape(n, n).T.copy() for m in (dq, dv, dt)]  # col-major -> [row, col]

    def frame(self, frame, q, mode=0, prec=0):
        p, J = np.zeros(3), np.zeros(3 * self.nq)
        lib().orc_frame(self.h, prec, frame, mode, _d(np.ascontiguousarray(q, dtype=float)), _d(p), _d(J))
        return p, J.reshape(self.nq, 3).T.copy()

    def first_order(self, x, u, prec=0):
        f, fx, fu = np.zeros(self.nx), np.zeros(self.ndx * self.ndx), np.zeros(self.ndx * self.nu)
        lib().orc_first_order(self.h, prec, _d(np.ascontiguousarray(x, dtype=float)), _d(np.ascontiguousarray(u, dtype=float)), _d(f), _d(fx), _d(fu))
        return f, fx.reshape(self.ndx, self.ndx).T.copy(), fu.reshape(self.nu, self.ndx).T.copy()

    def flops_first_order(self, x, u):
        return int(lib().orc_flops_first_order(self.h, _d(np.ascontiguousarray(x, dtype=float)), _d(np.ascontiguousarray(u, dtype=float))))

    def flops_eval(self, x, u):
        return int(lib().orc_flops_eval(self.h, _d(np.ascontiguousarray(x, dtype=float)), _d(np.ascontiguousarray(u, dtype=float))))

    def rollout(self, x0, us, prec=0):
        xs = np.zeros((self.T + 1) * self.nx)
        us = np.ascontiguousarray(us, dtype=float).ravel()
        lib().orc_rollout(self.h, prec, _d(np.ascontiguousarray(x0, dtype=float)), _d(us), _d(xs))
        return xs.reshape(self.T + 1, self.nx)

    def eval_eq(self, t, x, u, prec=0):
        o = np.zeros(max(int(self.ne[t]), 1))
        lib().orc_eval_eq(self.h, prec, t, _d(np.ascontiguousarray(x, dtype=float)), _d(np.ascontiguousarray(u, dtype=float)), _d(o))
        return o[: int(self.ne[t])]

    def compute_derivatives(self, xs, us, prec=0):
        d = {f: np.zeros(max(n, 1)) for f, n in self.deriv_sizes().items()}
        xs = np.ascontiguousarray(xs, dtype=float).ravel(); us = np.ascontiguousarray(us, dtype=float).ravel()
        dc = self._derivs_c(d)
        lib().orc_compute_derivatives(self.h, prec, _d(xs), _d(us), C.byref(dc))
        return {f: d[f][:n] for f, n in self.deriv_sizes().items()}

    def _pad(self, a):
        a = np.ascontiguousarray(a, dtype=float).ravel()
        return a if a.size else np.zeros(1)

    def backward_pass(self, derivs, xs, mult_val, mult_jac, reg, mu, prec=0):
        d = {f: self._pad(derivs[f]) for f in DERIV_FIELDS}
        dc = self._derivs_c(d)
        k = np.zeros(self.T * self.nu); K = np.zeros(self.T * self.nu * self.ndx)
        mu_o, reg_o, rs = C.c_double(), C.c_double(), C.c_int()
        xs = np.ascontiguousarray(xs, dtype=float).ravel()
        mv, mj = self._pad(mult_val), self._pad(mult_jac)
        lib().orc_backward_pass(self.h, prec, C.byref(dc), _d(xs), _d(mv), _d(mj), C.c_double(reg), C.c_double(mu), _d(k), _d(K),
                                C.byref(mu_o), C.byref(reg_o), C.byref(rs))
        return k.reshape(self.T, self.nu), K.reshape(self.T, self.ndx, self.nu), mu_o.value, reg_o.value, rs.value

    def cost_seq_aug(self, xs, us, m_origin, m_val, m_jac, mu, prec=0):
        c = np.zeros(self.T + 1)
        lib().orc_cost_seq_aug(self.h, prec, _d(self._pad(xs)), _d(self._pad(us)), _d(self._pad(m_origin)), _d(self._pad(m_val)),
                               _d(self._pad(m_jac)), C.c_double(mu), _d(c))
        return c

    def forward_pass(self, xs, us, m_origin, m_val, m_jac, k, K, mu, xs_new0=None, us_new0=None, prec=0):
        xs = self._pad(xs); us = self._pad(us)
        xn = xs.copy() if xs_new0 is None else self._pad(xs_new0).copy()
        un = us.copy() if us_new0 is None else self._pad(us_new0).copy()
        step, ro = C.c_double(), C.c_int()
        lib().orc_forward_pass(self.h, prec, _d(xs), _d(us), _d(self._pad(m_origin)), _d(self._pad(m_val)), _d(self._pad(m_jac)),
                               _d(self._pad(k)), _d(self._pad(K)), C.c_double(mu), _d(xn), _d(un), C.byref(step), C.byref(ro))
        return xn.reshape(self.T + 1, self.nx), un.reshape(self.T, self.nu), step.value, ro.value

    def optimality(self, derivs, xs, m_origin, m_val, m_jac, mu, prec=0):
        d = {f: self._pad(derivs[f]) for f in DERIV_FIELDS}
        dc = self._derivs_c(d)
        a, b = C.c_double(), C.c_double()
        lib().orc_optimality(self.h, prec, C.byref(dc), _d(self._pad(xs)), _d(self._pad(m_origin)), _d(self._pad(m_val)),
                             _d(self._pad(m_jac)), C.c_double(mu), C.byref(a), C.byref(b))
        return a.value, b.value

    def solve(self, xs0, us0, mult_jac0, max_iter, mu, reg=0.0, w=None, n=None, thr=0.0, prec=0):
        w = 1.0 / mu if w is None else w
        n = 1.0 / mu ** 0.1 if n is None else n
        T, nx, nu, nd = self.T, self.nx, self.nu, self.ndx
        xs = np.zeros((T + 1) * nx); us = np.zeros(T * nu); k = np.zeros(T * nu); K = np.zeros(T * nu * nd)
        mv = np.zeros(max(self.ne_total, 1)); mj = np.zeros(max(self.ne_total * nd, 1)); log = np.zeros((max(max_iter, 1), 10))
        it, st = C.c_int(), C.c_int()
        lib().orc_solve(self.h, prec, max_iter, C.c_double(thr), C.c_double(mu), C.c_double(reg), C.c_double(w), C.c_double(n),
                        _d(self._pad(xs0)), _d(self._pad(us0)), _d(self._pad(mult_jac0)), _d(xs), _d(us), _d(k), _d(K), _d(mv), _d(mj),
                        _d(log), C.byref(it), C.byref(st))
        return dict(xs=xs.reshape(T + 1, nx), us=us.reshape(T, nu), k=k.reshape(T, nu), K=K.reshape(T, nd, nu),
                    mult_val=mv[: self.ne_total], mult_jac=mj[: self.ne_total * nd], log=log[: it.value + (1 if st.value == 1 else 0)],
                    iterations=it.value, status=st.value)


def from_workload(w, targets, T=None):
    """Oracle instance for one problem of a synthetic Workload (ddp_pinocchio_b200.synthetic)."""
    return Oracle(w.model, w.T if T is None else T, dt=w.dt, c=w.c, hess_mode=w.hess_mode, eq_kind=w.eq_kind, frame=w.frame,
                  jac_mode=w.jac_mode, ne_per_knot=w.ne_per_knot(), targets=targets, max_restarts=w.max_restarts)


def solve_batch(w, data, nprob, nthreads, max_iter, mu, reg=0.0, thr=0.0, want_traj=False):
    """OpenMP batch of independent single-problem solves (the CPU baseline of bench.py). Returns (seconds, info)."""
    keep = []
    md = model_desc_c(w.model, keep)
    ne = np.ascontiguousarray(w.ne_per_knot(), dtype=np.int32)
    pd = ProblemDescC(w.T, w.dt, w.c, 1 if w.hess_mode else 0, w.eq_kind, w.frame, w.jac_mode, _i(ne), None, w.max_restarts)
    nx, nu, T = w.nx, w.nu, w.T
    x0 = data["x0"][:nprob]
    # zero-control initial trajectory per problem, rolled out by the oracle itself (make_trajectory)
    xs0 = np.zeros((nprob, T + 1, nx)); us0 = np.ascontiguousarray(data["us"][:nprob])
    o = Oracle(w.model, T, dt=w.dt, c=w.c, hess_mode=w.hess_mode)
    for b in range(nprob):
        xs0[b] = o.rollout(x0[b], us0[b])
    tg = np.ascontiguousarray(data["targets"][:nprob]); mj = np.ascontiguousarray(data["mult_jac"][:nprob])
    if tg.size == 0:
        tg = np.zeros(1); mj = np.zeros(1)
    xs_out = np.zeros_like(xs0) if want_traj else None
    us_out = np.zeros_like(us0) if want_traj else None
    ro, rs, dv, it = C.c_int(), C.c_int(), C.c_int(), C.c_longlong()
    sec = lib().orc_solve_batch(C.byref(md), C.byref(pd), nprob, nthreads, max_iter, C.c_double(thr), C.c_double(mu), C.c_double(reg),
                                C.c_double(1.0 / mu), C.c_double(1.0 / mu ** 0.1), _d(xs0), _d(us0), _d(tg), _d(mj),
                                _d(xs_out) if want_traj else None, _d(us_out) if want_traj else None, C.byref(ro), C.byref(rs), C.byref(it), C.byref(dv))
    return sec, dict(rollouts=ro.value, restarts=rs.value, iterations=it.value, diverged=dv.value, xs=xs_out, us=us_out)
