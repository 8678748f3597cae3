"""ctypes binding of the C ABI (include/ddpb.h -> libddpb.so).  Host-side plumbing only: every method
is one C-ABI call; there is no Python or CPU implementation of the path behind it, and a missing
library or a missing GPU raises."""
import ctypes as C
import os
import numpy as np

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DDPB_LIB", os.path.join(PKG_DIR, "libddpb.so"))  # DDPB_LIB: development override

DERIV_FIELDS = ["lfx", "lfxx", "lx", "lu", "lxx", "lux", "luu", "f_val", "fx", "fu", "fxx", "fux", "fuu",
                "eq_val", "eq_x", "eq_u", "eq_xx", "eq_ux", "eq_uu"]
SCALAR_FIELDS = ["mu", "reg", "w", "n", "step", "opt_obj", "opt_constr", "cost", "merit", "iterations", "status", "rollouts",
                 "restarts", "mult_updates"]
NSCALARS = 16
EXPORTS = ["ddpb_model_create", "ddpb_model_destroy", "ddpb_batch_create", "ddpb_batch_destroy", "ddpb_batch_set_stream",
           "ddpb_batch_sync", "ddpb_last_error", "ddpb_host_alloc", "ddpb_host_free", "ddpb_set_trajectory", "ddpb_get_trajectory",
           "ddpb_get_new_trajectory", "ddpb_make_trajectory", "ddpb_set_targets", "ddpb_set_multipliers", "ddpb_get_multipliers",
           "ddpb_set_feedback", "ddpb_get_feedback", "ddpb_set_derivatives", "ddpb_get_derivatives", "ddpb_set_scalar",
           "ddpb_get_scalars", "ddpb_scalars_device_ptr", "ddpb_compute_derivatives", "ddpb_backward_pass", "ddpb_forward_pass",
           "ddpb_cost_seq_aug", "ddpb_optimality", "ddpb_solve", "ddpb_solve_async", "ddpb_batch_dims", "ddpb_launch_count",
           "ddpb_phase_times", "ddpb_enable_phase_timing", "ddpb_snapshot_save", "ddpb_snapshot_restore",
           "ddpb_measure_fp64_peak", "ddpb_set_subbatches", "ddpb_model_create_from_urdf", "ddpb_model_dims", "ddpb_model_get",
           "ddpb_model_frame_name", "ddpb_model_joint_types", "ddpb_model_dynamics_aba", "ddpb_model_d_dynamics_aba",
           "ddpb_model_inverse_dynamics_rnea", "ddpb_model_frame_coordinates", "ddpb_model_d_frame_coordinates"]


class DdpbError(RuntimeError):
    pass


class ModelDescC(C.Structure):
    _fields_ = [("kind", C.c_int), ("n_joints", C.c_int), ("parent", _ip), ("placement_rot", _dp), ("placement_trans", _dp),
                ("axis", _dp), ("mass", _dp), ("com", _dp), ("inertia", _dp), ("gravity", C.c_double * 3), ("n_frames", C.c_int),
                ("frame_parent", _ip), ("frame_rot", _dp), ("frame_trans", _dp), ("pend_mass", C.c_double), ("pend_length", C.c_double),
                ("joint_type", _ip)]


class ParamsC(C.Structure):
    _fields_ = [("dt", C.c_double), ("c", C.c_double), ("hess_mode", C.c_int), ("eq_kind", C.c_int), ("frame_id", C.c_int),
                ("frame_jac_mode", C.c_int), ("max_backward_restarts", C.c_int)]


class SolverParamsC(C.Structure):
    _fields_ = [("max_iterations", C.c_int), ("optimality_stopping_threshold", C.c_double), ("mu", C.c_double), ("reg", C.c_double),
                ("w", C.c_double), ("n", C.c_double)]


class DerivsC(C.Structure):
    _fields_ = [(f, _dp) for f in DERIV_FIELDS]


_lib = None


def lib():
    """Load libddpb.so (built in-tree by __graft_entry__.build()); raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DdpbError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (no fallback exists)")
        _lib = C.CDLL(LIB_PATH)
        _lib.ddpb_last_error.restype = C.c_char_p
        _lib.ddpb_launch_count.restype = C.c_ulonglong
    return _lib


def _check(rc):
    if rc != 0:
        raise DdpbError(f"ddpb error {rc}: {lib().ddpb_last_error().decode()}")


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def model_desc_c(m, keep):
    md = ModelDescC()
    md.kind = m.kind
    md.n_joints = m.n_joints
    md.pend_mass = m.pend_mass
    md.pend_length = m.pend_length
    for k in range(3):
        md.gravity[k] = float(m.gravity[k])
    if m.kind != 0:
        arrs = dict(parent=np.ascontiguousarray(m.parent, dtype=np.int32), placement_rot=_c(m.placement_rot),
                    placement_trans=_c(m.placement_trans), axis=_c(m.axis), mass=_c(m.mass), com=_c(m.com), inertia=_c(m.inertia),
                    frame_parent=np.ascontiguousarray(m.frame_parent, dtype=np.int32), frame_rot=_c(m.frame_rot),
                    frame_trans=_c(m.frame_trans), joint_type=np.ascontiguousarray(m.joint_types, dtype=np.int32))
        keep.append(arrs)
        md.parent = _i(arrs["parent"]); md.placement_rot = _d(arrs["placement_rot"]); md.placement_trans = _d(arrs["placement_trans"])
        md.axis = _d(arrs["axis"]); md.mass = _d(arrs["mass"]); md.com = _d(arrs["com"]); md.inertia = _d(arrs["inertia"])
        md.n_frames = m.n_frames
        md.frame_parent = _i(arrs["frame_parent"]); md.frame_rot = _d(arrs["frame_rot"]); md.frame_trans = _d(arrs["frame_trans"])
        md.joint_type = _i(arrs["joint_type"])
    return md


def model_from_urdf(text):
    """URDF text -> models.ModelDesc through the library's own converter (ddpb_model_create_from_urdf: the replacement of
    pinocchio::urdf::buildModel, pinocchio_model.ipp:98-113); the arrays come back from the C ABI, nothing is parsed here."""
    from . import models
    L = lib()
    L.ddpb_model_frame_name.restype = C.c_char_p
    data = text.encode() if isinstance(text, str) else bytes(text)
    h = C.c_void_p()
    _check(L.ddpb_model_create_from_urdf(data, C.c_ulonglong(len(data)), C.byref(h)))
    try:
        dims = (C.c_int * 2)()
        _check(L.ddpb_model_dims(h, dims))
        n, f = dims[0], dims[1]
        R = np.zeros((n, 3, 3)); p = np.zeros((n, 3)); ax = np.zeros((n, 3)); m = np.zeros(n); c = np.zeros((n, 3)); I = np.zeros((n, 6))
        fpar = np.zeros(f, dtype=np.int32); fp = np.zeros((f, 3))
        _check(L.ddpb_model_get(h, _d(R), _d(p), _d(ax), _d(m), _d(c), _d(I), _i(fpar), _d(fp)))
        names = [L.ddpb_model_frame_name(h, i).decode() for i in range(f)]
        jt = np.zeros(n, dtype=np.int32)
        _check(L.ddpb_model_joint_types(h, _i(jt)))
    finally:
        L.ddpb_model_destroy(h)
    return models.ModelDesc(name="urdf", kind=models.KIND_REVOLUTE_CHAIN, n_joints=n, parent=np.arange(-1, n - 1, dtype=np.int32),
                            placement_rot=R, placement_trans=p, axis=ax, mass=m, com=c, inertia=I, frame_parent=fpar,
                            frame_rot=np.tile(np.eye(3), (f, 1, 1)), frame_trans=fp, frame_names=names,
                            joint_type=jt if jt.any() else None)


class BatchBase:
    """Shape bookkeeping shared by the GPU binding and the CPU test emulator (tests/hostemu_lib.py)."""

    def _init_dims(self, model, B, T, ne_per_knot, eq_kind):
        n = model.nq
        self.B, self.T, self.n, self.nx, self.ndx, self.nu = B, T, n, 2 * n, 2 * n, n
        self.ne = np.zeros(T, dtype=np.int32) if (ne_per_knot is None or eq_kind == 0) else np.ascontiguousarray(ne_per_knot, dtype=np.int32)
        self.ne_total = int(self.ne.sum())
        self.eq_off = np.concatenate([[0], np.cumsum(self.ne)]).astype(int)

    def deriv_sizes(self):
        T, nd, nu, nx, ne = self.T, self.ndx, self.nu, self.nx, self.ne_total
        return dict(lfx=nd, lfxx=nd * nd, lx=T * nd, lu=T * nu, lxx=T * nd * nd, lux=T * nu * nd, luu=T * nu * nu, f_val=T * nx,
                    fx=T * nd * nd, fu=T * nd * nu, fxx=T * nd ** 3, fux=T * nd * nu * nd, fuu=T * nd * nu * nu, eq_val=ne,
                    eq_x=ne * nd, eq_u=ne * nu, eq_xx=ne * nd * nd, eq_ux=ne * nu * nd, eq_uu=ne * nu * nu)


class ModelBase:
    """The Model concept's own functions for n samples (pinocchio_model.hpp:147-185): shapes and argument checks shared by
    the C-ABI binding below and the CPU emulation of the tests; subclasses provide _eval(**arrays)."""

    def _init_model(self, model):
        self.model = model
        self.nq = 1 if model.kind == 0 else model.n_joints

    def _arr(self, a):
        a = np.ascontiguousarray(np.atleast_2d(np.asarray(a, dtype=np.float64)))
        assert a.shape[1] == self.nq, "samples are rows of nq entries"
        return a

    def dynamics_aba(self, q, v, tau):
        q, v, tau = self._arr(q), self._arr(v), self._arr(tau)
        a = np.zeros_like(q)
        self._eval(n=len(q), q=q, v=v, tau=tau, a=a)
        return a

    def d_dynamics_aba(self, q, v, tau):
        """(da/dq, da/dv, da/dtau), each [n][nv][nv] with out[s][r][c] = d a_r / d x_c."""
        q, v, tau = self._arr(q), self._arr(v), self._arr(tau)
        n, nq = q.shape
        o = [np.zeros((n, nq * nq)) for _ in range(3)]
        self._eval(n=n, q=q, v=v, tau=tau, a_dq=o[0], a_dv=o[1], a_dtau=o[2])
        return tuple(x.reshape(n, nq, nq).transpose(0, 2, 1).copy() for x in o)  # column-major per sample -> [r][c]

    def inverse_dynamics_rnea(self, q, v, a):
        q, v, a = self._arr(q), self._arr(v), self._arr(a)
        tau = np.zeros_like(q)
        self._eval(n=len(q), q=q, v=v, acc=a, tau_out=tau)
        return tau

    def frame_coordinates(self, frame, q):
        q = self._arr(q)
        p = np.zeros((len(q), 3))
        self._eval(n=len(q), frame=frame, q=q, p3=p)
        return p

    def d_frame_coordinates(self, frame, q, jac_mode=0):
        """[n][3][nv]"""
        q = self._arr(q)
        n, nq = q.shape
        J = np.zeros((n, 3 * nq))
        self._eval(n=n, frame=frame, jac_mode=jac_mode, q=q, J=J)
        return J.reshape(n, nq, 3).transpose(0, 2, 1).copy()


class Model(ModelBase):
    """ddpb_model + the batched Model-concept entry points of include/ddpb.h (device evaluation, host arrays)."""

    def __init__(self, model, device=0):
        self._keep = []
        self._init_model(model)
        self.device = device
        md = model_desc_c(model, self._keep)
        self._m = C.c_void_p()
        _check(lib().ddpb_model_create(C.byref(md), C.byref(self._m)))

    def close(self):
        if getattr(self, "_m", None):
            lib().ddpb_model_destroy(self._m)
            self._m = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _eval(self, n, q, v=None, tau=None, acc=None, a=None, a_dq=None, a_dv=None, a_dtau=None, tau_out=None, p3=None, J=None, frame=0,
              jac_mode=0):
        L = lib()
        p = lambda x: _d(x) if x is not None else None
        if a is not None:
            _check(L.ddpb_model_dynamics_aba(self._m, self.device, n, p(q), p(v), p(tau), p(a)))
        if a_dq is not None or a_dv is not None or a_dtau is not None:
            _check(L.ddpb_model_d_dynamics_aba(self._m, self.device, n, p(q), p(v), p(tau), p(a_dq), p(a_dv), p(a_dtau)))
        if tau_out is not None:
            _check(L.ddpb_model_inverse_dynamics_rnea(self._m, self.device, n, p(q), p(v), p(acc), p(tau_out)))
        if p3 is not None:
            _check(L.ddpb_model_frame_coordinates(self._m, self.device, n, frame, p(q), p(p3)))
        if J is not None:
            _check(L.ddpb_model_d_frame_coordinates(self._m, self.device, n, frame, jac_mode, p(q), p(J)))


class Batch(BatchBase):
    """B independent DDP problems resident on one GPU (ddpb_batch)."""

    def __init__(self, model, B, T, ne_per_knot=None, dt=0.01, c=1.0, hess_mode=0, eq_kind=0, frame_id=0, jac_mode=0, device=0,
                 max_restarts=0):
        L = lib()
        self._keep = []
        self._init_dims(model, B, T, ne_per_knot, eq_kind)
        md = model_desc_c(model, self._keep)
        self._m = C.c_void_p()
        _check(L.ddpb_model_create(C.byref(md), C.byref(self._m)))
        pr = ParamsC(dt, c, hess_mode, eq_kind, frame_id, jac_mode, max_restarts)
        self._h = C.c_void_p()
        rc = L.ddpb_batch_create(self._m, B, T, _i(self.ne), C.byref(pr), device, C.byref(self._h))
        if rc != 0:
            L.ddpb_model_destroy(self._m)
            self._m = None
            _check(rc)

    @classmethod
    def from_workload(cls, w, B=None, device=0, max_restarts=None):
        return cls(w.model, w.B if B is None else B, w.T, w.ne_per_knot(), dt=w.dt, c=w.c, hess_mode=w.hess_mode, eq_kind=w.eq_kind,
                   frame_id=w.frame, jac_mode=w.jac_mode, device=device, max_restarts=w.max_restarts if max_restarts is None else max_restarts)

    def close(self):
        if getattr(self, "_h", None):
            lib().ddpb_batch_destroy(self._h)
            self._h = None
        if getattr(self, "_m", None):
            lib().ddpb_model_destroy(self._m)
            self._m = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- data ----
    def set_stream(self, cuda_stream_ptr):
        _check(lib().ddpb_batch_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def sync(self):
        _check(lib().ddpb_batch_sync(self._h))

    def set_trajectory(self, xs, us):
        xs, us = _c(xs), _c(us)
        assert xs.size == self.B * (self.T + 1) * self.nx and us.size == self.B * self.T * self.nu
        _check(lib().ddpb_set_trajectory(self._h, _d(xs), _d(us)))

    def get_trajectory(self, out=None):
        """out = (xs, us): caller-owned destination arrays (e.g. pinned, capi.host_alloc) instead of fresh pageable ones."""
        if out is None:
            xs = np.empty((self.B, self.T + 1, self.nx)); us = np.empty((self.B, self.T, self.nu))
        else:
            xs, us = out
            assert xs.size == self.B * (self.T + 1) * self.nx and us.size == self.B * self.T * self.nu and xs.flags.c_contiguous and us.flags.c_contiguous
        _check(lib().ddpb_get_trajectory(self._h, _d(xs), _d(us)))
        return xs, us

    def get_new_trajectory(self):
        xs = np.empty((self.B, self.T + 1, self.nx)); us = np.empty((self.B, self.T, self.nu))
        _check(lib().ddpb_get_new_trajectory(self._h, _d(xs), _d(us)))
        return xs, us

    def make_trajectory(self, x0, us=None):
        x0 = _c(x0)
        assert x0.size == self.B * self.nx
        if us is not None:
            us = _c(us)
            assert us.size == self.B * self.T * self.nu
        _check(lib().ddpb_make_trajectory(self._h, _d(x0), _d(us) if us is not None else None))

    def set_targets(self, targets):
        if self.ne_total == 0:
            return
        t = _c(targets)
        assert t.size == self.B * self.ne_total
        _check(lib().ddpb_set_targets(self._h, _d(t)))

    def set_multipliers(self, val=None, jac=None):
        v = None if val is None or self.ne_total == 0 else _c(val)
        j = None if jac is None or self.ne_total == 0 else _c(jac)
        _check(lib().ddpb_set_multipliers(self._h, _d(v) if v is not None else None, _d(j) if j is not None else None))

    def get_multipliers(self):
        o = np.empty((self.B, self.T, self.nx)); v = np.zeros((self.B, self.ne_total)); j = np.zeros((self.B, self.ne_total * self.ndx))
        _check(lib().ddpb_get_multipliers(self._h, _d(o), _d(v) if self.ne_total else None, _d(j) if self.ne_total else None))
        return o, v, j

    def set_feedback(self, k, K):
        k, K = _c(k), _c(K)
        _check(lib().ddpb_set_feedback(self._h, _d(k), _d(K)))

    def get_feedback(self):
        o = np.empty((self.B, self.T, self.nx)); k = np.empty((self.B, self.T, self.nu)); K = np.empty((self.B, self.T, self.ndx, self.nu))
        _check(lib().ddpb_get_feedback(self._h, _d(o), _d(k), _d(K)))
        return o, k, K

    def _derivs_c(self, d):
        dc = DerivsC()
        for f in DERIV_FIELDS:
            a = d.get(f)
            setattr(dc, f, _d(a) if a is not None and a.size else None)
        return dc

    def get_derivatives(self, fields=None):
        sizes = self.deriv_sizes()
        d = {f: np.zeros((self.B, sizes[f])) for f in (fields or DERIV_FIELDS)}
        dc = self._derivs_c(d)
        _check(lib().ddpb_get_derivatives(self._h, C.byref(dc)))
        return d

    def set_derivatives(self, d):
        d = {f: _c(a) for f, a in d.items()}
        dc = self._derivs_c(d)
        self._keep_d = d
        _check(lib().ddpb_set_derivatives(self._h, C.byref(dc)))

    def set_scalar(self, name, values):
        v = _c(np.broadcast_to(values, (self.B,)))
        _check(lib().ddpb_set_scalar(self._h, SCALAR_FIELDS.index(name), _d(v)))

    def get_scalars(self):
        out = np.empty((NSCALARS, self.B))
        _check(lib().ddpb_get_scalars(self._h, _d(out)))
        return {n: out[i].copy() for i, n in enumerate(SCALAR_FIELDS)}

    def scalars_device_ptr(self):
        p = C.c_void_p()
        _check(lib().ddpb_scalars_device_ptr(self._h, C.byref(p)))
        return p.value

    # ---- hot path ----
    def compute_derivatives(self):
        _check(lib().ddpb_compute_derivatives(self._h))

    def backward_pass(self, mu, reg):
        mu = _c(np.broadcast_to(mu, (self.B,))).copy(); reg = _c(np.broadcast_to(reg, (self.B,))).copy()
        _check(lib().ddpb_backward_pass(self._h, _d(mu), _d(reg)))
        return mu, reg

    def forward_pass(self):
        step = np.empty(self.B)
        _check(lib().ddpb_forward_pass(self._h, _d(step)))
        return step

    def cost_seq_aug(self):
        c = np.empty((self.B, self.T + 1))
        _check(lib().ddpb_cost_seq_aug(self._h, _d(c)))
        return c

    def optimality(self):
        a, b = np.empty(self.B), np.empty(self.B)
        _check(lib().ddpb_optimality(self._h, _d(a), _d(b)))
        return a, b

    def solve(self, max_iterations, mu, reg=0.0, w=None, n=None, thr=0.0, read_back=True, out=None):
        w = 1.0 / mu if w is None else w
        n = 1.0 / mu ** 0.1 if n is None else n
        sp = SolverParamsC(max_iterations, thr, mu, reg, w, n)
        if not read_back:
            _check(lib().ddpb_solve_async(self._h, C.byref(sp)))
            return None
        if out is None:
            out = np.empty((NSCALARS, self.B))
        _check(lib().ddpb_solve(self._h, C.byref(sp), _d(out)))
        return {nm: out[i].copy() for i, nm in enumerate(SCALAR_FIELDS)}

    def snapshot_save(self):
        _check(lib().ddpb_snapshot_save(self._h))

    def snapshot_restore(self):
        _check(lib().ddpb_snapshot_restore(self._h))

    # ---- introspection ----
    def launch_count(self):
        return int(lib().ddpb_launch_count(self._h))

    def set_subbatches(self, n):
        _check(lib().ddpb_set_subbatches(self._h, int(n)))

    def default_subbatches(self):
        """what ddpb_batch_create chooses: enough problems per sub-batch to keep whole CTAs busy"""
        return 4 if self.B >= 512 else (2 if self.B >= 128 else 1)

    def enable_phase_timing(self, on=True):
        _check(lib().ddpb_enable_phase_timing(self._h, 1 if on else 0))

    def phase_times(self):
        ms = (C.c_float * 4)()
        _check(lib().ddpb_phase_times(self._h, ms))
        return dict(linearize=ms[0], backward=ms[1], forward=ms[2], other=ms[3])


def measure_fp64_peak(device=0):
    v = C.c_double()
    _check(lib().ddpb_measure_fp64_peak(device, C.byref(v)))
    return v.value


def host_alloc(nbytes):
    """Pinned host buffer (numpy float64 view) for the H2D / D2H transfers."""
    p = C.c_void_p()
    _check(lib().ddpb_host_alloc(C.byref(p), C.c_ulonglong(nbytes)))
    arr = np.ctypeslib.as_array(C.cast(p, _dp), shape=(nbytes // 8,))
    return arr, p


def host_free(p):
    _check(lib().ddpb_host_free(p))
