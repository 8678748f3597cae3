"""CPU emulator of the CUDA kernels' CTA bodies (tests/hostemu/hostemu.cpp) behind the same Python
interface as ddp_pinocchio_b200.capi.Batch.  TEST INFRASTRUCTURE ONLY: lets `pytest -m "not gpu"` check the
kernel arithmetic against the oracle in a GPU-less container.  The product never loads it."""
import ctypes as C
import os
import subprocess
import numpy as np

import ddpb_loader

P = ddpb_loader.load()
from ddp_pinocchio_b200 import capi  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostemu", "hostemu.cpp")
LIB = os.path.join(HERE, "hostemu", "libhostemu.so")
_dp = C.POINTER(C.c_double)
_lib = None
XS, US, XS_NEW, US_NEW, TARGETS, M_ORIGIN, M_VAL, M_JAC, FB_ORIGIN, FB_K, FB_KK, SCAL, COSTS_OLD = range(13)
S_IDX = {n: i for i, n in enumerate(capi.SCALAR_FIELDS)}


def build(force=False):
    csrc = os.path.join(ddpb_loader.PKG_DIR, "csrc")
    deps = [SRC] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".hpp"))]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.check_call(["g++", "-O2", "-march=x86-64-v3", "-std=c++17", "-fPIC", "-shared", "-DDDPB_PRISMATIC=1", "-x", "c++", SRC, "-o", LIB])  # the variant with the joint-type tests (csrc/variant.cuh) serves every model


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.emu_create.restype = C.c_void_p
        _lib.emu_ptr.restype = _dp
    return _lib


class EmuBatch(capi.BatchBase):
    def __init__(self, model, B, T, ne_per_knot=None, dt=0.01, c=1.0, hess_mode=0, eq_kind=0, frame_id=0, jac_mode=0, device=0,
                 max_restarts=0):
        self._keep = []
        self._init_dims(model, B, T, ne_per_knot, eq_kind)
        md = capi.model_desc_c(model, self._keep)
        pr = capi.ParamsC(dt, c, hess_mode, eq_kind, frame_id, jac_mode, max_restarts)
        self._h = C.c_void_p(lib().emu_create(C.byref(md), B, T, self.ne.ctypes.data_as(C.POINTER(C.c_int)), C.byref(pr)))
        assert self._h.value, "emu_create failed"

    @classmethod
    def from_workload(cls, w, B=None, device=0, max_restarts=None):
        return cls(w.model, w.B if B is None else B, w.T, w.ne_per_knot(), dt=w.dt, c=w.c, hess_mode=w.hess_mode, eq_kind=w.eq_kind,
                   frame_id=w.frame, jac_mode=w.jac_mode, max_restarts=w.max_restarts if max_restarts is None else max_restarts)

    def close(self):
        if self._h:
            lib().emu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _arr(self, which, count):
        """numpy view [count][B] of a device-layout array"""
        p = lib().emu_ptr(self._h, which)
        return np.ctypeslib.as_array(p, shape=(max(count, 1), self.B))[:count]

    # host [B][C] <-> device [C][B]
    def _put(self, which, host, count):
        self._arr(which, count)[:] = np.asarray(host, dtype=float).reshape(self.B, count).T

    def _get(self, which, count):
        return np.ascontiguousarray(self._arr(which, count).T)

    def sync(self):
        pass

    def set_trajectory(self, xs, us):
        B, T, nx, nu = self.B, self.T, self.nx, self.nu
        self._put(XS, xs, (T + 1) * nx); self._put(US, us, T * nu)
        self._put(XS_NEW, xs, (T + 1) * nx); self._put(US_NEW, us, T * nu)

    def get_trajectory(self):
        B, T, nx, nu = self.B, self.T, self.nx, self.nu
        return self._get(XS, (T + 1) * nx).reshape(B, T + 1, nx), self._get(US, T * nu).reshape(B, T, nu)

    def get_new_trajectory(self):
        B, T, nx, nu = self.B, self.T, self.nx, self.nu
        return self._get(XS_NEW, (T + 1) * nx).reshape(B, T + 1, nx), self._get(US_NEW, T * nu).reshape(B, T, nu)

    def make_trajectory(self, x0, us=None):
        B, T, nx, nu = self.B, self.T, self.nx, self.nu
        self._arr(XS, (T + 1) * nx)[:nx] = np.asarray(x0, dtype=float).reshape(B, nx).T
        self._put(US, np.zeros((B, T * nu)) if us is None else us, T * nu)
        lib().emu_rollout(self._h)
        self._arr(XS_NEW, (T + 1) * nx)[:] = self._arr(XS, (T + 1) * nx)
        self._arr(US_NEW, T * nu)[:] = self._arr(US, T * nu)

    def set_targets(self, targets):
        if self.ne_total:
            self._put(TARGETS, targets, self.ne_total)

    def set_multipliers(self, val=None, jac=None):
        net, nd = self.ne_total, self.ndx
        if net:
            self._put(M_VAL, np.zeros((self.B, net)) if val is None else val, net)
            self._put(M_JAC, np.zeros((self.B, net * nd)) if jac is None else jac, net * nd)
        self._arr(M_ORIGIN, self.T * self.nx)[:] = self._arr(XS, (self.T + 1) * self.nx)[: self.T * self.nx]

    def get_multipliers(self):
        net, nd = self.ne_total, self.ndx
        return (self._get(M_ORIGIN, self.T * self.nx).reshape(self.B, self.T, self.nx), self._get(M_VAL, net), self._get(M_JAC, net * nd))

    def set_feedback(self, k, K):
        self._put(FB_K, k, self.T * self.nu); self._put(FB_KK, K, self.T * self.nu * self.ndx)
        self._arr(FB_ORIGIN, self.T * self.nx)[:] = self._arr(XS, (self.T + 1) * self.nx)[: self.T * self.nx]

    def get_feedback(self):
        B, T, nx, nu, nd = self.B, self.T, self.nx, self.nu, self.ndx
        return (self._get(FB_ORIGIN, T * nx).reshape(B, T, nx), self._get(FB_K, T * nu).reshape(B, T, nu),
                self._get(FB_KK, T * nu * nd).reshape(B, T, nd, nu))

    def get_derivatives(self, fields=None):
        sizes = self.deriv_sizes()
        out = {}
        for f in (fields or capi.DERIV_FIELDS):
            a = np.zeros((self.B, sizes[f]))
            if f not in ("lfx", "lfxx") and sizes[f]:
                lib().emu_field(self._h, a.ctypes.data_as(_dp), capi.DERIV_FIELDS.index(f) - 2, 0)
            out[f] = a
        return out

    def set_derivatives(self, d):
        for f, a in d.items():
            if f in ("lfx", "lfxx") or a.size == 0:
                continue
            a = np.ascontiguousarray(a, dtype=float)
            bad = lib().emu_field(self._h, a.ctypes.data_as(_dp), capi.DERIV_FIELDS.index(f) - 2, 1)
            if bad:
                raise RuntimeError(f"{f}: {bad} entries differ from the structural constants the slab does not store")

    def set_scalar(self, name, values):
        self._arr(SCAL, capi.NSCALARS)[S_IDX[name]] = np.broadcast_to(values, (self.B,))

    def get_scalars(self):
        s = self._arr(SCAL, capi.NSCALARS)
        return {n: s[i].copy() for i, n in enumerate(capi.SCALAR_FIELDS)}

    def compute_derivatives(self):
        lib().emu_linearize(self._h)

    def backward_pass(self, mu, reg):
        self.set_scalar("mu", mu); self.set_scalar("reg", reg)
        lib().emu_backward(self._h)
        s = self.get_scalars()
        return s["mu"], s["reg"]

    def forward_pass(self):
        lib().emu_forward(self._h)
        return self.get_scalars()["step"]

    def cost_seq_aug(self):
        lib().emu_cost(self._h)
        return self._get(COSTS_OLD, self.T + 1)

    def optimality(self):
        lib().emu_optimality(self._h)
        s = self.get_scalars()
        return s["opt_obj"], s["opt_constr"]

    def solve(self, max_iterations, mu, reg=0.0, w=None, n=None, thr=0.0, read_back=True):
        w = 1.0 / mu if w is None else w
        n = 1.0 / mu ** 0.1 if n is None else n
        sp = capi.SolverParamsC(max_iterations, thr, mu, reg, w, n)
        lib().emu_solve(self._h, C.byref(sp))
        return self.get_scalars()


class ModelEvalArgsC(C.Structure):  # csrc/model_eval.cuh ModelEvalArgs
    _fields_ = [("n", C.c_int), ("frame", C.c_int), ("jac_mode", C.c_int), ("q", _dp), ("v", _dp), ("tau", _dp), ("acc", _dp), ("a", _dp),
                ("a_dq", _dp), ("a_dv", _dp), ("a_dtau", _dp), ("tau_out", _dp), ("p3", _dp), ("J", _dp)]


class EmuModel(capi.ModelBase):
    """capi.Model on the CPU: the body of model_eval_kernel one sample at a time."""

    def __init__(self, model):
        self._keep = []
        self._init_model(model)
        self._md = capi.model_desc_c(model, self._keep)

    def _eval(self, n, q, frame=0, jac_mode=0, **arrs):
        A = ModelEvalArgsC()
        A.n = n; A.frame = frame; A.jac_mode = jac_mode
        A.q = q.ctypes.data_as(_dp)
        for k, x in arrs.items():
            if x is not None:
                setattr(A, k, x.ctypes.data_as(_dp))
        assert lib().emu_model_eval(C.byref(self._md), C.byref(A)) == 0
