"""Deterministic synthetic DDP problem batches.

Counter-based splitmix64 streams, seed 0xDD9 + problem_id, generated once on the host and
shared by the GPU path and the CPU oracle.  Replaces the reference's unseeded
`eq.jac().setRandom()` (/root/reference/include/ddp/ddp.hpp:762) by an explicit input.
"""
from dataclasses import dataclass
import numpy as np
from . import models

EQ_NONE, EQ_CONFIG, EQ_SPATIAL = 0, 1, 2
HESS_FD_JACOBIANS, HESS_FD_VALUES = 0, 1   # dynamics_t::second_order_finite_diff = false / true
FJ_WORLD, FJ_LOCAL_WORLD_ALIGNED = 0, 1

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(seed, count):
    """`count` uniform doubles in [0,1) from stream `seed` (vectorised over a seed array)."""
    seed = np.atleast_1d(np.asarray(seed, dtype=np.uint64))
    idx = np.arange(1, count + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = seed[:, None] + idx[None, :] * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


@dataclass
class Workload:
    name: str
    model: models.ModelDesc
    B: int
    T: int
    dt: float = 0.01
    c: float = 1.0
    hess_mode: int = HESS_FD_JACOBIANS
    eq_kind: int = EQ_CONFIG
    eq_every_knot: bool = False   # False: terminal constraint only (rows at knot T-2, test/pendulum_ddp.cpp:37-47)
    frame: int = 0
    jac_mode: int = FJ_WORLD
    # double-precision solver parameters (the reference's mu=1e20 / thr=1e-80 are MPFR-only)
    mu: float = 1e4
    reg: float = 0.0
    iterations: int = 10
    # initial control guess: "zero" (the reference tests' zero-control generator, test/pendulum_ddp.cpp:70-82)
    # or "hold" (static gravity compensation at q0, so a 1-2 s horizon does not start from a free fall)
    init_controls: str = "hold"
    v0_scale: float = 0.0
    # guarded mode (ddpb_params.max_backward_restarts): a problem whose backward pass needs more than this many
    # mu-doubling restarts is marked diverged instead of looping to mu = inf (0 = reference behaviour).  Measured on
    # C2 and C5 (tools/wl_sweep.py): restarts per pass are bimodal -- 0 or 1, or every one up to the limit -- and the
    # limits 3 and 10 freeze exactly the same problems; 3 detects a hopeless problem after 4 passes instead of 11.
    max_restarts: int = 3
    target_offset: float = 1.0

    @property
    def n(self):
        return self.model.nq

    @property
    def nx(self):
        return 2 * self.n

    ndx = nx

    @property
    def nu(self):
        return self.n

    def ne_per_knot(self):
        ne_rows = 3 if self.eq_kind == EQ_SPATIAL else (self.n if self.eq_kind == EQ_CONFIG else 0)
        ne = np.zeros(self.T, dtype=np.int32)
        if self.eq_kind != EQ_NONE:
            if self.eq_every_knot:
                ne[:] = ne_rows
            else:
                ne[self.T - 2] = ne_rows
        return ne

    def generate(self, B=None, first_problem=0):
        """Per-problem inputs: x0 [B,nx], us [B,T,nu], targets [B,sum ne], mult_jac [B, sum ne*ndx].

        Chains: a rest-to-rest reaching task -- q0 ~ U(-pi/2, pi/2)^n, v0 = v0_scale * U(-0.5, 0.5)^n
        (0 by default: a 1-2 s open-loop horizon of a free-falling arm is chaotic and sends a few percent of
        the reference algorithm's mu/reg restart loops to infinity, see DESIGN.md "Workloads"), controls =
        static gravity compensation at q0, goal configuration q* = q0 + target_offset * U(-1, 1)^n imposed on
        q_T (config constraint) or its end-effector position FK(q*) (spatial constraint).
        Pendulum: problem 0 is the reference scenario (x0 = 0, u = 0, q_T = 3.14, test/pendulum_ddp.cpp:50-58);
        further problems start at q0 ~ U(-pi/2, pi/2)."""
        B = self.B if B is None else B
        n, T = self.n, self.T
        ne = self.ne_per_knot()
        net = int(ne.sum())
        nknots = int((ne > 0).sum())
        cnt = 3 * n + net * self.ndx
        r = splitmix64(np.uint64(0xDD9) + np.arange(first_problem, first_problem + B, dtype=np.uint64), cnt)
        q0 = (r[:, :n] - 0.5) * np.pi
        v0 = (r[:, n:2 * n] - 0.5) * self.v0_scale
        dq = (2.0 * r[:, 2 * n:3 * n] - 1.0) * self.target_offset
        mj = 2.0 * r[:, 3 * n:3 * n + net * self.ndx] - 1.0
        us = np.zeros((B, T, self.nu))
        if self.model.kind == models.KIND_PENDULUM:
            qstar = np.full((B, 1), 3.14) + dq * (np.arange(first_problem, first_problem + B)[:, None] > 0)
            q0 = q0 * (np.arange(first_problem, first_problem + B)[:, None] > 0)
            v0 = v0 * 0.0
        else:
            qstar = q0 + dq
            if self.init_controls == "hold":
                for b in range(B):
                    us[b, :, :] = models.gravity_torque(self.model, q0[b])[None, :]
        if self.eq_kind == EQ_SPATIAL:
            one = np.stack([models.frame_position(self.model, qstar[b], self.frame) for b in range(B)])
        elif self.eq_kind == EQ_CONFIG:
            one = qstar
        else:
            one = np.zeros((B, 0))
        tg = np.tile(one, (1, nknots)) if net else np.zeros((B, 0))
        x0 = np.concatenate([q0, v0], axis=1)
        return dict(x0=np.ascontiguousarray(x0), us=us, targets=np.ascontiguousarray(tg), mult_jac=np.ascontiguousarray(mj), ne=ne)


def workload(name, B=None, T=None):
    """The five BASELINE.json configs (C1..C5) by short name."""
    if name == "c1":   # pendulum DDP, single problem (test/pendulum_ddp.cpp)
        w = Workload("c1_pendulum", models.pendulum(), 1, 200, hess_mode=HESS_FD_VALUES, init_controls="zero")
    elif name == "c2":  # synthetic 2R "double pendulum", 4096 x 200
        w = Workload("c2_chain2r_4096x200", models.chain_2r(), 4096, 200)
    elif name == "c3":  # 6R UR5-style arm, 1024 x 100, analytic first order + FD-of-Jacobians Hessians
        w = Workload("c3_arm6r_1024x100", models.arm_6r(), 1024, 100)
    elif name == "c4":  # same arm, spatial end-effector equality constraint
        m = models.arm_6r()
        w = Workload("c4_arm6r_spatial_1024x100", m, 1024, 100, eq_kind=EQ_SPATIAL, frame=m.n_frames - 1)
    elif name == "c5":  # 7R arm, value-based finite-difference Hessians, 8192 x 200 over 8 GPUs
        # value-based FD Hessians (eps = 2^-13) are noisier: goals within 0.25 rad and a 1 s horizon (dt = 0.005, as long
        # as C3's) keep the reference algorithm convergent -- with dt = 0.01 the undamped Riccati recursion over 200
        # knots of an arm held against gravity blows up for a third of the problems at the very first backward pass
        # (measured: 37 % frozen by the guard against 1 % here, tools/wl_sweep2.py)
        w = Workload("c5_arm7r_fd_8192x200", models.arm_7r(), 8192, 200, dt=0.005, hess_mode=HESS_FD_VALUES, target_offset=0.25)
    else:
        raise ValueError(name)
    if B is not None:
        w.B = B
    if T is not None:
        w.T = T
    return w
