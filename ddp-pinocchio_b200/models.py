"""Embedded robot parameter tables (there is no URDF parser and no URDF file offline).

The reference builds its model with pinocchio::urdf::buildModel
(/root/reference/include/ddp/pinocchio_model.ipp:98-113) from ur5_gripper.urdf
(/root/reference/test/pinocchio_ddp.cpp:30); here the same information -- kinematic tree,
joint placements and axes, link inertias, operational frames -- is given as plain arrays,
which is exactly what `ddpb_model_desc` (include/ddpb.h) carries across the C ABI.
"""
from dataclasses import dataclass, field
import numpy as np

KIND_PENDULUM = 0
KIND_REVOLUTE_CHAIN = 1   # serial chain of one-degree-of-freedom joints (the name predates the prismatic joints)
JOINT_REVOLUTE = 0
JOINT_PRISMATIC = 1


def _rpy(r, p, y):
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


@dataclass
class ModelDesc:
    name: str
    kind: int
    n_joints: int
    parent: np.ndarray = None          # int32 [n]
    placement_rot: np.ndarray = None   # [n,3,3] joint frame -> parent coordinates
    placement_trans: np.ndarray = None  # [n,3]
    axis: np.ndarray = None            # [n,3] unit axis in the joint frame
    mass: np.ndarray = None            # [n]
    com: np.ndarray = None             # [n,3]
    inertia: np.ndarray = None         # [n,6] xx,xy,xz,yy,yz,zz about the COM
    gravity: np.ndarray = field(default_factory=lambda: np.array([0.0, 0.0, -9.81]))
    frame_parent: np.ndarray = None    # int32 [f]
    frame_rot: np.ndarray = None       # [f,3,3]
    frame_trans: np.ndarray = None     # [f,3]
    frame_names: list = None
    pend_mass: float = 1.0
    pend_length: float = 1.0
    joint_type: np.ndarray = None      # int32 [n]: JOINT_REVOLUTE / JOINT_PRISMATIC (None = all revolute)

    @property
    def joint_types(self):
        return np.zeros(self.n_joints, dtype=np.int32) if self.joint_type is None else np.asarray(self.joint_type, dtype=np.int32)

    @property
    def nq(self):
        return 1 if self.kind == KIND_PENDULUM else self.n_joints

    nv = nq

    @property
    def n_frames(self):
        return 0 if self.frame_parent is None else len(self.frame_parent)


def pendulum(mass=1.0, length=1.0):
    """pendulum_model_t{1.0, 1.0} of /root/reference/test/pendulum_ddp.cpp:30."""
    return ModelDesc(name="pendulum", kind=KIND_PENDULUM, n_joints=1, pend_mass=mass, pend_length=length)


def _chain(name, joints, frames_extra):
    n = len(joints)
    R = np.zeros((n, 3, 3)); p = np.zeros((n, 3)); ax = np.zeros((n, 3))
    m = np.zeros(n); c = np.zeros((n, 3)); I = np.zeros((n, 6)); jt = np.zeros(n, dtype=np.int32)
    for i, j in enumerate(joints):
        R[i] = _rpy(*j["rpy"]); p[i] = j["xyz"]; ax[i] = j["axis"]; jt[i] = j.get("type", JOINT_REVOLUTE)
        m[i] = j["mass"]; c[i] = j["com"]; I[i] = j["inertia"]
    # one frame per joint (as Pinocchio adds a frame for every URDF joint/link) + extras
    fpar = list(range(n)); fR = [np.eye(3)] * n; fp = [np.zeros(3)] * n; names = [f"joint{i + 1}" for i in range(n)]
    for (nm, par, rpy, xyz) in frames_extra:
        fpar.append(par); fR.append(_rpy(*rpy)); fp.append(np.array(xyz, dtype=float)); names.append(nm)
    return ModelDesc(name=name, kind=KIND_REVOLUTE_CHAIN, n_joints=n, parent=np.arange(-1, n - 1, dtype=np.int32),
                     placement_rot=R, placement_trans=p, axis=ax, mass=m, com=c, inertia=I,
                     frame_parent=np.array(fpar, dtype=np.int32), frame_rot=np.array(fR), frame_trans=np.array(fp),
                     frame_names=names, joint_type=jt if jt.any() else None)


def chain_2r():
    """Synthetic planar double pendulum (BASELINE.json configs[1]; test/pinocchio_double.cpp is an empty file)."""
    link = dict(rpy=(0, 0, 0), axis=(0, 1, 0), mass=1.0, com=(0, 0, -0.5), inertia=(1 / 12, 0, 0, 1 / 12, 0, 1e-3))
    j1 = dict(link, xyz=(0, 0, 0))
    j2 = dict(link, xyz=(0, 0, -1.0))
    return _chain("chain_2r", [j1, j2], [("tip", 1, (0, 0, 0), (0, 0, -1.0))])


def arm_6r():
    """UR5-style 6R arm (kinematics/inertias after the public ur_description UR5 table; gripper omitted)."""
    h = np.pi / 2
    J = [
        dict(xyz=(0, 0, 0.089159), rpy=(0, 0, 0), axis=(0, 0, 1), mass=3.7, com=(0, 0, 0),
             inertia=(0.010267495893, 0, 0, 0.010267495893, 0, 0.00666)),
        dict(xyz=(0, 0.13585, 0), rpy=(0, h, 0), axis=(0, 1, 0), mass=8.393, com=(0, 0, 0.28),
             inertia=(0.22689067591, 0, 0, 0.22689067591, 0, 0.0151074)),
        dict(xyz=(0, -0.1197, 0.425), rpy=(0, 0, 0), axis=(0, 1, 0), mass=2.275, com=(0, 0, 0.25),
             inertia=(0.049443313556, 0, 0, 0.049443313556, 0, 0.004095)),
        dict(xyz=(0, 0, 0.39225), rpy=(0, h, 0), axis=(0, 1, 0), mass=1.219, com=(0, 0, 0),
             inertia=(0.111172755531, 0, 0, 0.111172755531, 0, 0.21942)),
        dict(xyz=(0, 0.093, 0), rpy=(0, 0, 0), axis=(0, 0, 1), mass=1.219, com=(0, 0, 0),
             inertia=(0.111172755531, 0, 0, 0.111172755531, 0, 0.21942)),
        dict(xyz=(0, 0, 0.09465), rpy=(0, 0, 0), axis=(0, 1, 0), mass=0.1879, com=(0, 0, 0),
             inertia=(0.0171364731454, 0, 0, 0.0171364731454, 0, 0.033822)),
    ]
    return _chain("arm_6r_ur5", J, [("ee_link", 5, (0, 0, h), (0, 0.0823, 0)), ("tool0", 5, (-h, 0, 0), (0, 0.0823, 0))])


def arm_7r():
    """iiwa-style 7R arm (all joints about local z, alternating +-90 deg twists)."""
    h, pi = np.pi / 2, np.pi
    J = [
        dict(xyz=(0, 0, 0.1575), rpy=(0, 0, 0), axis=(0, 0, 1), mass=4.0, com=(0, -0.03, 0.12), inertia=(0.1, 0, 0, 0.09, 0, 0.02)),
        dict(xyz=(0, 0, 0.2025), rpy=(h, 0, pi), axis=(0, 0, 1), mass=4.0, com=(0.0003, 0.059, 0.042), inertia=(0.05, 0, 0, 0.018, 0, 0.044)),
        dict(xyz=(0, 0.2045, 0), rpy=(h, 0, pi), axis=(0, 0, 1), mass=3.0, com=(0, 0.03, 0.13), inertia=(0.08, 0, 0, 0.075, 0, 0.01)),
        dict(xyz=(0, 0, 0.2155), rpy=(h, 0, 0), axis=(0, 0, 1), mass=2.7, com=(0, 0.067, 0.034), inertia=(0.03, 0, 0, 0.01, 0, 0.029)),
        dict(xyz=(0, 0.1845, 0), rpy=(-h, pi, 0), axis=(0, 0, 1), mass=1.7, com=(0.0001, 0.021, 0.076), inertia=(0.02, 0, 0, 0.018, 0, 0.005)),
        dict(xyz=(0, 0, 0.2155), rpy=(h, 0, 0), axis=(0, 0, 1), mass=1.8, com=(0, 0.0006, 0.0004), inertia=(0.005, 0, 0, 0.0036, 0, 0.0047)),
        dict(xyz=(0, 0.081, 0), rpy=(-h, pi, 0), axis=(0, 0, 1), mass=0.3, com=(0, 0, 0.02), inertia=(0.001, 0, 0, 0.001, 0, 0.001)),
    ]
    return _chain("arm_7r_iiwa", J, [("tool", 6, (0, 0, 0), (0, 0, 0.045))])


def chain_nr(n, seed=7):
    """Random but well-conditioned n-joint chain (general axes / placements / inertias) for parity sweeps."""
    rng = np.random.RandomState(seed)
    J = []
    for i in range(n):
        a = rng.randn(3); a /= np.linalg.norm(a)
        A = rng.randn(3, 3) * 0.1
        Ic = A @ A.T + 0.02 * np.eye(3)
        J.append(dict(xyz=tuple(rng.uniform(-0.3, 0.3, 3)), rpy=tuple(rng.uniform(-1, 1, 3)), axis=tuple(a),
                      mass=float(rng.uniform(0.5, 3.0)), com=tuple(rng.uniform(-0.1, 0.1, 3)),
                      inertia=(Ic[0, 0], Ic[0, 1], Ic[0, 2], Ic[1, 1], Ic[1, 2], Ic[2, 2])))
    return _chain(f"chain_{n}r_random", J, [("tip", n - 1, tuple(rng.uniform(-1, 1, 3)), tuple(rng.uniform(-0.2, 0.2, 3)))])


def chain_mixed(n, types, seed=11):
    """chain_nr with the given joint types ('R' / 'P' per joint): general axes, placements and inertias."""
    m = chain_nr(n, seed)
    m.name = "chain_" + types.lower()
    m.joint_type = np.array([JOINT_PRISMATIC if t == "P" else JOINT_REVOLUTE for t in types], dtype=np.int32)
    assert len(m.joint_type) == n
    return m


def scara_rrpr():
    """SCARA-style arm: two vertical revolute joints, a vertical prismatic quill and a wrist rotation (gravity along the
    prismatic axis; what pinocchio::urdf::buildModel makes of a URDF with revolute + prismatic joints)."""
    z = (0, 0, 1)
    J = [
        dict(xyz=(0, 0, 0.4), rpy=(0, 0, 0), axis=z, mass=6.0, com=(0.2, 0, 0.03), inertia=(0.02, 0, 0, 0.12, 0, 0.125)),
        dict(xyz=(0.4, 0, 0.05), rpy=(0, 0, 0), axis=z, mass=4.0, com=(0.15, 0, 0.02), inertia=(0.012, 0, 0, 0.05, 0, 0.055)),
        dict(xyz=(0.3, 0, 0), rpy=(np.pi, 0, 0), axis=z, mass=1.2, com=(0, 0, -0.1), inertia=(0.01, 0, 0, 0.01, 0, 0.0008),
             type=JOINT_PRISMATIC),
        dict(xyz=(0, 0, 0.05), rpy=(0, 0, 0), axis=z, mass=1.0, com=(0.01, 0, 0.03), inertia=(0.004, 0, 0, 0.004, 0, 0.003)),
    ]
    return _chain("scara_rrpr", J, [("flange", 3, (0, 0, 0), (0, 0, 0.06))])


def _rodrigues(a, q):
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(q) * K + (1 - np.cos(q)) * (K @ K)


def joint_step(m, i, qi, R, p):
    """World rotation / origin of joint frame i from those of its parent (revolute: rotation about the axis; prismatic:
    translation along it)."""
    p = p + R @ m.placement_trans[i]
    if m.joint_types[i] == JOINT_PRISMATIC:
        R = R @ m.placement_rot[i]
        return R, p + R @ (m.axis[i] * qi)
    return R @ m.placement_rot[i] @ _rodrigues(m.axis[i], qi), p


def potential_energy(m, q):
    """U(q) = -sum_i m_i g . p_com_i(q) of a serial chain (numpy forward kinematics; input generation only)."""
    R, p, U = np.eye(3), np.zeros(3), 0.0
    for i in range(m.n_joints):
        R, p = joint_step(m, i, q[i], R, p)
        U -= m.mass[i] * float(m.gravity @ (p + R @ m.com[i]))
    return U


def frame_position(m, q, frame):
    """World position of an operational frame (numpy forward kinematics; input generation only)."""
    R, p = np.eye(3), np.zeros(3)
    for i in range(int(m.frame_parent[frame]) + 1):
        R, p = joint_step(m, i, q[i], R, p)
    return p + R @ m.frame_trans[frame]


def gravity_torque(m, q, h=1e-6):
    """tau = dU/dq: the static torque that holds the chain at q (hold-still initial control guess of the
    synthetic workloads; central differences are plenty for an initial guess)."""
    q = np.asarray(q, dtype=float)
    g = np.zeros(m.n_joints)
    for i in range(m.n_joints):
        e = np.zeros(m.n_joints); e[i] = h
        g[i] = (potential_energy(m, q + e) - potential_energy(m, q - e)) / (2 * h)
    return g


BY_NAME = {"pendulum": pendulum, "chain_2r": chain_2r, "arm_6r": arm_6r, "arm_7r": arm_7r, "scara_rrpr": scara_rrpr}
