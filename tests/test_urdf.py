"""URDF text -> model through the C ABI (ddpb_model_create_from_urdf, csrc/urdf_host.hpp): the replacement of
pinocchio::urdf::buildModel (/root/reference/include/ddp/pinocchio_model.ipp:98-113) for serial chains of revolute / prismatic joints.  No GPU needed:
model construction is host code."""
import numpy as np
import pytest

import ddpb_loader

P = ddpb_loader.load()
from ddp_pinocchio_b200 import capi, models  # noqa: E402

H = "1.5707963267948966"

# the 6R arm of models.arm_6r() written as a URDF: root link on the world, six revolute joints, two fixed frames on the last link
UR5 = f"""<?xml version="1.0"?>
<!-- UR5-style arm -->
<robot name="ur5_like">
  <link name="base_link"><inertial><mass value="4.0"/><inertia ixx="0.0044" ixy="0" ixz="0" iyy="0.0044" iyz="0" izz="0.0072"/></inertial></link>
  <joint name="shoulder_pan_joint" type="revolute"><parent link="base_link"/><child link="shoulder_link"/>
    <origin xyz="0 0 0.089159" rpy="0 0 0"/><axis xyz="0 0 1"/><limit lower="-6.28" upper="6.28" effort="150" velocity="3.15"/></joint>
  <link name="shoulder_link"><inertial><mass value="3.7"/><origin xyz="0 0 0" rpy="0 0 0"/>
    <inertia ixx="0.010267495893" ixy="0" ixz="0" iyy="0.010267495893" iyz="0" izz="0.00666"/></inertial></link>
  <joint name="shoulder_lift_joint" type="revolute"><parent link="shoulder_link"/><child link="upper_arm_link"/>
    <origin xyz="0 0.13585 0" rpy="0 {H} 0"/><axis xyz="0 1 0"/></joint>
  <link name="upper_arm_link"><inertial><mass value="8.393"/><origin xyz="0 0 0.28"/>
    <inertia ixx="0.22689067591" ixy="0" ixz="0" iyy="0.22689067591" iyz="0" izz="0.0151074"/></inertial></link>
  <joint name="elbow_joint" type="revolute"><parent link="upper_arm_link"/><child link="forearm_link"/>
    <origin xyz="0 -0.1197 0.425" rpy="0 0 0"/><axis xyz="0 1 0"/></joint>
  <link name="forearm_link"><inertial><mass value="2.275"/><origin xyz="0 0 0.25"/>
    <inertia ixx="0.049443313556" ixy="0" ixz="0" iyy="0.049443313556" iyz="0" izz="0.004095"/></inertial></link>
  <joint name="wrist_1_joint" type="revolute"><parent link="forearm_link"/><child link="wrist_1_link"/>
    <origin xyz="0 0 0.39225" rpy="0 {H} 0"/><axis xyz="0 1 0"/></joint>
  <link name="wrist_1_link"><inertial><mass value="1.219"/>
    <inertia ixx="0.111172755531" ixy="0" ixz="0" iyy="0.111172755531" iyz="0" izz="0.21942"/></inertial></link>
  <joint name="wrist_2_joint" type="continuous"><parent link="wrist_1_link"/><child link="wrist_2_link"/>
    <origin xyz="0 0.093 0"/><axis xyz="0 0 2"/></joint>
  <link name="wrist_2_link"><inertial><mass value="1.219"/>
    <inertia ixx="0.111172755531" ixy="0" ixz="0" iyy="0.111172755531" iyz="0" izz="0.21942"/></inertial></link>
  <joint name="wrist_3_joint" type="revolute"><parent link="wrist_2_link"/><child link="wrist_3_link"/>
    <origin xyz="0 0 0.09465"/><axis xyz="0 1 0"/></joint>
  <link name="wrist_3_link"><inertial><mass value="0.1879"/>
    <inertia ixx="0.0171364731454" ixy="0" ixz="0" iyy="0.0171364731454" iyz="0" izz="0.033822"/></inertial></link>
  <joint name="ee_fixed_joint" type="fixed"><parent link="wrist_3_link"/><child link="ee_link"/><origin xyz="0 0.0823 0" rpy="0 0 {H}"/></joint>
  <link name="ee_link"/>
  <joint name="tool_fixed_joint" type='fixed'><parent link="wrist_3_link"/><child link="tool0"/><origin xyz="0 0.0823 0" rpy="-{H} 0 0"/></joint>
  <link name="tool0"/>
</robot>
"""


def test_ur5_like_urdf_equals_the_embedded_table():
    m = capi.model_from_urdf(UR5)
    ref = models.arm_6r()
    assert m.n_joints == 6
    for f in ("placement_rot", "placement_trans", "axis", "mass", "com", "inertia"):
        assert np.array_equal(getattr(m, f), getattr(ref, f)), f     # same numbers, to the bit
    # BODY frames: the six moving links, then the two fixed links of the flange (traversal order)
    assert m.frame_names == ["shoulder_link", "upper_arm_link", "forearm_link", "wrist_1_link", "wrist_2_link", "wrist_3_link", "ee_link", "tool0"]
    assert list(m.frame_parent) == [0, 1, 2, 3, 4, 5, 5, 5]
    assert np.array_equal(m.frame_trans[:6], np.zeros((6, 3))) and np.allclose(m.frame_trans[6:], [[0, 0.0823, 0]] * 2, atol=0, rtol=0)
    assert np.array_equal(m.frame_trans[6:], ref.frame_trans[6:]) and list(ref.frame_parent[6:]) == [5, 5]


def test_fixed_joint_merges_inertia_into_the_supporting_body():
    """A fixed child link with mass: composite rigid body (parallel-axis theorem), placement through the fixed joint."""
    text = """<robot name="r"><link name="world_link"/>
      <joint name="j1" type="revolute"><parent link="world_link"/><child link="a"/><origin xyz="0 0 0.1"/><axis xyz="0 1 0"/></joint>
      <link name="a"><inertial><mass value="2.0"/><origin xyz="0 0 -0.5"/><inertia ixx="0.1" ixy="0" ixz="0" iyy="0.2" iyz="0" izz="0.3"/></inertial></link>
      <joint name="f" type="fixed"><parent link="a"/><child link="payload"/><origin xyz="0.2 0 -1.0" rpy="0 0 1.5707963267948966"/></joint>
      <link name="payload"><inertial><mass value="1.0"/><origin xyz="0.1 0 0"/><inertia ixx="0.01" ixy="0" ixz="0" iyy="0.02" iyz="0" izz="0.03"/></inertial></link>
    </robot>"""
    m = capi.model_from_urdf(text)
    assert m.n_joints == 1 and m.frame_names == ["a", "payload"]
    Rz = np.array([[0.0, -1, 0], [1, 0, 0], [0, 0, 1]])
    c1, m1, I1 = np.array([0, 0, -0.5]), 2.0, np.diag([0.1, 0.2, 0.3])
    c2, m2, I2 = np.array([0.2, 0, -1.0]) + Rz @ np.array([0.1, 0, 0]), 1.0, Rz @ np.diag([0.01, 0.02, 0.03]) @ Rz.T
    mt = m1 + m2
    c = (m1 * c1 + m2 * c2) / mt
    par = lambda mm, d: mm * (d @ d * np.eye(3) - np.outer(d, d))
    It = I1 + par(m1, c1 - c) + I2 + par(m2, c2 - c)
    assert abs(m.mass[0] - mt) < 1e-15 and np.allclose(m.com[0], c, atol=1e-15)
    assert np.allclose(m.inertia[0], [It[0, 0], It[0, 1], It[0, 2], It[1, 1], It[1, 2], It[2, 2]], atol=1e-15)
    assert np.allclose(m.frame_trans[1], [0.2, 0, -1.0], atol=1e-16)


# the SCARA of models.scara_rrpr() as a URDF: two revolute joints, a prismatic quill (joint frame rolled by pi), a revolute wrist
SCARA = """<robot name="scara">
  <link name="base"/>
  <joint name="j1" type="revolute"><parent link="base"/><child link="arm1"/><origin xyz="0 0 0.4"/><axis xyz="0 0 1"/></joint>
  <link name="arm1"><inertial><mass value="6.0"/><origin xyz="0.2 0 0.03"/><inertia ixx="0.02" ixy="0" ixz="0" iyy="0.12" iyz="0" izz="0.125"/></inertial></link>
  <joint name="j2" type="revolute"><parent link="arm1"/><child link="arm2"/><origin xyz="0.4 0 0.05"/><axis xyz="0 0 1"/></joint>
  <link name="arm2"><inertial><mass value="4.0"/><origin xyz="0.15 0 0.02"/><inertia ixx="0.012" ixy="0" ixz="0" iyy="0.05" iyz="0" izz="0.055"/></inertial></link>
  <joint name="quill" type="prismatic"><parent link="arm2"/><child link="quill_link"/><origin xyz="0.3 0 0" rpy="3.141592653589793 0 0"/><axis xyz="0 0 3"/>
    <limit lower="0" upper="0.3" effort="100" velocity="1"/></joint>
  <link name="quill_link"><inertial><mass value="1.2"/><origin xyz="0 0 -0.1"/><inertia ixx="0.01" ixy="0" ixz="0" iyy="0.01" iyz="0" izz="0.0008"/></inertial></link>
  <joint name="wrist" type="revolute"><parent link="quill_link"/><child link="wrist_link"/><origin xyz="0 0 0.05"/><axis xyz="0 0 1"/></joint>
  <link name="wrist_link"><inertial><mass value="1.0"/><origin xyz="0.01 0 0.03"/><inertia ixx="0.004" ixy="0" ixz="0" iyy="0.004" iyz="0" izz="0.003"/></inertial></link>
  <joint name="flange_fixed" type="fixed"><parent link="wrist_link"/><child link="flange"/><origin xyz="0 0 0.06"/></joint>
  <link name="flange"/>
</robot>"""


def test_prismatic_joint_from_urdf():
    """<joint type="prismatic"> becomes a one-degree-of-freedom sliding joint (as Pinocchio's parser makes it): the URDF of
    the SCARA gives the embedded table, joint types included, and its forward kinematics slides along the rolled axis."""
    m = capi.model_from_urdf(SCARA)
    ref = models.scara_rrpr()
    assert m.n_joints == 4 and list(m.joint_types) == [0, 0, 1, 0] and list(ref.joint_types) == [0, 0, 1, 0]
    for f in ("placement_trans", "axis", "mass", "com", "inertia"):
        assert np.array_equal(getattr(m, f), getattr(ref, f)), f
    assert np.abs(m.placement_rot - ref.placement_rot).max() < 1e-15     # rpy = (pi, 0, 0): sin(pi) is not exactly 0 on either side
    assert m.frame_names == ["arm1", "arm2", "quill_link", "wrist_link", "flange"] and list(m.frame_parent) == [0, 1, 2, 3, 3]
    q = np.array([0.3, -0.7, 0.12, 0.5])
    tip0, tip1 = models.frame_position(m, q, 4), models.frame_position(m, q + np.array([0, 0, 0.05, 0]), 4)
    assert np.allclose(tip1 - tip0, [0, 0, -0.05], atol=1e-15)           # the quill's z axis points down


@pytest.mark.parametrize("text,code,what", [
    ("<robot name='r'><link name='a'/><link name='b'/><joint name='j' type='planar'><parent link='a'/><child link='b'/></joint></robot>", 2, "planar"),
    ("<robot name='r'><link name='a'/><link name='b'/><joint name='j' type='floating'><parent link='a'/><child link='b'/></joint></robot>", 2, "floating"),
    ("<robot name='r'><link name='a'/><link name='b'/><link name='c'/>"
     "<joint name='j1' type='revolute'><parent link='a'/><child link='b'/></joint>"
     "<joint name='j2' type='revolute'><parent link='a'/><child link='c'/></joint></robot>", 2, "serial chains"),
    ("<robot name='r'><link name='a'></robot>", 1, "mismatched"),
    ("<model/>", 1, "expected <robot>"),
    ("<robot name='r'><link name='a'/><joint name='j' type='revolute'><parent link='a'/><child link='zzz'/></joint></robot>", 1, "unknown link"),
    ("<robot name='r'><link name='a'/></robot>", 2, "no movable joint"),
])
def test_rejections(text, code, what):
    L = capi.lib()
    import ctypes as C
    L.ddpb_last_error.restype = C.c_char_p
    h = C.c_void_p()
    data = text.encode()
    assert L.ddpb_model_create_from_urdf(data, C.c_ulonglong(len(data)), C.byref(h)) == code
    assert what in L.ddpb_last_error().decode()


@pytest.mark.gpu
def test_urdf_model_linearizes_like_the_embedded_one():
    """The same arm from URDF text and from the embedded table: identical derivative records (bitwise)."""
    from ddp_pinocchio_b200 import synthetic
    w = synthetic.workload("c3", B=4, T=6)
    g = w.generate()
    outs = []
    for model in (w.model, capi.model_from_urdf(UR5)):
        w.model = model
        bt = capi.Batch.from_workload(w)
        bt.set_targets(g["targets"]); bt.make_trajectory(g["x0"], g["us"]); bt.compute_derivatives()
        outs.append(bt.get_derivatives(["f_val", "fx", "fu", "fxx"]))
        bt.close()
    for f in outs[0]:
        assert np.array_equal(outs[0][f], outs[1][f]), f
