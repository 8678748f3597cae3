"""The C++ host facade (ddp-pinocchio_b200/include/ddpb/ddpb.hpp) compiles as C++14 against nothing but the C ABI,
and the reference-style pendulum main runs (GPU) or fails loudly (no GPU)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "ddp-pinocchio_b200")


def _build(tmp_path):
    exe = str(tmp_path / "pendulum_ddp")
    subprocess.check_call(["g++", "-std=c++14", "-Wall", "-I", os.path.join(PKG, "include"), os.path.join(ROOT, "tests", "cpp", "pendulum_ddp.cpp"),
                           "-L", PKG, "-lddpb", f"-Wl,-rpath,{PKG}", "-o", exe])
    return exe


def test_facade_compiles_and_fails_loudly_without_gpu(tmp_path, pkg):
    import torch
    exe = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stdout


@pytest.mark.gpu
def test_reference_style_pendulum_main(tmp_path, pkg):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "problem 0: q_T = 3.1" in r.stdout


def test_model_from_urdf_path(tmp_path, pkg):
    """model_t{urdf_path} (pinocchio_model.ipp:155) through the facade: host code only, runs without a GPU."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_urdf import UR5
    exe = str(tmp_path / "urdf_model")
    subprocess.check_call(["g++", "-std=c++14", "-Wall", "-I", os.path.join(PKG, "include"), os.path.join(ROOT, "tests", "cpp", "urdf_model.cpp"),
                           "-L", PKG, "-lddpb", f"-Wl,-rpath,{PKG}", "-o", exe])
    path = tmp_path / "ur5_like.urdf"
    path.write_text(UR5)
    r = subprocess.run([exe, str(path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "nq 6 nv 6 frames 8" in r.stdout and "frame 7 tool0" in r.stdout and "moved nq 6" in r.stdout
    r = subprocess.run([exe, str(tmp_path / "missing.urdf")], capture_output=True, text=True)
    assert r.returncode == 1 and "cannot open" in r.stdout


def _build_model_concept(tmp_path):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_urdf import UR5
    exe = str(tmp_path / "model_concept")
    subprocess.check_call(["g++", "-std=c++14", "-Wall", "-I", os.path.join(PKG, "include"), os.path.join(ROOT, "tests", "cpp", "model_concept.cpp"),
                           "-L", PKG, "-lddpb", f"-Wl,-rpath,{PKG}", "-o", exe])
    path = tmp_path / "ur5_like.urdf"
    path.write_text(UR5)
    return exe, str(path)


def test_model_concept_host_functions_and_loud_failure_without_gpu(tmp_path, pkg):
    """integrate / difference / d_difference_dq_finish are host code (the identities of the reference's test/pinocchio.cpp);
    the dynamics functions need the device."""
    import torch
    exe, urdf = _build_model_concept(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([exe, urdf], capture_output=True, text=True)
    assert "round trip 0.0e+00, d_difference_dq_finish(0,0) 1.0 (1,0) 0.0" in r.stdout
    assert r.returncode == 1 and "cudaSetDevice" in r.stdout


@pytest.mark.gpu
def test_model_concept_main(tmp_path, pkg):
    exe, urdf = _build_model_concept(tmp_path)
    r = subprocess.run([exe, urdf], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "round trip 0.0e+00" in r.stdout
    err = float(r.stdout.split("rnea(aba) - tau ")[1].split(",")[0])
    assert err < 1e-12, r.stdout


def _build_knot_functions(tmp_path):
    exe = str(tmp_path / "knot_functions")
    subprocess.check_call(["g++", "-std=c++14", "-Wall", "-I", os.path.join(PKG, "include"), os.path.join(ROOT, "tests", "cpp", "knot_functions.cpp"),
                           "-L", PKG, "-lddpb", f"-Wl,-rpath,{PKG}", "-o", exe])
    return exe


def test_knot_functions_compile_and_fail_loudly_without_gpu(tmp_path, pkg):
    """dynamics_t::eval_to / first_order_deriv / second_order_deriv, problem_t::eval_f_to / eval_eq_to / l of the facade."""
    import torch
    exe = _build_knot_functions(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stdout


@pytest.mark.gpu
def test_knot_functions_main(tmp_path, pkg):
    """The per-call forms of the Dynamics / Problem concepts on the reference's pendulum against the closed forms (in the C++ main)."""
    r = subprocess.run([_build_knot_functions(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    first = float(r.stdout.split("first order ")[1].split()[0])
    second = float(r.stdout.split("second order ")[1].split()[0])
    eq = float(r.stdout.split("eval_eq_to / eval_f_to ")[1].split(",")[0])
    assert first < 1e-13 and second < 1e-6 and eq < 1e-13, r.stdout   # second order: forward differences of Jacobians, eps = 2^-26
    assert "l 0.125, eq_dim(T-2) 1 eq_dim(0) 0" in r.stdout


def test_reference_layout_kats_in_cpp(tmp_path, pkg):
    """ddpb/layout.hpp (indexers, mat_seq_t, tensor_view_t under the reference's names) passes the known answers of the
    reference's own unit tests, test/indexing.cpp and test/mat_seq.cpp.  Host code: runs without a GPU."""
    exe = str(tmp_path / "layout_kats")
    subprocess.check_call(["g++", "-std=c++14", "-Wall", "-Wextra", "-I", os.path.join(PKG, "include"), os.path.join(ROOT, "tests", "cpp", "layout_kats.cpp"),
                           "-L", PKG, "-lddpb", f"-Wl,-rpath,{PKG}", "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "0 failures" in r.stdout, r.stdout


def _build_pinocchio_ddp(tmp_path):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_urdf import UR5
    exe = str(tmp_path / "pinocchio_ddp")
    subprocess.check_call(["g++", "-std=c++14", "-Wall", "-I", os.path.join(PKG, "include"), os.path.join(ROOT, "tests", "cpp", "pinocchio_ddp.cpp"),
                           "-L", PKG, "-lddpb", f"-Wl,-rpath,{PKG}", "-o", exe])
    path = tmp_path / "ur5_like.urdf"
    path.write_text(UR5)
    return exe, str(path)


def test_reference_style_arm_mains_compile(tmp_path, pkg):
    import torch
    exe, urdf = _build_pinocchio_ddp(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([exe, urdf, "config", "3"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["config", "spatial"])
def test_reference_style_arm_mains(tmp_path, pkg, kind):
    """test/pinocchio_ddp.cpp and test/pinocchio_spatial_eq_ddp.cpp through the facade (UR5-style arm from URDF, horizon 10):
    the final configuration against the oracle's solve of the same scenario."""
    import sys
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    from test_urdf import UR5
    from ddp_pinocchio_b200 import capi, synthetic as S
    exe, urdf = _build_pinocchio_ddp(tmp_path)
    iters = 4
    r = subprocess.run([exe, urdf, kind, str(iters)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    qT = np.array([float(v) for v in r.stdout.split("q_T")[1].split()])
    m = capi.model_from_urdf(UR5)
    T = 10
    ne = np.full(T, 6, dtype=np.int32) if kind == "config" else np.array([0] * 8 + [3, 0], dtype=np.int32)
    o = ol.Oracle(m, T, dt=0.01, c=1.0, hess_mode=0, eq_kind=S.EQ_CONFIG if kind == "config" else S.EQ_SPATIAL, frame=m.n_frames - 1,
                  jac_mode=0, ne_per_knot=ne, targets=np.zeros(int(ne.sum())))
    us = np.zeros((T, 6))
    xs = o.rollout(np.zeros(12), us)
    seed = np.array([0.01 * (((i * 37) % 19) - 9.0) for i in range(int(ne.sum()) * 12)])
    ref = o.solve(xs, us, seed, iters, 1e4)
    assert f"ne {int(ne.sum())} iterations {iters}" in r.stdout
    # (4 iterations of forward-difference Hessians: the emulator is at 6e-12 (config) and 1e-8 (spatial) from the oracle here)
    assert np.abs(qT - ref["xs"][T, :6]).max() < 1e-6, (qT, ref["xs"][T, :6])


def test_reference_liegroup_tests_in_cpp(tmp_path, pkg):
    """ddpb/liegroup.hpp (free-flyer / spherical / planar configuration spaces under the Model concept's names) passes the
    reference's test/pinocchio.cpp -- integrate / difference round trips and the d_difference_dq_finish identity on
    all_joints_test_model() -- and central-difference checks of its four Jacobians.  Host code: runs without a GPU."""
    exe = str(tmp_path / "liegroup_kats")
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-Wall", "-Wextra", "-I", os.path.join(PKG, "include"),
                           os.path.join(ROOT, "tests", "cpp", "liegroup_kats.cpp"), "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "0 failures" in r.stdout, r.stdout[-2000:]
