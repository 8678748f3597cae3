// liegroup.hpp -- configuration-space operations of the Model concept for joints whose configuration is NOT a vector
// space: free-flyer SE(3), spherical SO(3), planar SE(2), unbounded revolute SO(2) (and the vector-space joints for completeness).
//
// The reference forwards these to Pinocchio (/root/reference/include/ddp/pinocchio_model.ipp:199-321:
//   neutral -> pinocchio::neutral, integrate -> pinocchio::integrate, d_integrate_dq / dv -> dIntegrate(ARG0 / ARG1),
//   difference -> pinocchio::difference, d_difference_dq_start / finish -> dDifference(ARG0 / ARG1))
// and tests them on `all_joints_test_model()` = free-flyer + spherical + planar (pinocchio_model.ipp:68-90,
// test/pinocchio.cpp:17-100).  Pinocchio is not available offline; this header restates the published Lie-group formulas with
// Pinocchio's conventions -- quaternions stored (x, y, z, w), free-flyer q = (p, quat), tangent (v, omega) in the LOCAL
// frame, q (+) v = q * exp(v), q1 (-) q0 = log(q0^-1 q1), planar q = (x, y, cos, sin) -- and tests/cpp/liegroup_kats.cpp
// repeats the reference's two tests on it plus finite-difference checks of all four Jacobians.
// Host code only: the device kernels of this repo support joints with nq = nv = 1; this is the
// configuration-space layer of SURVEY.md section 8(f)3, not yet a dynamics for such joints.
#pragma once
#include <cassert>
#include <cmath>
#include <vector>

namespace ddpb {
namespace liegroup {

// ---- 3-vectors / 3x3 (row-major) helpers -------------------------------------------------------
struct mat3 { double m[9]; };
inline mat3 identity3() { return mat3{{1, 0, 0, 0, 1, 0, 0, 0, 1}}; }
inline mat3 skew(const double* w) { return mat3{{0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0}}; }
inline mat3 mul(const mat3& a, const mat3& b) {
  mat3 c;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) c.m[i * 3 + j] = a.m[i * 3] * b.m[j] + a.m[i * 3 + 1] * b.m[3 + j] + a.m[i * 3 + 2] * b.m[6 + j];
  return c;
}
inline mat3 transpose(const mat3& a) { return mat3{{a.m[0], a.m[3], a.m[6], a.m[1], a.m[4], a.m[7], a.m[2], a.m[5], a.m[8]}}; }
inline mat3 lin(double a, const mat3& A, double b, const mat3& B, double c, const mat3& C) {
  mat3 r;
  for (int k = 0; k < 9; ++k) r.m[k] = a * A.m[k] + b * B.m[k] + c * C.m[k];
  return r;
}
inline void mulv(const mat3& a, const double* v, double* o) {
  const double x = a.m[0] * v[0] + a.m[1] * v[1] + a.m[2] * v[2], y = a.m[3] * v[0] + a.m[4] * v[1] + a.m[5] * v[2],
               z = a.m[6] * v[0] + a.m[7] * v[1] + a.m[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}

// ---- SO(3) ------------------------------------------------------------------------------------------
// coefficients of the exponential series, with their Taylor expansions near 0
struct so3_coeffs {
  double a;  // sin(t)/t
  double b;  // (1 - cos t)/t^2
  double c;  // (t - sin t)/t^3
  double d;  // 1/t^2 - (1 + cos t)/(2 t sin t)
};
inline so3_coeffs coeffs(double t) {
  const double t2 = t * t;
  if (t < 1e-3) return so3_coeffs{1 - t2 / 6 + t2 * t2 / 120, 0.5 - t2 / 24 + t2 * t2 / 720, 1.0 / 6 - t2 / 120 + t2 * t2 / 5040, 1.0 / 12 + t2 / 720 + t2 * t2 / 30240};
  const double s = std::sin(t), c = std::cos(t);
  return so3_coeffs{s / t, (1 - c) / t2, (t - s) / (t2 * t), 1 / t2 - (1 + c) / (2 * t * s)};
}
inline double norm3(const double* w) { return std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]); }
inline mat3 exp3(const double* w) {  // Rodrigues
  const so3_coeffs k = coeffs(norm3(w));
  const mat3 W = skew(w);
  return lin(1.0, identity3(), k.a, W, k.b, mul(W, W));
}
inline mat3 Jl3(const double* w) {  // left Jacobian of SO(3) (= V of the SE(3) exponential); the right one is Jl3(-w)
  const so3_coeffs k = coeffs(norm3(w));
  const mat3 W = skew(w);
  return lin(1.0, identity3(), k.b, W, k.c, mul(W, W));
}
inline mat3 Jl3_inv(const double* w) {
  const so3_coeffs k = coeffs(norm3(w));
  const mat3 W = skew(w);
  return lin(1.0, identity3(), -0.5, W, k.d, mul(W, W));
}
// unit quaternion (x, y, z, w) <-> rotation
inline void quat_mul(const double* a, const double* b, double* o) {
  const double x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1], y = a[3] * b[1] - a[0] * b[2] + a[1] * b[3] + a[2] * b[0],
               z = a[3] * b[2] + a[0] * b[1] - a[1] * b[0] + a[2] * b[3], w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  o[0] = x; o[1] = y; o[2] = z; o[3] = w;
}
inline void quat_normalize(double* q) {
  const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int k = 0; k < 4; ++k) q[k] /= n;
}
inline void quat_exp(const double* w, double* q) {
  const double t = norm3(w), h = 0.5 * t;
  const double s = h < 1e-4 ? 0.5 - t * t / 48 : std::sin(h) / t;
  q[0] = s * w[0]; q[1] = s * w[1]; q[2] = s * w[2]; q[3] = std::cos(h);
}
inline void quat_log(const double* qin, double* w) {
  double q[4] = {qin[0], qin[1], qin[2], qin[3]};
  if (q[3] < 0) for (int k = 0; k < 4; ++k) q[k] = -q[k];  // the shorter of the two rotations
  const double n = norm3(q);
  const double f = n < 1e-8 ? 2.0 / q[3] : 2.0 * std::atan2(n, q[3]) / n;
  w[0] = f * q[0]; w[1] = f * q[1]; w[2] = f * q[2];
}
inline mat3 quat_rot(const double* q) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  return mat3{{1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
               2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)}};
}

// ---- SE(3), tangent (v, omega) ---------------------------------------------------------------------
struct mat6 { double m[36]; };  // row-major
inline mat6 blocks(const mat3& A, const mat3& B, const mat3& D) {  // [[A, B], [0, D]]
  mat6 r{};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) { r.m[i * 6 + j] = A.m[i * 3 + j]; r.m[i * 6 + 3 + j] = B.m[i * 3 + j]; r.m[(3 + i) * 6 + 3 + j] = D.m[i * 3 + j]; }
  return r;
}
inline mat3 scaled(double a, const mat3& A) { return lin(a, A, 0, A, 0, A); }
// the coupling block Q(rho, phi) of the left Jacobian of SE(3) (Barfoot, State Estimation for Robotics, eq. 7.86)
inline mat3 Q_block(const double* rho, const double* phi) {
  const double t = norm3(phi), t2 = t * t;
  double c1, c2, c3;
  if (t < 1e-2) {
    c1 = 1.0 / 6 - t2 / 120 + t2 * t2 / 5040;
    c2 = 1.0 / 24 - t2 / 720 + t2 * t2 / 40320;
    c3 = 1.0 / 120 - t2 / 2520 + t2 * t2 / 120960;
  } else {
    const double s = std::sin(t), c = std::cos(t), t4 = t2 * t2;
    c1 = (t - s) / (t2 * t);
    c2 = -(1 - t2 / 2 - c) / t4;
    c3 = -0.5 * ((1 - t2 / 2 - c) / t4 - 3 * (t - s - t2 * t / 6) / (t4 * t));
  }
  const mat3 R = skew(rho), P = skew(phi);
  const mat3 PR = mul(P, R), RP = mul(R, P), PRP = mul(PR, P), PPR = mul(P, PR), RPP = mul(RP, P);
  mat3 q = scaled(0.5, R);
  const mat3 t1 = lin(1, PR, 1, RP, 1, PRP), t2m = lin(1, PPR, 1, RPP, -3, PRP), t3 = lin(1, mul(PRP, P), 1, mul(P, PRP), 0, P);
  for (int k = 0; k < 9; ++k) q.m[k] += c1 * t1.m[k] + c2 * t2m.m[k] + c3 * t3.m[k];
  return q;
}
inline mat6 Jl6(const double* xi) { const mat3 J = Jl3(xi + 3); return blocks(J, Q_block(xi, xi + 3), J); }
inline mat6 Jl6_inv(const double* xi) {
  const mat3 Ji = Jl3_inv(xi + 3);
  return blocks(Ji, scaled(-1.0, mul(mul(Ji, Q_block(xi, xi + 3)), Ji)), Ji);
}
inline void neg6(const double* xi, double* o) { for (int k = 0; k < 6; ++k) o[k] = -xi[k]; }
// exp6: rotation R and translation t of exp((v, omega))
inline void exp6(const double* xi, mat3* R, double* t) { *R = exp3(xi + 3); mulv(Jl3(xi + 3), xi, t); }
// Ad(T)^-1 for T = (R, t) on (v, omega): [[R^T, -R^T [t]x], [0, R^T]]
inline mat6 Ad_inv(const mat3& R, const double* t) {
  const mat3 Rt = transpose(R);
  return blocks(Rt, scaled(-1.0, mul(Rt, skew(t))), Rt);
}

// ---- joints -------------------------------------------------------------------------------------------
// revolute_unbounded: what Pinocchio's URDF parser makes of a `continuous` joint -- q = (cos, sin), one tangent coordinate
enum class joint_kind { revolute, prismatic, revolute_unbounded, spherical, planar, free_flyer };
inline int nq_of(joint_kind k) { return k == joint_kind::free_flyer ? 7 : (k == joint_kind::spherical || k == joint_kind::planar ? 4 : (k == joint_kind::revolute_unbounded ? 2 : 1)); }
inline int nv_of(joint_kind k) { return k == joint_kind::free_flyer ? 6 : (k == joint_kind::spherical || k == joint_kind::planar ? 3 : 1); }

inline void neutral(joint_kind k, double* q) {
  for (int i = 0; i < nq_of(k); ++i) q[i] = 0.0;
  if (k == joint_kind::free_flyer) q[6] = 1.0;
  if (k == joint_kind::spherical) q[3] = 1.0;
  if (k == joint_kind::planar) q[2] = 1.0;
  if (k == joint_kind::revolute_unbounded) q[0] = 1.0;
}
inline void integrate(joint_kind k, const double* q, const double* v, double* out) {
  switch (k) {
    case joint_kind::revolute:
    case joint_kind::prismatic: out[0] = q[0] + v[0]; return;
    case joint_kind::revolute_unbounded: {
      const double cw = std::cos(v[0]), sw = std::sin(v[0]), c = q[0] * cw - q[1] * sw, s = q[1] * cw + q[0] * sw, n = std::sqrt(c * c + s * s);
      out[0] = c / n; out[1] = s / n;
      return;
    }
    case joint_kind::spherical: {
      double e[4];
      quat_exp(v, e);
      quat_mul(q, e, out);
      quat_normalize(out);
      return;
    }
    case joint_kind::free_flyer: {
      mat3 R;
      double t[3], Rt[3], e[4];
      exp6(v, &R, t);
      mulv(quat_rot(q + 3), t, Rt);
      for (int i = 0; i < 3; ++i) out[i] = q[i] + Rt[i];
      quat_exp(v + 3, e);
      quat_mul(q + 3, e, out + 3);
      quat_normalize(out + 3);
      return;
    }
    case joint_kind::planar: {
      const double w = v[2], a = coeffs(std::fabs(w)).a, b = coeffs(std::fabs(w)).b * w;  // sin w / w, (1 - cos w) / w
      const double tx = a * v[0] - b * v[1], ty = b * v[0] + a * v[1];
      const double c0 = q[2], s0 = q[3], cw = std::cos(w), sw = std::sin(w);
      out[0] = q[0] + c0 * tx - s0 * ty;
      out[1] = q[1] + s0 * tx + c0 * ty;
      const double c = c0 * cw - s0 * sw, s = s0 * cw + c0 * sw, n = std::sqrt(c * c + s * s);
      out[2] = c / n; out[3] = s / n;
      return;
    }
  }
}
inline void difference(joint_kind k, const double* q0, const double* q1, double* out) {
  switch (k) {
    case joint_kind::revolute:
    case joint_kind::prismatic: out[0] = q1[0] - q0[0]; return;
    case joint_kind::revolute_unbounded: out[0] = std::atan2(q0[0] * q1[1] - q0[1] * q1[0], q0[0] * q1[0] + q0[1] * q1[1]); return;
    case joint_kind::spherical: {
      const double c[4] = {-q0[0], -q0[1], -q0[2], q0[3]};
      double r[4];
      quat_mul(c, q1, r);
      quat_log(r, out);
      return;
    }
    case joint_kind::free_flyer: {
      const double c[4] = {-q0[3], -q0[4], -q0[5], q0[6]};
      double r[4], d[3] = {q1[0] - q0[0], q1[1] - q0[1], q1[2] - q0[2]}, t[3];
      quat_mul(c, q1 + 3, r);
      quat_log(r, out + 3);
      mulv(transpose(quat_rot(q0 + 3)), d, t);
      mulv(Jl3_inv(out + 3), t, out);
      return;
    }
    case joint_kind::planar: {
      const double c0 = q0[2], s0 = q0[3], cd = c0 * q1[2] + s0 * q1[3], sd = c0 * q1[3] - s0 * q1[2];
      const double w = std::atan2(sd, cd);
      const double dx = q1[0] - q0[0], dy = q1[1] - q0[1], tx = c0 * dx + s0 * dy, ty = -s0 * dx + c0 * dy;
      const double a = coeffs(std::fabs(w)).a, b = coeffs(std::fabs(w)).b * w, n = a * a + b * b;
      out[0] = (a * tx + b * ty) / n;
      out[1] = (-b * tx + a * ty) / n;
      out[2] = w;
      return;
    }
  }
}
// the planar joint as the subgroup of SE(3) that moves in the x-y plane: tangent (vx, vy, w) -> (vx, vy, 0, 0, 0, w)
inline void planar_embed(const double* v, double* xi) { xi[0] = v[0]; xi[1] = v[1]; xi[2] = 0; xi[3] = 0; xi[4] = 0; xi[5] = v[2]; }
inline void planar_restrict(const mat6& J, double* out /*3x3 col-major*/) {
  const int id[3] = {0, 1, 5};
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) out[c * 3 + r] = J.m[id[r] * 6 + id[c]];
}
enum class jac_kind { integrate_dq, integrate_dv, difference_dq_start, difference_dq_finish };
// Jacobians in the tangent space, nv x nv column-major.  a, b: (q, v) for the integrate kinds, (q_start, q_finish) for the difference kinds.
//   d (q (+) v) / dq = Ad(exp(v))^-1 ,  d (q (+) v) / dv = Jr(v) = Jl(-v) ,
//   d (q1 (-) q0) / dq1 = Jr^-1(q1 (-) q0) ,  d (q1 (-) q0) / dq0 = -Jl^-1(q1 (-) q0)
inline void jacobian(joint_kind k, jac_kind what, const double* a, const double* b, double* out) {
  const int nv = nv_of(k);
  if (nv == 1) { out[0] = what == jac_kind::difference_dq_start ? -1.0 : 1.0; return; }  // commutative groups
  double v[6], xi[6], mxi[6];
  if (what == jac_kind::integrate_dq || what == jac_kind::integrate_dv) for (int i = 0; i < nv; ++i) v[i] = b[i];
  else difference(k, a, b, v);
  if (k == joint_kind::spherical) {
    const double mv[3] = {-v[0], -v[1], -v[2]};
    mat3 J;
    switch (what) {
      case jac_kind::integrate_dq: J = transpose(exp3(v)); break;
      case jac_kind::integrate_dv: J = Jl3(mv); break;
      case jac_kind::difference_dq_finish: J = Jl3_inv(mv); break;
      default: J = scaled(-1.0, Jl3_inv(v)); break;
    }
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) out[c * 3 + r] = J.m[r * 3 + c];
    return;
  }
  if (k == joint_kind::planar) planar_embed(v, xi);
  else for (int i = 0; i < 6; ++i) xi[i] = v[i];
  neg6(xi, mxi);
  mat6 J;
  switch (what) {
    case jac_kind::integrate_dq: { mat3 R; double t[3]; exp6(xi, &R, t); J = Ad_inv(R, t); break; }
    case jac_kind::integrate_dv: J = Jl6(mxi); break;
    case jac_kind::difference_dq_finish: J = Jl6_inv(mxi); break;
    default: { J = Jl6_inv(xi); for (double& x : J.m) x = -x; break; }
  }
  if (k == joint_kind::planar) { planar_restrict(J, out); return; }
  for (int r = 0; r < 6; ++r)
    for (int c = 0; c < 6; ++c) out[c * 6 + r] = J.m[r * 6 + c];
}

// ---- a configuration space = a list of joints; the member names of the reference's Model concept --------------------
class config_space_t {
 public:
  explicit config_space_t(std::vector<joint_kind> joints) : m_joints(std::move(joints)) {
    for (joint_kind k : m_joints) { m_qoff.push_back(m_nq); m_voff.push_back(m_nv); m_nq += nq_of(k); m_nv += nv_of(k); }
  }
  // model_t::all_joints_test_model(): free-flyer, spherical, planar in a chain (pinocchio_model.ipp:68-90)
  static config_space_t all_joints_test_model() { return config_space_t({joint_kind::free_flyer, joint_kind::spherical, joint_kind::planar}); }
  long configuration_dim() const { return m_nq; }
  long tangent_dim() const { return m_nv; }
  void neutral_configuration(double* out_q) const { each([&](joint_kind k, int qo, int) { neutral(k, out_q + qo); }); }
  // a configuration from `unit` in [0, 1)^nq (the caller's random numbers): positions in [-1, 1], uniform rotations
  void random_configuration(double* out_q, const double* unit) const {
    each([&](joint_kind k, int qo, int) {
      double* q = out_q + qo;
      const double* r = unit + qo;
      auto rot = [](const double* r4, double* quat) { for (int i = 0; i < 4; ++i) quat[i] = 2 * r4[i] - 1 + (i == 3 ? 1e-3 : 0); quat_normalize(quat); };
      switch (k) {
        case joint_kind::revolute:
        case joint_kind::prismatic: q[0] = 2 * r[0] - 1; break;
        case joint_kind::revolute_unbounded: { const double th = 6.283185307179586 * r[0]; q[0] = std::cos(th); q[1] = std::sin(th); break; }
        case joint_kind::spherical: rot(r, q); break;
        case joint_kind::free_flyer: for (int i = 0; i < 3; ++i) q[i] = 2 * r[i] - 1; rot(r + 3, q + 3); break;
        case joint_kind::planar: { const double th = 6.283185307179586 * r[2]; q[0] = 2 * r[0] - 1; q[1] = 2 * r[1] - 1; q[2] = std::cos(th); q[3] = std::sin(th); break; }
      }
    });
  }
  void integrate(double* out_q, const double* q, const double* v) const {
    each([&](joint_kind k, int qo, int vo) { liegroup::integrate(k, q + qo, v + vo, out_q + qo); });
  }
  void difference(double* out_v, const double* q_start, const double* q_finish) const {
    each([&](joint_kind k, int qo, int vo) { liegroup::difference(k, q_start + qo, q_finish + qo, out_v + vo); });
  }
  // nv x nv column-major, block diagonal over the joints
  void d_integrate_dq(double* out, const double* q, const double* v) const { jac(jac_kind::integrate_dq, out, q, v, true); }
  void d_integrate_dv(double* out, const double* q, const double* v) const { jac(jac_kind::integrate_dv, out, q, v, true); }
  void d_difference_dq_start(double* out, const double* q_start, const double* q_finish) const { jac(jac_kind::difference_dq_start, out, q_start, q_finish, false); }
  void d_difference_dq_finish(double* out, const double* q_start, const double* q_finish) const { jac(jac_kind::difference_dq_finish, out, q_start, q_finish, false); }

 private:
  template <class F> void each(F&& f) const { for (std::size_t j = 0; j < m_joints.size(); ++j) f(m_joints[j], m_qoff[j], m_voff[j]); }
  void jac(jac_kind what, double* out, const double* a, const double* b, bool b_is_tangent) const {
    for (long i = 0; i < (long)m_nv * m_nv; ++i) out[i] = 0.0;
    each([&](joint_kind k, int qo, int vo) {
      double blk[36];
      jacobian(k, what, a + qo, b + (b_is_tangent ? vo : qo), blk);
      const int n = nv_of(k);
      for (int c = 0; c < n; ++c)
        for (int r = 0; r < n; ++r) out[(std::size_t)(vo + c) * m_nv + vo + r] = blk[c * n + r];
    });
  }
  std::vector<joint_kind> m_joints;
  std::vector<int> m_qoff, m_voff;
  int m_nq = 0, m_nv = 0;
};

}  // namespace liegroup
}  // namespace ddpb
