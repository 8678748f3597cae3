// The reference's configuration-space tests (/root/reference/test/pinocchio.cpp) on ddpb/liegroup.hpp, in double:
//   :17-57   "integrate-difference": difference(q0, integrate(q0, v)) == v and integrating back returns q0, on
//            all_joints_test_model() = free-flyer + spherical + planar
//   :59-100  "affine-function": d_difference_dq_finish against a finite step (the re-centring identity of update_origin)
// plus central-difference checks of all four Jacobians, per joint kind, at generic and at small tangent vectors.
// The reference runs them at 1000 digits with dh = 1e-50; here dh = 1e-6 and the bounds are those of double arithmetic.
#include <cmath>
#include <cstdio>
#include <vector>

#include "ddpb/liegroup.hpp"

using namespace ddpb::liegroup;

static int failures = 0;
#define CHECK(c) do { if (!(c)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); ++failures; } } while (0)

static unsigned long long rng_state = 0x9E3779B97F4A7C15ull;
static double unit() {  // splitmix64 -> [0, 1)
  unsigned long long z = (rng_state += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
  return double(z >> 11) / 9007199254740992.0;
}
static double maxdiff(const std::vector<double>& a, const std::vector<double>& b) {
  double e = 0;
  for (std::size_t i = 0; i < a.size(); ++i) e = std::fmax(e, std::fabs(a[i] - b[i]));
  return e;
}
static std::vector<double> random_q(const config_space_t& m) {
  std::vector<double> u(m.configuration_dim()), q(m.configuration_dim());
  for (double& x : u) x = unit();
  m.random_configuration(q.data(), u.data());
  return q;
}

// all four Jacobians of `m` against central differences in the tangent space
static void check_jacobians(const config_space_t& m, double vscale, double tol) {
  const long nq = m.configuration_dim(), nv = m.tangent_dim();
  const double h = 1e-6;
  std::vector<double> q0 = random_q(m), v(nv), q1(nq), J(nv * nv), e(nv), qa(nq), qb(nq), ra(nq), rb(nq), da(nv), db(nv);
  for (double& x : v) x = vscale * (2 * unit() - 1);
  m.integrate(q1.data(), q0.data(), v.data());
  for (int which = 0; which < 4; ++which) {
    if (which == 0) m.d_integrate_dq(J.data(), q0.data(), v.data());
    if (which == 1) m.d_integrate_dv(J.data(), q0.data(), v.data());
    if (which == 2) m.d_difference_dq_start(J.data(), q0.data(), q1.data());
    if (which == 3) m.d_difference_dq_finish(J.data(), q0.data(), q1.data());
    double worst = 0;
    for (long c = 0; c < nv; ++c) {
      for (long i = 0; i < nv; ++i) e[i] = i == c ? h : 0.0;
      std::vector<double> ep = e, em = e;
      for (double& x : em) x = -x;
      if (which < 2) {
        // integrate(q (+) e, v) resp. integrate(q, v + e), measured from the unperturbed result
        if (which == 0) { m.integrate(qa.data(), q0.data(), ep.data()); m.integrate(qb.data(), q0.data(), em.data()); m.integrate(ra.data(), qa.data(), v.data()); m.integrate(rb.data(), qb.data(), v.data()); }
        else { std::vector<double> vp = v, vm = v; vp[c] += h; vm[c] -= h; m.integrate(ra.data(), q0.data(), vp.data()); m.integrate(rb.data(), q0.data(), vm.data()); }
        m.difference(da.data(), q1.data(), ra.data());
        m.difference(db.data(), q1.data(), rb.data());
      } else {
        const std::vector<double>& base = which == 2 ? q0 : q1;
        m.integrate(qa.data(), base.data(), ep.data());
        m.integrate(qb.data(), base.data(), em.data());
        if (which == 2) { m.difference(da.data(), qa.data(), q1.data()); m.difference(db.data(), qb.data(), q1.data()); }
        else { m.difference(da.data(), q0.data(), qa.data()); m.difference(db.data(), q0.data(), qb.data()); }
      }
      for (long r = 0; r < nv; ++r) worst = std::fmax(worst, std::fabs((da[r] - db[r]) / (2 * h) - J[c * nv + r]));
    }
    if (!(worst < tol)) std::printf("jacobian %d: worst %.3e (vscale %.1e)\n", which, worst, vscale);
    CHECK(worst < tol);
  }
}

int main() {
  auto model = config_space_t::all_joints_test_model();
  const long nq = model.configuration_dim(), nv = model.tangent_dim();
  CHECK(nq == 15 && nv == 12);
  const double eps = 2.220446049250313e-16;
  for (int rep = 0; rep < 20; ++rep) {  // "integrate-difference"
    std::vector<double> q0 = random_q(model), q2 = q0, q1(nq), v0(nv), v1(nv);
    for (double& x : v0) x = 2 * unit() - 1;
    model.integrate(q1.data(), q0.data(), v0.data());
    model.difference(v1.data(), q0.data(), q1.data());
    CHECK(maxdiff(v0, v1) < eps * 1e3);
    for (double& x : v0) x = -x;
    model.integrate(q0.data(), q1.data(), v0.data());
    CHECK(maxdiff(q0, q2) < eps * 1e3);
  }
  for (int rep = 0; rep < 20; ++rep) {  // "affine-function"
    std::vector<double> q0 = random_q(model), q1 = random_q(model), v1(nv), v2(nv), J(nv * nv), dv(nv), q2(nq), j0(nv);
    const double f0 = 2 * unit() - 1, dh = 1e-6;
    for (double& x : j0) x = 2 * unit() - 1;
    model.difference(v1.data(), q0.data(), q1.data());
    model.d_difference_dq_finish(J.data(), q0.data(), q1.data());
    for (double& x : dv) x = (2 * unit() - 1) * dh;
    model.integrate(q2.data(), q1.data(), dv.data());
    model.difference(v2.data(), q0.data(), q2.data());
    double f1 = f0, lhs = f0, rhs, err = 0;
    for (long i = 0; i < nv; ++i) { f1 += j0[i] * v1[i]; lhs += j0[i] * v2[i]; }
    rhs = f1;
    for (long c = 0; c < nv; ++c) {
      double j1c = 0;
      for (long r = 0; r < nv; ++r) j1c += j0[r] * J[c * nv + r];
      rhs += j1c * dv[c];
    }
    for (long r = 0; r < nv; ++r) {
      double Jdv = 0;
      for (long c = 0; c < nv; ++c) Jdv += J[c * nv + r] * dv[c];
      err = std::fmax(err, std::fabs(v2[r] - v1[r] - Jdv));
    }
    const double eps2 = dh * dh * 1e2 + 1e-13;  // second-order remainder of a step of size dh, rotations up to pi
    CHECK(std::fabs(lhs - rhs) < eps2 * nv);
    CHECK(err < eps2);
  }
  {  // neutral configuration: identity placements, zero difference to itself
    std::vector<double> q(nq), v(nv);
    model.neutral_configuration(q.data());
    CHECK(q[6] == 1.0 && q[10] == 1.0 && q[13] == 1.0 && q[14] == 0.0);
    model.difference(v.data(), q.data(), q.data());
    CHECK(maxdiff(v, std::vector<double>(nv, 0.0)) == 0.0);
  }
  // Jacobians per joint kind and for the composite model; small tangents exercise the series branches
  for (joint_kind k : {joint_kind::revolute, joint_kind::prismatic, joint_kind::revolute_unbounded, joint_kind::spherical, joint_kind::planar, joint_kind::free_flyer}) {
    config_space_t one({k});
    for (int rep = 0; rep < 5; ++rep) { check_jacobians(one, 1.0, 2e-8); check_jacobians(one, 1e-4, 2e-8); check_jacobians(one, 1e-9, 2e-8); }
  }
  for (int rep = 0; rep < 5; ++rep) check_jacobians(model, 0.7, 2e-8);
  {  // a chain as Pinocchio's URDF parser builds it when a joint is `continuous`: round trip through the (cos, sin) parametrisation
    config_space_t arm({joint_kind::revolute, joint_kind::revolute_unbounded, joint_kind::prismatic});
    CHECK(arm.configuration_dim() == 4 && arm.tangent_dim() == 3);
    std::vector<double> q0 = random_q(arm), q1(4), v = {0.3, 2.5, -0.7}, d(3);
    arm.integrate(q1.data(), q0.data(), v.data());
    arm.difference(d.data(), q0.data(), q1.data());
    CHECK(maxdiff(v, d) < 1e-14);
    CHECK(std::fabs(q1[1] * q1[1] + q1[2] * q1[2] - 1.0) < 1e-15);
  }
  std::printf("%d failures\n", failures);
  return failures != 0;
}
