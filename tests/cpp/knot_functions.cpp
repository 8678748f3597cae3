// The per-call forms of the Dynamics / Problem concepts through the C++ facade (ddpb.hpp) on the reference's pendulum
// (/root/reference/test/pendulum_ddp.cpp:30-66: m = l = 1, dt = 0.01), checked in place against the closed forms of
// pendulum_model.hpp:105-130 and the explicit Euler scheme of problem.hpp:441-503:
//   f(x, u) = [q + dt v ; v + dt (-g/l sin q + u/m)],  f_x = [[1, dt], [-dt g/l cos q, 1]],  f_u = [0 ; dt/m],
//   f_xx(v-row, q, q) = dt g/l sin q,   eq(t, x, u) = q_{t+2} - target with q_{t+2} = q_{t+1} + dt v_{t+1}.
#include <cmath>
#include <cstdio>
#include <vector>

#include "ddpb/ddpb.hpp"

int main() {
  try {
    auto model = ddpb::model_t::pendulum(1.0, 1.0);
    const double dt = 0.01, g = 9.81;
    ddpb::dynamics_t dy{model, dt, false};
    const int n = 3;
    const double x[n * 2] = {0.3, -0.2, 1.1, 0.7, -2.0, 0.05}, u[n] = {0.5, -1.0, 2.0};
    double f[n * 2], fx[n * 4], fu[n * 2], fxx[n * 8], fux[n * 4], fuu[n * 2];
    dy.second_order_deriv(fxx, fux, fuu, fx, fu, f, 0, x, u, n);
    double e1 = 0.0, e2 = 0.0;
    for (int s = 0; s < n; ++s) {
      const double q = x[2 * s], v = x[2 * s + 1];
      const double a = -g * std::sin(q) + u[s];
      const double fr[2] = {q + dt * v, v + dt * a}, fxr[4] = {1.0, -dt * g * std::cos(q), dt, 1.0}, fur[2] = {0.0, dt};
      for (int k = 0; k < 2; ++k) e1 = std::fmax(e1, std::fabs(f[2 * s + k] - fr[k]) + std::fabs(fu[2 * s + k] - fur[k]));
      for (int k = 0; k < 4; ++k) e1 = std::fmax(e1, std::fabs(fx[4 * s + k] - fxr[k]));
      // T(k, i, j) at k + 2 i + 4 j: the only non-zero second derivative is (k = 1, i = 0, j = 0)
      for (int k = 0; k < 8; ++k) e2 = std::fmax(e2, std::fabs(fxx[8 * s + k] - (k == 1 ? dt * g * std::sin(q) : 0.0)));
      for (int k = 0; k < 4; ++k) e2 = std::fmax(e2, std::fabs(fux[4 * s + k]));
      for (int k = 0; k < 2; ++k) e2 = std::fmax(e2, std::fabs(fuu[2 * s + k]));
    }
    std::printf("first order %.1e second order %.1e\n", e1, e2);
    // the reference's problem: terminal configuration constraint q_T = 3.14 (test/pendulum_ddp.cpp:37-47)
    const ddpb::index_t T = 200;
    ddpb::problem_t prob{0, T, 1.0, dy, ddpb::constraint_t::config_terminal(T, 1)};
    const double target[n] = {3.14, 3.14, 3.14};
    double eq[n], f1[n * 2];
    prob.eval_eq_to(eq, T - 2, x, u, target, n);
    prob.eval_f_to(f1, 0, x, u, n);
    double e3 = 0.0;
    for (int s = 0; s < n; ++s) {
      const double q = x[2 * s], v = x[2 * s + 1];
      const double v1 = v + dt * (-g * std::sin(q) + u[s]), q1 = q + dt * v;
      e3 = std::fmax(e3, std::fabs(eq[s] - ((q1 + dt * v1) - 3.14)) + std::fabs(f1[2 * s] - q1) + std::fabs(f1[2 * s + 1] - v1));
    }
    std::printf("eval_eq_to / eval_f_to %.1e, l %.3f, eq_dim(T-2) %d eq_dim(0) %d\n", e3, prob.l(0, x, u), prob.m_constraint.eq_dim(T - 2),
                prob.m_constraint.eq_dim(0));
  } catch (ddpb::error const& e) {
    std::printf("%s (code %d)\n", e.what(), e.code);
    return 1;
  }
  return 0;
}
