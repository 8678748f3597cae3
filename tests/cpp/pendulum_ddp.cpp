// Batched counterpart of the reference's test/pendulum_ddp.cpp (print-only main there): same model
// (pendulum_model_t{1.0, 1.0}), horizon 200, dt = 0.01, c = 1.0, terminal configuration constraint q_T = 3.14,
// zero-control initial trajectory from x = 0, solve<primal_dual_affine_multipliers>.  Differences forced by
// double precision: mu = 1e4 instead of the MPFR-only 1e20, 15 iterations, and the multiplier Jacobian seed
// is explicit (the reference calls setRandom, ddp.hpp:762).  Problem 1 of the batch uses a different seed.
#include <cmath>
#include <cstdio>

#include "ddpb/ddpb.hpp"

using namespace ddpb;

auto main() -> int {
  auto model = model_t::pendulum(1.0, 1.0);
  constexpr index_t horizon = 200;
  dynamics_t dy{model, 0.01};  // second_order_finite_diff defaults to true, as in the reference test
  problem_t prob{0, horizon, 1.0, dy, constraint_t::config_terminal(horizon, 1)};
  const int B = 2;
  try {
    ddp_solver_t solver{prob, B};
    std::vector<double> x_init(B * 2, 0.0), target(B, 3.14), jac_seed = {0.25, -0.5, -0.125, 0.75};
    solver.set_targets(target.data());
    solver.make_trajectory(x_init.data());           // zero-control generator
    solver.set_multipliers(nullptr, jac_seed.data());
    const double mu_init = 1e4;
    auto res = solver.solve<method::primal_dual_affine_multipliers>({15, 0.0, mu_init, 0.0, 1 / mu_init, 1 / std::pow(mu_init, 0.1)});
    std::vector<double> xs, us;
    solver.trajectory(xs, us);
    for (int b = 0; b < B; ++b)
      std::printf("problem %d: q_T = %.12f v_T = %.12f opt_constr = %.3e iterations = %.0f rollouts = %.0f\n", b, xs[(b * (horizon + 1) + horizon) * 2],
                  xs[(b * (horizon + 1) + horizon) * 2 + 1], res.at(DDPB_S_OPT_CONSTR, b), res.at(DDPB_S_ITERATIONS, b), res.at(DDPB_S_ROLLOUTS, b));
    const double qT = xs[horizon * 2];
    return std::fabs(qT - 3.14) < 0.1 ? 0 : 2;
  } catch (const error& e) {
    std::printf("%s\n", e.what());
    return 1;
  }
}
