// Counterparts of the reference's arm mains (print-only there), through the C++ facade:
//   test/pinocchio_ddp.cpp             UR5 from URDF, horizon 10, dt = 0.01, c = 1, second_order_finite_diff = false,
//                                      configuration constraint q = neutral at every base index in [2, 12)
//   test/pinocchio_spatial_eq_ddp.cpp  same arm, end-effector (last frame) at (0, 0, 0) at the terminal knot
// both from the neutral configuration at rest with the zero-control generator, solve<primal_dual_affine_multipliers>.
// Differences forced by double precision: mu = 1e4 instead of the MPFR-only 1e20, a fixed number of iterations, and an explicit
// multiplier-Jacobian seed (the reference calls setRandom, ddp.hpp:762).  usage: pinocchio_ddp <urdf> config|spatial <iterations>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "ddpb/ddpb.hpp"
#include "ddpb/layout.hpp"

using namespace ddpb;

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  try {
    model_t model{std::string(argv[1])};
    const bool spatial = std::strcmp(argv[2], "spatial") == 0;
    const int iterations = std::atoi(argv[3]);
    const int nq = int(model.configuration_dim());
    constexpr index_t horizon = 10;
    dynamics_t dy{model, 0.01, false};
    // the constraint's base indexer, as the reference's tests build it; the batch takes the rows per knot of the doubly advanced constraint
    auto eq_idx = spatial ? indexing::range_row_filter(indexing::vec_regular_indexer(2, horizon + 2, 3), horizon, horizon + 1)
                          : indexing::vec_regular_indexer(2, horizon + 2, nq);
    constraint_t con;
    con.kind = spatial ? DDPB_EQ_SPATIAL : DDPB_EQ_CONFIG;
    con.rows = eq_rows_per_knot(horizon, eq_idx);
    con.frame_id = int(model.n_frames()) - 1;
    problem_t prob{0, horizon, 1.0, dy, con};
    ddp_solver_t solver{prob, 1};
    const int ne = solver.ne_total();
    std::vector<double> x_init(2 * nq, 0.0), target(ne, 0.0), jac_seed(static_cast<std::size_t>(ne) * 2 * nq);
    model.neutral_configuration(x_init.data());
    for (std::size_t i = 0; i < jac_seed.size(); ++i) jac_seed[i] = 0.01 * (double((i * 37) % 19) - 9.0);
    solver.set_targets(target.data());
    solver.make_trajectory(x_init.data());
    solver.set_multipliers(nullptr, jac_seed.data());
    const double mu_init = 1e4;
    auto res = solver.solve<method::primal_dual_affine_multipliers>({iterations, 0.0, mu_init, 0.0, 1 / mu_init, 1 / std::pow(mu_init, 0.1)});
    std::vector<double> xs, us;
    solver.trajectory(xs, us);
    mat_seq_t<double> x_seq(indexing::vec_regular_indexer(0, horizon + 1, 2 * nq), xs.data());  // trajectory_t's state sequence of problem 0
    std::printf("ne %d iterations %.0f rollouts %.0f restarts %.0f opt_constr %.6e\nq_T", ne, res.at(DDPB_S_ITERATIONS, 0), res.at(DDPB_S_ROLLOUTS, 0),
                res.at(DDPB_S_RESTARTS, 0), res.at(DDPB_S_OPT_CONSTR, 0));
    for (int i = 0; i < nq; ++i) std::printf(" %.12e", x_seq[horizon](i, 0));
    std::printf("\n");
    return std::isfinite(x_seq[horizon](0, 0)) ? 0 : 3;
  } catch (const error& e) {
    std::printf("%s (code %d)\n", e.what(), e.code);
    return 1;
  }
}
