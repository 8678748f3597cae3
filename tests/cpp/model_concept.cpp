// The Model concept through the C++ facade (ddpb.hpp): the functions the reference's dynamics_t / constraints call on
// pinocchio::model_t (/root/reference/include/ddp/pinocchio_model.hpp:96-185), same names and argument order.
// integrate / difference and their Jacobians are host code; dynamics_aba, inverse_dynamics_rnea and the frame functions
// run on the device and fail loudly without one.
#include <cmath>
#include <cstdio>
#include <vector>

#include "ddpb/ddpb.hpp"

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  try {
    ddpb::model_t model{std::string(argv[1])};
    const int n = int(model.tangent_dim());
    std::vector<double> q0(n), q(n), v(n, 0.25), d(n), J(n * n);
    model.neutral_configuration(q0.data());
    model.integrate(q.data(), q0.data(), v.data());
    model.difference(d.data(), q0.data(), q.data());          // test/pinocchio.cpp:17-57: difference(q, integrate(q, v)) == v
    double err = 0.0;
    for (int i = 0; i < n; ++i) err = std::fmax(err, std::fabs(d[i] - v[i]));
    model.d_difference_dq_finish(J.data(), q0.data(), q.data());
    std::printf("round trip %.1e, d_difference_dq_finish(0,0) %.1f (1,0) %.1f\n", err, J[0], J[1]);
    std::vector<double> tau(n, 0.5), a(n), tau_back(n), p(3), Jf(3 * n);
    model.dynamics_aba(a.data(), q.data(), v.data(), tau.data());
    model.inverse_dynamics_rnea(tau_back.data(), q.data(), v.data(), a.data());
    double e2 = 0.0;
    for (int i = 0; i < n; ++i) e2 = std::fmax(e2, std::fabs(tau_back[i] - tau[i]));
    model.frame_coordinates(p.data(), model.n_frames() - 1, q.data());
    model.d_frame_coordinates(Jf.data(), model.n_frames() - 1, q.data());
    std::printf("rnea(aba) - tau %.1e, last frame at (%.3f, %.3f, %.3f)\n", e2, p[0], p[1], p[2]);
  } catch (ddpb::error const& e) {
    std::printf("%s (code %d)\n", e.what(), e.code);
    return 1;
  }
  return 0;
}
