// ddpb.hpp -- C++ host-side mirror of the reference's operator interface for the batched hot path.
//
// The reference is a header-only C++14 template library; its "plugin API" for this path is the trio of
// duck-typed concepts Model / Dynamics+Constraint / Problem consumed by ddp_solver_t
// (/root/reference/include/ddp/problem.hpp:343-525, 527-864, 872-1150; ddp.hpp:300-869).  This header keeps
// the same names, argument meaning and defaults, but every object describes a BATCH of independent
// problems and every hot call forwards to the C ABI of include/ddpb.h (libddpb.so, CUDA, sm_100a).
// It depends on nothing but the C ABI and the standard library (no Eigen / fmt / Boost).
//
// Error convention: the reference terminates on precondition violations (DDP_ASSERT -> std::terminate,
// src/lib.cpp:107-110); here a failing C-ABI call throws ddpb::error carrying ddpb_last_error().
#pragma once
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <cstdlib>
#include <utility>
#include <vector>

#include "../../../include/ddpb.h"

namespace ddpb {

using index_t = long;  // ddp::index_t

struct error : std::runtime_error {
  int code;
  error(int c, const char* what) : std::runtime_error(std::string("ddpb: ") + what), code(c) {}
};
inline void check(int rc) {
  if (rc != DDPB_OK) throw error(rc, ddpb_last_error());
}

enum struct method { primal, primal_dual_constant_multipliers, primal_dual_affine_multipliers };  // ddp.hpp:18-22

// solver_parameters_t (ddp.hpp:42-50)
template <typename Scalar> struct solver_parameters_t {
  index_t max_iterations;
  Scalar optimality_stopping_threshold;
  Scalar mu, reg, w, n;
};

// Model concept: pinocchio::model_t / pendulum_model_t.  Owns a ddpb_model (move-only, like the reference's
// pimpl, pinocchio_model.hpp:71-75).
struct model_t {
  using scalar_t = double;
  ddpb_model* h = nullptr;
  int nq = 0;
  int frames = 0;
  explicit model_t(const ddpb_model_desc& d) {
    check(ddpb_model_create(&d, &h));
    nq = d.kind == DDPB_MODEL_PENDULUM ? 1 : d.n_joints;
    frames = d.n_frames;
  }
  // model_t{urdf_path} (pinocchio_model.ipp:155: pinocchio::urdf::buildModel with a fixed base; a leading '~' is the home
  // directory as in builder_from_urdf_t, :98-113).  Serial chains of revolute / continuous / prismatic joints, fixed joints merged.
  explicit model_t(const std::string& urdf_path) {
    std::string path = urdf_path;
    if (!path.empty() && path[0] == '~') { const char* home = std::getenv("HOME"); path = std::string(home ? home : "") + path.substr(1); }
    std::ifstream in(path, std::ios::binary);
    if (!in) throw error(DDPB_ERR_INVALID, ("cannot open " + path).c_str());
    std::ostringstream ss;
    ss << in.rdbuf();
    *this = from_urdf_text(ss.str());
  }
  static model_t from_urdf_text(const std::string& xml) {
    model_t m;
    check(ddpb_model_create_from_urdf(xml.data(), xml.size(), &m.h));
    int dims[2];
    check(ddpb_model_dims(m.h, dims));
    m.nq = dims[0]; m.frames = dims[1];
    return m;
  }
  const char* frame_name(index_t i) const { return ddpb_model_frame_name(h, (int)i); }  // pinocchio_model.hpp frame_name(i)
  // pendulum_model_t{mass, length} (pendulum_model.hpp:29)
  static model_t pendulum(double mass, double length) {
    ddpb_model_desc d{};
    d.kind = DDPB_MODEL_PENDULUM; d.pend_mass = mass; d.pend_length = length; d.gravity[2] = -9.81;
    return model_t(d);
  }
  model_t(model_t&& o) noexcept : h(o.h), nq(o.nq), frames(o.frames) { o.h = nullptr; }
  model_t& operator=(model_t&& o) noexcept {
    if (this != &o) { ddpb_model_destroy(h); h = o.h; nq = o.nq; frames = o.frames; o.h = nullptr; }
    return *this;
  }
  model_t(const model_t&) = delete;
  model_t& operator=(const model_t&) = delete;

 private:
  model_t() = default;

 public:
  ~model_t() { ddpb_model_destroy(h); }
  index_t configuration_dim() const { return nq; }
  index_t tangent_dim() const { return nq; }
  index_t n_frames() const { return frames; }

  // ---- the Model concept's functions (pinocchio_model.hpp:96-185, pendulum_model.hpp:35-132).  Same names and argument
  // order; the views are plain pointers into caller-owned storage, matrices column-major like the reference's Eigen maps.
  // Every joint of the supported models has nq = nv = 1 (revolute on the plain angle, prismatic, the pendulum), so the
  // configuration space is a vector space: integrate is q + v, difference is q1 - q0, their Jacobians the identity --
  // exactly what pinocchio::integrate / difference / dIntegrate / dDifference return for such joints.
  void neutral_configuration(double* out_q) const { for (int i = 0; i < nq; ++i) out_q[i] = 0.0; }
  void integrate(double* out_q, const double* q, const double* v) const { for (int i = 0; i < nq; ++i) out_q[i] = q[i] + v[i]; }
  void difference(double* out_v, const double* q_start, const double* q_finish) const { for (int i = 0; i < nq; ++i) out_v[i] = q_finish[i] - q_start[i]; }
  void d_integrate_dq(double* out, const double*, const double*) const { identity(out, 1.0); }
  void d_integrate_dv(double* out, const double*, const double*) const { identity(out, 1.0); }
  void d_difference_dq_start(double* out, const double*, const double*) const { identity(out, -1.0); }
  void d_difference_dq_finish(double* out, const double*, const double*) const { identity(out, 1.0); }
  // The dynamics and frame functions run on the device (one thread per sample, csrc/model_eval.cuh): the single-sample
  // forms keep the reference's signatures, the *_batch forms take n samples (sample-major arrays) in one launch.
  void dynamics_aba(double* out_acceleration, const double* q, const double* v, const double* tau, int device = 0) const {
    check(ddpb_model_dynamics_aba(h, device, 1, q, v, tau, out_acceleration));
  }
  void d_dynamics_aba(double* out_dq, double* out_dv, double* out_dtau, const double* q, const double* v, const double* tau, int device = 0) const {
    check(ddpb_model_d_dynamics_aba(h, device, 1, q, v, tau, out_dq, out_dv, out_dtau));
  }
  void inverse_dynamics_rnea(double* out_tau, const double* q, const double* v, const double* a, int device = 0) const {
    check(ddpb_model_inverse_dynamics_rnea(h, device, 1, q, v, a, out_tau));
  }
  void frame_coordinates(double* out_p3, index_t i, const double* q, int device = 0) const {
    check(ddpb_model_frame_coordinates(h, device, 1, int(i), q, out_p3));
  }
  void d_frame_coordinates(double* out_3xnv, index_t i, const double* q, int frame_jac_mode = DDPB_FJ_WORLD, int device = 0) const {
    check(ddpb_model_d_frame_coordinates(h, device, 1, int(i), frame_jac_mode, q, out_3xnv));
  }
  void dynamics_aba_batch(int n, double* out_acceleration, const double* q, const double* v, const double* tau, int device = 0) const {
    check(ddpb_model_dynamics_aba(h, device, n, q, v, tau, out_acceleration));
  }
  void d_dynamics_aba_batch(int n, double* out_dq, double* out_dv, double* out_dtau, const double* q, const double* v, const double* tau,
                            int device = 0) const {
    check(ddpb_model_d_dynamics_aba(h, device, n, q, v, tau, out_dq, out_dv, out_dtau));
  }

 private:
  void identity(double* out, double d) const {
    for (int c = 0; c < nq; ++c)
      for (int r = 0; r < nq; ++r) out[c * nq + r] = r == c ? d : 0.0;
  }
};

namespace detail {
// The per-call forms of the Dynamics / Constraint / Problem concepts (eval_to, first_order_deriv, second_order_deriv,
// eval_eq_to ...) evaluate ONE knot of n independent (x, u) samples: a batch of n problems with horizon 1 goes through
// ddpb_compute_derivatives, i.e. through the same linearization kernel as a solve.  A call creates and destroys its batch
// (device allocations: milliseconds) -- bulk work belongs to ddp_solver_t, these exist so that code written against the
// reference's concepts finds the functions it calls.
inline void knot_derivatives(const ddpb_model* model, int nq, double dt, double c, bool second_order_finite_diff, int eq_kind, int ne,
                             int frame_id, int frame_jac_mode, const double* targets, int n, const double* x, const double* u,
                             const ddpb_derivs& out, int device) {
  ddpb_params pr{};
  pr.dt = dt; pr.c = c;
  pr.hess_mode = second_order_finite_diff ? DDPB_HESS_FD_VALUES : DDPB_HESS_FD_JACOBIANS;
  pr.eq_kind = ne > 0 ? eq_kind : DDPB_EQ_NONE; pr.frame_id = frame_id; pr.frame_jac_mode = frame_jac_mode;
  ddpb_batch* b = nullptr;
  check(ddpb_batch_create(model, n, 1, ne > 0 ? &ne : nullptr, &pr, device, &b));
  struct guard { ddpb_batch* b; ~guard() { ddpb_batch_destroy(b); } } g{b};
  const std::size_t nx = 2 * static_cast<std::size_t>(nq);
  std::vector<double> xs(static_cast<std::size_t>(n) * 2 * nx, 0.0);  // [n][(T+1) nx]: x_0 = x; x_1 is not read by knot 0
  for (int s = 0; s < n; ++s)
    for (std::size_t k = 0; k < nx; ++k) xs[static_cast<std::size_t>(s) * 2 * nx + k] = x[static_cast<std::size_t>(s) * nx + k];
  check(ddpb_set_trajectory(b, xs.data(), u));
  if (ne > 0) check(ddpb_set_targets(b, targets));
  check(ddpb_compute_derivatives(b));
  check(ddpb_get_derivatives(b, &out));
}
}  // namespace detail

// dynamics_t{model, dt, second_order_finite_diff = true} (problem.hpp:522-524)
struct dynamics_t {
  model_t const& m_model;
  double dt;
  bool second_order_finite_diff = true;
  index_t state_dim() const { return 2 * m_model.nq; }
  index_t dstate_dim() const { return 2 * m_model.nq; }
  index_t control_dim(index_t) const { return m_model.nq; }

  // Dynamics concept, per call (problem.hpp:441-521).  out: x_next [nx]; out_x: f_x [ndx x ndx], out_u: f_u [ndx x nu]
  // (column-major); out_xx / out_ux / out_uu: the tensors T(k, i, j) at k + i O + j O L (detail/tensor.hpp:146).  The time index
  // is ignored as in the reference.  `n` samples per call (sample-major arrays), default one.
  void eval_to(double* out, index_t, const double* x, const double* u, int n = 1, int device = 0) const {
    ddpb_derivs d{}; d.f_val = out;
    run(d, x, u, n, device);
  }
  void first_order_deriv(double* out_x, double* out_u, double* out, index_t, const double* x, const double* u, int n = 1, int device = 0) const {
    ddpb_derivs d{}; d.f_val = out; d.fx = out_x; d.fu = out_u;
    run(d, x, u, n, device);
  }
  void second_order_deriv(double* out_xx, double* out_ux, double* out_uu, double* out_x, double* out_u, double* out, index_t, const double* x,
                          const double* u, int n = 1, int device = 0) const {
    ddpb_derivs d{}; d.f_val = out; d.fx = out_x; d.fu = out_u; d.fxx = out_xx; d.fux = out_ux; d.fuu = out_uu;
    run(d, x, u, n, device);
  }
  // integrate_x / integrate_u / difference_out of the concept (problem.hpp:404-439) on the vector-space state of these models
  void integrate_x(double* out, const double* x, const double* dx) const { for (index_t i = 0; i < state_dim(); ++i) out[i] = x[i] + dx[i]; }
  void integrate_u(double* out, index_t, const double* u, const double* du) const { for (int i = 0; i < m_model.nq; ++i) out[i] = u[i] + du[i]; }
  void difference_out(double* out, const double* start, const double* finish) const { for (index_t i = 0; i < state_dim(); ++i) out[i] = finish[i] - start[i]; }

 private:
  void run(const ddpb_derivs& d, const double* x, const double* u, int n, int device) const {
    detail::knot_derivatives(m_model.h, m_model.nq, dt, 0.0, second_order_finite_diff, DDPB_EQ_NONE, 0, 0, DDPB_FJ_WORLD, nullptr, n, x, u, d, device);
  }
};

// Constraint concept.  The reference composes constraint_advance_time(constraint_advance_time(c)) around a
// config_constraint_t or spatial_constraint_t (test/pinocchio_ddp.cpp:61-68); the batch path implements exactly
// that composition, so the constraint is described by its kind, its target rows per BASE index t in
// [2, T+2) (the reference's eq_idx, test/pendulum_ddp.cpp:37-47) and, for the spatial one, the frame.
struct constraint_t {
  int kind = DDPB_EQ_NONE;          // DDPB_EQ_CONFIG | DDPB_EQ_SPATIAL
  std::vector<int> rows;            // rows[t - 2] for base index t in [2, T+2)
  int frame_id = 0;
  int frame_jac_mode = DDPB_FJ_WORLD;
  static constraint_t config_terminal(index_t horizon, int nq) {  // range_row_filter(idx, T, T+1)
    constraint_t c; c.kind = DDPB_EQ_CONFIG; c.rows.assign(horizon, 0); c.rows[horizon - 2] = nq; return c;
  }
  static constraint_t config_every_knot(index_t horizon, int nq) {  // vec_regular_indexer(2, T+2, nq)
    constraint_t c; c.kind = DDPB_EQ_CONFIG; c.rows.assign(horizon, nq); return c;
  }
  static constraint_t spatial_terminal(index_t horizon, int frame) {
    constraint_t c; c.kind = DDPB_EQ_SPATIAL; c.rows.assign(horizon, 0); c.rows[horizon - 2] = 3; c.frame_id = frame; return c;
  }
  // eq_dim(t) of the doubly advanced constraint (problem.hpp:553): rows the base indexer has at t + 2
  int eq_dim(index_t t) const { return rows[static_cast<std::size_t>(t)]; }
};

// problem_t{begin, end, c, dynamics, constraint} (problem.hpp:1145-1149); c defaults to 1e2 there.
struct problem_t {
  index_t m_begin, m_end;
  double c = 1e2;
  dynamics_t m_dynamics;
  constraint_t m_constraint;
  index_t horizon() const { return m_end - m_begin; }

  // Problem concept, per call (problem.hpp:932-955): lf(x) = 0, l(t, x, u) = c/2 |u|^2, eval_f_to = the dynamics,
  // eval_eq_to = the doubly advanced constraint at knot t (rows eq_dim(t), `target` = those rows of the target),
  // and one knot of compute_derivatives for n samples (any pointer of `out` may be null).
  double lf(const double*) const { return 0.0; }
  double l(index_t, const double*, const double* u) const {
    double s = 0.0;
    for (int i = 0; i < m_dynamics.m_model.nq; ++i) s += u[i] * u[i];
    return c / 2 * s;
  }
  void eval_f_to(double* x_out, index_t t, const double* x, const double* u, int n = 1, int device = 0) const { m_dynamics.eval_to(x_out, t, x, u, n, device); }
  void eval_eq_to(double* eq_out, index_t t, const double* x, const double* u, const double* target, int n = 1, int device = 0) const {
    ddpb_derivs d{}; d.eq_val = eq_out;
    knot_derivatives(t, x, u, target, d, n, device);
  }
  void knot_derivatives(index_t t, const double* x, const double* u, const double* target, const ddpb_derivs& out, int n = 1, int device = 0) const {
    const int ne = m_constraint.kind == DDPB_EQ_NONE ? 0 : m_constraint.eq_dim(t - m_begin);
    detail::knot_derivatives(m_dynamics.m_model.h, m_dynamics.m_model.nq, m_dynamics.dt, c, m_dynamics.second_order_finite_diff, m_constraint.kind, ne,
                             m_constraint.frame_id, m_constraint.frame_jac_mode, target, n, x, u, out, device);
  }
};

// per-problem scalars returned by solve (replaces the reference's stdout / log files, ddp.hpp:651-671)
struct scalars_t {
  int B = 0;
  std::vector<double> data;  // [DDPB_NSCALARS][B]
  double at(int field, int b) const { return data[static_cast<std::size_t>(field) * B + b]; }
};

// ddp_solver_t<Problem> for a batch of B problems that share `prob`'s structure (ddp.hpp:300-869).
struct ddp_solver_t {
  problem_t const& prob;
  int B;
  ddpb_batch* h = nullptr;
  int nx, nu, T;

  ddp_solver_t(problem_t const& p, int batch, int device = 0, int max_backward_restarts = 0)
      : prob(p), B(batch), nx(int(p.m_dynamics.state_dim())), nu(p.m_dynamics.m_model.nq), T(int(p.horizon())) {
    ddpb_params pr{};
    pr.dt = p.m_dynamics.dt; pr.c = p.c;
    pr.hess_mode = p.m_dynamics.second_order_finite_diff ? DDPB_HESS_FD_VALUES : DDPB_HESS_FD_JACOBIANS;
    pr.eq_kind = p.m_constraint.kind; pr.frame_id = p.m_constraint.frame_id; pr.frame_jac_mode = p.m_constraint.frame_jac_mode;
    pr.max_backward_restarts = max_backward_restarts;
    check(ddpb_batch_create(p.m_dynamics.m_model.h, B, T, p.m_constraint.kind == DDPB_EQ_NONE ? nullptr : p.m_constraint.rows.data(), &pr,
                            device, &h));
  }
  ddp_solver_t(const ddp_solver_t&) = delete;             // non-copyable, non-movable like the reference (ddp.hpp:737-741)
  ddp_solver_t& operator=(const ddp_solver_t&) = delete;
  ~ddp_solver_t() { ddpb_batch_destroy(h); }

  int ne_total() const { int s = 0; for (int r : prob.m_constraint.rows) s += r; return prob.m_constraint.kind == DDPB_EQ_NONE ? 0 : s; }

  // make_trajectory(control generator) (ddp.hpp:392-402): x0 [B][nx]; us [B][T*nu] or nullptr for the zero-control generator
  void make_trajectory(const double* x0, const double* us = nullptr) { check(ddpb_make_trajectory(h, x0, us)); }
  void set_targets(const double* targets) { check(ddpb_set_targets(h, targets)); }
  // zero_multipliers<M>() + the (seeded) jac initialisation of solve (ddp.hpp:759-764)
  void set_multipliers(const double* val, const double* jac) { check(ddpb_set_multipliers(h, val, jac)); }

  // problem_t::compute_derivatives (problem.hpp:956)
  void compute_derivatives() { check(ddpb_compute_derivatives(h)); }
  // backward_pass<M>(ctrl_fb, traj, mults, reg, mu, derivs) -> {feedback, mu, reg}  (ddp.hpp:845-853); per-problem mu / reg
  void backward_pass(std::vector<double>& mu, std::vector<double>& reg) { check(ddpb_backward_pass(h, mu.data(), reg.data())); }
  // forward_pass<M>(new_traj, traj, mults, bres, do_linesearch = true) -> step (ddp.hpp:855-862)
  std::vector<double> forward_pass() { std::vector<double> s(B); check(ddpb_forward_pass(h, s.data())); return s; }

  // solve<M>(solver_parameters, initial_trajectory) (ddp.hpp:744-748); the trajectory / feedback stay on the device,
  // read them with trajectory() / feedback()
  template <method M = method::primal_dual_affine_multipliers> scalars_t solve(solver_parameters_t<double> sp) {
    static_assert(M == method::primal_dual_affine_multipliers, "only the method the reference instantiates is provided");
    ddpb_solver_params p{int(sp.max_iterations), sp.optimality_stopping_threshold, sp.mu, sp.reg, sp.w, sp.n};
    scalars_t s; s.B = B; s.data.resize(static_cast<std::size_t>(DDPB_NSCALARS) * B);
    check(ddpb_solve(h, &p, s.data.data()));
    return s;
  }
  void trajectory(std::vector<double>& xs, std::vector<double>& us) {
    xs.resize(static_cast<std::size_t>(B) * (T + 1) * nx); us.resize(static_cast<std::size_t>(B) * T * nu);
    check(ddpb_get_trajectory(h, xs.data(), us.data()));
  }
  void feedback(std::vector<double>& k, std::vector<double>& K) {
    k.resize(static_cast<std::size_t>(B) * T * nu); K.resize(static_cast<std::size_t>(B) * T * nu * nx);
    check(ddpb_get_feedback(h, nullptr, k.data(), K.data()));
  }
};

}  // namespace ddpb
