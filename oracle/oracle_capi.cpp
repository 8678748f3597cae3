// oracle_capi.cpp -- C ABI over the CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
// Loaded with ctypes by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference leg only.  See ddp_oracle.hpp for the parity status ("parity unpinned"
// for the Pinocchio arithmetic).
//
// All arrays are double at the boundary.  prec = 0 computes in double, prec = 1 computes in
// __float128 (stand-in for the reference's MPFR build, test/pinocchio_mpfr_500.cpp) and rounds
// the results to double.
#define ORC_HAVE_QUAD 1
#include "ddp_oracle.hpp"

#include <chrono>
#include <cstdio>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace orc;
typedef __float128 quad;

struct orc_model_desc {
  int kind, n_joints;
  const int* parent;
  const double* placement_rot;    // 9 n, row-major
  const double* placement_trans;  // 3 n
  const double* axis;             // 3 n
  const double* mass;             // n
  const double* com;              // 3 n
  const double* inertia;          // 6 n
  double gravity[3];
  int n_frames;
  const int* frame_parent;
  const double* frame_rot;        // 9 f
  const double* frame_trans;      // 3 f
  double pend_mass, pend_length;
  const int* joint_type;          // n: 0 revolute, 1 prismatic (NULL = all revolute)
};
struct orc_problem_desc {
  int T;
  double dt, c;
  int second_order_finite_diff;   // 1: value-based FD Hessians (problem.hpp:152), 0: FD of Jacobians (:67)
  int eq_kind, frame, jac_mode;
  const int* ne_per_knot;         // T (knot t constrains the state two steps later)
  const double* targets;          // sum(ne) packed per knot
  int max_restarts;               // guarded mode (0 = reference behaviour)
};
struct orc_derivs {
  double *lfx, *lfxx, *lx, *lu, *lxx, *lux, *luu, *f_val, *fx, *fu, *fxx, *fux, *fuu, *eq_val, *eq_x, *eq_u, *eq_xx, *eq_ux, *eq_uu;
};

struct Handle {
  Problem<double> pd;
  Problem<quad> pq;
  Problem<Counted> pc;
};

template <class S> static void fill_problem(Problem<S>& p, const orc_model_desc* m, const orc_problem_desc* d) {
  Model<S>& M = p.model;
  M.kind = m->kind;
  M.n = m->kind == KIND_PENDULUM ? 1 : m->n_joints;
  if (m->kind == KIND_REVOLUTE_TREE) {
    for (int i = 0; i < M.n; ++i) {
      M.parent[i] = m->parent[i];
      M.jtype[i] = m->joint_type ? m->joint_type[i] : JOINT_REVOLUTE;
      for (int k = 0; k < 9; ++k) M.Rp[i][k] = S(m->placement_rot[i * 9 + k]);
      for (int k = 0; k < 3; ++k) {
        M.pp[i][k] = S(m->placement_trans[i * 3 + k]);
        M.axis[i][k] = S(m->axis[i * 3 + k]);
        M.com[i][k] = S(m->com[i * 3 + k]);
      }
      M.mass[i] = S(m->mass[i]);
      for (int k = 0; k < 6; ++k) M.Ic[i][k] = S(m->inertia[i * 6 + k]);
    }
    M.nframes = m->n_frames;
    for (int f = 0; f < m->n_frames; ++f) {
      M.fparent[f] = m->frame_parent[f];
      for (int k = 0; k < 9; ++k) M.fR[f][k] = S(m->frame_rot[f * 9 + k]);
      for (int k = 0; k < 3; ++k) M.fp[f][k] = S(m->frame_trans[f * 3 + k]);
    }
  }
  for (int k = 0; k < 3; ++k) M.grav[k] = S(m->gravity[k]);
  M.pend_m = S(m->pend_mass);
  M.pend_l = S(m->pend_length);
  p.T = d->T;
  p.c = S(d->c);
  p.dyn.model = &p.model;
  p.dyn.dt = S(d->dt);
  p.dyn.second_order_finite_diff = d->second_order_finite_diff != 0;
  p.eq_kind = d->eq_kind;
  p.frame = d->frame;
  p.jac_mode = d->jac_mode;
  p.max_restarts = d->max_restarts;
  p.targets.t_begin = 2;
  p.targets.t_end = d->T + 2;
  p.targets.ne.assign(d->T, 0);
  p.targets.off.assign(d->T + 1, 0);
  int tot = 0;
  for (int t = 0; t < d->T; ++t) {
    p.targets.ne[t] = d->eq_kind == EQ_NONE ? 0 : d->ne_per_knot[t];
    p.targets.off[t] = tot;
    tot += p.targets.ne[t];
  }
  p.targets.off[d->T] = tot;
  p.targets.data.resize(tot);
  for (int k = 0; k < tot; ++k) p.targets.data[k] = S(d->targets[k]);
  p.finalize();
}

template <class S> static Problem<S>& prob(Handle* h);
template <> Problem<double>& prob<double>(Handle* h) { return h->pd; }
template <> Problem<quad>& prob<quad>(Handle* h) { return h->pq; }
template <> Problem<Counted>& prob<Counted>(Handle* h) { return h->pc; }

template <class S> static std::vector<S> in(const double* p, size_t n) {
  std::vector<S> v(n);
  for (size_t i = 0; i < n; ++i) v[i] = S(p[i]);
  return v;
}
template <class S> static void out(double* p, const S* v, size_t n) {
  if (!p) return;
  for (size_t i = 0; i < n; ++i) p[i] = ToDouble(v[i]);
}
template <class S> static void out(double* p, const std::vector<S>& v) { out(p, v.data(), v.size()); }

template <class S> static void derivs_in(const Problem<S>& p, const orc_derivs* a, Derivs<S>& d) {
  d.resize(p);
#define CP(f) for (size_t i = 0; i < d.f.size(); ++i) d.f[i] = S(a->f[i]);
  CP(lfx) CP(lfxx) CP(lx) CP(lu) CP(lxx) CP(lux) CP(luu) CP(f_val) CP(fx) CP(fu) CP(fxx) CP(fux) CP(fuu)
  CP(eq_val) CP(eq_x) CP(eq_u) CP(eq_xx) CP(eq_ux) CP(eq_uu)
#undef CP
}
template <class S> static void derivs_out(const Derivs<S>& d, const orc_derivs* a) {
#define CP(f) out(a->f, d.f);
  CP(lfx) CP(lfxx) CP(lx) CP(lu) CP(lxx) CP(lux) CP(luu) CP(f_val) CP(fx) CP(fu) CP(fxx) CP(fux) CP(fuu)
  CP(eq_val) CP(eq_x) CP(eq_u) CP(eq_xx) CP(eq_ux) CP(eq_uu)
#undef CP
}

template <class S> static void affine_in(const Problem<S>& p, AffineSeq<S>& a, const std::vector<int>& rows, const double* origin,
                                         const double* val, const double* jac) {
  a.init(p.T, rows, p.nx(), p.ndx());
  for (size_t i = 0; i < a.origin.size(); ++i) a.origin[i] = S(origin[i]);
  for (size_t i = 0; i < a.val.size(); ++i) a.val[i] = S(val[i]);
  for (size_t i = 0; i < a.jac.size(); ++i) a.jac[i] = S(jac[i]);
}

#define DISPATCH(prec, CALL)                 \
  do {                                       \
    if ((prec) == 0) { typedef double S; CALL; } \
    else { typedef quad S; CALL; }           \
  } while (0)

extern "C" void* orc_create(const orc_model_desc* m, const orc_problem_desc* d) {
  Handle* h = new Handle();
  fill_problem(h->pd, m, d);
  fill_problem(h->pq, m, d);
  fill_problem(h->pc, m, d);
  return h;
}
extern "C" void orc_destroy(void* h) { delete (Handle*)h; }

extern "C" void orc_dims(void* hv, int* o /* nq, nx, ndx, nu, ne_total, T */) {
  Handle* h = (Handle*)hv;
  o[0] = h->pd.model.nq(); o[1] = h->pd.nx(); o[2] = h->pd.ndx(); o[3] = h->pd.nu(); o[4] = h->pd.ne_total(); o[5] = h->pd.T;
}

template <class S> static void t_aba(Handle* h, const double* q, const double* v, const double* tau, double* a) {
  auto& p = prob<S>(h);
  int n = p.model.nv();
  auto Q = in<S>(q, n), V = in<S>(v, n), U = in<S>(tau, n);
  S A[MAXJ];
  dynamics_aba(p.model, Q.data(), V.data(), U.data(), A);
  out(a, A, n);
}
extern "C" void orc_aba(void* h, int prec, const double* q, const double* v, const double* tau, double* a) {
  DISPATCH(prec, t_aba<S>((Handle*)h, q, v, tau, a));
}
template <class S> static void t_rnea(Handle* h, const double* q, const double* v, const double* a, double* tau) {
  auto& p = prob<S>(h);
  int n = p.model.nv();
  auto Q = in<S>(q, n), V = in<S>(v, n), A = in<S>(a, n);
  S Tq[MAXJ];
  rnea(p.model, Q.data(), V.data(), A.data(), Tq);
  out(tau, Tq, n);
}
extern "C" void orc_rnea(void* h, int prec, const double* q, const double* v, const double* a, double* tau) {
  DISPATCH(prec, t_rnea<S>((Handle*)h, q, v, a, tau));
}
template <class S> static void t_daba(Handle* h, const double* q, const double* v, const double* tau, double* dq, double* dv, double* dt) {
  auto& p = prob<S>(h);
  int n = p.model.nv();
  auto Q = in<S>(q, n), V = in<S>(v, n), U = in<S>(tau, n);
  S a[MAXJ * MAXJ], b[MAXJ * MAXJ], c[MAXJ * MAXJ];
  d_dynamics_aba(p.model, Q.data(), V.data(), U.data(), a, b, c);
  out(dq, a, n * n); out(dv, b, n * n); out(dt, c, n * n);
}
extern "C" void orc_d_aba(void* h, int prec, const double* q, const double* v, const double* tau, double* dq, double* dv, double* dtau) {
  DISPATCH(prec, t_daba<S>((Handle*)h, q, v, tau, dq, dv, dtau));
}
template <class S> static void t_frame(Handle* h, int frame, int mode, const double* q, double* p3, double* J) {
  auto& p = prob<S>(h);
  int n = p.model.nv();
  auto Q = in<S>(q, n);
  S P[3], Jm[3 * MAXJ];
  frame_coordinates(p.model, frame, Q.data(), P);
  d_frame_coordinates(p.model, frame, Q.data(), mode, Jm);
  out(p3, P, 3); out(J, Jm, 3 * n);
}
extern "C" void orc_frame(void* h, int prec, int frame, int mode, const double* q, double* p3, double* J) {
  DISPATCH(prec, t_frame<S>((Handle*)h, frame, mode, q, p3, J));
}
template <class S> static void t_first_order(Handle* h, const double* x, const double* u, double* f, double* fx, double* fu) {
  auto& p = prob<S>(h);
  int nx = p.nx(), nd = p.ndx(), nu = p.nu();
  auto X = in<S>(x, nx), U = in<S>(u, nu);
  std::vector<S> F(nx), FX(nd * nd), FU(nd * nu);
  p.dyn.first_order_deriv(FX.data(), FU.data(), F.data(), X.data(), U.data());
  out(f, F); out(fx, FX); out(fu, FU);
}
extern "C" void orc_first_order(void* h, int prec, const double* x, const double* u, double* f, double* fx, double* fu) {
  DISPATCH(prec, t_first_order<S>((Handle*)h, x, u, f, fx, fu));
}
// flops of one dynamics first_order_deriv (counting scalar); pins W_fo of BASELINE.md section 3
extern "C" uint64_t orc_flops_first_order(void* hv, const double* x, const double* u) {
  Handle* h = (Handle*)hv;
  auto& p = h->pc;
  int nx = p.nx(), nd = p.ndx(), nu = p.nu();
  auto X = in<Counted>(x, nx), U = in<Counted>(u, nu);
  std::vector<Counted> F(nx), FX(nd * nd), FU(nd * nu);
  FlopCounter::n() = 0;
  p.dyn.first_order_deriv(FX.data(), FU.data(), F.data(), X.data(), U.data());
  return FlopCounter::n();
}
extern "C" uint64_t orc_flops_eval(void* hv, const double* x, const double* u) {
  Handle* h = (Handle*)hv;
  auto& p = h->pc;
  auto X = in<Counted>(x, p.nx()), U = in<Counted>(u, p.nu());
  std::vector<Counted> F(p.nx());
  FlopCounter::n() = 0;
  p.dyn.eval_to(F.data(), X.data(), U.data());
  return FlopCounter::n();
}
// make_trajectory (ddp.hpp:392-415): roll the controls out from x0
template <class S> static void t_rollout(Handle* h, const double* x0, const double* us, double* xs) {
  auto& p = prob<S>(h);
  int nx = p.nx(), nu = p.nu(), T = p.T;
  std::vector<S> X((T + 1) * nx);
  auto U = in<S>(us, T * nu);
  for (int k = 0; k < nx; ++k) X[k] = S(x0[k]);
  for (int t = 0; t < T; ++t) p.dyn.eval_to(&X[(t + 1) * nx], &X[t * nx], &U[t * nu]);
  out(xs, X);
}
extern "C" void orc_rollout(void* h, int prec, const double* x0, const double* us, double* xs) { DISPATCH(prec, t_rollout<S>((Handle*)h, x0, us, xs)); }

template <class S> static void t_eval_eq(Handle* h, int t, const double* x, const double* u, double* o) {
  auto& p = prob<S>(h);
  auto X = in<S>(x, p.nx()), U = in<S>(u, p.nu());
  S ce[MAXE];
  p.eval_eq_to(ce, t, X.data(), U.data());
  out(o, ce, p.ne[t]);
}
extern "C" void orc_eval_eq(void* h, int prec, int t, const double* x, const double* u, double* o) { DISPATCH(prec, t_eval_eq<S>((Handle*)h, t, x, u, o)); }

template <class S> static void t_compute_derivatives(Handle* h, const double* xs, const double* us, const orc_derivs* o, int nthreads) {
  auto& p = prob<S>(h);
  auto X = in<S>(xs, (p.T + 1) * p.nx()), U = in<S>(us, p.T * p.nu());
  Derivs<S> d;
  d.resize(p);
  (void)nthreads;
  compute_derivatives(p, d, X.data(), U.data());
  derivs_out(d, o);
}
extern "C" void orc_compute_derivatives(void* h, int prec, const double* xs, const double* us, const orc_derivs* o) {
  DISPATCH(prec, t_compute_derivatives<S>((Handle*)h, xs, us, o, 1));
}

template <class S>
static void t_backward(Handle* h, const orc_derivs* din, const double* xs, const double* mval, const double* mjac, double reg, double mu,
                       double* k, double* K, double* mu_out, double* reg_out, int* restarts) {
  auto& p = prob<S>(h);
  Derivs<S> d;
  derivs_in(p, din, d);
  auto X = in<S>(xs, (p.T + 1) * p.nx());
  AffineSeq<S> mults, fb;
  affine_in(p, mults, p.ne, xs, mval, mjac);
  fb.init(p.T, std::vector<int>(p.T, p.nu()), p.nx(), p.ndx());
  BwdResult<S> r = backward_pass(p, d, X.data(), mults, fb, S(reg), S(mu));
  out(k, fb.val); out(K, fb.jac);
  *mu_out = ToDouble(r.mu); *reg_out = ToDouble(r.reg); *restarts = r.restarts;
}
extern "C" void orc_backward_pass(void* h, int prec, const orc_derivs* din, const double* xs, const double* mval, const double* mjac, double reg,
                       double mu, double* k, double* K, double* mu_out, double* reg_out, int* restarts) {
  DISPATCH(prec, t_backward<S>((Handle*)h, din, xs, mval, mjac, reg, mu, k, K, mu_out, reg_out, restarts));
}

template <class S>
static void t_cost(Handle* h, const double* xs, const double* us, const double* morig, const double* mval, const double* mjac, double mu,
                   double* costs) {
  auto& p = prob<S>(h);
  auto X = in<S>(xs, (p.T + 1) * p.nx()), U = in<S>(us, p.T * p.nu());
  AffineSeq<S> mults;
  affine_in(p, mults, p.ne, morig, mval, mjac);
  std::vector<S> c(p.T + 1);
  cost_seq_aug(p, c.data(), X.data(), U.data(), mults, S(mu));
  out(costs, c);
}
extern "C" void orc_cost_seq_aug(void* h, int prec, const double* xs, const double* us, const double* morig, const double* mval, const double* mjac,
                      double mu, double* costs) {
  DISPATCH(prec, t_cost<S>((Handle*)h, xs, us, morig, mval, mjac, mu, costs));
}

template <class S>
static void t_forward(Handle* h, const double* xs, const double* us, const double* morig, const double* mval, const double* mjac,
                      const double* k, const double* K, double mu, double* xs_new, double* us_new, double* step, int* rollouts) {
  auto& p = prob<S>(h);
  auto X = in<S>(xs, (p.T + 1) * p.nx()), U = in<S>(us, p.T * p.nu());
  AffineSeq<S> mults, fb;
  affine_in(p, mults, p.ne, morig, mval, mjac);
  affine_in(p, fb, std::vector<int>(p.T, p.nu()), xs, k, K);
  // the new-trajectory storage starts as a clone of the reference trajectory (ddp.hpp:752)
  std::vector<S> XN = in<S>(xs_new, X.size()), UN = in<S>(us_new, U.size());
  FwdResult<S> r = forward_pass(p, XN.data(), UN.data(), X.data(), U.data(), mults, fb, S(mu), true);
  out(xs_new, XN); out(us_new, UN);
  *step = ToDouble(r.step); *rollouts = r.rollouts;
}
extern "C" void orc_forward_pass(void* h, int prec, const double* xs, const double* us, const double* morig, const double* mval, const double* mjac,
                      const double* k, const double* K, double mu, double* xs_new, double* us_new, double* step, int* rollouts) {
  DISPATCH(prec, t_forward<S>((Handle*)h, xs, us, morig, mval, mjac, k, K, mu, xs_new, us_new, step, rollouts));
}

template <class S>
static void t_optimality(Handle* h, const orc_derivs* din, const double* xs, const double* morig, const double* mval, const double* mjac,
                         double mu, double* opt_obj, double* opt_constr) {
  auto& p = prob<S>(h);
  Derivs<S> d;
  derivs_in(p, din, d);
  auto X = in<S>(xs, (p.T + 1) * p.nx());
  AffineSeq<S> mults;
  affine_in(p, mults, p.ne, morig, mval, mjac);
  *opt_obj = ToDouble(optimality_obj(p, X.data(), mults, S(mu), d));
  *opt_constr = ToDouble(optimality_constr(p, d));
}
extern "C" void orc_optimality(void* h, int prec, const orc_derivs* din, const double* xs, const double* morig, const double* mval, const double* mjac,
                    double mu, double* opt_obj, double* opt_constr) {
  DISPATCH(prec, t_optimality<S>((Handle*)h, din, xs, morig, mval, mjac, mu, opt_obj, opt_constr));
}

// log row: mu, reg, w, n, step, opt_obj, opt_constr, rollouts, restarts, mult_update  (10 doubles per iteration)
template <class S>
static void t_solve(Handle* h, int max_iter, double thr, double mu, double reg, double w, double n, const double* xs0, const double* us0,
                    const double* mjac0, double* xs, double* us, double* k, double* K, double* mval, double* mjac, double* log,
                    int* iterations, int* status) {
  auto& p = prob<S>(h);
  auto X = in<S>(xs0, (p.T + 1) * p.nx()), U = in<S>(us0, p.T * p.nu());
  std::vector<S> MJ = in<S>(mjac0, (size_t)p.ne_total() * p.ndx());
  SolverParams<S> sp{max_iter, S(thr), S(mu), S(reg), S(w), S(n)};
  SolveResult<S> R;
  solve(p, sp, X.data(), U.data(), MJ.data(), R);
  out(xs, R.xs); out(us, R.us); out(k, R.fb.val); out(K, R.fb.jac); out(mval, R.mults.val); out(mjac, R.mults.jac);
  if (log)
    for (size_t i = 0; i < R.log.size(); ++i) {
      const auto& l = R.log[i];
      double row[10] = {ToDouble(l.mu), ToDouble(l.reg), ToDouble(l.w), ToDouble(l.n), ToDouble(l.step), ToDouble(l.opt_obj),
                        ToDouble(l.opt_constr), (double)l.rollouts, (double)l.restarts, (double)l.mult_update};
      for (int c = 0; c < 10; ++c) log[i * 10 + c] = row[c];
    }
  *iterations = R.iterations;
  *status = R.status;
}
extern "C" void orc_solve(void* h, int prec, int max_iter, double thr, double mu, double reg, double w, double n, const double* xs0, const double* us0,
               const double* mjac0, double* xs, double* us, double* k, double* K, double* mval, double* mjac, double* log, int* iterations,
               int* status) {
  DISPATCH(prec, t_solve<S>((Handle*)h, max_iter, thr, mu, reg, w, n, xs0, us0, mjac0, xs, us, k, K, mval, mjac, log, iterations, status));
}

// CPU baseline for bench.py: solve `nprob` independent problems (same model / horizon, per-problem
// x0, targets and multiplier seeds) with `nthreads` OpenMP threads; returns wall seconds.
// xs0: [nprob][(T+1) nx], us0: [nprob][T nu], targets: [nprob][sum ne], mjac0: [nprob][sum ne * ndx]
extern "C" double orc_solve_batch(const orc_model_desc* m, const orc_problem_desc* d, int nprob, int nthreads, int max_iter, double thr, double mu,
                       double reg, double w, double n, const double* xs0, const double* us0, const double* targets, const double* mjac0,
                       double* xs_out, double* us_out, int* rollouts_total, int* restarts_total, long long* iterations_total, int* diverged_total) {
  int ro = 0, rs = 0, dv = 0;
  long long it = 0;
  auto t0 = std::chrono::steady_clock::now();
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 1) reduction(+ : ro, rs, dv, it)
#endif
  for (int b = 0; b < nprob; ++b) {
    Problem<double> p;
    orc_problem_desc db = *d;
    int netot = 0;
    for (int t = 0; t < d->T; ++t) netot += d->eq_kind == EQ_NONE ? 0 : d->ne_per_knot[t];
    db.targets = targets + (size_t)b * netot;
    fill_problem(p, m, &db);
    const int nx = p.nx(), nu = p.nu(), T = p.T;
    SolverParams<double> sp{max_iter, thr, mu, reg, w, n};
    SolveResult<double> R;
    solve(p, sp, xs0 + (size_t)b * (T + 1) * nx, us0 + (size_t)b * T * nu, mjac0 + (size_t)b * netot * p.ndx(), R);
    if (xs_out) out(xs_out + (size_t)b * (T + 1) * nx, R.xs);
    if (us_out) out(us_out + (size_t)b * T * nu, R.us);
    for (auto& l : R.log) { ro += l.rollouts; rs += l.restarts; }
    // iterations carried through: a problem frozen by the guard (status 2) does not complete the iteration it stopped in
    it += (long long)R.log.size() - (R.status == 2 && !R.log.empty() ? 1 : 0);
    dv += R.status == 2;
  }
  auto t1 = std::chrono::steady_clock::now();
  if (rollouts_total) *rollouts_total = ro;
  if (restarts_total) *restarts_total = rs;
  if (iterations_total) *iterations_total = it;
  if (diverged_total) *diverged_total = dv;
  return std::chrono::duration<double>(t1 - t0).count();
}

extern "C" int orc_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

