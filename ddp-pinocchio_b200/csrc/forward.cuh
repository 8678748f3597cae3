// forward.cuh -- closed-loop rollouts with backtracking line search on the augmented-Lagrangian merit,
// plus the per-problem sequential sweeps of the outer loop (re-centering, optimality measures,
// multiplier update).  One thread per problem (the device versions of the line search and of the sweeps, with their
// own work split, are in outer_warp.cuh; the bodies here are also what tests/hostemu runs on the CPU); all per-knot
// vectors are SoA-by-batch so every access of a warp is one coalesced span.
//
// Restates (all under /root/reference/include/ddp/):
//   forward_pass                ddp_fwd.ipp:10-67
//   cost_seq_aug                ddp.hpp:699-735
//   make_trajectory             ddp.hpp:392-415
//   update_origin               detail/mat_seq_common.hpp:62-89
//   optimality_obj / _constr    ddp.hpp:516-627
//   update_derivatives (after compute_derivatives) and the mu / w / n state machine of solve
//                               ddp.hpp:657-696, 786-802
#pragma once
#include "batch.cuh"
#include "linearize.cuh"

DDPB_VARIANT_BEGIN

// x_{t+1} = f(x_t, u_t): explicit Euler on forward dynamics (problem.hpp:441-461)
// per-thread scratch of the fused forward dynamics: element e at p[e * stride] (rbd.cuh aba_fused)
struct Scratch {
  double* p;
  int stride;
};
template <int NJ> DDPB_HD void dyn_eval(const ModelDev& m, double dt, const double* x, const double* u, double* xn, const Scratch& sc) {
  double qdd[NJ];
  if (NJ <= 3) {
    // small models: the whole q-record fits the register file (same arithmetic as aba_fused)
    double qrec[QRec<NJ>::SIZE];
    q_task<NJ>(m, x, qrec);
    v_task<NJ, false>(m, qrec, x + NJ, u, qdd, nullptr);
  } else {
    aba_fused<NJ>(m, x, x + NJ, u, qdd, sc.p, sc.stride);
  }
#pragma unroll
  for (int k = 0; k < NJ; ++k) {
    const double vo = dt * x[NJ + k];
    xn[k] = x[k] + vo;
    xn[NJ + k] = x[NJ + k] + qdd[k] * dt;
  }
}
template <int NJ> struct DynOps {
  static constexpr int SCR = NJ <= 3 ? 1 : ABA_SCR_PER_BODY * NJ;  // scratch doubles per thread
  DDPB_HD static void eval(const ModelDev& m, double dt, const double* x, const double* u, double* xn, const Scratch& sc) { dyn_eval<NJ>(m, dt, x, u, xn, sc); }
  DDPB_HD static void frame(const ModelDev& m, const double* q, int frame, double* p3) { frame_task<NJ>(m, q, frame, 0, p3, nullptr); }
};
// pendulum: NJ = 0 selects the analytic model
template <> struct DynOps<0> {
  static constexpr int SCR = 1;
  DDPB_HD static void eval(const ModelDev& m, double dt, const double* x, const double* u, double* xn, const Scratch&) { pend_eval(m, dt, x, u[0], xn); }
  DDPB_HD static void frame(const ModelDev&, const double*, int, double* p3) { p3[0] = p3[1] = p3[2] = 0.0; }
};
template <int NJ> struct Dim { static constexpr int n = NJ == 0 ? 1 : NJ; };

// equality value of knot t given x_{t+1} (the doubly advanced constraint only sees q_{t+2} = q_{t+1} + dt v_{t+1})
template <int NJ> DDPB_HD void eq_value(const ModelDev& m, const BatchDev& bd, int b, int t, const double* xn, double* ce) {
  constexpr int n = Dim<NJ>::n;
  const int ne = bd.ne[t];
  double qnn[n];
#pragma unroll
  for (int k = 0; k < n; ++k) { const double vo = bd.dt * xn[n + k]; qnn[k] = xn[k] + vo; }
  if (bd.eq_kind == EQ_CONFIG) {
    for (int r = 0; r < ne; ++r) ce[r] = qnn[r] - bd.targets[((size_t)bd.eq_off[t] + r) * bd.B + b];
  } else {
    double p3[3];
    DynOps<NJ>::frame(m, qnn, bd.frame_id, p3);
    for (int r = 0; r < ne; ++r) ce[r] = p3[r] - bd.targets[((size_t)bd.eq_off[t] + r) * bd.B + b];
  }
}

// one term of cost_seq_aug (ddp.hpp:727-731): l + pe(x) . ce + mu/2 |ce|^2
template <int NJ>
DDPB_HD double merit_term(const ModelDev& m, const BatchDev& bd, int b, int t, const double* x, const double* u, const double* xn, double mu) {
  constexpr int n = Dim<NJ>::n, nd = 2 * n, nu = n, nx = 2 * n;
  const int B = bd.B, ne = bd.ne[t];
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < nu; ++k) s += u[k] * u[k];
  const double l = 0.5 * bd.c * s;
  if (ne == 0) return l + 0.0 + (mu / 2.0) * 0.0;
  double ce[MAXE], dot = 0.0, sq = 0.0;
  eq_value<NJ>(m, bd, b, t, xn, ce);
  const size_t eo = bd.eq_off[t];
  for (int i = 0; i < ne; ++i) {
    double pe = bd.m_val[(eo + i) * B + b];
    double a = 0.0;
#pragma unroll
    for (int k = 0; k < nd; ++k) a += bd.m_jac[(eo * nd + k * ne + i) * B + b] * (x[k] - bd.m_origin[((size_t)t * nx + k) * B + b]);
    pe += a;
    dot += pe * ce[i];
    sq += ce[i] * ce[i];
  }
  return l + dot + (mu / 2.0) * sq;
}

// make_trajectory: roll the stored controls out from x_0 (xs[0] must be set)
template <int NJ> DDPB_HD void rollout_problem(const ModelDev& m, const BatchDev& bd, int b, const Scratch& sc) {
  constexpr int n = Dim<NJ>::n, nx = 2 * n, nu = n;
  const int B = bd.B;
  double x[nx], u[nu], xn[nx];
#pragma unroll
  for (int k = 0; k < nx; ++k) x[k] = bd.xs[(size_t)k * B + b];
  for (int t = 0; t < bd.T; ++t) {
#pragma unroll
    for (int k = 0; k < nu; ++k) u[k] = bd.us[((size_t)t * nu + k) * B + b];
    DynOps<NJ>::eval(m, bd.dt, x, u, xn, sc);
#pragma unroll
    for (int k = 0; k < nx; ++k) { x[k] = xn[k]; bd.xs[((size_t)(t + 1) * nx + k) * B + b] = xn[k]; }
  }
}

// one term of cost_seq_aug of the current trajectory: costs[t][b] (t == T: the terminal cost lf = 0)
template <int NJ> DDPB_HD void cost_term_knot(const ModelDev& m, const BatchDev& bd, int b, int t, double* costs, double mu, const Scratch& sc) {
  constexpr int n = Dim<NJ>::n, nx = 2 * n, nu = n;
  const int B = bd.B;
  if (t >= bd.T) { costs[(size_t)bd.T * B + b] = 0.0; return; }
  double x[nx], u[nu], xn[nx];
#pragma unroll
  for (int k = 0; k < nx; ++k) x[k] = bd.xs[((size_t)t * nx + k) * B + b];
#pragma unroll
  for (int k = 0; k < nu; ++k) u[k] = bd.us[((size_t)t * nu + k) * B + b];
  // the reference re-evaluates f(x_t,u_t) inside eval_eq_to; only needed where the knot has rows
  if (bd.ne[t] > 0) DynOps<NJ>::eval(m, bd.dt, x, u, xn, sc);
  costs[(size_t)t * B + b] = merit_term<NJ>(m, bd, b, t, x, u, xn, mu);
}

// cost_seq_aug of the current trajectory into costs[(T+1)][B]; returns the plain cost sum_t l
template <int NJ> DDPB_HD double cost_seq_problem(const ModelDev& m, const BatchDev& bd, int b, double* costs, double mu, const Scratch& sc) {
  constexpr int n = Dim<NJ>::n, nx = 2 * n, nu = n;
  const int B = bd.B;
  double plain = 0.0;
  for (int t = 0; t < bd.T; ++t) {
    double x[nx], u[nu], xn[nx];
#pragma unroll
    for (int k = 0; k < nx; ++k) x[k] = bd.xs[((size_t)t * nx + k) * B + b];
#pragma unroll
    for (int k = 0; k < nu; ++k) u[k] = bd.us[((size_t)t * nu + k) * B + b];
    // the reference re-evaluates f(x_t,u_t) inside eval_eq_to; only needed where the knot has rows
    if (bd.ne[t] > 0) DynOps<NJ>::eval(m, bd.dt, x, u, xn, sc);
    costs[(size_t)t * B + b] = merit_term<NJ>(m, bd, b, t, x, u, xn, mu);
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < nu; ++k) s += u[k] * u[k];
    plain += 0.5 * bd.c * s;
  }
  costs[(size_t)bd.T * B + b] = 0.0;  // lf = 0
  return plain;
}

DDPB_HD void l2_prefetch(const void* p) {
#if defined(__CUDA_ARCH__)
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}
// rows of knot t that a rollout step reads (x_t, u_t, k_t, K_t, old cost) -> L2, one request per row and warp
template <int NJ> DDPB_HD void prefetch_knot(const BatchDev& bd, int b, int t, const double* costs_old) {
  constexpr int n = Dim<NJ>::n, nx = 2 * n, nu = n, nd = 2 * n;
  const int B = bd.B;
  if (t >= bd.T) return;
#pragma unroll
  for (int k = 0; k < nx; ++k) l2_prefetch(bd.xs + ((size_t)t * nx + k) * B + b);
#pragma unroll
  for (int k = 0; k < nu; ++k) { l2_prefetch(bd.us + ((size_t)t * nu + k) * B + b); l2_prefetch(bd.fb_k + ((size_t)t * nu + k) * B + b); }
#pragma unroll
  for (int k = 0; k < nu * nd; ++k) l2_prefetch(bd.fb_K + ((size_t)t * nu * nd + k) * B + b);
  l2_prefetch(costs_old + (size_t)t * B + b);
}

// one trial of the line search (ddp_fwd.ipp:39-56): closed-loop rollout with step size `step` from x_0, written
// to (xs_out, us_out) (same SoA layout as the trajectory), and sum_t (cost_new - cost_old) in the reference's order
template <int NJ>
DDPB_HD void rollout_candidate(const ModelDev& m, const BatchDev& bd, int b, double step, double mu, const double* costs_old, double* xs_out,
                               double* us_out, double* sum_out, double* merit_out, double* plain_out, const Scratch& sc, bool do_prefetch = false) {
  constexpr int n = Dim<NJ>::n, nx = 2 * n, nu = n, nd = 2 * n;
  const int B = bd.B, T = bd.T;
  double x[nx], xn[nx], u[nu];
#pragma unroll
  for (int k = 0; k < nx; ++k) x[k] = bd.xs_new[(size_t)k * B + b];  // x_0 of the new-trajectory storage (= x_0)
  double sum = 0.0, merit = 0.0, plain = 0.0;
  // The rows a step reads (x_t, u_t, k_t, K_t, old cost) do not depend on the rollout itself.  Small models keep the
  // NEXT knot's rows in registers (loaded one knot ahead); larger ones send the rows of knot t+2 to L2 while knot t
  // is evaluated.
  constexpr bool REGPRE = Dim<NJ>::n <= 3;
  double r_x[nd], r_K[nu * nd], r_k[nu], r_u[nu], r_c = 0.0;
  auto load_rows = [&](int t) {
#pragma unroll
    for (int k = 0; k < nd; ++k) r_x[k] = bd.xs[((size_t)t * nx + k) * B + b];
#pragma unroll
    for (int k = 0; k < nu * nd; ++k) r_K[k] = bd.fb_K[((size_t)t * nu * nd + k) * B + b];
#pragma unroll
    for (int k = 0; k < nu; ++k) { r_k[k] = bd.fb_k[((size_t)t * nu + k) * B + b]; r_u[k] = bd.us[((size_t)t * nu + k) * B + b]; }
    r_c = costs_old[(size_t)t * B + b];
  };
  if (REGPRE) load_rows(0);
  else if (do_prefetch) prefetch_knot<NJ>(bd, b, 1, costs_old);
  for (int t = 0; t < T; ++t) {
    if (!REGPRE) {
      if (do_prefetch) prefetch_knot<NJ>(bd, b, t + 2, costs_old);
      load_rows(t);
    }
    double dx[nd], c_old = r_c;
#pragma unroll
    for (int k = 0; k < nd; ++k) dx[k] = x[k] - r_x[k];
#pragma unroll
    for (int i = 0; i < nu; ++i) {
      double a = 0.0;
#pragma unroll
      for (int k = 0; k < nd; ++k) a += r_K[k * nu + i] * dx[k];
      u[i] = r_u[i] + step * r_k[i] + a;
      us_out[((size_t)t * nu + i) * B + b] = u[i];
    }
    if (REGPRE && t + 1 < T) load_rows(t + 1);  // in flight during the dynamics of knot t
    DynOps<NJ>::eval(m, bd.dt, x, u, xn, sc);
    const double cn = merit_term<NJ>(m, bd, b, t, x, u, xn, mu);
    sum += cn - c_old;
    merit += cn;
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < nu; ++k) s += u[k] * u[k];
    plain += 0.5 * bd.c * s;
#pragma unroll
    for (int k = 0; k < nx; ++k) { x[k] = xn[k]; xs_out[((size_t)(t + 1) * nx + k) * B + b] = xn[k]; }
  }
  sum += 0.0 - costs_old[(size_t)T * B + b];
  *sum_out = sum; *merit_out = merit; *plain_out = plain;
}

// forward_pass with sequential step halving for one problem (the CPU-emulated statement of the line search; the
// device runs the same trials concurrently, see linesearch_block in outer_warp.cuh)
template <int NJ> DDPB_HD void forward_problem(const ModelDev& m, const BatchDev& bd, int b, double* costs_old) {
  constexpr int n = Dim<NJ>::n, nx = 2 * n, nu = n;
  const int B = bd.B, T = bd.T;
  if (bd.scal[(size_t)S_STATUS * B + b] != 0.0) {
    // frozen problem: keep its trajectory through the pointer swap
    for (int t = 0; t <= T; ++t)
      for (int k = 0; k < nx; ++k) bd.xs_new[((size_t)t * nx + k) * B + b] = bd.xs[((size_t)t * nx + k) * B + b];
    for (int t = 0; t < T; ++t)
      for (int k = 0; k < nu; ++k) bd.us_new[((size_t)t * nu + k) * B + b] = bd.us[((size_t)t * nu + k) * B + b];
    return;
  }
  const double mu = bd.scal[(size_t)S_MU * B + b];
  double scr_[DynOps<NJ>::SCR];
  const Scratch sc{scr_, 1};
  cost_seq_problem<NJ>(m, bd, b, costs_old, mu, sc);
  double step = 1.0;
  bool success = false;
  int rollouts = 0;
  double merit = 0.0, plain = 0.0;
  while (!success) {
    if (step < 1e-10) break;
    ++rollouts;
    double sum;
    rollout_candidate<NJ>(m, bd, b, step, mu, costs_old, bd.xs_new, bd.us_new, &sum, &merit, &plain, sc);
    if (sum <= 0.0) success = true;
    else step *= 0.5;
  }
  bd.scal[(size_t)S_STEP * B + b] = step;
  bd.scal[(size_t)S_ROLLOUTS * B + b] += (double)rollouts;
  bd.scal[(size_t)S_MERIT * B + b] = merit;
  bd.scal[(size_t)S_COST * B + b] = plain;
}

// ---- outer-loop sweeps ------------------------------------------------------------------------------
// adjoint recursion of optimality_obj (ddp.hpp:575-627) for one problem (mults already re-centred: pe = val)
template <int NJ> DDPB_HD double optimality_obj_problem(const BatchDev& bd, int b, double mu) {
  constexpr int n = Dim<NJ>::n, nx = 2 * n, nu = n, nd = 2 * n;
  const int B = bd.B;
  double adj[nd], nadj[nd];
#pragma unroll
  for (int k = 0; k < nd; ++k) adj[k] = 0.0;  // lfx = 0
  double retval = 0.0;
  for (int t = bd.T - 1; t >= 0; --t) {
    const int ne = bd.ne[t];
    const size_t eo = bd.eq_off[t];
    const RecLayout L = rec_layout(n, ne, bd.hess_mode);
    const RecView rec = rec_view(bd, t, b, L.size);
    double pe[MAXE];
    for (int i = 0; i < ne; ++i) {
      // p_eq(pe, x_t): val + jac (x_t - origin)
      double v = bd.m_val[(eo + i) * B + b], a = 0.0;
#pragma unroll
      for (int k = 0; k < nd; ++k)
        a += bd.m_jac[(eo * nd + k * ne + i) * B + b] * (bd.xs[((size_t)t * nx + k) * B + b] - bd.m_origin[((size_t)t * nx + k) * B + b]);
      pe[i] = v + a;
    }
    double sq = 0.0;
#pragma unroll
    for (int i = 0; i < nu; ++i) {
      double v = rec[L.lu + i], a = 0.0;
      for (int k = 0; k < ne; ++k) a += pe[k] * rec[L.eq_u + i * ne + k];
      v += a;
      a = 0.0;
      for (int k = 0; k < ne; ++k) a += (mu * rec[L.eq + k]) * rec[L.eq_u + i * ne + k];
      v += a;
      a = 0.0;  // adj f_u: rows k < n of f_u are structural zeros (batch.cuh)
#pragma unroll
      for (int k = 0; k < n; ++k) a += adj[n + k] * rec[L.fu + i * n + k];
      v += a;
      sq += v * v;
    }
    const double nrm = sqrt(sq);
    if (retval < nrm) retval = nrm;
#pragma unroll
    for (int i = 0; i < nd; ++i) {
      // adj f_x: rows k < n of f_x are the constants [I, dt I] (batch.cuh); l_x = 0
      double a = i < n ? adj[i] : adj[i - n] * bd.dt;
#pragma unroll
      for (int k = 0; k < n; ++k) a += adj[n + k] * rec[L.fx + i * n + k];
      double v = a;
      a = 0.0;
      for (int k = 0; k < ne; ++k) a += (mu * rec[L.eq + k]) * rec[L.eq_x + i * ne + k];
      v += a;
      a = 0.0;
      for (int k = 0; k < ne; ++k) a += pe[k] * rec[L.eq_x + i * ne + k];
      v += a;
      a = 0.0;
      for (int k = 0; k < ne; ++k) a += rec[L.eq + k] * bd.m_jac[(eo * nd + i * ne + k) * B + b];
      v += a;
      nadj[i] = v;
    }
#pragma unroll
    for (int k = 0; k < nd; ++k) adj[k] = nadj[k];
  }
  return retval;
}
template <int NJ> DDPB_HD double optimality_constr_problem(const BatchDev& bd, int b) {
  constexpr int n = Dim<NJ>::n, nx = 2 * n, nu = n, nd = 2 * n;
  double r = 0.0;
  for (int t = 0; t < bd.T; ++t) {
    const int ne = bd.ne[t];
    if (ne == 0) continue;
    const RecLayout L = rec_layout(n, ne, bd.hess_mode);
    const RecView rec = rec_view(bd, t, b, L.size);
    double s = 0.0;
    for (int k = 0; k < ne; ++k) s += rec[L.eq + k] * rec[L.eq + k];
    const double nrm = sqrt(s);
    if (r < nrm) r = nrm;
  }
  return r;
}

// update_derivatives after compute_derivatives (ddp.hpp:657-695) + the switch of solve (ddp.hpp:786-802)
template <int NJ> DDPB_HD void al_update_problem(const BatchDev& bd, int b, double thr) {
  constexpr int n = Dim<NJ>::n, nx = 2 * n, nu = n, nd = 2 * n;
  const int B = bd.B, T = bd.T;
  if (bd.scal[(size_t)S_STATUS * B + b] != 0.0) return;
  // mults.eq.update_origin(traj.x) ; fb_seq.update_origin(traj.x)
  for (int t = 0; t < T; ++t) {
    const int ne = bd.ne[t];
    const size_t eo = bd.eq_off[t];
    double dxm[nd], dxf[nd];
#pragma unroll
    for (int k = 0; k < nd; ++k) {
      const double xv = bd.xs[((size_t)t * nx + k) * B + b];
      dxm[k] = xv - bd.m_origin[((size_t)t * nx + k) * B + b];
      dxf[k] = xv - bd.fb_origin[((size_t)t * nx + k) * B + b];
      bd.m_origin[((size_t)t * nx + k) * B + b] = xv;
      bd.fb_origin[((size_t)t * nx + k) * B + b] = xv;
    }
    for (int i = 0; i < ne; ++i) {
      double a = 0.0;
#pragma unroll
      for (int k = 0; k < nd; ++k) a += bd.m_jac[(eo * nd + k * ne + i) * B + b] * dxm[k];
      bd.m_val[(eo + i) * B + b] += a;
    }
#pragma unroll
    for (int i = 0; i < nu; ++i) {
      double a = 0.0;
#pragma unroll
      for (int k = 0; k < nd; ++k) a += bd.fb_K[((size_t)t * nu * nd + k * nu + i) * B + b] * dxf[k];
      bd.fb_k[((size_t)t * nu + i) * B + b] += a;
    }
  }
  double mu = bd.scal[(size_t)S_MU * B + b];
  double w = bd.scal[(size_t)S_W * B + b];
  double nn = bd.scal[(size_t)S_N * B + b];
  const double opt_obj = optimality_obj_problem<NJ>(bd, b, mu);
  const double opt_constr = optimality_constr_problem<NJ>(bd, b);
  bd.scal[(size_t)S_OPT_OBJ * B + b] = opt_obj;
  bd.scal[(size_t)S_OPT_CONSTR * B + b] = opt_constr;
  if (opt_constr < thr && opt_obj < thr) {
    bd.scal[(size_t)S_STATUS * B + b] = 1.0;  // optimum attained
    return;
  }
  if (opt_obj < w) {
    if (opt_constr < nn) {
      // p += mu (eq + eq_u k) ; P += mu (eq_x + eq_u K)   (ddp.hpp:686-687)
      for (int t = 0; t < T; ++t) {
        const int ne = bd.ne[t];
        if (ne == 0) continue;
        const size_t eo = bd.eq_off[t];
        const RecLayout L = rec_layout(n, ne, bd.hess_mode);
        const RecView rec = rec_view(bd, t, b, L.size);
        for (int i = 0; i < ne; ++i) {
          double a = 0.0;
#pragma unroll
          for (int k = 0; k < nu; ++k) a += rec[L.eq_u + k * ne + i] * bd.fb_k[((size_t)t * nu + k) * B + b];
          bd.m_val[(eo + i) * B + b] += mu * (rec[L.eq + i] + a);
        }
        for (int c = 0; c < nd; ++c)
          for (int i = 0; i < ne; ++i) {
            double a = 0.0;
#pragma unroll
            for (int k = 0; k < nu; ++k) a += rec[L.eq_u + k * ne + i] * bd.fb_K[((size_t)t * nu * nd + c * nu + k) * B + b];
            bd.m_jac[(eo * nd + c * ne + i) * B + b] += mu * (rec[L.eq_x + c * ne + i] + a);
          }
      }
      const double oo = optimality_obj_problem<NJ>(bd, b, mu);
      nn = oo / pow(mu, 0.1);
      w /= mu;
      bd.scal[(size_t)S_MULT_UPDATES * B + b] += 1.0;
    } else {
      mu *= 10.0;
    }
  }
  bd.scal[(size_t)S_MU * B + b] = mu;
  bd.scal[(size_t)S_W * B + b] = w;
  bd.scal[(size_t)S_N * B + b] = nn;
}

// after forward_pass inside the loop (ddp.hpp:819-826)
DDPB_HD void post_forward_problem(const BatchDev& bd, int b) {
  const int B = bd.B;
  if (bd.scal[(size_t)S_STATUS * B + b] != 0.0) return;
  double reg = bd.scal[(size_t)S_REG * B + b];
  if (bd.max_restarts > 0) {
    const double merit = bd.scal[(size_t)S_MERIT * B + b], mu = bd.scal[(size_t)S_MU * B + b];
    if (!(merit == merit) || !(fabs(merit) < 1.79e308) || !(mu < 1.79e308)) { bd.scal[(size_t)S_STATUS * B + b] = 2.0; return; }
  }
  if (bd.scal[(size_t)S_STEP * B + b] >= 0.5) {
    reg /= 2.0;
    if (reg < 1e-5) reg = 0.0;
  }
  bd.scal[(size_t)S_REG * B + b] = reg;
  bd.scal[(size_t)S_ITERATIONS * B + b] += 1.0;
}

DDPB_VARIANT_END
