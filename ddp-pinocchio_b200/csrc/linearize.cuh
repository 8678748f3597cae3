// linearize.cuh -- per-knot linearization: one CTA computes the complete derivative records of a
// group of (problem, knot) pairs of one time index: cost derivatives, Euler-discretised dynamics value / Jacobians / finite-
// difference Hessians, and the doubly time-advanced equality constraint with its derivatives.
//
// Restates problem_t::compute_derivatives (/root/reference/include/ddp/problem.hpp:956-997) with
//   dynamics_t::eval_to / first_order_deriv            (problem.hpp:441-503)
//   finite_diff_hessian_compute::second_order_deriv_1  (problem.hpp:67-150, eps = sqrt(eps_mach))
//   finite_diff_hessian_compute::second_order_deriv_2  (problem.hpp:152-298, eps = eps_mach^(1/4))
//   constraint_advance_time_t o constraint_advance_time_t o {config,spatial}_constraint_t
//                                                      (problem.hpp:527-864)
// The finite-difference stencils are evaluated on exactly the reference's points, but work that the
// reference repeats is shared: all evaluation points with the same q share one q_task (articulated
// inertias, M^-1), and the advanced constraint reuses the dynamics already evaluated at the point
// (c depends on q_{t+2} = q_{t+1} + dt v_{t+1} only, so the second ABA of the reference's
// eval_to chain never influences the result).
#pragma once
#include "batch.cuh"

// development switches (geometry sweeps, tools/lin_variants.sh)
#ifndef DDPB_COL_MERGED
#define DDPB_COL_MERGED -1  // both derivative kinds of a column in one task: -1 = by model size; 0 / 1 force one form
#endif
#ifndef DDPB_FUSE_PROD
#define DDPB_FUSE_PROD 1    // M^-1 products inside the column tasks where M^-1 is ready by then
#endif

DDPB_VARIANT_BEGIN

DDPB_HD int pair_rank(int i, int j, int N) { return i * N - (i * (i + 1)) / 2 + (j - i - 1); }
DDPB_HD void pair_unrank(int r, int N, int* i, int* j) {
  int ii = 0;
  while (r >= N - 1 - ii) { r -= N - 1 - ii; ++ii; }
  *i = ii; *j = ii + 1 + r;
}

template <int NJ> struct LinPlan {
  static constexpr int N = 3 * NJ;
  int mode, P, NQ, NT;  // evaluation points, distinct-q sets, points that need first-order data
  int ne;               // equality rows at this knot
  bool u_points;        // whether the u-perturbed evaluation points exist
  // shared-memory offsets (doubles)
  // Shared memory of a CTA that owns kpb knots: [kpb][NQ] q-records | [kpb][NT] v-records | [kpb] everything else.  The
  // two record arrays are region-major with one odd stride each, so the tasks of a phase (consecutive threads ->
  // consecutive records, whatever knot they belong to) hit different banks.
  int o_base, o_tgt, o_fval, o_eqv, o_J, o_eqxu, o_cterm, o_dg, o_minv0, o_qdd, q_sz, v_sz, rest, total;
  // value-based mode: the evaluation points whose perturbation touches no control index (the only ones that run the
  // articulated-body sweeps, see "torque-affine points" in linearize_group), ND = 2 NJ tangent indices of the state
  static constexpr int ND = 2 * NJ;
  static constexpr int PA = 1 + ND + ND * (ND - 1) / 2;
  DDPB_HD static int aba_point(int a, int* i, int* j) {  // a-th such point -> its index among all P points, its perturbation
    *i = -1; *j = -1;
    if (a == 0) return 0;
    if (a <= ND) { *i = a - 1; return a; }
    pair_unrank(a - 1 - ND, ND, i, j);  // (state tangent indices are the same numbers in both enumerations)
    return 1 + N + pair_rank(*i, *j, N);
  }
  // stride of the q-records (M^-1 aliases the articulated inertias inside a record; the value-based mode needs M^-1 of
  // the base and the singly perturbed configurations only, while its value-only points still read IA, so those live in
  // their own slots, o_minv0)
  DDPB_HD static int qstride(int) { return QRec<NJ>::SIZE; }
  DDPB_HD static int points(int mode) { return mode == HESS_FD_JACOBIANS ? 1 + N : 1 + N + N * (N - 1) / 2; }
  DDPB_HD static int qsets(int mode) { return mode == HESS_FD_JACOBIANS ? 1 + NJ : 1 + NJ + NJ * (NJ - 1) / 2; }
  // FD-of-Jacobians mode: the reference evaluates (f_x, f_u) at the u-perturbed points only to difference f_u
  // (problem.hpp:138-145), and f_u = dt M^-1(q) cannot change when only u moves: f_uu (and eq_uu of a configuration
  // constraint) are exact zeros there too.  Those N/3 points are therefore dropped unless a spatial constraint needs
  // the frame Jacobian at their q_{t+2}.
  DDPB_HD void init(int mode_, int ne_max, int ne_, int eq_kind) {
    mode = mode_; ne = ne_;
    P = points(mode); NQ = qsets(mode);
    u_points = !(mode == HESS_FD_JACOBIANS && !(eq_kind == EQ_SPATIAL && ne_ > 0));
    if (!u_points) P = 1 + 2 * NJ;
    NT = mode == HESS_FD_JACOBIANS ? P : 1;
    int o = 0;
    o_base = o; o += 4 * NJ;
    o_tgt = o; o += (ne_max + 1) & ~1;
    q_sz = NQ * qstride(mode);
    v_sz = NT * VRec<NJ>::SIZE;
    o_fval = o; o += P * 2 * NJ;
    o_eqv = o; o += P * ne_max;
    o_J = o; o += (eq_kind == EQ_SPATIAL && ne_max > 0) ? NT * 3 * NJ : 0;  // frame Jacobians of a spatial constraint (knots with rows only)
    o_eqxu = o; o += NT * ne_max * N;
    // value-based mode only: per-body column terms of the base point, diagonal second differences of f and eq
    o_cterm = o; o += mode == HESS_FD_VALUES ? 2 * NJ * NJ * 6 : 0;
    o_dg = o; o += mode == HESS_FD_VALUES ? 2 * N * (2 * NJ + ne_max) : 0;  // second differences | Jacobian entries
    o_minv0 = o; o += mode == HESS_FD_VALUES ? (1 + NJ) * NJ * NJ : 0;  // M^-1 of the base configuration and of the NJ singly perturbed ones
    o_qdd = o; o += mode == HESS_FD_VALUES ? (1 + ND) * NJ : 0;          // accelerations at the base point and the ND state-perturbed single points
    rest = o;
    total = q_sz + v_sz + rest;
  }
  // which tangent index/indices a point perturbs (-1 = none)
  DDPB_HD void point_pert(int p, int* i, int* j) const {
    *i = -1; *j = -1;
    if (p == 0) return;
    if (p <= N) { *i = p - 1; return; }
    pair_unrank(p - 1 - N, N, i, j);
  }
  DDPB_HD static int pert_qset(int i, int j) {  // configuration of the point that perturbs tangent indices i < j (-1 = none)
    if (i < 0 || i >= NJ) return 0;          // i < j, so j >= NJ as well
    if (j < 0 || j >= NJ) return 1 + i;
    return 1 + NJ + pair_rank(i, j, NJ);
  }
  DDPB_HD int point_qset(int p) const {
    int i, j;
    point_pert(p, &i, &j);
    return pert_qset(i, j);
  }
  DDPB_HD void qset_pert(int s, int* i, int* j) const {
    *i = -1; *j = -1;
    if (s == 0) return;
    if (s <= NJ) { *i = s - 1; return; }
    pair_unrank(s - 1 - NJ, NJ, i, j);
  }
  DDPB_HD double eps() const { return mode == HESS_FD_JACOBIANS ? 1.4901161193847656e-08 /* 2^-26 */ : 1.220703125e-04 /* 2^-13 */; }
};

// One CTA linearizes `nk` consecutive problems b0 .. b0+nk-1 of knot t (same t => same number of equality
// rows, adjacent records in HBM).  `ne_sm` is the row count the shared-memory plan was sized for (>= ne(t)).
// MODE = HESS_FD_JACOBIANS / HESS_FD_VALUES at compile time: the two modes are separate kernels, so the register
// allocation of one does not pay for the code paths of the other.
template <int NJ, int MODE, class Ex>
DDPB_HD void linearize_group(const ModelDev& m, const BatchDev& bd, int t, int b0, int nk, int kpb, int ne_sm, double* sm_all, const Ex& ex) {
  constexpr int n = NJ, nd = 2 * NJ, nu = NJ, nx = 2 * NJ, N = 3 * NJ;
  using Q = QRec<NJ>;
  using V = VRec<NJ>;
  const int B = bd.B;
  const int ne = bd.ne[t];
  LinPlan<NJ> pl;
  pl.init(MODE, ne_sm, ne, bd.eq_kind);
  const double eps = pl.eps();
  const double dt = bd.dt;
  const int tid = ex.tid(), nthr = ex.nthr();
  double* const qbase = sm_all;                                        // [kpb][NQ] records of QS doubles
  double* const vbase = sm_all + (size_t)kpb * pl.q_sz;                // [kpb][NT] records of V::SIZE doubles
  double* const rbase = vbase + (size_t)kpb * pl.v_sz;                 // [kpb] x pl.rest
  constexpr int QS = Q::SIZE;  // stride between q-records (LinPlan::qstride)
  // M^-1 of q-set s of knot kk (value-based mode: the base configuration's only)
#define SM(kk, off) (rbase + (size_t)(kk) * pl.rest + (off))
#define QR(kk, s) (qbase + ((size_t)(kk) * pl.NQ + (s)) * QS)
#define VR(kk, p) (vbase + ((size_t)(kk) * pl.NT + (p)) * V::SIZE)
  auto minv_of = [&](int kk, int s) -> double* { return MODE == HESS_FD_VALUES ? SM(kk, pl.o_minv0) : QR(kk, s) + Q::MINV; };

  DDPB_PHASE_START(bd, pc_);
  // phase 0: stage x_t, u_t and the target
  for (int idx = tid; idx < nk * (N + ne); idx += nthr) {
    const int kk = idx / (N + ne), e = idx - kk * (N + ne), b = b0 + kk;
    if (e < nx) SM(kk, pl.o_base)[e] = bd.xs[((size_t)t * nx + e) * B + b];
    else if (e < N) SM(kk, pl.o_base)[e] = bd.us[((size_t)t * nu + (e - nx)) * B + b];
    else SM(kk, pl.o_tgt)[e - N] = bd.targets[((size_t)bd.eq_off[t] + (e - N)) * B + b];
  }
  ex.sync();
  DDPB_PHASE_MARK(bd, 0, pc_);

  // state of evaluation point p of knot kk: the base (x, u) with its tangent perturbation(s)
  auto pert_state = [&](int kk, int i, int j, double* q, double* qd, double* tau) {
    const double* base = SM(kk, pl.o_base);
#pragma unroll
    for (int k = 0; k < NJ; ++k) {
      q[k] = base[k] + ((k == i || k == j) ? eps : 0.0);
      qd[k] = base[n + k] + ((n + k == i || n + k == j) ? eps : 0.0);
      tau[k] = base[2 * n + k] + ((2 * n + k == i || 2 * n + k == j) ? eps : 0.0);
    }
  };
  auto point_state = [&](int kk, int p, double* q, double* qd, double* tau) {
    int i, j;
    pl.point_pert(p, &i, &j);
    pert_state(kk, i, j, q, qd, tau);
  };
  // Short tasks (<= 3 joints) run each narrow phase as one task per configuration / point; longer chains split them
  // into parallel and serial sub-phases (more barriers, much shorter serial parts).
  constexpr bool SPLIT = NJ >= 4;
  constexpr bool COL_MERGED = DDPB_COL_MERGED < 0 ? NJ <= 3 : DDPB_COL_MERGED != 0;
  if (!SPLIT) {
    for (int idx = tid; idx < nk * pl.NQ; idx += nthr) {
      const int kk = idx / pl.NQ, s = idx - kk * pl.NQ;
      const double* base = SM(kk, pl.o_base);
      int i, j;
      pl.qset_pert(s, &i, &j);
      double q[NJ];
#pragma unroll
      for (int k = 0; k < NJ; ++k) q[k] = base[k] + ((k == i || k == j) ? eps : 0.0);
      q_task<NJ>(m, q, QR(kk, s));
    }
  } else {
    // phase 1: q-only data for every distinct configuration, as four sub-phases (rbd.cuh "q_task"): the local joint
    // rotations and the per-body world inertias are parallel over (configuration, joint / body); only the kinematic chain
    // and the articulated-inertia recursion are serial per configuration
    // (record index fastest over the lanes in the per-(record, body) loops: records have odd strides, bodies even ones)
    for (int idx = tid; idx < nk * pl.NQ * NJ; idx += nthr) {
      const int jn = idx / (nk * pl.NQ), r = idx - jn * (nk * pl.NQ), kk = r / pl.NQ, s = r - kk * pl.NQ;
      int i, j;
      pl.qset_pert(s, &i, &j);
      const double qj = SM(kk, pl.o_base)[jn] + ((jn == i || jn == j) ? eps : 0.0);
      double Rl[9], plc[3];
      joint_local_rot(m, jn, qj, Rl);
      joint_local_trans(m, jn, qj, plc);
      double* dst = QR(kk, s) + Q::SCR + QScratch<NJ>::RL + 12 * jn;
  #pragma unroll
      for (int k = 0; k < 9; ++k) dst[k] = Rl[k];
      if (HAS_PRISMATIC) {
  #pragma unroll
        for (int k = 0; k < 3; ++k) dst[9 + k] = plc[k];
      }
    }
    ex.sync();
    DDPB_PHASE_MARK(bd, 8, pc_);
    for (int idx = tid; idx < nk * pl.NQ; idx += nthr) {
      const int kk = idx / pl.NQ, s = idx - kk * pl.NQ;
      q_chain<NJ>(m, QR(kk, s));
    }
    ex.sync();
    DDPB_PHASE_MARK(bd, 9, pc_);
    for (int idx = tid; idx < nk * pl.NQ * NJ; idx += nthr) {
      const int jn = idx / (nk * pl.NQ), r = idx - jn * (nk * pl.NQ), kk = r / pl.NQ, s = r - kk * pl.NQ;
      q_body<NJ>(m, jn, QR(kk, s));
    }
    ex.sync();
    DDPB_PHASE_MARK(bd, 10, pc_);
    for (int idx = tid; idx < nk * pl.NQ; idx += nthr) {
      const int kk = idx / pl.NQ, s = idx - kk * pl.NQ;
      q_articulate<NJ>(QR(kk, s));
    }
    // the body velocities of every recorded point (v_prefix: S and qd only) are as serial and as narrow as the
    // articulated-inertia recursion and independent of it: they run beside it, on threads taken from the top of the CTA
    for (int idx = nthr - 1 - tid; idx < nk * pl.NT; idx += nthr) {
      const int kk = idx / pl.NT, p = idx - kk * pl.NT;
      double q[NJ], qd[NJ], tau[NJ];
      point_state(kk, p, q, qd, tau);
      v_prefix<NJ>(QR(kk, pl.point_qset(p)), qd, VR(kk, p));
    }
  }
  ex.sync();
  DDPB_PHASE_MARK(bd, 1, pc_);

  // phase 2: accelerations (and body data where derivative columns are needed) at every evaluation point.  Points with
  // a record (all of them in FD-of-Jacobians mode, the base point otherwise) run as sub-phases (rbd.cuh "v_task as
  // sub-phases"); value-only points run the whole of v_task<false> in the serial sub-phase.
  // value-based mode: M^-1 of the base configuration and of the NJ singly perturbed ones (S, U, 1/D only; what the
  // torque-affine points below need), threads taken from the top of the CTA
  auto minv_value_tasks = [&]() {
    for (int idx = nthr - 1 - tid; idx < nk * (1 + NJ) * NJ; idx += nthr) {
      const int kk = idx / ((1 + NJ) * NJ), r2 = idx - kk * (1 + NJ) * NJ;
      const int c = r2 % NJ, s = r2 / NJ;
      double col[NJ];
      minv_col<NJ>(QR(kk, s), c, col);
      double* mi = SM(kk, pl.o_minv0) + (size_t)s * NJ * NJ + c * NJ;
#pragma unroll
      for (int k = 0; k < NJ; ++k) mi[k] = col[k];
    }
  };
  if (SPLIT) {
    for (int idx = tid; idx < nk * pl.NT * NJ; idx += nthr) {
      const int jn = idx / (nk * pl.NT), r = idx - jn * (nk * pl.NT), kk = r / pl.NT, p = r - kk * pl.NT;
      int i, j;
      pl.point_pert(p, &i, &j);
      const double qdj = SM(kk, pl.o_base)[n + jn] + ((n + jn == i || n + jn == j) ? eps : 0.0);
      v_bias<NJ>(QR(kk, pl.point_qset(p)), jn, qdj, VR(kk, p));
    }
    // (value-based mode: this phase has NJ bias tasks per knot; the M^-1 columns fill it instead of lengthening the sweeps)
    if (MODE == HESS_FD_VALUES) minv_value_tasks();
    ex.sync();
    DDPB_PHASE_MARK(bd, 12, pc_);
  }
  // Torque-affine points (value-based mode).  The forward dynamics is affine in the torque for fixed (q, v):
  // qdd(q, v, tau + d) = qdd(q, v, tau) + M^-1(q) d.  More than half of the reference's stencil (problem.hpp:192-296:
  // 3 NJ single and 3 NJ (3 NJ - 1) / 2 pair perturbations) touches a control index; those points are not run through the
  // articulated-body sweeps again but derived from the point that has the same state perturbation (the base point or a
  // single state-perturbed point) and the M^-1 columns of its configuration -- the same evaluation points, the same
  // values up to rounding, PA = 1 + 2 NJ + NJ (2 NJ - 1) sweeps per knot instead of P (7R arm: 106 instead of 232, one
  // round of a 128-thread CTA instead of two).
  constexpr bool ULIN = MODE == HESS_FD_VALUES;
  constexpr int PA = LinPlan<NJ>::PA;
  const int n_eval = ULIN ? PA : pl.P;
  for (int idx = tid; idx < nk * n_eval; idx += nthr) {
    const int kk = idx / n_eval, a = idx - kk * n_eval;
    int pi, pj, p = a;  // the point and the tangent indices it perturbs (decoded once)
    if (ULIN) p = LinPlan<NJ>::aba_point(a, &pi, &pj);
    else pl.point_pert(p, &pi, &pj);
    double* fval = SM(kk, pl.o_fval);
    double q[NJ], qd[NJ], tau[NJ], qdd[NJ];
    pert_state(kk, pi, pj, q, qd, tau);
    const double* qr = QR(kk, LinPlan<NJ>::pert_qset(pi, pj));
    if (MODE == HESS_FD_JACOBIANS || p < pl.NT) {  // every point has a record in FD-of-Jacobians mode (NT == P)
      if (SPLIT) v_sweeps<NJ>(m, qr, tau, qdd, VR(kk, p));
      else v_task<NJ, true>(m, qr, qd, tau, qdd, VR(kk, p));
    } else {
      v_task<NJ, false>(m, qr, qd, tau, qdd, nullptr);
    }
    // dynamics_t::eval_to (problem.hpp:441-461): v_out = dt v ; q_out = q + v_out ; v_out = v + acc dt
#pragma unroll
    for (int k = 0; k < NJ; ++k) {
      const double vo = dt * qd[k];
      fval[p * nx + k] = q[k] + vo;
      fval[p * nx + n + k] = qd[k] + qdd[k] * dt;
    }
    if (ULIN && p <= nd) {  // the base point and the single state perturbations are what the control-perturbed points start from
#pragma unroll
      for (int k = 0; k < NJ; ++k) SM(kk, pl.o_qdd)[p * NJ + k] = qdd[k];
    }
  }
  if (ULIN && !SPLIT) minv_value_tasks();  // short chains: beside the sweeps, from the top of the CTA
  // the control-perturbed points of knot kk: d enumerates NJ singles u_j, 2 NJ * NJ pairs (x_i, u_j), NJ (NJ - 1) / 2 pairs (u_i, u_j)
  auto derived_task = [&](int r) {
    constexpr int NDER = NJ + nd * NJ + NJ * (NJ - 1) / 2;
    const int kk = r / NDER, d = r - kk * NDER;
    int i, j;
    if (d < NJ) { i = -1; j = nd + d; }
    else if (d < NJ + nd * NJ) { const int d2 = d - NJ; i = d2 / NJ; j = nd + (d2 - i * NJ); }
    else { pair_unrank(d - NJ - nd * NJ, NJ, &i, &j); i += nd; j += nd; }
    const int p = i < 0 ? 1 + j : 1 + N + pair_rank(i, j, N);
    const int src = (i < 0 || i >= nd) ? 0 : 1 + i;        // the point with the same state perturbation
    const int s = (i >= 0 && i < NJ) ? 1 + i : 0;          // its configuration
    const double* Mi = SM(kk, pl.o_minv0) + (size_t)s * NJ * NJ;
    const double* qdd0 = SM(kk, pl.o_qdd) + src * NJ;
    const double* base = SM(kk, pl.o_base);
    double* fval = SM(kk, pl.o_fval);
#pragma unroll
    for (int k = 0; k < NJ; ++k) {
      double qdd = qdd0[k] + eps * Mi[(j - nd) * NJ + k];
      if (i >= nd) qdd += eps * Mi[(i - nd) * NJ + k];
      const double qdk = base[n + k] + (n + k == i ? eps : 0.0);
      fval[p * nx + k] = fval[src * nx + k];
      fval[p * nx + n + k] = qdk + qdd * dt;
    }
  };
  constexpr int NDER = NJ + nd * NJ + NJ * (NJ - 1) / 2;
  // M^-1 columns need S, U and 1/D of their configuration only, and their destination (aliased onto the articulated
  // inertias, QRec::MINV) has had its last reader in v_bias: in split FD-of-Jacobians mode they run here, on the threads
  // the serial sweeps above leave idle (taken from the top of the CTA), instead of lengthening the column phase.
  constexpr bool MINV_EARLY = SPLIT && MODE == HESS_FD_JACOBIANS;
  auto minv_task = [&](int r) {
    const int n_minv = pl.NQ * NJ;
    const int kk = r / n_minv, r2 = r - kk * n_minv;
    const int c = r2 % NJ, s = r2 / NJ;
    double* qr = QR(kk, s);
    double col[NJ];
    minv_col<NJ>(qr, c, col);
#pragma unroll
    for (int k = 0; k < NJ; ++k) qr[Q::MINV + c * NJ + k] = col[k];
  };
  if (MINV_EARLY)
    for (int idx = nthr - 1 - tid; idx < nk * pl.NQ * NJ; idx += nthr) minv_task(idx);
  if (SPLIT) {
    ex.sync();
    DDPB_PHASE_MARK(bd, 13, pc_);
    for (int idx = tid; idx < nk * pl.NT * NJ; idx += nthr) {
      const int jn = idx / (nk * pl.NT), r = idx - jn * (nk * pl.NT), kk = r / pl.NT, p = r - kk * pl.NT;
      v_force<NJ>(QR(kk, pl.point_qset(p)), jn, VR(kk, p));
    }
    if (ULIN)  // beside the (few) force tasks of the base point, from the top of the CTA
      for (int idx = nthr - 1 - tid; idx < nk * NDER; idx += nthr) derived_task(idx);
    ex.sync();
    DDPB_PHASE_MARK(bd, 14, pc_);
    // (folding these suffix sums into the column tasks was measured slower: 2.63 -> 2.74 ms on C3)
    for (int idx = tid; idx < nk * pl.NT; idx += nthr) {
      const int kk = idx / pl.NT, p = idx - kk * pl.NT;
      v_suffix<NJ>(VR(kk, p));
    }
  } else if (ULIN) {
    ex.sync();
    for (int idx = tid; idx < nk * NDER; idx += nthr) derived_task(idx);
  }
  ex.sync();
  DDPB_PHASE_MARK(bd, 2, pc_);

  // phase 3: tangent columns of inverse dynamics and columns of M^-1
  if (MODE == HESS_FD_VALUES) {
    // only the base point needs columns: too few tasks to fill even one warp, so every column is split by body
    // (col_term, NJ * NJ slots per derivative kind of which k >= j are used) and reduced afterwards (col_reduce)
    // (M^-1 of the base configuration is already there: phase 2)
    for (int idx = tid; idx < nk * 2 * NJ * NJ; idx += nthr) {
      const int kk = idx / (2 * NJ * NJ), r = idx - kk * (2 * NJ * NJ);
      double* qr = QR(kk, 0);
      const int wrt = r / (NJ * NJ), r2 = r - wrt * NJ * NJ, j = r2 / NJ, k = r2 - j * NJ;
      if (k >= j) {
        double g[6];
        col_term<NJ>(qr, VR(kk, 0), j, k, wrt, g);
        double* dst = SM(kk, pl.o_cterm) + (size_t)(wrt * NJ + j) * NJ * 6 + 6 * k;
#pragma unroll
        for (int e = 0; e < 6; ++e) dst[e] = g[e];
      }
    }
    ex.sync();
    for (int idx = tid; idx < nk * 2 * NJ; idx += nthr) {
      const int kk = idx / (2 * NJ), r = idx - kk * 2 * NJ, wrt = r / NJ, j = r - wrt * NJ;
      double col[NJ];
      double* vr = VR(kk, 0);
      col_reduce<NJ>(QR(kk, 0), vr, SM(kk, pl.o_cterm) + (size_t)(wrt * NJ + j) * NJ * 6, j, wrt, col);
      double* dst = vr + (wrt ? V::DV : V::DQ) + j * NJ;
#pragma unroll
      for (int k = 0; k < NJ; ++k) dst[k] = col[k];
    }
    // (the base point's suffix sums beside the column terms and the M^-1 products fused into this reduction -- two barriers
    // fewer per knot -- measured no gain: 7R 18.70 -> 18.80 ms)
  } else {

    // short chains: one tangent task = both derivative kinds of column j at point p (col_task2 loads the body data
    // once: 2R chain 1.62 -> 1.56 ms).  From 4 joints on the two kinds stay separate tasks: together they spill at 128
    // registers (6R: 656 B of spill stores per thread) and measured slower (3.04 -> 3.30 ms).
    constexpr bool MERGED = COL_MERGED;
    constexpr bool FUSE_PROD = DDPB_FUSE_PROD && MINV_EARLY && !MERGED;
    constexpr int KINDS = MERGED ? 1 : 2;
    const int n_tan = pl.NT * KINDS * NJ;
    const int n_minv = MINV_EARLY ? 0 : pl.NQ * NJ;
    // tangent tasks of all knots first (j-major inside a knot), then the M^-1 tasks.  A column costs one term per body
    // k >= j, so the task list runs from expensive to cheap: every other round walks it backwards, which hands the
    // cheap tail to the threads that had the expensive head (critical path ~ the mean instead of head + tail).
    const int n_all = nk * (n_tan + n_minv);
    for (int base = 0, rnd = 0; base < n_all; base += nthr, ++rnd) {
      const int idx = base + ((rnd & 1) ? nthr - 1 - tid : tid);
      if (idx >= n_all) continue;
      if (idx < nk * n_tan) {
        // order: j slowest so that the threads of a warp share the loop bounds of col_task
        const int j = idx / (nk * KINDS * pl.NT), rem = idx - j * nk * KINDS * pl.NT;
        const int kk = rem / (KINDS * pl.NT), rem2 = rem - kk * KINDS * pl.NT;
        const int wrt = rem2 / pl.NT, p = rem2 - wrt * pl.NT;
        double* vr = VR(kk, p);
        if (MERGED) {
          double cq[NJ], cv[NJ];
          col_task2<NJ>(QR(kk, pl.point_qset(p)), vr, j, cq, cv);
#pragma unroll
          for (int k = 0; k < NJ; ++k) { vr[V::DQ + j * NJ + k] = cq[k]; vr[V::DV + j * NJ + k] = cv[k]; }
        } else {
          double col[NJ];
          const double* qr = QR(kk, pl.point_qset(p));
          col_task<NJ>(qr, vr, j, wrt, col);
          double* dst = vr + (wrt ? V::DV : V::DQ) + j * NJ;
          if (FUSE_PROD) {
            // phase 4 fused: M^-1 of the configuration is ready (MINV_EARLY), so the column goes straight to
            // dt * da/dq resp. dt * da/dv + I (problem.hpp:495-502) without a round trip through shared memory and a barrier
            const double* Mi = qr + Q::MINV;
#pragma unroll
            for (int r = 0; r < NJ; ++r) {
              double acc = 0.0;
#pragma unroll
              for (int k = 0; k < NJ; ++k) acc += Mi[k * NJ + r] * col[k];
              double d = (-acc) * dt;
              if (wrt && r == j) d += 1.0;
              dst[r] = d;
            }
          } else {
#pragma unroll
            for (int k = 0; k < NJ; ++k) dst[k] = col[k];
          }
        }
      } else {
        minv_task(idx - nk * n_tan);
      }
    }
  }
  ex.sync();
  DDPB_PHASE_MARK(bd, 3, pc_);

  // phase 4: da/dq = -M^-1 dtau/dq, da/dv = -M^-1 dtau/dv, already scaled into the bottom blocks of
  // fx (problem.hpp:495-502): DQ <- dt * da/dq ; DV <- dt * da/dv + I   (done inside phase 3 where M^-1 is ready by then)
  constexpr bool PROD_DONE = MODE == HESS_FD_JACOBIANS && MINV_EARLY && DDPB_FUSE_PROD && !COL_MERGED;
  for (int idx = tid; idx < (PROD_DONE ? 0 : nk * pl.NT * 2 * NJ); idx += nthr) {
    const int wc = idx / (nk * pl.NT), task = idx - wc * (nk * pl.NT);  // (derivative kind, column) slowest, point fastest
    const int kk = task / pl.NT, p = task - kk * pl.NT;
    const int wrt = wc / NJ, c = wc - wrt * NJ;
    double* vr = VR(kk, p);
    double* colp = vr + (wrt ? V::DV : V::DQ) + c * NJ;
    const double* Mi = minv_of(kk, pl.point_qset(p));
    double src[NJ], dst[NJ];
#pragma unroll
    for (int k = 0; k < NJ; ++k) src[k] = colp[k];
#pragma unroll
    for (int r = 0; r < NJ; ++r) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < NJ; ++k) acc += Mi[k * NJ + r] * src[k];
      dst[r] = (-acc) * dt;
      if (wrt && r == c) dst[r] += 1.0;
    }
#pragma unroll
    for (int k = 0; k < NJ; ++k) colp[k] = dst[k];
  }
  if (!PROD_DONE) ex.sync();
  DDPB_PHASE_MARK(bd, 4, pc_);

  // element accessors of fx / fu at point p (p < NT) of knot kk
  auto fx_at = [&](int kk, int p, int r, int c) -> double {
    if (r < n) return c < n ? (r == c ? 1.0 : 0.0) : (r == c - n ? dt : 0.0);
    const double* vr = VR(kk, p);
    return c < n ? vr[V::DQ + c * NJ + (r - n)] : vr[V::DV + (c - n) * NJ + (r - n)];
  };
  auto fu_at = [&](int kk, int p, int r, int c) -> double {
    if (r < n) return 0.0;
    return minv_of(kk, pl.point_qset(p))[c * NJ + (r - n)] * dt;
  };

  // phase 5: the doubly advanced equality constraint at every point (empty loops when the knot has no rows)
  for (int idx = tid; idx < (ne > 0 ? nk * pl.P : 0); idx += nthr) {
    const int kk = idx / pl.P, p = idx - kk * pl.P;
    const double* fval = SM(kk, pl.o_fval);
    const double* tgt = SM(kk, pl.o_tgt);
    double* eqv = SM(kk, pl.o_eqv);
    // x_{t+1} = f(x,u) is fval[p]; inner eval_to: v_out = dt v ; q_out = q + v_out
    double qnn[NJ];
#pragma unroll
    for (int k = 0; k < NJ; ++k) { const double vo = dt * fval[p * nx + n + k]; qnn[k] = fval[p * nx + k] + vo; }
    if (bd.eq_kind == EQ_CONFIG) {
      for (int r = 0; r < ne; ++r) eqv[p * ne_sm + r] = qnn[r] - tgt[r];
    } else {
      double p3[3];
      frame_task<NJ>(m, qnn, bd.frame_id, bd.jac_mode, p3, p < pl.NT ? SM(kk, pl.o_J) + p * 3 * NJ : nullptr);
      for (int r = 0; r < ne; ++r) eqv[p * ne_sm + r] = p3[r] - tgt[r];
    }
  }
  if (ne > 0) ex.sync();  // (uniform over the CTA: all knots of a launch have the same number of rows)
  DDPB_PHASE_MARK(bd, 5, pc_);
  // eq_x = c_x [I, dt I] fx ; eq_u = c_x [I, dt I] fu   (problem.hpp:603-604, applied twice)
  for (int idx = tid; idx < (ne > 0 ? nk * pl.NT * ne * N : 0); idx += nthr) {
    const int kk = idx / (pl.NT * ne * N), task = idx - kk * pl.NT * ne * N;
    const int p = task / (ne * N), rem = task - p * ne * N;
    const int c = rem / ne, r = rem - c * ne;
    double acc = 0.0;
    if (bd.eq_kind == EQ_CONFIG) {
      if (c < nd) { acc = fx_at(kk, p, r, c); acc += dt * fx_at(kk, p, n + r, c); }
      else { acc = fu_at(kk, p, r, c - nd); acc += dt * fu_at(kk, p, n + r, c - nd); }
    } else {
      const double* J = SM(kk, pl.o_J) + p * 3 * NJ;
      for (int k = 0; k < nd; ++k) {
        const double e = k < n ? J[k * 3 + r] : J[(k - n) * 3 + r] * dt;
        acc += e * (c < nd ? fx_at(kk, p, k, c) : fu_at(kk, p, k, c - nd));
      }
    }
    SM(kk, pl.o_eqxu)[(size_t)p * ne_sm * N + c * ne + r] = acc;
  }
  if (ne > 0) ex.sync();
  DDPB_PHASE_MARK(bd, 6, pc_);

  // phase 6: assemble the stored part of the records (batch.cuh: the structural entries -- cost blocks except lu, rows
  // k < n of fx / fu and, with FD-of-Jacobians Hessians, of the tensors, v-slices of f_ux, f_uu -- are never written)
  // and stream it out block by block: inside a block the element index runs fastest over consecutive threads, so
  // consecutive threads write consecutive doubles.
  const RecLayout L = rec_layout(NJ, ne, MODE);
  double* const out0 = slab_rec(bd, t, b0, L.size);
  const size_t ks = (size_t)L.size;  // knot stride
  const double inv_eps = 1.0 / eps, inv_eps2 = 1.0 / (eps * eps);
  const size_t ps = (size_t)ne_sm * N;  // eqxu stride per point
  constexpr bool fdj = MODE == HESS_FD_JACOBIANS;
#define REC(kk, off) out0[(size_t)(kk) * ks + (size_t)(off)]
  // fn(kk, e) for every (knot of the group, element of a block of SIZE elements)
  auto for_block = [&](int SIZE, auto&& fn) {
#pragma unroll 4
    for (int idx_ = tid; idx_ < nk * SIZE; idx_ += nthr) {
      const int kk = idx_ / SIZE, e = idx_ - kk * SIZE;
      fn(kk, e);
    }
  };
  {
    for_block(nx, [&](int kk, int e) { REC(kk, L.f + e) = SM(kk, pl.o_fval)[e]; });
    for_block(n * nd, [&](int kk, int e) { REC(kk, L.fx + e) = fx_at(kk, 0, n + e % n, e / n); });
    for_block(n * nu, [&](int kk, int e) { REC(kk, L.fu + e) = fu_at(kk, 0, n + e % n, e / n); });
    for_block(nu, [&](int kk, int e) { REC(kk, L.lu + e) = bd.c * SM(kk, pl.o_base)[2 * n + e]; });  // lu = c u
    if (L.size > L.eq_uu + ne * nu * nu) for_block(1, [&](int kk, int) { REC(kk, L.size - 1) = 0.0; });  // padding double
  }
  if (fdj) {
    // problem.hpp:128-145: T(k,j,idx) = (J'(k,j) - J(k,j)) / eps ; rows k < n are exact constants (zero difference, not
    // stored).  Stored rows: fx(k,j) of point p is DQ/DV[(k-n) + j'*n] of that point's record; fu(k,j) is dt * M^-1 of
    // its q-set.  Two-level loops: a thread owns a (row k, column j) pair of the Jacobian and walks the perturbation
    // index i, so the row / column / offset arithmetic is done once per pair and the inner loop is load - subtract -
    // scale - store with constant strides (consecutive threads still write consecutive doubles for every i).
    for_block(n * nd, [&](int kk, int r) {
      const int kr = r % n, j = r / n;
      const int off = (j < n ? V::DQ + j * NJ : V::DV + (j - n) * NJ) + kr;
      const double* vr = VR(kk, 0) + off;
      const double base = vr[0];
#pragma unroll
      for (int i = 0; i < nd; ++i) REC(kk, L.fxx + r + i * n * nd) = (vr[(size_t)(1 + i) * V::SIZE] - base) * inv_eps;
    });
    // only a q-perturbation changes M^-1; v-perturbed points share the base q-set (exact zeros, not stored), and so
    // do the u-perturbed ones (f_uu: exact zeros, not stored)
    for_block(n * nu, [&](int kk, int r) {
      const int kr = r % n, j = r / n;
      const double* qr = QR(kk, 0) + Q::MINV + j * NJ + kr;
      const double base = qr[0] * dt;
#pragma unroll
      for (int i = 0; i < n; ++i) REC(kk, L.fux + r + i * n * nu) = (qr[(size_t)(1 + i) * QS] * dt - base) * inv_eps;
    });
    if (ne > 0) {
      for_block(ne, [&](int kk, int e) { REC(kk, L.eq + e) = SM(kk, pl.o_eqv)[e]; });
      for_block(ne * N, [&](int kk, int e) { REC(kk, L.eq_x + e) = SM(kk, pl.o_eqxu)[e]; });  // eq_x | eq_u adjacent on both sides
      for_block(ne * nd * nd, [&](int kk, int e) {
        const int k = e % ne, j = (e / ne) % nd, i = e / (ne * nd);
        const double* x = SM(kk, pl.o_eqxu);
        REC(kk, L.eq_xx + e) = (x[(1 + i) * ps + j * ne + k] - x[j * ne + k]) * inv_eps;
      });
      for_block(ne * nu * nd, [&](int kk, int e) {
        const int k = e % ne, j = (e / ne) % nu, i = e / (ne * nu);
        const double* x = SM(kk, pl.o_eqxu);
        REC(kk, L.eq_ux + e) = (x[(1 + i) * ps + (nd + j) * ne + k] - x[(nd + j) * ne + k]) * inv_eps;
      });
      for_block(ne * nu * nu, [&](int kk, int e) {
        const int k = e % ne, j = (e / ne) % nu, i = e / (ne * nu);
        const double* x = SM(kk, pl.o_eqxu);
        REC(kk, L.eq_uu + e) = pl.u_points ? (x[(1 + nd + i) * ps + (nd + j) * ne + k] - x[(nd + j) * ne + k]) * inv_eps : 0.0;
      });
    }
  } else {
    // problem.hpp:192-296 (value-based FD): row k of function `which` (0 = f, 1 = eq), global tangent indices ga, gb.
    // Here every tensor entry is data (rows k < n carry the reference's rounding noise): dense blocks.
    auto diag = [&](double f1, double f0, double col) -> double {
      double df = f1 - f0;
      df -= eps * col;
      df *= 2.0;
      return df * inv_eps2;
    };
    auto offd = [&](double f1, double f0, double col1, double col2, double d1, double d2) -> double {
      double df = f1 - f0;
      df -= eps * col1;
      df -= eps * col2;
      df *= 2.0;
      return 0.5 * (df * inv_eps2 - d1 - d2);
    };
    // Jacobian entry (row k of function `which`, tangent index g) and function value at point p
    auto jcol = [&](int kk, int which, int k, int g) -> double {
      if (which == 0) return g < nd ? fx_at(kk, 0, k, g) : fu_at(kk, 0, k, g - nd);
      return SM(kk, pl.o_eqxu)[g * ne + k];
    };
    auto val = [&](int kk, int which, int k, int p) -> double {
      return which == 0 ? SM(kk, pl.o_fval)[p * nx + k] : SM(kk, pl.o_eqv)[p * ne_sm + k];
    };
    // Two tables per knot, indexed [tangent index g][row: f rows, then eq rows]: the Jacobian entries and the diagonal
    // second differences (shared by every off-diagonal entry of their row / column, so computed once).  The tensors are
    // then filled by two-level loops: a thread owns (row k, left index a) and walks the right index b -- three
    // shared-memory loads and a handful of flops per entry, consecutive threads writing consecutive doubles.
    const int dgs = nd + ne_sm;  // stride of the tables per tangent index
    for (int idx = tid; idx < nk * N * (nd + ne); idx += nthr) {
      const int kk = idx / (N * (nd + ne)), r = idx - kk * (N * (nd + ne));
      const int g = r / (nd + ne), k0 = r - g * (nd + ne);
      const int which = k0 < nd ? 0 : 1, k = which ? k0 - nd : k0;
      const double jc = jcol(kk, which, k, g);
      SM(kk, pl.o_dg)[g * dgs + k0] = diag(val(kk, which, k, 1 + g), val(kk, which, k, 0), jc);
      SM(kk, pl.o_dg)[(N + g) * dgs + k0] = jc;
    }
    ex.sync();
    // block of a tensor: rows R (f: nd, eq: ne) of function `which`, left indices a0 .. a0+LA-1, right indices b0 .. b0+LB-1,
    // element (k, a, b) at off + k + R * a + R * LA * b
    auto fill = [&](int which, int off, int R, int a0, int LA, int b0_, int LB) {
      for_block(R * LA, [&](int kk, int r) {
        const int k = r % R, ga = a0 + r / R;
        const int row = which ? nd + k : k;
        const double* dg = SM(kk, pl.o_dg) + row;
        const double* jc = dg + (size_t)N * dgs;
        const double* vals = which == 0 ? SM(kk, pl.o_fval) + k : SM(kk, pl.o_eqv) + k;
        const int vs = which == 0 ? nx : ne_sm;
        const double f0 = vals[0], ca = jc[ga * dgs], da = dg[ga * dgs];
        // (measured slower on the 7R arm, 18.8 ms per pass as it is: a select-free form with tables pre-scaled by powers of
        // two and the index pair by min / max +0.5 ms; 7 entries in flight instead of 4, 21.5 ms; one copy of this code looped
        // over the six blocks instead of six inlined copies -- 11.7 k instead of 18.3 k instructions -- 21.0 ms)
#pragma unroll 4
        for (int bl = 0; bl < LB; ++bl) {  // independent entries: unrolled so that their loads and stores overlap
          const int gb = b0_ + bl;
          double v;
          if (gb == ga) {
            v = da;
          } else {
            const double cb = jc[gb * dgs], db = dg[gb * dgs];
            const bool lo = gb < ga;  // the smaller tangent index is subtracted first (problem.hpp:226-296)
            const int i = lo ? gb : ga, j = lo ? ga : gb;
            v = offd(vals[(size_t)(1 + N + pair_rank(i, j, N)) * vs], f0, lo ? cb : ca, lo ? ca : cb, lo ? db : da, lo ? da : db);
          }
          REC(kk, off + r + bl * R * LA) = v;
        }
      });
    };
    if (ne > 0) {
      for_block(ne, [&](int kk, int e) { REC(kk, L.eq + e) = SM(kk, pl.o_eqv)[e]; });
      for_block(ne * N, [&](int kk, int e) { REC(kk, L.eq_x + e) = SM(kk, pl.o_eqxu)[e]; });
    }
    fill(0, L.fxx, nd, 0, nd, 0, nd);
    fill(0, L.fux, nd, nd, nu, 0, nd);
    fill(0, L.fuu, nd, nd, nu, nd, nu);
    if (ne > 0) {
      fill(1, L.eq_xx, ne, 0, nd, 0, nd);
      fill(1, L.eq_ux, ne, nd, nu, 0, nd);
      fill(1, L.eq_uu, ne, nd, nu, nd, nu);
    }
  }
#undef REC
  DDPB_PHASE_MARK(bd, 7, pc_);
#undef SM
#undef QR
#undef VR
}

// ---- pendulum (nx = 2, nu = 1): one thread per knot, following the oracle statement by statement ----
DDPB_HD void pend_eval(const ModelDev& m, double dt, const double* x, double u, double* f) {
  const double vo = dt * x[1];
  f[0] = x[0] + vo;
  f[1] = x[1] + pend_acc(m, x[0], u) * dt;
}
DDPB_HD void pend_first(const ModelDev& m, double dt, const double* x, double u, double* f, double* fx /*2x2 col-major*/, double* fu /*2*/) {
  pend_eval(m, dt, x, u, f);
  fx[0] = 1.0; fx[2] = dt;
  fx[1] = pend_dacc_dq(m, x[0]) * dt;
  fx[3] = 0.0 * dt + 1.0;
  fu[0] = 0.0;
  fu[1] = (1.0 / m.pend_m) * dt;
}
// eq(t,x,u) = q_{t+2} - target with x_{t+1} = f(x,u), q_{t+2} = q_{t+1} + dt v_{t+1}
DDPB_HD double pend_eq(const ModelDev& m, double dt, const double* x, double u, double tgt) {
  double f[2];
  pend_eval(m, dt, x, u, f);
  const double vo = dt * f[1];
  return (f[0] + vo) - tgt;
}
DDPB_HD void pend_eq_first(const ModelDev& m, double dt, const double* x, double u, double tgt, double* e, double* ex /*2*/, double* eu) {
  double f[2], fx[4], fu[2];
  pend_first(m, dt, x, u, f, fx, fu);
  const double vo = dt * f[1];
  *e = (f[0] + vo) - tgt;
  // eq_n_x = [1, dt] ; out_x = eq_n_x fx ; out_u = eq_n_x fu
  ex[0] = fx[0]; ex[0] += dt * fx[1];
  ex[1] = fx[2]; ex[1] += dt * fx[3];
  *eu = fu[0]; *eu += dt * fu[1];
}

DDPB_HD void linearize_pendulum(const ModelDev& m, const BatchDev& bd, int b, int t) {
  const int B = bd.B;
  const int ne = bd.ne[t];
  const double dt = bd.dt;
  const double x[2] = {bd.xs[((size_t)t * 2 + 0) * B + b], bd.xs[((size_t)t * 2 + 1) * B + b]};
  const double u = bd.us[(size_t)t * B + b];
  const double tgt = ne ? bd.targets[(size_t)bd.eq_off[t] * B + b] : 0.0;
  const RecLayout L = rec_layout(1, ne, bd.hess_mode);
  const RecView rv = rec_view(bd, t, b, L.size);
  // dense entry e of field F (reference order); structural entries (batch.cuh) are not stored
  enum { F_lx, F_lu, F_lxx, F_lux, F_luu, F_f, F_fx, F_fu, F_fxx, F_fux, F_fuu, F_eq, F_eq_x, F_eq_u, F_eq_xx, F_eq_ux, F_eq_uu };
  auto put = [&](int field, int e, double v) {
    double sv;
    const int off = stored_offset(L, 1, field, e, dt, bd.c, &sv);
    if (off >= 0) rv[off] = v;
  };
  put(F_lx, 0, 0.0); put(F_lx, 1, 0.0);
  put(F_lu, 0, bd.c * u);
  for (int e = 0; e < 4; ++e) put(F_lxx, e, 0.0);
  put(F_lux, 0, 0.0); put(F_lux, 1, 0.0);
  put(F_luu, 0, bd.c);
  double f0[2], fx0[4], fu0[2];
  pend_first(m, dt, x, u, f0, fx0, fu0);
  put(F_f, 0, f0[0]); put(F_f, 1, f0[1]);
  for (int e = 0; e < 4; ++e) put(F_fx, e, fx0[e]);
  put(F_fu, 0, fu0[0]); put(F_fu, 1, fu0[1]);
  double e0 = 0.0, ex0[2] = {0.0, 0.0}, eu0 = 0.0;
  if (ne) {
    pend_eq_first(m, dt, x, u, tgt, &e0, ex0, &eu0);
    put(F_eq, 0, e0); put(F_eq_x, 0, ex0[0]); put(F_eq_x, 1, ex0[1]); put(F_eq_u, 0, eu0);
  }
  if (bd.hess_mode == HESS_FD_JACOBIANS) {
    const double eps = 1.4901161193847656e-08, ie = 1.0 / eps;
    for (int i = 0; i < 3; ++i) {
      double xp[2] = {x[0] + (i == 0 ? eps : 0.0), x[1] + (i == 1 ? eps : 0.0)};
      double up = u + (i == 2 ? eps : 0.0);
      double f1[2], fx1[4], fu1[2];
      pend_first(m, dt, xp, up, f1, fx1, fu1);
      double e1 = 0, ex1[2] = {0, 0}, eu1 = 0;
      if (ne) pend_eq_first(m, dt, xp, up, tgt, &e1, ex1, &eu1);
      if (i < 2) {
        for (int k = 0; k < 2; ++k) {
          for (int j = 0; j < 2; ++j) put(F_fxx, k + j * 2 + i * 4, (fx1[j * 2 + k] - fx0[j * 2 + k]) * ie);
          put(F_fux, k + i * 2, (fu1[k] - fu0[k]) * ie);
        }
        if (ne) {
          for (int j = 0; j < 2; ++j) put(F_eq_xx, j + i * 2, (ex1[j] - ex0[j]) * ie);
          put(F_eq_ux, i, (eu1 - eu0) * ie);
        }
      } else {
        for (int k = 0; k < 2; ++k) put(F_fuu, k, (fu1[k] - fu0[k]) * ie);
        if (ne) put(F_eq_uu, 0, (eu1 - eu0) * ie);
      }
    }
  } else {
    const double eps = 1.220703125e-04, ie2 = 1.0 / (eps * eps);
    // points: 3 singles, 3 pairs (0,1), (0,2), (1,2) over the tangent indices (q, v, u)
    double fs[3][2], es[3] = {0, 0, 0};
    double col[3][2] = {{fx0[0], fx0[1]}, {fx0[2], fx0[3]}, {fu0[0], fu0[1]}};
    double ecol[3] = {ex0[0], ex0[1], eu0};
    double dg[3][2], edg[3] = {0, 0, 0};
    for (int i = 0; i < 3; ++i) {
      double xp[2] = {x[0] + (i == 0 ? eps : 0.0), x[1] + (i == 1 ? eps : 0.0)};
      double up = u + (i == 2 ? eps : 0.0);
      pend_eval(m, dt, xp, up, fs[i]);
      for (int k = 0; k < 2; ++k) {
        double df = fs[i][k] - f0[k];
        df -= eps * col[i][k];
        df *= 2.0;
        dg[i][k] = df * ie2;
      }
      if (ne) {
        es[i] = pend_eq(m, dt, xp, up, tgt);
        double df = es[i] - e0;
        df -= eps * ecol[i];
        df *= 2.0;
        edg[i] = df * ie2;
      }
    }
    for (int k = 0; k < 2; ++k) {
      put(F_fxx, k + 0 * 2 + 0 * 4, dg[0][k]);
      put(F_fxx, k + 1 * 2 + 1 * 4, dg[1][k]);
      put(F_fuu, k, dg[2][k]);
    }
    if (ne) { put(F_eq_xx, 0, edg[0]); put(F_eq_xx, 3, edg[1]); put(F_eq_uu, 0, edg[2]); }
    for (int i = 0; i < 3; ++i)
      for (int j = i + 1; j < 3; ++j) {
        double xp[2] = {x[0] + ((i == 0 || j == 0) ? eps : 0.0), x[1] + ((i == 1 || j == 1) ? eps : 0.0)};
        double up = u + ((i == 2 || j == 2) ? eps : 0.0);
        double f1[2];
        pend_eval(m, dt, xp, up, f1);
        for (int k = 0; k < 2; ++k) {
          double df = f1[k] - f0[k];
          df -= eps * col[i][k];
          df -= eps * col[j][k];
          df *= 2.0;
          const double v = 0.5 * (df * ie2 - dg[i][k] - dg[j][k]);
          if (j < 2) { put(F_fxx, k + j * 2 + i * 4, v); put(F_fxx, k + i * 2 + j * 4, v); }
          else put(F_fux, k + i * 2, v);  // ux(k, u, x_i)
        }
        if (ne) {
          double df = pend_eq(m, dt, xp, up, tgt) - e0;
          df -= eps * ecol[i];
          df -= eps * ecol[j];
          df *= 2.0;
          const double v = 0.5 * (df * ie2 - edg[i] - edg[j]);
          if (j < 2) { put(F_eq_xx, j + i * 2, v); put(F_eq_xx, i + j * 2, v); }
          else put(F_eq_ux, i, v);
        }
      }
  }
}

DDPB_VARIANT_END
