// linearize_pt.cuh -- linearization of chains of <= 2 joints with FD-of-Jacobians Hessians at knots without equality rows:
// one THREAD per evaluation point, everything in registers.
//
// Restates the same reference functions as linearize.cuh (problem_t::compute_derivatives problem.hpp:956-997,
// dynamics_t::first_order_deriv :463-503, finite_diff_hessian_compute::second_order_deriv_1 :67-150) on exactly the
// reference's evaluation points; what differs is the work split.  A 2R knot has 1 + 2n = 5 points (the u-perturbed
// points only difference f_u = dt M^-1(q), which u cannot change -- exact zeros, see linearize.cuh) of ~3 kflop each.
// The CTA formulation of linearize.cuh (phases separated by barriers, records in shared memory) runs that at 12 warps
// per SM limited by shared memory, with most lanes idle in the serial sub-phases: 1.40 ms for 4096 x 200 knots, FP64 pipe
// 31 % busy.  Here a thread evaluates ONE point from start to finish -- world-frame axes / inertias, articulated inertias,
// ABA sweeps, the 2n derivative columns, M^-1 and the products -- as straight-line code on register-resident records
// (no shared memory, no barriers, no idle lanes: a warp holds floor(32 / (1 + 2n)) whole knots), the lanes of a knot then
// exchange the base point's Jacobian with shuffles, difference, and the warp writes its knots' records through a
// shared-memory stage with fully coalesced 16-byte stores.  Points with the same q recompute the q-only data (two extra
// q_task per 2R knot): cheaper than sharing it.
#pragma once
#include "batch.cuh"

DDPB_VARIANT_BEGIN

template <int NJ> struct PtGeom {
  static constexpr int n = NJ, nd = 2 * NJ, nu = NJ, nx = 2 * NJ;
  static constexpr int P = 1 + 2 * NJ;   // evaluation points per knot: base, q_i + eps, v_i + eps
  static constexpr int KPW = 32 / P;     // whole knots per warp
  // stored record of a knot without equality rows (batch.cuh rec_layout(n, 0, HESS_FD_JACOBIANS))
  static constexpr int o_f = 0, o_fx = nx, o_fu = o_fx + n * nd, o_lu = o_fu + n * nu, o_fxx = o_lu + nu, o_fux = o_fxx + n * nd * nd,
                       o_end = o_fux + n * nu * n, SIZE = (o_end + 1) & ~1;
};

// what one evaluation point contributes
template <int NJ> struct PtOut {
  double f[2 * NJ];          // f(x, u)                                  (base point only)
  double fxB[NJ * 2 * NJ];   // rows k >= n of f_x, (k - n) + n j        (problem.hpp:495-502)
  double fuB[NJ * NJ];       // rows k >= n of f_u = dt M^-1, (k - n) + n j
};

// first-order data of the Euler-discretised dynamics at evaluation point p of (x, u): p = 0 base, 1 + i: q_i + eps,
// 1 + n + i: v_i + eps.  Same arithmetic, in the same order, as phases 1-4 of linearize_group for chains of <= 3 joints.
template <int NJ> DDPB_HD void lin_point(const ModelDev& m, double dt, const double* x, const double* u, int p, PtOut<NJ>& o) {
  constexpr int n = NJ;
  using Q = QRec<NJ>;
  using V = VRec<NJ>;
  const double eps = 1.4901161193847656e-08;  // 2^-26 = sqrt(eps_mach)  (problem.hpp:107)
  double q[n], qd[n], tau[n], qdd[n];
#pragma unroll
  for (int k = 0; k < n; ++k) {
    q[k] = x[k] + (p == 1 + k ? eps : 0.0);
    qd[k] = x[n + k] + (p == 1 + n + k ? eps : 0.0);
    tau[k] = u[k];
  }
  double qrec[Q::SIZE], vrec[V::SIZE];
  q_task<NJ>(m, q, qrec);
  v_task<NJ, true>(m, qrec, qd, tau, qdd, vrec);
  // dynamics_t::eval_to (problem.hpp:441-461): v_out = dt v ; q_out = q + v_out ; v_out = v + acc dt
#pragma unroll
  for (int k = 0; k < n; ++k) {
    const double vo = dt * qd[k];
    o.f[k] = q[k] + vo;
    o.f[n + k] = qd[k] + qdd[k] * dt;
  }
  double Mi[n * n];
#pragma unroll
  for (int c = 0; c < n; ++c) minv_col<NJ>(qrec, c, Mi + c * n);
  double DQ[n * n], DV[n * n];
#pragma unroll
  for (int j = 0; j < n; ++j) col_task2<NJ>(qrec, vrec, j, DQ + j * n, DV + j * n);
  // da/dq = -M^-1 dtau/dq, da/dv = -M^-1 dtau/dv, scaled into the bottom blocks of fx: dt da/dq ; dt da/dv + I
#pragma unroll
  for (int wrt = 0; wrt < 2; ++wrt)
#pragma unroll
    for (int c = 0; c < n; ++c) {
      const double* src = (wrt ? DV : DQ) + c * n;
#pragma unroll
      for (int r = 0; r < n; ++r) {
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < n; ++k) acc += Mi[k * n + r] * src[k];
        double d = (-acc) * dt;
        if (wrt && r == c) d += 1.0;
        o.fxB[r + n * (wrt * n + c)] = d;
      }
    }
#pragma unroll
  for (int e = 0; e < n * n; ++e) o.fuB[e] = Mi[e] * dt;
}

// the part of the stored record that point p writes, given the base point's Jacobians
template <int NJ> DDPB_HD void lin_point_write(double* rec, int p, const PtOut<NJ>& o, const double* base_fxB, const double* base_fuB, const double* u, double c_cost) {
  using G = PtGeom<NJ>;
  constexpr int n = NJ, nd = 2 * NJ, nu = NJ, nx = 2 * NJ;
  const double inv_eps = 1.0 / 1.4901161193847656e-08;
  if (p == 0) {
#pragma unroll
    for (int e = 0; e < nx; ++e) rec[G::o_f + e] = o.f[e];
#pragma unroll
    for (int e = 0; e < n * nd; ++e) rec[G::o_fx + e] = o.fxB[e];
#pragma unroll
    for (int e = 0; e < n * nu; ++e) rec[G::o_fu + e] = o.fuB[e];
#pragma unroll
    for (int e = 0; e < nu; ++e) rec[G::o_lu + e] = c_cost * u[e];  // lu = c u
    if (G::SIZE > G::o_end) rec[G::SIZE - 1] = 0.0;                   // padding double
  } else {
    // problem.hpp:128-145: T(k, j, i) = (J'(k, j) - J(k, j)) / eps for the stored rows; M^-1 only moves with q
    const int i = p - 1;
#pragma unroll
    for (int e = 0; e < n * nd; ++e) rec[G::o_fxx + e + n * nd * i] = (o.fxB[e] - base_fxB[e]) * inv_eps;
    if (i < n) {
#pragma unroll
      for (int e = 0; e < n * nu; ++e) rec[G::o_fux + e + n * nu * i] = (o.fuB[e] - base_fuB[e]) * inv_eps;
    }
  }
}

// CPU emulation / plain statement: one knot
template <int NJ> inline void linearize_pt_knot_plain(const ModelDev& m, const BatchDev& bd, int t, int b) {
  using G = PtGeom<NJ>;
  constexpr int nx = 2 * NJ, nu = NJ;
  double x[nx], u[nu];
  for (int k = 0; k < nx; ++k) x[k] = bd.xs[((size_t)t * nx + k) * bd.B + b];
  for (int k = 0; k < nu; ++k) u[k] = bd.us[((size_t)t * nu + k) * bd.B + b];
  double* rec = slab_rec(bd, t, b, G::SIZE);
  PtOut<NJ> base, o;
  lin_point<NJ>(m, bd.dt, x, u, 0, base);
  lin_point_write<NJ>(rec, 0, base, base.fxB, base.fuB, u, bd.c);
  for (int p = 1; p < G::P; ++p) {
    lin_point<NJ>(m, bd.dt, x, u, p, o);
    lin_point_write<NJ>(rec, p, o, base.fxB, base.fuB, u, bd.c);
  }
}

#if defined(__CUDACC__)
// knots [t0, t0 + nt) x problems [bd.b0, bd.b0 + bd.nb): knot id = (t - t0) * nb + (b - b0), KPW consecutive ids per warp.
// stage: per-warp shared memory of KPW * PtGeom::SIZE doubles.
template <int NJ> __device__ __forceinline__ void linearize_pt_warp(const ModelDev& m, const BatchDev& bd, int t0, int nt, double* stage) {
  using G = PtGeom<NJ>;
  constexpr int nx = 2 * NJ, nu = NJ, n = NJ, nd = 2 * NJ;
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  // 32-bit knot ids (a launch covers at most T x B < 2^31 knots); the warp's first knot is split into (knot, problem)
  // once, the others follow by increments -- no division per thread or per stored chunk
  const unsigned warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned total = (unsigned)nt * (unsigned)bd.nb, nb = (unsigned)bd.nb;
  const int kk = lane / G::P, p = lane - kk * G::P;
  const unsigned kid_first = warp_global * G::KPW;
  const unsigned tt_first = kid_first / nb, bi_first = kid_first - tt_first * nb;
  auto split = [&](int k, int& tq, int& bq) {  // knot kid_first + k, k < KPW <= nb + ... : at most a few wraps
    unsigned tt = tt_first, bi = bi_first + (unsigned)k;
    while (bi >= nb) { bi -= nb; ++tt; }
    tq = t0 + (int)tt; bq = bd.b0 + (int)bi;
  };
  const unsigned kid = kid_first + (unsigned)kk;
  const bool inrange = kk < G::KPW && kid < total;
  int t = t0, b = bd.b0;
  if (inrange) split(kk, t, b);
  const int B = bd.B;
  const bool act = inrange && bd.scal[(size_t)S_STATUS * B + b] == 0.0;  // frozen problems keep their records
  PtOut<NJ> o;
  double u[nu];
  if (act) {
    double x[nx];
#pragma unroll
    for (int k = 0; k < nx; ++k) x[k] = bd.xs[((size_t)t * nx + k) * B + b];
#pragma unroll
    for (int k = 0; k < nu; ++k) u[k] = bd.us[((size_t)t * nu + k) * B + b];
    lin_point<NJ>(m, bd.dt, x, u, p, o);
  } else {
#pragma unroll
    for (int e = 0; e < n * nd; ++e) o.fxB[e] = 0.0;
#pragma unroll
    for (int e = 0; e < n * nu; ++e) o.fuB[e] = 0.0;
#pragma unroll
    for (int e = 0; e < nu; ++e) u[e] = 0.0;
  }
  // the base point's Jacobians to every lane of its knot
  double bfx[n * nd], bfu[n * nu];
  const int base_lane = (kk < G::KPW ? kk : 0) * G::P;
#pragma unroll
  for (int e = 0; e < n * nd; ++e) bfx[e] = __shfl_sync(FULL, o.fxB[e], base_lane);
#pragma unroll
  for (int e = 0; e < n * nu; ++e) bfu[e] = __shfl_sync(FULL, o.fuB[e], base_lane);
  if (act) lin_point_write<NJ>(stage + kk * G::SIZE, p, o, bfx, bfu, u, bd.c);
  __syncwarp();
  // the warp's records, 16 bytes per lane and store: consecutive lanes -> consecutive chunks of a record
  constexpr int CH = G::SIZE / 2;
  for (int idx = lane; idx < G::KPW * CH; idx += 32) {
    const int k2 = idx / CH, c = idx - k2 * CH;   // CH is a compile-time constant
    if (kid_first + (unsigned)k2 >= total) break;
    int t2, b2;
    split(k2, t2, b2);
    if (bd.scal[(size_t)S_STATUS * B + b2] != 0.0) continue;
    reinterpret_cast<double2*>(slab_rec(bd, t2, b2, G::SIZE))[c] = reinterpret_cast<const double2*>(stage + k2 * G::SIZE)[c];
  }
}
#endif

DDPB_VARIANT_END
