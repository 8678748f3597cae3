// backward.cuh -- Riccati-like backward recursion, one CTA per problem, sequential over knots.
//
// Restates ddp_solver_t::backward_pass (/root/reference/include/ddp/ddp_bwd.ipp:9-155):
//   tmp = p + mu eq ; tmp2 = P + mu eq_x                                   (:46-47)
//   Q_x, Q_u, Q_xx, Q_uu, Q_ux incl. the six tensor contractions           (:61-87)
//   LL^T = Q_uu + reg I ; failure -> reg = max(reg, mu), mu *= 2, reg *= 2, restart from t = T-1 (:102-132)
//   k = -(LL^T)^-1 Q_u ; K = -(LL^T)^-1 Q_ux ; origin = x_t                (:134-136)
//   V_x = Q_x + Q_ux^T k ; V_xx = Q_xx + Q_ux^T K                          (:142-146)
// Small matrices (V, Q, f_x, f_u, eq blocks) live in shared memory; the third-order tensors are
// used exactly once, so every thread streams the contiguous k-run of "its" output element straight
// from HBM (the record is [k fastest], so a warp reads one contiguous span).
#pragma once
#include "batch.cuh"

namespace ddpb {

template <int NJ> struct BwdSmem {
  static constexpr int nd = 2 * NJ, nu = NJ;
  static constexpr int Vxx = 0, Vx = Vxx + nd * nd, Qx = Vx + nd, Qu = Qx + nd, Qxx = Qu + nu, Quu = Qxx + nd * nd,
                       Qux = Quu + nu * nu, VF = Qux + nu * nd, VU = VF + nd * nd, fx = VU + nd * nu, fu = fx + nd * nd,
                       L = fu + nd * nu, kk = L + nu * nu, KK = kk + nu, small = KK + nu * nd,  // lx | lu
                       eqv = small + nd + nu, eqx = eqv + MAXE, equ = eqx + MAXE * nd, pe = equ + MAXE * nu,
                       pex = pe + MAXE, tmp = pex + MAXE * nd, tmp2 = tmp + MAXE, flags = tmp2 + MAXE * nd, SIZE = flags + 4;
};

template <int NJ, class Ex>
DDPB_HD void backward_problem(const BatchDev& bd, int b, double* sm, const Ex& ex) {
  constexpr int nd = 2 * NJ, nu = NJ, nx = 2 * NJ;
  using S = BwdSmem<NJ>;
  const int B = bd.B, T = bd.T;
  const int tid = ex.tid(), nthr = ex.nthr();
  double mu = bd.scal[(size_t)S_MU * B + b];
  double reg = bd.scal[(size_t)S_REG * B + b];
  int restarts = 0;
  bool success = false;
  while (!success) {
    // V_xx = lfxx = 0, V_x = lfx = 0  (problem.hpp:958-959)
    for (int e = tid; e < nd * nd + nd; e += nthr) sm[S::Vxx + e] = 0.0;
    ex.sync();
    bool failed = false;
    for (int t = T - 1; t >= 0; --t) {
      const int ne = bd.ne[t], eo = bd.eq_off[t];
      const RecLayout L = rec_layout(nd, nu, nx, ne);
      const double* rec = slab_rec(bd, t, b, L.size);
      // stage the small blocks
      for (int e = tid; e < nd * nd; e += nthr) sm[S::fx + e] = rec[L.fx + e];
      for (int e = tid; e < nd * nu; e += nthr) sm[S::fu + e] = rec[L.fu + e];
      for (int e = tid; e < nd + nu; e += nthr) sm[S::small + e] = rec[L.lx + e];  // lx | lu are adjacent
      if (ne > 0) {
        for (int e = tid; e < ne; e += nthr) {
          sm[S::eqv + e] = rec[L.eq + e];
          sm[S::pe + e] = bd.m_val[((size_t)eo + e) * B + b];
        }
        for (int e = tid; e < ne * nd; e += nthr) {
          sm[S::eqx + e] = rec[L.eq_x + e];
          sm[S::pex + e] = bd.m_jac[((size_t)eo * nd + e) * B + b];
        }
        for (int e = tid; e < ne * nu; e += nthr) sm[S::equ + e] = rec[L.eq_u + e];
      }
      ex.sync();
      for (int e = tid; e < ne; e += nthr) sm[S::tmp + e] = sm[S::pe + e] + mu * sm[S::eqv + e];
      for (int e = tid; e < ne * nd; e += nthr) sm[S::tmp2 + e] = sm[S::pex + e] + mu * sm[S::eqx + e];
      // VF = V_xx f_x ; VU = V_xx f_u
      for (int e = tid; e < nd * nd + nd * nu; e += nthr) {
        const bool isF = e < nd * nd;
        const int ee = isF ? e : e - nd * nd;
        const int r = ee % nd, c = ee / nd;
        const double* M = isF ? sm + S::fx : sm + S::fu;
        double a = 0.0;
#pragma unroll
        for (int k = 0; k < nd; ++k) a += sm[S::Vxx + k * nd + r] * M[c * nd + k];
        sm[(isF ? S::VF : S::VU) + ee] = a;
      }
      ex.sync();
      const double* fx = sm + S::fx;
      const double* fu = sm + S::fu;
      const double* Vx = sm + S::Vx;
      const double* eqv = sm + S::eqv;
      const double* eqx = sm + S::eqx;
      const double* equ = sm + S::equ;
      const double* pex = sm + S::pex;
      const double* tmp = sm + S::tmp;
      const double* tmp2 = sm + S::tmp2;
      // all Q elements as one task list: Q_xx | Q_ux | Q_uu | Q_x | Q_u
      constexpr int nQxx = nd * nd, nQux = nu * nd, nQuu = nu * nu;
      for (int e = tid; e < nQxx + nQux + nQuu + nd + nu; e += nthr) {
        if (e < nQxx) {
          const int r = e % nd, c = e / nd;
          double q = rec[L.lxx + e], a = 0.0;
#pragma unroll
          for (int k = 0; k < nd; ++k) a += fx[r * nd + k] * sm[S::VF + c * nd + k];
          q += a;
          a = 0.0;
          for (int k = 0; k < ne; ++k) a += eqx[r * ne + k] * tmp2[c * ne + k];
          q += a;
          a = 0.0;
          for (int k = 0; k < ne; ++k) a += pex[r * ne + k] * eqx[c * ne + k];
          q += a;
          a = 0.0;
          const double* t3 = rec + L.eq_xx + r * ne + (size_t)c * ne * nd;
          for (int k = 0; k < ne; ++k) a += tmp[k] * t3[k];
          q += a;
          a = 0.0;
          const double* f3 = rec + L.fxx + r * nd + (size_t)c * nd * nd;
#pragma unroll
          for (int k = 0; k < nd; ++k) a += Vx[k] * f3[k];
          q += a;
          sm[S::Qxx + e] = q;
        } else if (e < nQxx + nQux) {
          const int ee = e - nQxx;
          const int r = ee % nu, c = ee / nu;
          double q = rec[L.lux + ee], a = 0.0;
#pragma unroll
          for (int k = 0; k < nd; ++k) a += fu[r * nd + k] * sm[S::VF + c * nd + k];
          q += a;
          a = 0.0;
          for (int k = 0; k < ne; ++k) a += equ[r * ne + k] * tmp2[c * ne + k];
          q += a;
          a = 0.0;
          const double* t3 = rec + L.eq_ux + r * ne + (size_t)c * ne * nu;
          for (int k = 0; k < ne; ++k) a += tmp[k] * t3[k];
          q += a;
          a = 0.0;
          const double* f3 = rec + L.fux + r * nd + (size_t)c * nd * nu;
#pragma unroll
          for (int k = 0; k < nd; ++k) a += Vx[k] * f3[k];
          q += a;
          sm[S::Qux + ee] = q;
        } else if (e < nQxx + nQux + nQuu) {
          const int ee = e - nQxx - nQux;
          const int r = ee % nu, c = ee / nu;
          double q = rec[L.luu + ee], a = 0.0;
#pragma unroll
          for (int k = 0; k < nd; ++k) a += fu[r * nd + k] * sm[S::VU + c * nd + k];
          q += a;
          a = 0.0;
          for (int k = 0; k < ne; ++k) a += equ[r * ne + k] * equ[c * ne + k];
          q += a * mu;
          a = 0.0;
          const double* t3 = rec + L.eq_uu + r * ne + (size_t)c * ne * nu;
          for (int k = 0; k < ne; ++k) a += tmp[k] * t3[k];
          q += a;
          a = 0.0;
          const double* f3 = rec + L.fuu + r * nd + (size_t)c * nd * nu;
#pragma unroll
          for (int k = 0; k < nd; ++k) a += Vx[k] * f3[k];
          q += a;
          sm[S::Quu + ee] = q;
        } else if (e < nQxx + nQux + nQuu + nd) {
          const int i = e - nQxx - nQux - nQuu;
          double q = sm[S::small + i], a = 0.0;
#pragma unroll
          for (int k = 0; k < nd; ++k) a += fx[i * nd + k] * Vx[k];
          q += a;
          a = 0.0;
          for (int k = 0; k < ne; ++k) a += eqx[i * ne + k] * tmp[k];
          q += a;
          a = 0.0;
          for (int k = 0; k < ne; ++k) a += pex[i * ne + k] * eqv[k];
          q += a;
          sm[S::Qx + i] = q;
        } else {
          const int i = e - nQxx - nQux - nQuu - nd;
          double q = sm[S::small + nd + i], a = 0.0;
#pragma unroll
          for (int k = 0; k < nd; ++k) a += fu[i * nd + k] * Vx[k];
          q += a;
          a = 0.0;
          for (int k = 0; k < ne; ++k) a += equ[i * ne + k] * tmp[k];
          q += a;
          sm[S::Qu + i] = q;
        }
      }
      ex.sync();
      // Cholesky of the lower triangle of Q_uu + reg I (Eigen::LLT semantics: fail when a pivot <= 0)
      if (tid == 0) {
        double* A = sm + S::L;
        for (int k = 0; k < nu * nu; ++k) A[k] = sm[S::Quu + k];
        for (int k = 0; k < nu; ++k) A[k * nu + k] += reg;
        double ok = 1.0;
        for (int k = 0; k < nu && ok != 0.0; ++k) {
          double x = A[k * nu + k];
          for (int j = 0; j < k; ++j) x -= A[j * nu + k] * A[j * nu + k];
          if (x <= 0.0) { ok = 0.0; break; }
          x = sqrt(x);
          A[k * nu + k] = x;
          for (int i = k + 1; i < nu; ++i) {
            double s = A[k * nu + i];
            for (int j = 0; j < k; ++j) s -= A[j * nu + i] * A[j * nu + k];
            A[k * nu + i] = s / x;
          }
        }
        sm[S::flags] = ok;
      }
      ex.sync();
      if (sm[S::flags] == 0.0) {
        if (reg < mu) reg = mu;
        mu *= 2.0;
        reg *= 2.0;
        ++restarts;
        failed = true;
        ex.sync();
        break;
      }
      // k = -(LL^T)^-1 Q_u, K = -(LL^T)^-1 Q_ux: one right-hand side per thread
      for (int c = tid; c < nd + 1; c += nthr) {
        const double* A = sm + S::L;
        double rhs[nu];
#pragma unroll
        for (int i = 0; i < nu; ++i) rhs[i] = -(c == nd ? sm[S::Qu + i] : sm[S::Qux + c * nu + i]);
#pragma unroll
        for (int i = 0; i < nu; ++i) {
          double s = rhs[i];
#pragma unroll
          for (int j = 0; j < nu; ++j)
            if (j < i) s -= A[j * nu + i] * rhs[j];
          rhs[i] = s / A[i * nu + i];
        }
#pragma unroll
        for (int i = nu - 1; i >= 0; --i) {
          double s = rhs[i];
#pragma unroll
          for (int j = 0; j < nu; ++j)
            if (j > i) s -= A[i * nu + j] * rhs[j];
          rhs[i] = s / A[i * nu + i];
        }
#pragma unroll
        for (int i = 0; i < nu; ++i) {
          if (c == nd) {
            sm[S::kk + i] = rhs[i];
            bd.fb_k[((size_t)t * nu + i) * B + b] = rhs[i];
          } else {
            sm[S::KK + c * nu + i] = rhs[i];
            bd.fb_K[((size_t)t * nu * nd + c * nu + i) * B + b] = rhs[i];
          }
        }
      }
      for (int e = tid; e < nx; e += nthr) bd.fb_origin[((size_t)t * nx + e) * B + b] = bd.xs[((size_t)t * nx + e) * B + b];
      ex.sync();
      for (int e = tid; e < nd * nd + nd; e += nthr) {
        if (e < nd * nd) {
          const int r = e % nd, c = e / nd;
          double a = 0.0;
#pragma unroll
          for (int k = 0; k < nu; ++k) a += sm[S::Qux + r * nu + k] * sm[S::KK + c * nu + k];
          sm[S::Vxx + e] = sm[S::Qxx + e] + a;
        } else {
          const int i = e - nd * nd;
          double a = 0.0;
#pragma unroll
          for (int k = 0; k < nu; ++k) a += sm[S::Qux + i * nu + k] * sm[S::kk + k];
          sm[S::Vx + i] = sm[S::Qx + i] + a;
        }
      }
      ex.sync();
    }
    if (!failed) success = true;
    else if (bd.max_restarts > 0 && (restarts > bd.max_restarts || !(mu < 1.79e308))) {
      // guarded mode: this problem's Q_uu cannot be made positive definite by the reference's mu/reg schedule
      if (tid == 0) bd.scal[(size_t)S_STATUS * B + b] = 2.0;
      break;
    }
  }
  if (tid == 0) {
    bd.scal[(size_t)S_MU * B + b] = mu;
    bd.scal[(size_t)S_REG * B + b] = reg;
    bd.scal[(size_t)S_RESTARTS * B + b] += (double)restarts;
  }
}

}  // namespace ddpb
