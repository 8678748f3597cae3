// backward.cuh -- Riccati-like backward recursion, ONE WARP per problem, sequential over knots.
//
// Restates ddp_solver_t::backward_pass (/root/reference/include/ddp/ddp_bwd.ipp:9-155):
//   tmp = p + mu eq ; tmp2 = P + mu eq_x                                   (:46-47)
//   Q_x, Q_u, Q_xx, Q_uu, Q_ux incl. the six tensor contractions           (:61-87)
//   LL^T = Q_uu + reg I ; failure -> reg = max(reg, mu), mu *= 2, reg *= 2, restart from t = T-1 (:102-132)
//   k = -(LL^T)^-1 Q_u ; K = -(LL^T)^-1 Q_ux ; origin = x_t                (:134-136)
//   V_x = Q_x + Q_ux^T k ; V_xx = Q_xx + Q_ux^T K                          (:142-146)
//
// The recursion is a latency chain (T dependent knots), so the design minimises the dependent path of one knot
// instead of occupancy: a single warp owns a problem (phase boundaries are __syncwarp, not CTA barriers), every
// element loop has a compile-time trip count and is fully unrolled so each lane runs several independent FMA
// chains, all global loads of a knot are issued back to back at its start (one memory latency per knot), and the
// Cholesky factorisation is spread over lanes.  Small matrices live in shared memory; the third-order tensors are
// used once and go HBM -> registers -> V_x-contraction, each lane reading the contiguous k-run of its output element.
// Every sum keeps the reference's term order, so results do not depend on the thread count.
#pragma once
#include "batch.cuh"

namespace ddpb {

template <int NJ> struct BwdSmem {
  static constexpr int nd = 2 * NJ, nu = NJ;
  static constexpr int NTEN = nd * nd + nu * nd + nu * nu;  // Q_xx | Q_ux | Q_uu task list
  static constexpr int Vxx = 0, Vx = Vxx + nd * nd, Qx = Vx + nd, Qu = Qx + nd, Qxx = Qu + nu, Qux = Qxx + nd * nd,
                       Quu = Qux + nu * nd, VF = Quu + nu * nu, VU = VF + nd * nd, fx = VU + nd * nu, fu = fx + nd * nd,
                       L = fu + nd * nu, kk = L + nu * nu, KK = kk + nu, small = KK + nu * nd,  // lx | lu
                       eqv = small + nd + nu, eqx = eqv + MAXE, equ = eqx + MAXE * nd, pe = equ + MAXE * nu,
                       pex = pe + MAXE, tmp = pex + MAXE * nd, tmp2 = tmp + MAXE, flags = tmp2 + MAXE * nd,
                       Wf = flags + 4,  // V_x-contraction of fxx | fux | fuu, task-list order
                       SIZE = Wf + NTEN;
  static constexpr int BT = 32;  // threads per problem on the device
};

// BT = compile-time thread count of the executor (32 on the device, 1 in the CPU emulation)
template <int NJ, int BT, class Ex>
DDPB_HD void backward_problem(const BatchDev& bd, int b, double* sm, const Ex& ex) {
  constexpr int nd = 2 * NJ, nu = NJ, nx = 2 * NJ;
  using S = BwdSmem<NJ>;
  constexpr int nQxx = nd * nd, nQux = nu * nd, nQuu = nu * nu, NTEN = S::NTEN;
  const int B = bd.B, T = bd.T;
  const int tid = ex.tid();
  double mu = bd.scal[(size_t)S_MU * B + b];
  double reg = bd.scal[(size_t)S_REG * B + b];
  int restarts = 0;
  bool success = false;
#define FOR_ELEMS(e, N) _Pragma("unroll") for (int u_ = 0; u_ < ((N) + BT - 1) / BT; ++u_) if (const int e = tid + u_ * BT; e < (N))
  while (!success) {
    // V_xx = lfxx = 0, V_x = lfx = 0  (problem.hpp:958-959)
    FOR_ELEMS(e, nd * nd + nd) sm[S::Vxx + e] = 0.0;
    ex.sync();
    bool failed = false;
    for (int t = T - 1; t >= 0; --t) {
      const int ne = bd.ne[t], eo = bd.eq_off[t];
      const RecLayout L = rec_layout(nd, nu, nx, ne);
      const double* rec = slab_rec(bd, t, b, L.size);
      // ---- phase A: every global load of the knot, back to back -------------------------------------------
      double xorg = 0.0;
      if (tid < nx) xorg = bd.xs[((size_t)t * nx + tid) * B + b];
      {
        constexpr int EPT = (NTEN + BT - 1) / BT;
        constexpr int CH = EPT < 4 ? EPT : 4;  // tensor elements in flight per lane (CH * nd doubles of registers)
#pragma unroll
        for (int u0 = 0; u0 < EPT; u0 += CH) {
          double ten[CH][nd];
#pragma unroll
          for (int u = 0; u < CH; ++u) {
            const int e = tid + (u0 + u) * BT;
            if (u0 + u < EPT && e < NTEN) {
              // element e of Q_xx | Q_ux | Q_uu  <->  k-run of fxx | fux | fuu
              const double* f3 = e < nQxx ? rec + L.fxx + (size_t)e * nd
                                          : (e < nQxx + nQux ? rec + L.fux + (size_t)(e - nQxx) * nd : rec + L.fuu + (size_t)(e - nQxx - nQux) * nd);
#pragma unroll
              for (int k = 0; k < nd; ++k) ten[u][k] = f3[k];
            }
          }
          if (u0 == 0) {
            FOR_ELEMS(e, nd * nd) sm[S::fx + e] = rec[L.fx + e];
            FOR_ELEMS(e, nd * nu) sm[S::fu + e] = rec[L.fu + e];
            FOR_ELEMS(e, nd + nu) sm[S::small + e] = rec[L.lx + e];  // lx | lu are adjacent
            FOR_ELEMS(e, nQxx) sm[S::Qxx + e] = rec[L.lxx + e];      // Q_* start from l_*
            FOR_ELEMS(e, nQux) sm[S::Qux + e] = rec[L.lux + e];
            FOR_ELEMS(e, nQuu) sm[S::Quu + e] = rec[L.luu + e];
            if (ne > 0) {
              for (int e = tid; e < ne; e += BT) {
                sm[S::eqv + e] = rec[L.eq + e];
                sm[S::pe + e] = bd.m_val[((size_t)eo + e) * B + b];
              }
              for (int e = tid; e < ne * nd; e += BT) {
                sm[S::eqx + e] = rec[L.eq_x + e];
                sm[S::pex + e] = bd.m_jac[((size_t)eo * nd + e) * B + b];
              }
              for (int e = tid; e < ne * nu; e += BT) sm[S::equ + e] = rec[L.eq_u + e];
            }
          }
          double vx[nd], wacc[CH];
#pragma unroll
          for (int k = 0; k < nd; ++k) vx[k] = sm[S::Vx + k];
#pragma unroll
          for (int u = 0; u < CH; ++u) {
            double a = 0.0;
#pragma unroll
            for (int k = 0; k < nd; ++k) a += vx[k] * ten[u][k];
            wacc[u] = a;
          }
#pragma unroll
          for (int u = 0; u < CH; ++u) {
            const int e = tid + (u0 + u) * BT;
            if (u0 + u < EPT && e < NTEN) sm[S::Wf + e] = wacc[u];
          }
        }
      }
      ex.sync();
      // ---- phase B: tmp, tmp2, VF = V_xx f_x, VU = V_xx f_u ----------------------------------------------------
      for (int e = tid; e < ne; e += BT) sm[S::tmp + e] = sm[S::pe + e] + mu * sm[S::eqv + e];
      for (int e = tid; e < ne * nd; e += BT) sm[S::tmp2 + e] = sm[S::pex + e] + mu * sm[S::eqx + e];
      {
        // results are kept in registers until every chain of the lane is done: a shared-memory store between two
        // chains would order them (possible aliasing) and serialise the FMA latencies
        constexpr int NB = nd * nd + nd * nu;
        double acc[(NB + BT - 1) / BT];
        FOR_ELEMS(e, NB) {
          const bool isF = e < nd * nd;
          const int ee = isF ? e : e - nd * nd;
          const int r = ee % nd, c = ee / nd;
          const double* M = isF ? sm + S::fx : sm + S::fu;
          double a = 0.0;
#pragma unroll
          for (int k = 0; k < nd; ++k) a += sm[S::Vxx + k * nd + r] * M[c * nd + k];
          acc[u_] = a;
        }
        FOR_ELEMS(e, NB) sm[S::VF + e] = acc[u_];  // VF | VU are adjacent
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
      // ---- phase C: all Q elements as one task list: Q_xx | Q_ux | Q_uu | Q_x | Q_u ------------------------------
      double qacc[(NTEN + nd + nu + BT - 1) / BT];
      FOR_ELEMS(e, NTEN + nd + nu) {
        if (e < nQxx) {
          const int r = e % nd, c = e / nd;
          double q = sm[S::Qxx + e], a = 0.0;
#pragma unroll
          for (int k = 0; k < nd; ++k) a += fx[r * nd + k] * sm[S::VF + c * nd + k];
          q += a;
          if (ne > 0) {
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
          }
          q += sm[S::Wf + e];
          qacc[u_] = q;
        } else if (e < nQxx + nQux) {
          const int ee = e - nQxx;
          const int r = ee % nu, c = ee / nu;
          double q = sm[S::Qux + ee], a = 0.0;
#pragma unroll
          for (int k = 0; k < nd; ++k) a += fu[r * nd + k] * sm[S::VF + c * nd + k];
          q += a;
          if (ne > 0) {
            a = 0.0;
            for (int k = 0; k < ne; ++k) a += equ[r * ne + k] * tmp2[c * ne + k];
            q += a;
            a = 0.0;
            const double* t3 = rec + L.eq_ux + r * ne + (size_t)c * ne * nu;
            for (int k = 0; k < ne; ++k) a += tmp[k] * t3[k];
            q += a;
          }
          q += sm[S::Wf + e];
          qacc[u_] = q;
        } else if (e < NTEN) {
          const int ee = e - nQxx - nQux;
          const int r = ee % nu, c = ee / nu;
          double q = sm[S::Quu + ee], a = 0.0;
#pragma unroll
          for (int k = 0; k < nd; ++k) a += fu[r * nd + k] * sm[S::VU + c * nd + k];
          q += a;
          if (ne > 0) {
            a = 0.0;
            for (int k = 0; k < ne; ++k) a += equ[r * ne + k] * equ[c * ne + k];
            q += a * mu;
            a = 0.0;
            const double* t3 = rec + L.eq_uu + r * ne + (size_t)c * ne * nu;
            for (int k = 0; k < ne; ++k) a += tmp[k] * t3[k];
            q += a;
          }
          q += sm[S::Wf + e];
          qacc[u_] = q;
        } else if (e < NTEN + nd) {
          const int i = e - NTEN;
          double q = sm[S::small + i], a = 0.0;
#pragma unroll
          for (int k = 0; k < nd; ++k) a += fx[i * nd + k] * Vx[k];
          q += a;
          if (ne > 0) {
            a = 0.0;
            for (int k = 0; k < ne; ++k) a += eqx[i * ne + k] * tmp[k];
            q += a;
            a = 0.0;
            for (int k = 0; k < ne; ++k) a += pex[i * ne + k] * eqv[k];
            q += a;
          }
          qacc[u_] = q;
        } else {
          const int i = e - NTEN - nd;
          double q = sm[S::small + nd + i], a = 0.0;
#pragma unroll
          for (int k = 0; k < nd; ++k) a += fu[i * nd + k] * Vx[k];
          q += a;
          if (ne > 0) {
            a = 0.0;
            for (int k = 0; k < ne; ++k) a += equ[i * ne + k] * tmp[k];
            q += a;
          }
          qacc[u_] = q;
        }
      }
      // Q_xx | Q_ux | Q_uu are adjacent in shared memory and in task-list order; Q_x | Q_u sit before them
      FOR_ELEMS(e, NTEN + nd + nu) sm[e < NTEN ? S::Qxx + e : S::Qx + (e - NTEN)] = qacc[u_];
      ex.sync();
      // ---- phase D: Cholesky of the lower triangle of Q_uu + reg I (Eigen::LLT semantics: fail when a pivot <= 0).
      // Left-looking, column by column; the rows below the pivot are spread over lanes and every sum keeps the
      // serial order, so the factor equals the one-thread result bit for bit.
      {
        double* A = sm + S::L;
        FOR_ELEMS(e, nu * nu) A[e] = sm[S::Quu + e] + ((e / nu == e % nu) ? reg : 0.0);
        if (tid == 0) sm[S::flags] = 1.0;
        ex.sync();
#pragma unroll
        for (int k = 0; k < nu; ++k) {
          if (tid == 0) {
            double x = A[k * nu + k];
#pragma unroll
            for (int j = 0; j < nu; ++j)
              if (j < k) x -= A[j * nu + k] * A[j * nu + k];
            if (x <= 0.0) sm[S::flags] = 0.0;
            A[k * nu + k] = sqrt(x);
          }
          ex.sync();
          if (sm[S::flags] == 0.0) break;
          for (int i = k + 1 + tid; i < nu; i += BT) {
            double s = A[k * nu + i];
#pragma unroll
            for (int j = 0; j < nu; ++j)
              if (j < k) s -= A[j * nu + i] * A[j * nu + k];
            A[k * nu + i] = s / A[k * nu + k];
          }
          ex.sync();
        }
      }
      if (sm[S::flags] == 0.0) {
        if (reg < mu) reg = mu;
        mu *= 2.0;
        reg *= 2.0;
        ++restarts;
        failed = true;
        ex.sync();
        break;
      }
      // ---- phase E: k = -(LL^T)^-1 Q_u, K = -(LL^T)^-1 Q_ux: one right-hand side per lane ----------------------
      FOR_ELEMS(c, nd + 1) {
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
      if (tid < nx) bd.fb_origin[((size_t)t * nx + tid) * B + b] = xorg;
      ex.sync();
      // ---- phase F: V_x = Q_x + Q_ux^T k ; V_xx = Q_xx + Q_ux^T K ---------------------------------------------------
      {
        double vacc[(nd * nd + nd + BT - 1) / BT];
        FOR_ELEMS(e, nd * nd + nd) {
          if (e < nd * nd) {
            const int r = e % nd, c = e / nd;
            double a = 0.0;
#pragma unroll
            for (int k = 0; k < nu; ++k) a += sm[S::Qux + r * nu + k] * sm[S::KK + c * nu + k];
            vacc[u_] = sm[S::Qxx + e] + a;
          } else {
            const int i = e - nd * nd;
            double a = 0.0;
#pragma unroll
            for (int k = 0; k < nu; ++k) a += sm[S::Qux + i * nu + k] * sm[S::kk + k];
            vacc[u_] = sm[S::Qx + i] + a;
          }
        }
        FOR_ELEMS(e, nd * nd + nd) sm[S::Vxx + e] = vacc[u_];  // V_xx | V_x are adjacent
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
#undef FOR_ELEMS
  if (tid == 0) {
    bd.scal[(size_t)S_MU * B + b] = mu;
    bd.scal[(size_t)S_REG * B + b] = reg;
    bd.scal[(size_t)S_RESTARTS * B + b] += (double)restarts;
  }
}

}  // namespace ddpb
