// backward_tpp.cuh -- Riccati-like backward recursion, one THREAD per problem, for models of <= 2 joints (and the
// pendulum's 1-joint layout) on the element-major slab ([knot][element][problem], batch.cuh).
//
// Restates ddp_solver_t::backward_pass (/root/reference/include/ddp/ddp_bwd.ipp:9-155) exactly as backward.cuh does
// (same terms, same order of the reference's additions); what differs is the work split.  With n_dx <= 4 a knot is a
// few hundred flops: a CTA (or even a warp) per problem spends its instructions on index arithmetic, shuffles and
// idle lanes (the one-warp-per-problem kernel of round 1 issued ~870 warp instructions per knot for ~10 useful ones per
// lane).  Here V_xx, V_x and every Q block of a problem live in the registers of ONE thread as straight-line fully
// unrolled code, a warp advances 32 problems per instruction, and all slab / gain accesses are coalesced because the
// problem index runs fastest in memory.
//
//   * every lane carries ITS OWN knot index: a lane whose Cholesky fails (ddp_bwd.ipp:105-132) updates its mu / reg and
//     restarts from knot T-1 on the spot while the other lanes of the warp keep going -- a restart costs the warp only
//     the knots that lane had already done, not a whole extra pass;
//   * the stored record of the lane's next knot is fetched with per-thread cp.async (8 bytes each, consecutive lanes ->
//     consecutive addresses) into a double-buffered shared-memory stage [element][lane] while the current knot is
//     computed, so the dependent chain of a knot never waits for HBM;
//   * knots with equality rows (rare: the terminal ones) read their extra blocks straight from global memory.
#pragma once
#include "backward.cuh"

namespace ddpb {

// compile-time offsets of the state-dependent part of a stored record (batch.cuh rec_layout with ne = 0)
template <int NJ, bool DENSE> struct TppRec {
  static constexpr int n = NJ, nd = 2 * NJ, nu = NJ, nx = 2 * NJ;
  static constexpr int RT = DENSE ? nd : n;   // stored rows per tensor column
  static constexpr int R0 = DENSE ? 0 : n;    // first stored row
  static constexpr int RI = DENSE ? nd : n;   // stored x-perturbation slices of f_ux
  static constexpr int o_f = 0, o_fx = nx, o_fu = o_fx + n * nd, o_lu = o_fu + n * nu, o_fxx = o_lu + nu, o_fux = o_fxx + RT * nd * nd,
                       o_fuu = o_fux + RT * nu * RI, o_end = o_fuu + (DENSE ? nd * nu * nu : 0);
  static constexpr int NST = o_end - o_fx;    // staged elements per knot (f is not used by the backward pass)
};

// One knot of the recursion for one problem.  R(e) = stored element e of the knot's record (staged), G = RecView of the
// same record in global memory (equality blocks).  Vxx is row-major [r * nd + c].  Returns false when the Cholesky
// factorisation of Q_uu + reg I meets a non-positive pivot (nothing is written then).
// EQ = false: the knot has no equality rows -- the hot path, compiled without any of the equality code (a few hundred
// straight-line instructions that stay in the instruction cache); EQ = true: the general knot.
template <int NJ, bool DENSE, bool EQ, class Src>
DDPB_HD bool bwd_knot_thread(const BatchDev& bd, int b, int t, const Src& R, double mu, double reg, double* Vxx, double* Vx) {
  using O = TppRec<NJ, DENSE>;
  constexpr int n = NJ, nd = 2 * NJ, nu = NJ, nx = 2 * NJ, RT = O::RT, R0 = O::R0, RI = O::RI;
  const int B = bd.B;
  const double dt = bd.dt;
  const int ne = EQ ? bd.ne[t] : 0;
  // origin = x_t (ddp_bwd.ipp:136): loaded first so that the store at the end never waits for memory
  double xo[nx];
#pragma unroll
  for (int k = 0; k < nx; ++k) xo[k] = bd.xs[((size_t)t * nx + k) * B + b];
  double F[n][nd], G[n][nu], lu[nu];
#pragma unroll
  for (int c = 0; c < nd; ++c)
#pragma unroll
    for (int k = 0; k < n; ++k) F[k][c] = R(O::o_fx + k + n * c);
#pragma unroll
  for (int c = 0; c < nu; ++c)
#pragma unroll
    for (int k = 0; k < n; ++k) G[k][c] = R(O::o_fu + k + n * c);
#pragma unroll
  for (int i = 0; i < nu; ++i) lu[i] = R(O::o_lu + i);
  // VF = V_xx f_x, VU = V_xx f_u: rows k < n of f_x are [I, dt I], those of f_u zero (batch.cuh)
  double VF[nd][nd], VU[nd][nu];
#pragma unroll
  for (int r = 0; r < nd; ++r) {
#pragma unroll
    for (int c = 0; c < nd; ++c) {
      double s = c < n ? Vxx[r * nd + c] : dt * Vxx[r * nd + (c - n)];
#pragma unroll
      for (int k = 0; k < n; ++k) s += Vxx[r * nd + n + k] * F[k][c];
      VF[r][c] = s;
    }
#pragma unroll
    for (int c = 0; c < nu; ++c) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < n; ++k) s += Vxx[r * nd + n + k] * G[k][c];
      VU[r][c] = s;
    }
  }
  // equality blocks of this knot (ne > 0 only), read from global memory where they are used
  const RecLayout L = rec_layout(n, ne, bd.hess_mode);
  const RecView GR = rec_view(bd, t, b, L.size);
  const size_t eo = bd.eq_off[t];
  auto eqv = [&](int k) -> double { return GR[L.eq + k]; };
  auto eqx = [&](int k, int c) -> double { return GR[L.eq_x + c * ne + k]; };
  auto equ = [&](int k, int c) -> double { return GR[L.eq_u + c * ne + k]; };
  auto pe = [&](int k) -> double { return bd.m_val[(eo + k) * B + b]; };
  auto pex = [&](int k, int c) -> double { return bd.m_jac[(eo * nd + c * ne + k) * B + b]; };
  auto tmp = [&](int k) -> double { return pe(k) + mu * eqv(k); };                  // p + mu eq       (ddp_bwd.ipp:46)
  auto tmp2 = [&](int k, int c) -> double { return pex(k, c) + mu * eqx(k, c); };   // P + mu eq_x     (ddp_bwd.ipp:47)
  // Q = l + f^T (V f) [+ equality terms] + tensor term, in the reference's order of additions (ddp_bwd.ipp:61-87)
  double Qxx[nd][nd], Qux[nu][nd], Quu[nu][nu], Qx[nd], Qu[nu];
#pragma unroll
  for (int c = 0; c < nd; ++c) {
#pragma unroll
    for (int r = 0; r < nd; ++r) {
      double q = r < n ? VF[r][c] : dt * VF[r - n][c];
#pragma unroll
      for (int k = 0; k < n; ++k) q += F[k][r] * VF[n + k][c];
      if (EQ && ne > 0) {
        double a = 0.0;
        for (int k = 0; k < ne; ++k) a += eqx(k, r) * tmp2(k, c);
        q += a;
        a = 0.0;
        for (int k = 0; k < ne; ++k) a += pex(k, r) * eqx(k, c);
        q += a;
        a = 0.0;
        for (int k = 0; k < ne; ++k) a += tmp(k) * GR[L.eq_xx + k + ne * (r + nd * c)];
        q += a;
      }
      double a0 = 0.0, a1 = 0.0;
#pragma unroll
      for (int k = 0; k < RT; ++k) {
        const double v = Vx[R0 + k] * R(O::o_fxx + k + RT * (r + nd * c));
        if (k & 1) a1 += v; else a0 += v;
      }
      Qxx[r][c] = q + (a0 + a1);
    }
#pragma unroll
    for (int r = 0; r < nu; ++r) {
      double q = 0.0;
#pragma unroll
      for (int k = 0; k < n; ++k) q += G[k][r] * VF[n + k][c];
      if (EQ && ne > 0) {
        double a = 0.0;
        for (int k = 0; k < ne; ++k) a += equ(k, r) * tmp2(k, c);
        q += a;
        a = 0.0;
        for (int k = 0; k < ne; ++k) a += tmp(k) * GR[L.eq_ux + k + ne * (r + nu * c)];
        q += a;
      }
      if (c < RI) {
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int k = 0; k < RT; ++k) {
          const double v = Vx[R0 + k] * R(O::o_fux + k + RT * (r + nu * (c < RI ? c : 0)));
          if (k & 1) a1 += v; else a0 += v;
        }
        q += a0 + a1;
      }
      Qux[r][c] = q;
    }
  }
#pragma unroll
  for (int c = 0; c < nu; ++c)
#pragma unroll
    for (int r = 0; r < nu; ++r) {
      double q = r == c ? bd.c : 0.0;
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < n; ++k) s += G[k][r] * VU[n + k][c];
      q += s;
      if (EQ && ne > 0) {
        double a = 0.0;
        for (int k = 0; k < ne; ++k) a += equ(k, r) * equ(k, c);
        q += a * mu;
        a = 0.0;
        for (int k = 0; k < ne; ++k) a += tmp(k) * GR[L.eq_uu + k + ne * (r + nu * c)];
        q += a;
      }
      if (DENSE) {
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int k = 0; k < RT; ++k) {
          const double v = Vx[R0 + k] * R(O::o_fuu + k + RT * (r + nu * c));
          if (k & 1) a1 += v; else a0 += v;
        }
        q += a0 + a1;
      }
      Quu[r][c] = q;
    }
#pragma unroll
  for (int i = 0; i < nd; ++i) {
    double q = i < n ? Vx[i] : dt * Vx[i - n];
#pragma unroll
    for (int k = 0; k < n; ++k) q += F[k][i] * Vx[n + k];
    if (EQ && ne > 0) {
      double a = 0.0;
      for (int k = 0; k < ne; ++k) a += eqx(k, i) * tmp(k);
      q += a;
      a = 0.0;
      for (int k = 0; k < ne; ++k) a += pex(k, i) * eqv(k);
      q += a;
    }
    Qx[i] = q;
  }
#pragma unroll
  for (int i = 0; i < nu; ++i) {
    double q = lu[i];
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < n; ++k) s += G[k][i] * Vx[n + k];
    q += s;
    if (EQ && ne > 0) {
      double a = 0.0;
      for (int k = 0; k < ne; ++k) a += equ(k, i) * tmp(k);
      q += a;
    }
    Qu[i] = q;
  }
  // LL^T = Q_uu + reg I (lower triangle, Eigen::LLT semantics: fails on a pivot <= 0; ddp_bwd.ipp:102-104)
  double Lc[nu][nu], dinv[nu];
  bool ok = true;
#pragma unroll
  for (int k = 0; k < nu; ++k) {
    double x = Quu[k][k] + reg;
#pragma unroll
    for (int j = 0; j < nu; ++j)
      if (j < k) x -= Lc[k][j] * Lc[k][j];
    if (x <= 0.0) ok = false;
#if defined(__CUDA_ARCH__)
    const double linv = rsqrt(x);
#else
    const double linv = 1.0 / sqrt(x);
#endif
    Lc[k][k] = x * linv;
    dinv[k] = linv;
#pragma unroll
    for (int i = 0; i < nu; ++i)
      if (i > k) {
        double s = Quu[i][k];
#pragma unroll
        for (int j = 0; j < nu; ++j)
          if (j < k) s -= Lc[i][j] * Lc[k][j];
        Lc[i][k] = s * linv;
      }
  }
  if (!ok) return false;
  // k = -(LL^T)^-1 Q_u ; K = -(LL^T)^-1 Q_ux   (ddp_bwd.ipp:134-135)
  double kk[nu], KK[nu][nd];
#pragma unroll
  for (int c = 0; c <= nd; ++c) {
    double rhs[nu];
#pragma unroll
    for (int i = 0; i < nu; ++i) rhs[i] = -(c == nd ? Qu[i] : Qux[i][c < nd ? c : 0]);
#pragma unroll
    for (int i = 0; i < nu; ++i) {
      double s = rhs[i];
#pragma unroll
      for (int j = 0; j < nu; ++j)
        if (j < i) s -= Lc[i][j] * rhs[j];
      rhs[i] = s * dinv[i];
    }
#pragma unroll
    for (int i = nu - 1; i >= 0; --i) {
      double s = rhs[i];
#pragma unroll
      for (int j = 0; j < nu; ++j)
        if (j > i) s -= Lc[j][i] * rhs[j];
      rhs[i] = s * dinv[i];
    }
#pragma unroll
    for (int i = 0; i < nu; ++i) {
      if (c == nd) {
        kk[i] = rhs[i];
        bd.fb_k[((size_t)t * nu + i) * B + b] = rhs[i];
      } else {
        KK[i][c < nd ? c : 0] = rhs[i];
        bd.fb_K[((size_t)t * nu * nd + c * nu + i) * B + b] = rhs[i];
      }
    }
  }
  // origin = x_t (ddp_bwd.ipp:136)
#pragma unroll
  for (int k = 0; k < nx; ++k) bd.fb_origin[((size_t)t * nx + k) * B + b] = xo[k];
  // V_x = Q_x + Q_ux^T k ; V_xx = Q_xx + Q_ux^T K   (ddp_bwd.ipp:142-146)
#pragma unroll
  for (int i = 0; i < nd; ++i) {
    double a = 0.0;
#pragma unroll
    for (int k = 0; k < nu; ++k) a += Qux[k][i] * kk[k];
    Vx[i] = Qx[i] + a;
  }
#pragma unroll
  for (int r = 0; r < nd; ++r)
#pragma unroll
    for (int c = 0; c < nd; ++c) {
      double a = 0.0;
#pragma unroll
      for (int k = 0; k < nu; ++k) a += Qux[k][r] * KK[k][c];
      Vxx[r * nd + c] = Qxx[r][c] + a;
    }
  return true;
}

// per-problem state machine shared by the device kernel and the CPU emulation: what happens after a knot
struct TppState {
  double mu, reg;
  int t, restarts;
  bool active;
};
// returns true when the lane must restart from knot T-1 (its staged next record is then stale)
template <int ND> DDPB_HD bool tpp_advance(const BatchDev& bd, int b, bool ok, TppState& st, double* Vxx, double* Vx) {
  if (ok) {
    if (--st.t < 0) st.active = false;
    return false;
  }
  // failure -> reg = max(reg, mu), mu *= 2, reg *= 2, restart from t = T-1   (ddp_bwd.ipp:105-132)
  if (st.reg < st.mu) st.reg = st.mu;
  st.mu *= 2.0;
  st.reg *= 2.0;
  ++st.restarts;
  if (bd.max_restarts > 0 && (st.restarts > bd.max_restarts || !(st.mu < 1.79e308))) {
    bd.scal[(size_t)S_STATUS * bd.B + b] = 2.0;  // guarded mode: see backward.cuh
    st.active = false;
    return false;
  }
#pragma unroll
  for (int e = 0; e < ND * ND; ++e) Vxx[e] = 0.0;  // V_xx = lfxx = 0, V_x = lfx = 0  (problem.hpp:958-959)
#pragma unroll
  for (int e = 0; e < ND; ++e) Vx[e] = 0.0;
  st.t = bd.T - 1;
  return true;
}
DDPB_HD void tpp_finish(const BatchDev& bd, int b, const TppState& st) {
  bd.scal[(size_t)S_MU * bd.B + b] = st.mu;
  bd.scal[(size_t)S_REG * bd.B + b] = st.reg;
  bd.scal[(size_t)S_RESTARTS * bd.B + b] += (double)st.restarts;
}

// CPU emulation / reference statement of the per-problem loop (no staging: elements are read from the slab)
template <int NJ, bool DENSE> DDPB_HD void backward_thread_plain(const BatchDev& bd, int b) {
  constexpr int nd = 2 * NJ;
  double Vxx[nd * nd], Vx[nd];
  for (int e = 0; e < nd * nd; ++e) Vxx[e] = 0.0;
  for (int e = 0; e < nd; ++e) Vx[e] = 0.0;
  TppState st{bd.scal[(size_t)S_MU * bd.B + b], bd.scal[(size_t)S_REG * bd.B + b], bd.T - 1, 0, true};
  while (st.active) {
    const RecLayout L = rec_layout(NJ, bd.ne[st.t], bd.hess_mode);
    const RecView rv = rec_view(bd, st.t, b, L.size);
    auto src = [&](int e) -> double { return rv[e]; };
    const bool ok = bd.ne[st.t] == 0 ? bwd_knot_thread<NJ, DENSE, false>(bd, b, st.t, src, st.mu, st.reg, Vxx, Vx)
                                     : bwd_knot_thread<NJ, DENSE, true>(bd, b, st.t, src, st.mu, st.reg, Vxx, Vx);
    tpp_advance<nd>(bd, b, ok, st, Vxx, Vx);
  }
  tpp_finish(bd, b, st);
}

#if defined(__CUDACC__)
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
// One warp = 32 problems; dynamic shared memory = 2 stages x NST elements x 32 lanes doubles per warp.
template <int NJ, bool DENSE> __device__ __forceinline__ void backward_tpp_warp(const BatchDev& bd, double* stage_warp) {
  using O = TppRec<NJ, DENSE>;
  constexpr int nd = 2 * NJ, NST = O::NST;
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int i = (int)((blockIdx.x * blockDim.x + threadIdx.x));
  const int b = bd.b0 + i;
  const bool valid = i < bd.nb;
  const int B = bd.B;
  double Vxx[nd * nd], Vx[nd];
#pragma unroll
  for (int e = 0; e < nd * nd; ++e) Vxx[e] = 0.0;
#pragma unroll
  for (int e = 0; e < nd; ++e) Vx[e] = 0.0;
  TppState st{0.0, 0.0, bd.T - 1, 0, false};
  if (valid && bd.scal[(size_t)S_STATUS * B + b] == 0.0) {
    st.mu = bd.scal[(size_t)S_MU * B + b];
    st.reg = bd.scal[(size_t)S_REG * B + b];
    st.active = true;
  }
  const bool started = st.active;
  double* stg = stage_warp + lane;  // element e of stage s at stg[(s * NST + e) * 32]
  auto fetch = [&](int t, int s) {
    const double* src = bd.slab + ((size_t)bd.slab_off[t] + O::o_fx) * B + b;
    double* dst = stg + (size_t)s * NST * 32;
#pragma unroll
    for (int e = 0; e < NST; ++e) cp_async8(dst + e * 32, src + (size_t)e * B);
  };
  int cur = 0;
  if (st.active) fetch(st.t, cur);
  asm volatile("cp.async.commit_group;" ::: "memory");
  while (__any_sync(FULL, st.active)) {
    // the record of the lane's next knot goes into the other stage while this one is computed
    if (st.active && st.t > 0) fetch(st.t - 1, cur ^ 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");  // the current stage has landed
    bool restart = false;
    if (st.active) {
      const double* sc = stg + (size_t)cur * NST * 32;
      auto src = [&](int e) -> double { return sc[(e - O::o_fx) * 32]; };
      bool ok;
      if (bd.ne[st.t] == 0) ok = bwd_knot_thread<NJ, DENSE, false>(bd, b, st.t, src, st.mu, st.reg, Vxx, Vx);
      else ok = bwd_knot_thread<NJ, DENSE, true>(bd, b, st.t, src, st.mu, st.reg, Vxx, Vx);
      if (!ok && bd.dbg) atomicAdd(bd.dbg + 8 + (st.t * 8) / bd.T, 1ull);  // development: where in the horizon passes fail (slots 8..15)
      restart = tpp_advance<nd>(bd, b, ok, st, Vxx, Vx);
    }
    if (__any_sync(FULL, restart)) {
      // a restarting lane's staged record (knot t_failed - 1) is stale: fetch knot T-1 into the same stage once the
      // copy in flight has landed (rare: ~0.1 restarts per problem and pass)
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      if (restart) fetch(st.t, cur ^ 1);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    cur ^= 1;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (started) tpp_finish(bd, b, st);
}
#endif

}  // namespace ddpb
