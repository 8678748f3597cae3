// backward_lg.cuh -- Riccati-like backward recursion for models of <= 2 joints: one LANE GROUP of ndx^2 lanes per problem
// (16 lanes for the 2R chain: two problems per warp), everything in registers, lanes exchange operands with shuffles.
//
// Restates ddp_solver_t::backward_pass (/root/reference/include/ddp/ddp_bwd.ipp:9-155) exactly as backward.cuh does
// (same terms, the reference's order of additions); what differs is the work split.  With n_dx <= 4 a knot is a few
// hundred flops.  The recursion is a latency chain -- T dependent knots, and a problem whose Cholesky fails restarts its
// whole pass (ddp_bwd.ipp:105-132) -- so the time of a launch is (longest chain of restarts) x T x (latency of one knot).
// Measured alternatives: one warp per problem working through shared memory (round 1: ~870 warp
// instructions and 4.8 us per knot under load) and one thread per problem (no idle lanes, 32 problems per instruction,
// but ~1300 dependent instructions = 2.4 us per knot and a warp that waits for its slowest lane).  Here lane (r, c) of a
// group owns V_xx(r, c), VF(r, c), Q_xx(r, c): every phase of a knot is 2-3 FMAs per lane plus a few shuffles, ~150 warp
// instructions per knot for two problems, and the dependent path of a knot is ~400 cycles.
//
//   * records are streamed by the group's own lanes with 16-byte cp.async into a 4-deep per-group ring (three knots ahead
//     of the one being computed), so the chain never waits for HBM;
//   * every group carries its own knot index: a failing problem restarts on the spot, the other group of the warp keeps going;
//   * knots with equality rows add their terms per lane (runtime loops over the rows, read from global memory).
//
// The same body runs on the CPU for tests/hostemu: LG_FOR(g) is "this lane" on the device and a loop over the group's
// lanes on the host, LV(x) the lane's copy of a per-lane variable, XL(x, src) lane src's copy (a shuffle on the device).
// Discipline: a value is produced in one LG_FOR block and read across lanes only in a later one.
#pragma once
#include "backward.cuh"

namespace ddpb {

// compile-time offsets of the state-dependent part of a stored record (batch.cuh rec_layout with ne = 0)
template <int NJ, bool DENSE> struct LgRec {
  static constexpr int n = NJ, nd = 2 * NJ, nu = NJ, nx = 2 * NJ;
  static constexpr int LG = nd * nd;          // lanes per problem
  static constexpr int RT = DENSE ? nd : n;   // stored rows per tensor column
  static constexpr int R0 = DENSE ? 0 : n;    // first stored row
  static constexpr int RI = DENSE ? nd : n;   // stored x-perturbation slices of f_ux
  static constexpr int o_f = 0, o_fx = nx, o_fu = o_fx + n * nd, o_lu = o_fu + n * nu, o_fxx = o_lu + nu, o_fux = o_fxx + RT * nd * nd,
                       o_fuu = o_fux + RT * nu * RI, o_end = o_fuu + (DENSE ? nd * nu * nu : 0);
  static constexpr int NCH = (o_end - o_fx + 1) / 2;  // 16-byte chunks staged per knot (f is not used by the backward pass)
  static constexpr int NS = 8;                        // ring depth (knots in flight: HBM latency / time of one knot)
  static constexpr int GPW = 32 / LG;                 // groups (problems) per warp
  static constexpr int STAGE = 2 * NCH;               // doubles per ring slot
  static constexpr int smem_per_warp = GPW * NS * STAGE;
};

#if defined(__CUDA_ARCH__)
#define LG_NL 1
#define LG_FOR(g) for (int g = lgx_, once_ = 1; once_; once_ = 0)
#define LV(x) x[0]
#define XL(x, src) __shfl_sync(0xffffffffu, x[0], gbase_ + (src))
#else
#define LG_NL (LgRec<NJ, DENSE>::LG)
#define LG_FOR(g) for (int g = 0; g < LgRec<NJ, DENSE>::LG; ++g)
#define LV(x) x[g]
#define XL(x, src) x[src]
#endif

// One knot for one problem, executed by its lane group.  R(e) = stored element e of the knot's record, V / X = the lanes'
// V_xx(r, c) / V_x(g) (lanes g < nd).  Returns false when the Cholesky factorisation of Q_uu + reg I meets a
// non-positive pivot (nothing is written then); the result is uniform over the group.
template <int NJ, bool DENSE, class Rd>
DDPB_HD bool lg_knot(const BatchDev& bd, int b, int t, bool live, const Rd& R, double mu, double reg, double* V, double* X, int lgx_, int gbase_) {
  using O = LgRec<NJ, DENSE>;
  constexpr int n = NJ, nd = 2 * NJ, nu = NJ, nx = 2 * NJ, RT = O::RT, R0 = O::R0, RI = O::RI;
  (void)lgx_; (void)gbase_;
  const int B = bd.B;
  const double dt = bd.dt;
  const int ne = live ? bd.ne[t] : 0;
  // equality blocks (ne > 0 only: their offsets are not even computed otherwise), read from global memory where they are used
  RecLayout L;
  const double* GR = nullptr;
  size_t eo = 0;
  if (ne > 0) { L = rec_layout(n, ne, bd.hess_mode); GR = slab_rec(bd, t, b, L.size); eo = bd.eq_off[t]; }
  auto eqv = [&](int k) -> double { return GR[L.eq + k]; };
  auto eqx = [&](int k, int c) -> double { return GR[L.eq_x + c * ne + k]; };
  auto equ = [&](int k, int c) -> double { return GR[L.eq_u + c * ne + k]; };
  auto pe = [&](int k) -> double { return bd.m_val[(eo + k) * B + b]; };
  auto pex = [&](int k, int c) -> double { return bd.m_jac[(eo * nd + c * ne + k) * B + b]; };
  auto tmp = [&](int k) -> double { return pe(k) + mu * eqv(k); };                  // p + mu eq       (ddp_bwd.ipp:46)
  auto tmp2 = [&](int k, int c) -> double { return pex(k, c) + mu * eqx(k, c); };   // P + mu eq_x     (ddp_bwd.ipp:47)

  // ---- block 1: VF = V_xx f_x, VU = V_xx f_u, tensor contractions with V_x, Q_x, Q_u ------------------------------------
  // rows k < n of f_x are [I, dt I], those of f_u zero (batch.cuh)
  double VF[LG_NL], VU[LG_NL], W[LG_NL], WU[LG_NL], WUU[LG_NL], QX[LG_NL], FR[LG_NL][n], GRr[LG_NL][n];
  LG_FOR(g) {
    const int r = g % nd, c = g / nd;
    double vrow[n], xk[RT], xn[n];
#pragma unroll
    for (int k = 0; k < n; ++k) vrow[k] = XL(V, r + nd * (n + k));
    const double vtop = XL(V, r + nd * (c < n ? c : c - n));
#pragma unroll
    for (int k = 0; k < RT; ++k) xk[k] = XL(X, R0 + k);
#pragma unroll
    for (int k = 0; k < n; ++k) xn[k] = DENSE ? XL(X, n + k) : xk[k];
    const double xtop = XL(X, r < n ? r : r - n);
    double vf = c < n ? vtop : dt * vtop;
    double vu = 0.0;
#pragma unroll
    for (int k = 0; k < n; ++k) {
      vf += vrow[k] * R(O::o_fx + k + n * c);
      vu += vrow[k] * R(O::o_fu + k + n * (c < nu ? c : 0));
      LV(FR)[k] = R(O::o_fx + k + n * r);
      LV(GRr)[k] = R(O::o_fu + k + n * (r < nu ? r : 0));
    }
    LV(VF) = vf;
    LV(VU) = vu;  // meaningful for c < nu
    // tensor term of Q_xx(r, c); for lanes r < nu also of Q_ux(r, c) [c < RI] and Q_uu(r, c) [value-based mode, c < nu]
    double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0, c0 = 0.0, c1 = 0.0;
#pragma unroll
    for (int k = 0; k < RT; ++k) {
      const double w = xk[k] * R(O::o_fxx + k + RT * (r + nd * c));
      const double wu = xk[k] * R(O::o_fux + k + RT * ((r < nu ? r : 0) + nu * (c < RI ? c : 0)));
      const double wuu = DENSE ? xk[k] * R(O::o_fuu + k + RT * ((r < nu ? r : 0) + nu * (c < nu ? c : 0))) : 0.0;
      if (k & 1) { a1 += w; b1 += wu; c1 += wuu; } else { a0 += w; b0 += wu; c0 += wuu; }
    }
    LV(W) = a0 + a1;
    LV(WU) = (c < RI) ? (b0 + b1) : 0.0;   // tensor term of Q_ux(r, c): slices c >= RI are structural zeros
    LV(WUU) = c0 + c1;
    // Q_x(r) (used from the lanes c = 0) and Q_u(r), r < nu (used from the lanes c = nu)
    double qx = r < n ? xtop : dt * xtop;
    double qu = 0.0;
#pragma unroll
    for (int k = 0; k < n; ++k) { qx += LV(FR)[k] * xn[k]; qu += LV(GRr)[k] * xn[k]; }
    qu = R(O::o_lu + (r < nu ? r : 0)) + qu;
    if (ne > 0) {
      double a = 0.0;
      for (int k = 0; k < ne; ++k) a += eqx(k, r) * tmp(k);
      qx += a;
      a = 0.0;
      for (int k = 0; k < ne; ++k) a += pex(k, r) * eqv(k);
      qx += a;
      if (r < nu) {
        a = 0.0;
        for (int k = 0; k < ne; ++k) a += equ(k, r) * tmp(k);
        qu += a;
      }
    }
    LV(QX) = qx;
    LV(VU) = (c < nu) ? vu : qu;  // lanes c >= nu have no VU: lane (r < nu, c = nu) carries Q_u(r) instead
  }
  // ---- block 2: Q_xx, Q_ux, Q_uu = l + f^T (V f) [+ equality terms] + tensor term, in the reference's order of additions ----
  double QXX[LG_NL], QUX[LG_NL], QUU[LG_NL];
  LG_FOR(g) {
    const int r = g % nd, c = g / nd;
    double vfk[n], vuk[n];
#pragma unroll
    for (int k = 0; k < n; ++k) { vfk[k] = XL(VF, (n + k) + nd * c); vuk[k] = XL(VU, (n + k) + nd * (c < nu ? c : 0)); }
    const double vftop = XL(VF, (r < n ? r : r - n) + nd * c);
    double qxx = r < n ? vftop : dt * vftop;
    double qux = 0.0, quu = 0.0;
#pragma unroll
    for (int k = 0; k < n; ++k) { qxx += LV(FR)[k] * vfk[k]; qux += LV(GRr)[k] * vfk[k]; quu += LV(GRr)[k] * vuk[k]; }
    quu = (r == c ? bd.c : 0.0) + quu;
    if (ne > 0) {
      double a = 0.0;
      for (int k = 0; k < ne; ++k) a += eqx(k, r) * tmp2(k, c);
      qxx += a;
      a = 0.0;
      for (int k = 0; k < ne; ++k) a += pex(k, r) * eqx(k, c);
      qxx += a;
      a = 0.0;
      for (int k = 0; k < ne; ++k) a += tmp(k) * GR[L.eq_xx + k + ne * (r + nd * c)];
      qxx += a;
      if (r < nu) {
        a = 0.0;
        for (int k = 0; k < ne; ++k) a += equ(k, r) * tmp2(k, c);
        qux += a;
        a = 0.0;
        for (int k = 0; k < ne; ++k) a += tmp(k) * GR[L.eq_ux + k + ne * (r + nu * c)];
        qux += a;
        if (c < nu) {
          a = 0.0;
          for (int k = 0; k < ne; ++k) a += equ(k, r) * equ(k, c);
          quu += a * mu;
          a = 0.0;
          for (int k = 0; k < ne; ++k) a += tmp(k) * GR[L.eq_uu + k + ne * (r + nu * c)];
          quu += a;
        }
      }
    }
    LV(QXX) = qxx + LV(W);
    LV(QUX) = qux + LV(WU);    // lanes r < nu
    LV(QUU) = quu + LV(WUU);   // lanes r < nu, c < nu
  }
  // ---- block 3: LL^T = Q_uu + reg I (Eigen::LLT semantics: fails on a pivot <= 0), k = -(LL^T)^-1 Q_u, K = -(LL^T)^-1 Q_ux ----
  // every lane factorises the same nu x nu matrix; lane (r, c) solves for column c of K, and all lanes for k
  double KC[LG_NL][nu], KK[LG_NL][nu];
  bool ok = true;
  LG_FOR(g) {
    const int r = g % nd, c = g / nd;
    double A[nu][nu], rhsK[nu], rhsk[nu];
#pragma unroll
    for (int i = 0; i < nu; ++i) {
#pragma unroll
      for (int j = 0; j < nu; ++j)
        if (j <= i) A[i][j] = XL(QUU, i + nd * j);
      rhsK[i] = -XL(QUX, i + nd * c);
      rhsk[i] = -XL(VU, i + nd * nu);  // Q_u(i), parked on lane (i, nu) by block 1
    }
    double Lc[nu][nu], dinv[nu];
    bool okl = true;
#pragma unroll
    for (int k = 0; k < nu; ++k) {
      double x = A[k][k] + reg;
#pragma unroll
      for (int j = 0; j < nu; ++j)
        if (j < k) x -= Lc[k][j] * Lc[k][j];
      if (x <= 0.0) okl = false;
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
          double s = A[i][k];
#pragma unroll
          for (int j = 0; j < nu; ++j)
            if (j < k) s -= Lc[i][j] * Lc[k][j];
          Lc[i][k] = s * linv;
        }
    }
    if (!okl) ok = false;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      double* rhs = which ? rhsk : rhsK;
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
    }
#pragma unroll
    for (int i = 0; i < nu; ++i) { LV(KC)[i] = rhsK[i]; LV(KK)[i] = rhsk[i]; }
    if (okl && live) {
      // lane (r < nu, c) owns K(r, c), lane (r < nu, c = 0) also k(r)  (ddp_bwd.ipp:134-135; the origin x_t of :136 is
      // written by fb_origin_knot before the pass, see backward.cuh)
      double kv = rhsK[0], kkv = rhsk[0];
#pragma unroll
      for (int i = 1; i < nu; ++i)
        if (r == i) { kv = rhsK[i]; kkv = rhsk[i]; }
      if (r < nu) {
        bd.fb_K[((size_t)t * nu * nd + c * nu + r) * B + b] = kv;
        if (c == 0) bd.fb_k[((size_t)t * nu + r) * B + b] = kkv;
      }
    }
  }
  // (no early return on failure: the other group of the warp still needs this group's lanes at the shuffles below)
  // ---- block 4: V_x = Q_x + Q_ux^T k ; V_xx = Q_xx + Q_ux^T K   (ddp_bwd.ipp:142-146) ---------------------------------------
  double Vn[LG_NL], Xn[LG_NL];
  LG_FOR(g) {
    const int r = g % nd;
    double a = 0.0, ax = 0.0;
#pragma unroll
    for (int k = 0; k < nu; ++k) {
      const double q = XL(QUX, k + nd * r);  // Q_ux(k, r)
      a += q * LV(KC)[k];
      ax += q * LV(KK)[k];
    }
    LV(Vn) = LV(QXX) + a;
    LV(Xn) = LV(QX) + ax;   // V_x(r), kept by the lanes c = 0
  }
  if (ok) {
    LG_FOR(g) {
      LV(V) = LV(Vn);
      LV(X) = LV(Xn);
    }
  }
  return ok;
}

// per-problem state of the recursion (uniform over the lanes of a group)
struct LgState {
  double mu, reg;
  int t, restarts;
  bool active;
};
// what happens after a knot; returns true when the problem restarts from knot T-1 (its staged records are then stale)
DDPB_HD bool lg_advance(const BatchDev& bd, int b, bool ok, LgState& st, bool writer) {
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
    if (writer) bd.scal[(size_t)S_STATUS * bd.B + b] = 2.0;  // guarded mode: see backward.cuh
    st.active = false;
    return false;
  }
  st.t = bd.T - 1;
  return true;
}
DDPB_HD void lg_finish(const BatchDev& bd, int b, const LgState& st) {
  bd.scal[(size_t)S_MU * bd.B + b] = st.mu;
  bd.scal[(size_t)S_REG * bd.B + b] = st.reg;
  bd.scal[(size_t)S_RESTARTS * bd.B + b] += (double)st.restarts;
}

// CPU emulation of the per-problem loop: the lanes of the group one after the other, records read from the slab
template <int NJ, bool DENSE> inline void backward_lg_plain(const BatchDev& bd, int b) {
  using O = LgRec<NJ, DENSE>;
  double V[O::LG], X[O::LG];
  for (int g = 0; g < O::LG; ++g) V[g] = X[g] = 0.0;   // V_xx = lfxx = 0, V_x = lfx = 0  (problem.hpp:958-959)
  LgState st{bd.scal[(size_t)S_MU * bd.B + b], bd.scal[(size_t)S_REG * bd.B + b], bd.T - 1, 0, true};
  while (st.active) {
    const RecLayout L = rec_layout(NJ, bd.ne[st.t], bd.hess_mode);
    const double* rec = slab_rec(bd, st.t, b, L.size);
    const bool ok = lg_knot<NJ, DENSE>(bd, b, st.t, true, [&](int e) -> double { return rec[e]; }, st.mu, st.reg, V, X, 0, 0);
    if (lg_advance(bd, b, ok, st, true))
      for (int g = 0; g < O::LG; ++g) V[g] = X[g] = 0.0;
  }
  lg_finish(bd, b, st);
}

#if defined(__CUDACC__)
// One warp = GPW problems; dynamic shared memory = LgRec::smem_per_warp doubles per warp.
template <int NJ, bool DENSE> __device__ __forceinline__ void backward_lg_warp(const BatchDev& bd, double* smem_warp) {
  using O = LgRec<NJ, DENSE>;
  constexpr int LG = O::LG, NS = O::NS, NCH = O::NCH, STAGE = O::STAGE;
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int grp = lane / LG, lgx = lane % LG, gbase = grp * LG;
  const int warp_global = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int p = warp_global * O::GPW + grp;
  const bool valid = (32 % LG == 0 || lane < O::GPW * LG) && p < bd.nb;
  const int b = bd.b0 + (valid ? p : 0);
  const int B = bd.B;
  double V[1] = {0.0}, X[1] = {0.0};
  LgState st{0.0, 0.0, bd.T - 1, 0, false};
  if (valid && bd.scal[(size_t)S_STATUS * B + b] == 0.0) {
    st.mu = bd.scal[(size_t)S_MU * B + b];
    st.reg = bd.scal[(size_t)S_REG * B + b];
    st.active = true;
  }
  const bool started = st.active;
  double* ring = smem_warp + (size_t)grp * NS * STAGE;
  // stage the record of knot tk of this group's problem into ring slot tk % NS: chunks lgx, lgx + LG, ...
  auto fetch = [&](int tk) {
    if (tk < 0 || !st.active) return;
    const double* src = slab_rec_t(bd, tk, b) + O::o_fx;
    double* dst = ring + (size_t)(tk % NS) * STAGE;
#pragma unroll
    for (int c = 0; c < (NCH + LG - 1) / LG; ++c) {
      const int ch = lgx + c * LG;
      if (ch < NCH) cp_async16(dst + 2 * ch, src + 2 * ch);
    }
  };
  auto commit = [&]() { asm volatile("cp.async.commit_group;" ::: "memory"); };
  // NS - 1 knots in flight ahead of the one being computed
#pragma unroll
  for (int s = 0; s < NS - 1; ++s) { fetch(st.t - s); commit(); }
  while (__any_sync(FULL, st.active)) {
    fetch(st.t - (NS - 1));
    commit();
    asm volatile("cp.async.wait_group %0;" ::"n"(NS - 1) : "memory");  // the current knot's record has landed ...
    __syncwarp();                                                        // ... and is visible to the whole group
    const double* sc = ring + (size_t)((st.t < 0 ? 0 : st.t) % NS) * STAGE;
    const bool ok = lg_knot<NJ, DENSE>(bd, b, st.t < 0 ? 0 : st.t, st.active, [&](int e) -> double { return sc[e - O::o_fx]; }, st.mu, st.reg, V, X, lgx, gbase);
    bool restart = false;
    if (st.active) {
      if (!ok && bd.dbg && lgx == 0) atomicAdd(bd.dbg + 8 + (st.t * 8) / bd.T, 1ull);  // development: where passes fail (slots 8..15)
      restart = lg_advance(bd, b, ok, st, lgx == 0);
      if (restart) { V[0] = 0.0; X[0] = 0.0; }  // V_xx = lfxx = 0, V_x = lfx = 0  (problem.hpp:958-959)
    }
    __syncwarp();  // every lane is past its reads of the current slot before a later fetch overwrites it
    if (__any_sync(FULL, restart)) {
      // a restarting problem's staged records (knots below the failed one) are stale: refill its ring from knot T-1 once
      // its copies in flight have landed (rare: ~0.1 restarts per problem and pass)
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
#pragma unroll
      for (int s = 0; s < NS - 1; ++s) { if (restart) fetch(st.t - s); commit(); }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (started && lgx == 0) lg_finish(bd, b, st);
}
#endif

}  // namespace ddpb
