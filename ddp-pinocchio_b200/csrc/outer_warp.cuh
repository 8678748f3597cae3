// outer_warp.cuh -- device-only, latency-oriented versions of the per-problem sweeps of the outer loop.
// The thread-per-problem bodies in forward.cuh stay as the CPU-emulated reference of the same arithmetic
// (tests/hostemu); these kernels split them by their parallelism:
//   origin_update_knot   : mults/feedback re-centering (mat_seq_common.hpp:62-89) -- independent per (knot, problem)
//   optimality_warp      : adjoint recursion of optimality_obj (ddp.hpp:575-627) -- sequential in t, one WARP per
//                          problem: lane i < ndx owns adjoint component i, lanes ndx..ndx+nu-1 own l_u rows, the
//                          f_x / f_u columns are read straight from the slab record (contiguous 8*ndx-byte runs) one
//                          knot ahead, the new adjoint is exchanged with shuffles
//   mult_update_knot     : multiplier update (ddp.hpp:686-687) -- independent per (knot, problem)
#pragma once
#if defined(__CUDACC__)
#include "forward.cuh"

namespace ddpb {

template <int NJ> __device__ __forceinline__ void origin_update_knot(const BatchDev& bd, int b, int t) {
  constexpr int n = Dim<NJ>::n, nx = 2 * n, nu = n, nd = 2 * n;
  const int B = bd.B;
  if (bd.scal[(size_t)S_STATUS * B + b] != 0.0) return;
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

template <int NJ> __device__ __forceinline__ void mult_update_knot(const BatchDev& bd, int b, int t) {
  constexpr int n = Dim<NJ>::n, nx = 2 * n, nu = n, nd = 2 * n;
  const int B = bd.B;
  if (bd.scal[(size_t)S_STATUS * B + b] != 0.0 || bd.scal[(size_t)S_FLAG * B + b] != 1.0) return;
  const int ne = bd.ne[t];
  if (ne == 0) return;
  const double mu = bd.scal[(size_t)S_MU * B + b];
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

// one warp: opt_obj of problem b (all lanes return the same value)
template <int NJ> __device__ __forceinline__ double optimality_obj_warp(const BatchDev& bd, int b, double mu) {
  constexpr int n = Dim<NJ>::n, nx = 2 * n, nu = n, nd = 2 * n;
  static_assert(nd + nu <= 32, "one lane per adjoint / control component");
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int B = bd.B;
  const bool isx = lane < nd, isu = lane >= nd && lane < nd + nu;
  const int col = isx ? lane : lane - nd;
  double adj[nd];
#pragma unroll
  for (int k = 0; k < nd; ++k) adj[k] = 0.0;  // lfx = 0
  double retval = 0.0;
  double atop = 0.0;  // the adjoint component that meets this lane's constant row of f_x: adj[col] or adj[col - n]
  // stored rows (k >= n, batch.cuh) of column `col` of f_x (lanes < nd) or f_u (next nu lanes) and the matching l_u
  // entry (l_x = 0), loaded one knot ahead
  double cur[n], nxt[n], lcur = 0.0, lnxt = 0.0;
  auto load = [&](int t, double* dst, double& l) {
    const RecLayout L = rec_layout(n, bd.ne[t], bd.hess_mode);
    const RecView rec = rec_view(bd, t, b, L.size);
    if (isx || isu) {
      const int c0 = (isx ? L.fx : L.fu) + col * n;
#pragma unroll
      for (int k = 0; k < n; ++k) dst[k] = rec[c0 + k];
      l = isx ? 0.0 : rec[L.lu + col];
    }
  };
  load(bd.T - 1, cur, lcur);
  for (int t = bd.T - 1; t >= 0; --t) {
    if (t > 0) load(t - 1, nxt, lnxt);
    const int ne = bd.ne[t];
    // adj . column: rows k < n of f_x are the constants [I, dt I], those of f_u zeros
    double dotc = isx ? (col < n ? atop : atop * bd.dt) : 0.0;
#pragma unroll
    for (int k = 0; k < n; ++k) dotc += adj[n + k] * cur[k];
    double v;
    if (ne == 0) {
      // lu = l_u + adj f_u ; adj' = adj f_x   (l_x = 0, the eq terms are empty sums)
      v = isx ? dotc : lcur + dotc;
    } else {
      const size_t eo = bd.eq_off[t];
      const RecLayout L = rec_layout(n, ne, bd.hess_mode);
      const RecView rec = rec_view(bd, t, b, L.size);
      double pe[MAXE];
      for (int i = 0; i < ne; ++i) {
        double pv = bd.m_val[(eo + i) * B + b], a = 0.0;
#pragma unroll
        for (int k = 0; k < nd; ++k)
          a += bd.m_jac[(eo * nd + k * ne + i) * B + b] * (bd.xs[((size_t)t * nx + k) * B + b] - bd.m_origin[((size_t)t * nx + k) * B + b]);
        pe[i] = pv + a;
      }
      if (isu) {
        v = lcur;
        double a = 0.0;
        for (int k = 0; k < ne; ++k) a += pe[k] * rec[L.eq_u + col * ne + k];
        v += a;
        a = 0.0;
        for (int k = 0; k < ne; ++k) a += (mu * rec[L.eq + k]) * rec[L.eq_u + col * ne + k];
        v += a;
        v += dotc;
      } else if (isx) {
        v = dotc;
        double a = 0.0;
        for (int k = 0; k < ne; ++k) a += (mu * rec[L.eq + k]) * rec[L.eq_x + col * ne + k];
        v += a;
        a = 0.0;
        for (int k = 0; k < ne; ++k) a += pe[k] * rec[L.eq_x + col * ne + k];
        v += a;
        a = 0.0;
        for (int k = 0; k < ne; ++k) a += rec[L.eq + k] * bd.m_jac[(eo * nd + col * ne + k) * B + b];
        v += a;
      } else {
        v = 0.0;
      }
    }
    // |lu|^2 in the reference's order (i = 0 .. nu-1), then the new adjoint to every lane
    double sq = 0.0;
#pragma unroll
    for (int i = 0; i < nu; ++i) { const double li = __shfl_sync(FULL, v, nd + i); sq += li * li; }
    const double nrm = sqrt(sq);
    if (retval < nrm) retval = nrm;
#pragma unroll
    for (int k = 0; k < nd; ++k) adj[k] = __shfl_sync(FULL, v, k);
    atop = __shfl_sync(FULL, v, isx ? (col < n ? col : col - n) : 0);
#pragma unroll
    for (int k = 0; k < n; ++k) cur[k] = nxt[k];
    lcur = lnxt;
  }
  return retval;
}

// optimality measures + the decision of update_derivatives / solve's switch (ddp.hpp:660-695, 786-802) for one problem
// per warp.  phase 0: first sweep and decision (S_FLAG: 1 = multipliers will be updated); phase 1: after the
// multiplier update, second sweep for the flagged problems (n = opt_obj / mu^0.1, w /= mu).
template <int NJ> __device__ __forceinline__ void al_decide_warp(const BatchDev& bd, int b, double thr, int phase) {
  const int B = bd.B;
  const int lane = threadIdx.x & 31;
  if (bd.scal[(size_t)S_STATUS * B + b] != 0.0) return;
  double mu = bd.scal[(size_t)S_MU * B + b];
  if (phase == 0) {
    const double opt_obj = optimality_obj_warp<NJ>(bd, b, mu);
    if (lane != 0) return;
    const double opt_constr = optimality_constr_problem<NJ>(bd, b);
    bd.scal[(size_t)S_OPT_OBJ * B + b] = opt_obj;
    bd.scal[(size_t)S_OPT_CONSTR * B + b] = opt_constr;
    double flag = 0.0;
    if (opt_constr < thr && opt_obj < thr) {
      bd.scal[(size_t)S_STATUS * B + b] = 1.0;  // optimum attained
    } else if (opt_obj < bd.scal[(size_t)S_W * B + b]) {
      if (opt_constr < bd.scal[(size_t)S_N * B + b]) flag = 1.0;
      else bd.scal[(size_t)S_MU * B + b] = mu * 10.0;
    }
    bd.scal[(size_t)S_FLAG * B + b] = flag;
  } else {
    if (bd.scal[(size_t)S_FLAG * B + b] != 1.0) return;
    const double oo = optimality_obj_warp<NJ>(bd, b, mu);
    if (lane != 0) return;
    bd.scal[(size_t)S_N * B + b] = oo / pow(mu, 0.1);
    bd.scal[(size_t)S_W * B + b] = bd.scal[(size_t)S_W * B + b] / mu;
    bd.scal[(size_t)S_MULT_UPDATES * B + b] += 1.0;
  }
}

// ---- line search with concurrent trials ---------------------------------------------------------------------
// The reference tries step = 1, 1/2, 1/4, ... one rollout after the other and takes the first step whose merit does not
// increase (ddp_fwd.ipp:29-60).  The trials are independent, so a block evaluates LS_NC of them at once: block =
// LS_NC warps, warp w = trial slot w of 32 consecutive problems (lane = problem => coalesced SoA accesses).  Every trial
// rolls out into its own scratch trajectory; the smallest accepted trial index wins exactly as in the sequential scan
// (34 trials at most, last one kept on give-up, ddp_fwd.ipp:35-37), and the block copies the winner into the
// new-trajectory storage.  The `rollouts` scalar keeps counting what the sequential algorithm would have run.
// Trials per round: 8 for models of <= 2 joints (a 4096-problem batch is then 1024 warps, ~2 per SM sub-partition; almost
// every launch has some problem that needs one of the trials 4..7, which used to cost the whole block a second round),
// 4 for the arms (their per-thread dynamics scratch fills shared memory).
template <int NJ> struct LsGeom { static constexpr int NC = Dim<NJ>::n <= 2 ? 8 : 4; };
constexpr int LS_MAX_TRIALS = 34;  // 2^-33 >= 1e-10 > 2^-34

template <int NJ>
__device__ __forceinline__ void linesearch_block(const ModelDev& m, const BatchDev& bd, const double* costs_old, double* xs_c, double* us_c, const Scratch& sc) {
  constexpr int n = Dim<NJ>::n, nx = 2 * n, nu = n, LS_NC = LsGeom<NJ>::NC;
  const int B = bd.B, T = bd.T;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int b = bd.b0 + blockIdx.x * 32 + lane;
  const bool valid = b < bd.b0 + bd.nb;
  const size_t xs_slot = (size_t)(T + 1) * nx * B, us_slot = (size_t)T * nu * B;
  __shared__ double s_sum[LS_NC][32], s_merit[LS_NC][32], s_plain[LS_NC][32];
  __shared__ int s_win[32];
  const bool frozen = valid && bd.scal[(size_t)S_STATUS * B + b] != 0.0;
  const double mu = valid ? bd.scal[(size_t)S_MU * B + b] : 0.0;
  if (w == 0) s_win[lane] = (valid && !frozen) ? -1 : -2;  // -1 unresolved, -2 nothing to do
  __syncthreads();
  int win = -1;
  for (int r = 0; r * LS_NC < LS_MAX_TRIALS; ++r) {
    const int c = r * LS_NC + w;
    const bool run = valid && !frozen && s_win[lane] == -1 && c < LS_MAX_TRIALS;
    double sum = 1.0, merit = 0.0, plain = 0.0;
    // trial slot 0 rolls out straight into the new-trajectory storage: when the first trial of a round wins (the
    // usual case: the full step) nothing has to be copied
    double* xo = w == 0 ? bd.xs_new : xs_c + w * xs_slot;
    double* uo = w == 0 ? bd.us_new : us_c + w * us_slot;
    if (run) rollout_candidate<NJ>(m, bd, b, ldexp(1.0, -c), mu, costs_old, xo, uo, &sum, &merit, &plain, sc, w == 0);
    s_sum[w][lane] = run ? sum : 1.0;
    s_merit[w][lane] = merit; s_plain[w][lane] = plain;
    __syncthreads();
    if (w == 0 && valid && !frozen && s_win[lane] == -1) {
      int found = -1;
      for (int k = 0; k < LS_NC; ++k)
        if (found < 0 && r * LS_NC + k < LS_MAX_TRIALS && s_sum[k][lane] <= 0.0) found = k;
      const bool last_round = (r + 1) * LS_NC >= LS_MAX_TRIALS;
      if (found < 0 && last_round) found = (LS_MAX_TRIALS - 1) - r * LS_NC;  // give up: keep the last trial
      if (found >= 0) {
        const int cstar = r * LS_NC + found;
        const bool accepted = s_sum[found][lane] <= 0.0;
        bd.scal[(size_t)S_STEP * B + b] = accepted ? ldexp(1.0, -cstar) : ldexp(1.0, -LS_MAX_TRIALS);
        bd.scal[(size_t)S_ROLLOUTS * B + b] += (double)(accepted ? cstar + 1 : LS_MAX_TRIALS);
        bd.scal[(size_t)S_MERIT * B + b] = s_merit[found][lane];
        bd.scal[(size_t)S_COST * B + b] = s_plain[found][lane];
        s_win[lane] = 1000 + found;  // resolved in this round: slot `found`
      }
    }
    __syncthreads();
    // problems resolved in this round: copy the winning scratch trajectory now (a later round would overwrite the slot)
    if (valid && s_win[lane] > 1000) {
      const int slot = s_win[lane] - 1000;
      const double* xw = xs_c + slot * xs_slot;
      const double* uw = us_c + slot * us_slot;
#pragma unroll 8
      for (int e = nx + w; e < (T + 1) * nx; e += LS_NC) bd.xs_new[(size_t)e * B + b] = xw[(size_t)e * B + b];
#pragma unroll 8
      for (int e = w; e < T * nu; e += LS_NC) bd.us_new[(size_t)e * B + b] = uw[(size_t)e * B + b];
    }
    __syncthreads();
    if (w == 0 && valid && s_win[lane] >= 1000) s_win[lane] = 0;  // done
    win = 0;
    for (int k = 0; k < 32; ++k) win |= (s_win[k] == -1);
    __syncthreads();
    if (!win) break;
  }
  if (frozen) {  // keep a frozen problem's trajectory through the pointer swap
    for (int e = w; e < (T + 1) * nx; e += LS_NC) bd.xs_new[(size_t)e * B + b] = bd.xs[(size_t)e * B + b];
    for (int e = w; e < T * nu; e += LS_NC) bd.us_new[(size_t)e * B + b] = bd.us[(size_t)e * B + b];
  }
}

}  // namespace ddpb
#endif
