// outer_warp.cuh -- device-only, latency-oriented versions of the per-problem sweeps of the outer loop.
// The thread-per-problem bodies in forward.cuh stay as the CPU-emulated reference of the same arithmetic
// (tests/hostemu); these kernels split them by their parallelism:
//   origin_update_knot   : mults/feedback re-centering (mat_seq_common.hpp:62-89) -- independent per (knot, problem)
//   optimality_warp      : adjoint recursion of optimality_obj (ddp.hpp:575-627) -- sequential in t, one WARP per
//                          problem: lane i < ndx owns adjoint component i, lanes ndx..ndx+nu-1 own l_u rows, the
//                          f_x / f_u columns are read straight from the slab record (contiguous 8*ndx-byte runs) one
//                          knot ahead, the new adjoint is exchanged with shuffles
//   mult_update_knot     : multiplier update (ddp.hpp:686-687) -- independent per (knot, problem)
//   ls_trials / ls_select: the backtracking line search (ddp_fwd.ipp:10-67) with all its trials at once
#pragma once
#if defined(__CUDACC__)
#include "forward.cuh"

DDPB_VARIANT_BEGIN

// fb_all = false: the feedback sequence is re-centred only at knots with equality rows.  Its re-centred k is read once,
// by the multiplier update at such knots (ddp.hpp:686-687); at every other knot k, K and the origin are overwritten by
// the backward pass of the same iteration before anything reads them, so the reference's update there (8 n^2 + 6 n doubles
// of traffic per knot) is dead -- unless a problem can stop on the optimality threshold in this iteration and hand its
// feedback back as the result of solve (ddp.hpp:744-748), which needs threshold > 0: the host passes fb_all = true then.
template <int NJ> __device__ __forceinline__ void origin_update_knot(const BatchDev& bd, int b, int t, bool fb_all) {
  constexpr int n = Dim<NJ>::n, nx = 2 * n, nu = n, nd = 2 * n;
  const int B = bd.B;
  if (bd.scal[(size_t)S_STATUS * B + b] != 0.0) return;
  const int ne = bd.ne[t];
  if (ne == 0 && !fb_all) {
#pragma unroll
    for (int k = 0; k < nd; ++k) bd.m_origin[((size_t)t * nx + k) * B + b] = bd.xs[((size_t)t * nx + k) * B + b];
    return;
  }
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

// lanes per problem of the adjoint sweeps: the smallest power of two holding one lane per adjoint / control component
// (8 for the 2R chain: four problems per warp, a quarter of the warp instructions of one warp per problem)
template <int NJ> struct AdjGeom {
  static constexpr int need = 3 * Dim<NJ>::n;
  static constexpr int G = need <= 4 ? 4 : (need <= 8 ? 8 : (need <= 16 ? 16 : 32));
};
// one lane group: opt_obj of problem b (all lanes of the group return the same value); shuffles are group-relative
template <int NJ> __device__ __forceinline__ double optimality_obj_warp(const BatchDev& bd, int b, double mu) {
  constexpr int n = Dim<NJ>::n, nx = 2 * n, nu = n, nd = 2 * n, G = AdjGeom<NJ>::G;
  static_assert(nd + nu <= 32, "one lane per adjoint / control component");
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & (G - 1);
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
  // offsets of the head blocks (batch.cuh rec_layout: independent of the knot's equality rows)
  constexpr int o_fx = nx, o_fu = o_fx + n * nd, o_lu = o_fu + n * nu;
  auto load = [&](int t, double* dst, double& l) {
    if (isx || isu) {
      const double* rec = slab_rec_t(bd, t, b);
      const int c0 = (isx ? o_fx : o_fu) + col * n;
#pragma unroll
      for (int k = 0; k < n; ++k) dst[k] = rec[c0 + k];
      l = isx ? 0.0 : rec[o_lu + col];
    }
  };
  // the sweep is a latency chain of T steps fed by ~150 bytes of the slab each: the column it needs is in registers one
  // knot ahead, and the record PF knots ahead is pulled into L2 so that this register load never waits for HBM
  constexpr int PF = 6;
  auto prefetch = [&](int t) {
    if (t < 0 || !(isx || isu)) return;
    l2_prefetch(slab_rec_t(bd, t, b) + (isx ? o_fx : o_fu) + col * n);
  };
  for (int k = 2; k <= PF; ++k) prefetch(bd.T - k);
  load(bd.T - 1, cur, lcur);
  for (int t = bd.T - 1; t >= 0; --t) {
    prefetch(t - 1 - PF);
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
    for (int i = 0; i < nu; ++i) { const double li = __shfl_sync(FULL, v, nd + i, G); sq += li * li; }
    const double nrm = sqrt(sq);
    if (retval < nrm) retval = nrm;
#pragma unroll
    for (int k = 0; k < nd; ++k) adj[k] = __shfl_sync(FULL, v, k, G);
    atop = __shfl_sync(FULL, v, isx ? (col < n ? col : col - n) : 0, G);
#pragma unroll
    for (int k = 0; k < n; ++k) cur[k] = nxt[k];
    lcur = lnxt;
  }
  return retval;
}

// optimality measures + the decision of update_derivatives / solve's switch (ddp.hpp:660-695, 786-802) for one problem
// per warp.  phase 0: first sweep and decision (S_FLAG: 1 = multipliers will be updated); phase 1: after the
// multiplier update, second sweep for the flagged problems (n = opt_obj / mu^0.1, w /= mu).
// (b, live): the group's problem; groups past the end of the sub-batch still take part in the warp's shuffles
template <int NJ> __device__ __forceinline__ void al_decide_warp(const BatchDev& bd, int b, bool live, double thr, int phase) {
  const int B = bd.B;
  const int lane = threadIdx.x & (AdjGeom<NJ>::G - 1);
  const bool act = live && bd.scal[(size_t)S_STATUS * B + b] == 0.0;
  const bool flagged = act && (phase == 0 || bd.scal[(size_t)S_FLAG * B + b] == 1.0);
  // a whole warp without work leaves together; otherwise every lane runs the sweep (idle groups on problem b, harmlessly)
  if (!__any_sync(0xffffffffu, flagged)) return;
  double mu = bd.scal[(size_t)S_MU * B + b];
  if (phase == 0) {
    const double opt_obj = optimality_obj_warp<NJ>(bd, b, mu);
    if (lane != 0 || !act) return;
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
    const double oo = optimality_obj_warp<NJ>(bd, b, mu);
    if (lane != 0 || !flagged) return;
    bd.scal[(size_t)S_N * B + b] = oo / pow(mu, 0.1);
    bd.scal[(size_t)S_W * B + b] = bd.scal[(size_t)S_W * B + b] / mu;
    bd.scal[(size_t)S_MULT_UPDATES * B + b] += 1.0;
  }
}

// ---- line search with concurrent trials ---------------------------------------------------------------------
// The reference tries step = 1, 1/2, 1/4, ... one rollout after the other and takes the first step whose merit does not
// increase (ddp_fwd.ipp:29-60).  A rollout is a latency chain of T dynamics evaluations, so a problem that needs the c-th
// halving used to cost its whole launch c rollouts -- and some problem of a large batch always does.  The trials are
// independent, so they run AT ONCE, spread over the grid: trial c of problem b is one thread (warp = trial c of 32
// consecutive problems => coalesced SoA accesses), every trial rolls out into its own scratch trajectory (trial 0 straight
// into the new-trajectory storage: the usual case, the full step, copies nothing), and a selection kernel then takes the
// smallest accepted trial index exactly as the sequential scan would (34 trials at most, the last one kept on give-up,
// ddp_fwd.ipp:35-37) and copies the winner.  Two waves: trials [0, NC1) for everybody, trials [NC1, 34) for the problems
// still unresolved (almost always none: those threads exit at once).  A block of 32 problems that needed the second wave in
// the previous pass runs ALL 34 trials in its first wave of this pass (`wide` flags, ping-pong between passes): workloads
// whose line searches regularly go deep (C4) then pay one rollout latency per pass instead of two; which wave computes a
// trial does not change its result.  The `rollouts` scalar keeps counting what the sequential algorithm would have run.
//   NC1: 8 for models of <= 2 joints (a 4096-problem batch is then ~2 warps per SM sub-partition: the FP64 pipe is the
//   limit, more speculative trials would cost time), 16 for the arms (their batches leave most of the machine idle).
template <int NJ> struct LsGeom {
  static constexpr int NC1 = Dim<NJ>::n <= 2 ? 8 : 16;
#ifndef DDPB_LS_WPB_ARMS
#define DDPB_LS_WPB_ARMS 4
#endif
  static constexpr int WPB = Dim<NJ>::n <= 3 ? 4 : DDPB_LS_WPB_ARMS;  // trial warps per block
};
constexpr int LS_MAX_TRIALS = 34;  // 2^-33 >= 1e-10 > 2^-34
// scratch slots: trial c uses slot c (slot 0 is the new-trajectory storage itself)
template <int NJ> constexpr int ls_slots() { return LS_MAX_TRIALS; }

// per-problem line-search state in `win` [B] ints: -1 unresolved, >= 0 resolved.  res: [LS_MAX_TRIALS][3][B] (sum, merit, plain)
template <int NJ>
__device__ __forceinline__ void ls_trials_block(const ModelDev& m, const BatchDev& bd, const double* costs_old, double* xs_c, double* us_c, double* res,
                                                const int* win, const int* wide, int c0, int c1, const Scratch& sc) {
  constexpr int n = Dim<NJ>::n, nx = 2 * n, nu = n;
  const int B = bd.B, T = bd.T;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int b = bd.b0 + blockIdx.x * 32 + lane;
  const int c = c0 + blockIdx.y * LsGeom<NJ>::WPB + w;
  if (b >= bd.b0 + bd.nb || c >= c1) return;
  // first wave: the trials past NC1 only run for blocks of problems that needed them in the previous pass
  if (c0 == 0 && c >= LsGeom<NJ>::NC1 && !wide[(bd.b0 >> 5) + blockIdx.x]) return;
  if (bd.scal[(size_t)S_STATUS * B + b] != 0.0 || (c0 > 0 && win[b] != -1)) return;  // frozen, or resolved by the first wave
  const size_t xs_slot = (size_t)(T + 1) * nx * B, us_slot = (size_t)T * nu * B;
  const int slot = c;
  double* xo = c == 0 ? bd.xs_new : xs_c + slot * xs_slot;
  double* uo = c == 0 ? bd.us_new : us_c + slot * us_slot;
  double sum, merit, plain;
  rollout_candidate<NJ>(m, bd, b, ldexp(1.0, -c), bd.scal[(size_t)S_MU * B + b], costs_old, xo, uo, &sum, &merit, &plain, sc, c == 0);
  res[((size_t)c * 3 + 0) * B + b] = sum;
  res[((size_t)c * 3 + 1) * B + b] = merit;
  res[((size_t)c * 3 + 2) * B + b] = plain;
}

// one warp per problem: the smallest accepted trial of [c0, c1) wins; last wave: give up on the last trial.  The winner's
// scratch trajectory is copied into the new-trajectory storage by the warp (element index strided by the batch size).
template <int NJ>
__device__ __forceinline__ void ls_select_warp(const BatchDev& bd, const double* xs_c, const double* us_c, const double* res, int* win, const int* wide,
                                               int* wide_next, int c0, int c1, bool last) {
  constexpr int n = Dim<NJ>::n, nx = 2 * n, nu = n;
  const int B = bd.B, T = bd.T;
  const int lane = threadIdx.x & 31;
  const int i = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), b = bd.b0 + i;
  if (i >= bd.nb) return;
  const size_t xs_slot = (size_t)(T + 1) * nx * B, us_slot = (size_t)T * nu * B;
  if (bd.scal[(size_t)S_STATUS * B + b] != 0.0) {
    if (c0 == 0) {  // keep a frozen problem's trajectory through the pointer swap
      for (int e = lane; e < (T + 1) * nx; e += 32) bd.xs_new[(size_t)e * B + b] = bd.xs[(size_t)e * B + b];
      for (int e = lane; e < T * nu; e += 32) bd.us_new[(size_t)e * B + b] = bd.us[(size_t)e * B + b];
      if (lane == 0) win[b] = 0;
    }
    return;
  }
  if (c0 > 0 && win[b] != -1) return;
  const int blk = b >> 5;
  // first wave: trials [0, NC1), or all of them where the block ran wide
  const int c_hi = (c0 == 0 && wide[blk]) ? LS_MAX_TRIALS : c1;
  int found = -1;
  for (int c = c0; c < c_hi; ++c)
    if (found < 0 && res[((size_t)c * 3 + 0) * B + b] <= 0.0) found = c;
  if (c0 == 0 && c_hi == LS_MAX_TRIALS) last = true;
  const bool accepted = found >= 0;
  if (!accepted && last) found = LS_MAX_TRIALS - 1;  // give up: keep the last trial
  if (found < 0) {
    if (lane == 0) win[b] = -1;
    return;
  }
  if (lane == 0 && found >= LsGeom<NJ>::NC1) wide_next[blk] = 1;  // this block goes deep: all trials at once next time
  if (found != 0) {
    const double* xw = xs_c + (size_t)found * xs_slot;
    const double* uw = us_c + (size_t)found * us_slot;
    for (int e = nx + lane; e < (T + 1) * nx; e += 32) bd.xs_new[(size_t)e * B + b] = xw[(size_t)e * B + b];
    for (int e = lane; e < T * nu; e += 32) bd.us_new[(size_t)e * B + b] = uw[(size_t)e * B + b];
  }
  if (lane == 0) {
    bd.scal[(size_t)S_STEP * B + b] = accepted ? ldexp(1.0, -found) : ldexp(1.0, -LS_MAX_TRIALS);
    bd.scal[(size_t)S_ROLLOUTS * B + b] += (double)(accepted ? found + 1 : LS_MAX_TRIALS);
    bd.scal[(size_t)S_MERIT * B + b] = res[((size_t)found * 3 + 1) * B + b];
    bd.scal[(size_t)S_COST * B + b] = res[((size_t)found * 3 + 2) * B + b];
    win[b] = found;
  }
}

DDPB_VARIANT_END
#endif
