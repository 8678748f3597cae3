// backward.cuh -- Riccati-like backward recursion, one group of BT threads (1..4 warps) per problem, sequential over knots.
//
// Restates ddp_solver_t::backward_pass (/root/reference/include/ddp/ddp_bwd.ipp:9-155):
//   tmp = p + mu eq ; tmp2 = P + mu eq_x                                   (:46-47)
//   Q_x, Q_u, Q_xx, Q_uu, Q_ux incl. the six tensor contractions           (:61-87)
//   LL^T = Q_uu + reg I ; failure -> reg = max(reg, mu), mu *= 2, reg *= 2, restart from t = T-1 (:102-132)
//   k = -(LL^T)^-1 Q_u ; K = -(LL^T)^-1 Q_ux ; origin = x_t                (:134-136)
//   V_x = Q_x + Q_ux^T k ; V_xx = Q_xx + Q_ux^T K                          (:142-146)
//
// The recursion is a latency chain (T dependent knots) fed by 8 * rec_size bytes per knot, so the design keeps the
// dependent path of one knot short and the memory system busy ahead of it:
//   * while knot t is being processed, the head of the record of knot t-1 goes into the idle head buffer and round 0
//     of every warp's tensor share into its staging buffer by TMA (cp.async.bulk global -> shared completing on an
//     mbarrier; one thread issues 1 + #warps bulk copies, no registers held), so a knot starts with its first operands
//     already in shared memory; later rounds are loaded one round ahead.  (A bulk L2 prefetch of
//     the record after next was measured to re-fetch ~50 % of the bytes from DRAM without shortening the pass and
//     was removed: DRAM reads now equal the algorithmic bytes.)
//   * the third-order tensors are used once: every warp streams its share with fully coalesced 16-byte loads through
//     a small warp-private staging buffer and contracts the k-run of one output element per lane with V_x;
//   * every matrix product is a list of inner products over CONTIGUOUS shared-memory runs (V_xx is kept row-major,
//     everything else column-major as in the record), read with 16-byte loads;
//   * the Cholesky factorisation runs in one warp with the rows of Q_uu in registers and shuffles instead of barriers
//     (Eigen::LLT semantics: left-looking, fails on a pivot <= 0); k and K are one right-hand side per lane.
// Every sum keeps the reference's term order, so results do not depend on the thread count.
#pragma once
#include "batch.cuh"

namespace ddpb {

template <int NJ> struct BwdSmem {
  static constexpr int n = NJ, nd = 2 * NJ, nu = NJ;
  static constexpr int NTEN = nd * nd + nu * nd + nu * nu;  // Q_xx | Q_ux | Q_uu task list
  // stored head of the slab record without f, copied verbatim: fx_B[n x nd] | fu_B[n x nu] | lu[nu]  (batch.cuh rec_layout)
  static constexpr int h_fx = 0, h_fu = h_fx + n * nd, h_lu = h_fu + n * nu, NHEAD = h_lu + nu;  // NHEAD is even
  static constexpr int even(int x) { return (x + 1) & ~1; }
  static constexpr int RS = NJ <= 6 ? 32 : 16;  // dense tensors: elements per staging round and warp
  static constexpr int head = 0, head2 = even(head + NHEAD), VxxR = even(head2 + NHEAD), Vx = VxxR + nd * nd, VF = even(Vx + nd), VU = VF + nd * nd,
                       Qx = even(VU + nd * nu), Qu = Qx + nd, Qxx = even(Qu + nu), Qux = Qxx + nd * nd, Quu = Qux + nu * nd,
                       Wf = even(Quu + nu * nu),  // V_x-contraction of fxx | fux | fuu, task-list order
                       L = even(Wf + NTEN), dinv = L + nu * nu, kk = even(dinv + nu), KK = even(kk + nu),
                       eqv = even(KK + nu * nd), eqx = eqv + MAXE, equ = eqx + MAXE * nd, pe = equ + MAXE * nu,
                       pex = pe + MAXE, tmp = pex + MAXE * nd, tmp2 = tmp + MAXE, flags = tmp2 + MAXE * nd,
                       cst = even(flags + 2),  // constants: 0.0, c, 1.0, dt (operands of the structural entries)
                       mbar = cst + 4,         // the mbarrier of the record staging (8 bytes, 16-byte slot)
                       stage = even(mbar + 2),  // device: NW warp-private staging buffers of RS * nd doubles
                       SIZE = stage;
  static constexpr int c_zero = cst, c_c = cst + 1, c_one = cst + 2, c_dt = cst + 3;
  // tensor geometry of a mode for NW warps: run length RL, element count NEL, elements per warp EW (even: every warp's
  // share starts 16-byte aligned), elements per staging round RSM.  The compact tensors of the FD-of-Jacobians mode are
  // staged in ONE round (the whole share of a warp is fetched by TMA during the previous knot).
  template <bool DENSE, int NW> struct Ten {
    static constexpr int RL = DENSE ? nd : n, NEL = DENSE ? NTEN : nd * nd + nu * n;
    static constexpr int EW = (((NEL + NW - 1) / NW) + 1) & ~1;
    static constexpr int RSM = DENSE ? RS : EW;
    static constexpr int STG = RSM * RL;  // staging doubles per warp
  };
  template <int NW> static constexpr int stage_doubles(bool dense) { return dense ? Ten<true, NW>::STG : Ten<false, NW>::STG; }
  static constexpr int size_for(int threads, bool dense) {
    return SIZE + (threads >= 128 ? 4 * stage_doubles<4>(dense) : threads >= 64 ? 2 * stage_doubles<2>(dense) : threads >= 32 ? stage_doubles<1>(dense) : 0);
  }
  // 16-byte inner products need even run lengths and even offsets: the stored columns have n rows
  static constexpr bool head_aligned = (n % 2) == 0;
};

// threads per problem on the device
#ifndef DDPB_BWD_BT_LARGE
#define DDPB_BWD_BT_LARGE 128
#endif
template <int NJ> struct BwdGeom { static constexpr int BT = NJ <= 3 ? 32 : (NJ <= 5 ? 64 : DDPB_BWD_BT_LARGE); };

// inner product over N contiguous doubles with four interleaved partial sums (the dependent chain is N/4 FMAs instead
// of N); 16-byte loads on the device when both runs are 16-byte aligned (ALIGNED) and N is even.  Host and device use
// the same association, so the CPU emulation of this kernel rounds identically.
template <int N, bool ALIGNED> DDPB_HD double dot_run(const double* a, const double* b) {
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
#if defined(__CUDA_ARCH__)
  if (ALIGNED && (N % 2) == 0) {
    const double2* a2 = reinterpret_cast<const double2*>(a);
    const double2* b2 = reinterpret_cast<const double2*>(b);
#pragma unroll
    for (int k = 0; k < N / 2; ++k) {
      const double2 x = a2[k], y = b2[k];
      acc[(2 * k) & 3] += x.x * y.x;
      acc[(2 * k + 1) & 3] += x.y * y.y;
    }
  } else
#endif
  {
#pragma unroll
    for (int k = 0; k < N; ++k) acc[k & 3] += a[k] * b[k];
  }
  return (acc[0] + acc[1]) + (acc[2] + acc[3]);
}

// ---- TMA (bulk asynchronous copy) staging of a slab record: cp.async.bulk global -> shared, completion counted in bytes
// on an mbarrier.  One thread issues the copies of a whole knot (record head + every warp's first staging round); all
// threads wait on the barrier's phase.  The wait is bounded: a byte-count bug traps instead of hanging the device.
#ifndef DDPB_BWD_TMA
#define DDPB_BWD_TMA 1
#endif
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void mbar_init(void* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, void* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  unsigned done = 0;
  for (unsigned spin = 0; !done; ++spin) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(a), "r"(parity) : "memory");
    if (spin > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

// 16-byte asynchronous copy global -> shared (LDGSTS) and its completion fence
DDPB_HD void cp_async16(void* smem_dst, const void* gmem_src) {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
#else
  (void)smem_dst; (void)gmem_src;
#endif
}
DDPB_HD void cp_async_wait_all() {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
#endif
}

// Cholesky of the lower triangle of A (nu x nu, column-major) in place + 1/diag; returns false on a non-positive pivot
// (Eigen::LLT semantics).  CPU emulation of phase D of backward_problem; the device version works on registers there.
template <int NU> inline bool chol_host(double* A, double* dinv) {
  for (int k = 0; k < NU; ++k) {
    double x = A[k * NU + k];
    for (int j = 0; j < k; ++j) x -= A[j * NU + k] * A[j * NU + k];
    if (x <= 0.0) return false;
    const double linv = 1.0 / sqrt(x);
    A[k * NU + k] = x * linv;
    dinv[k] = linv;
    for (int i = k + 1; i < NU; ++i) {
      double s = A[k * NU + i];
      for (int j = 0; j < k; ++j) s -= A[j * NU + i] * A[j * NU + k];
      A[k * NU + i] = s * linv;
    }
  }
  return true;
}

// origin = x_t of the feedback sequence (ddp_bwd.ipp:136) for the problems a backward pass is about to process: one
// element per thread before the pass (a plain copy: keeping it out of the recursion removes a global load -> store from
// every knot's dependent path).  i in [0, T * nx * nb).
DDPB_HD void fb_origin_knot(const BatchDev& bd, size_t i) {
  const size_t e = i / bd.nb;
  const int b = bd.b0 + (int)(i - e * bd.nb);
  if (bd.scal[(size_t)S_STATUS * bd.B + b] != 0.0) return;
  bd.fb_origin[e * bd.B + b] = bd.xs[e * bd.B + b];
}

template <int V> struct IntC { static constexpr int value = V; };

// sum_k vx[k] * f3[k] over a run of RL doubles with two interleaved partial sums (even / odd k); 16-byte loads on the
// device when the run is even-sized and 16-byte aligned (ALIGNED).  Host and device use the same association.
template <int RL, bool ALIGNED> DDPB_HD double contract_run(const double* vx, const double* f3) {
  double a0 = 0.0, a1 = 0.0;
#if defined(__CUDA_ARCH__)
  if (ALIGNED && (RL % 2) == 0) {
    const double2* f2 = reinterpret_cast<const double2*>(f3);
    const double2* x2 = reinterpret_cast<const double2*>(vx);
#pragma unroll
    for (int k = 0; k < RL / 2; ++k) {
      const double2 v = f2[k], x = x2[k];
      a0 += x.x * v.x;
      a1 += x.y * v.y;
    }
    return a0 + a1;
  }
#endif
#pragma unroll
  for (int k = 0; k < RL; k += 2) {
    a0 += vx[k] * f3[k];
    if (k + 1 < RL) a1 += vx[k + 1] * f3[k + 1];
  }
  return a0 + a1;
}

// BT = compile-time thread count of the executor (BwdGeom<NJ>::BT on the device, 1 in the CPU emulation).
// Record-major slabs only (chains of >= 3 joints on the device; backward_tpp.cuh covers the element-major ones).
template <int NJ, int BT, bool DENSE, class Ex>
DDPB_HD void backward_problem(const BatchDev& bd, int b, double* sm, const Ex& ex) {
  constexpr int n = NJ, nd = 2 * NJ, nu = NJ, nx = 2 * NJ;
  using S = BwdSmem<NJ>;
  constexpr int nQxx = nd * nd, nQux = nu * nd, NTEN = S::NTEN;
  constexpr bool AL = S::head_aligned;  // stored columns (n rows) start on 16-byte boundaries and have even length
  const int B = bd.B, T = bd.T;
  const int tid = ex.tid();
  // stored tensors (batch.cuh): FD-of-Jacobians mode keeps rows k >= n of f_xx | f_ux(:, :, i < n) only -- n doubles per
  // (j, i) pair, contracted with V_x[n..] -- value-based mode keeps f_xx | f_ux | f_uu dense
  // (DENSE: a compile-time copy of bd.hess_mode == HESS_FD_VALUES -- one kernel per mode, neither pays for the other's registers)
  constexpr bool dense = DENSE;
  constexpr int NEL = dense ? NTEN : nd * nd + nu * n;  // tensor elements (k-runs) of a record, contiguous from L.fxx
  // a one-warp group needs no CTA barrier: warp-level synchronisation orders its shared-memory traffic
  auto bsync = [&]() { if (BT == 32) ex.sync_warp(); else ex.sync(); };
  double mu = bd.scal[(size_t)S_MU * B + b];
  double reg = bd.scal[(size_t)S_REG * B + b];
  int restarts = 0;
  bool success = false;
#define FOR_ELEMS(e, N) _Pragma("unroll") for (int u_ = 0; u_ < ((N) + BT - 1) / BT; ++u_) if (const int e = tid + u_ * BT; e < (N))
  // Per-thread task descriptors (shared-memory offsets of the operands of every inner product this thread owns),
  // computed once: the knot loop then runs one branch-free code path per phase instead of re-deriving row / column /
  // block from the task index at every knot.  Every product with f_x / f_u is "constant rows + stored rows":
  //   sum_k f(k, r) y(k) = y(r) [r < n]  or  dt y(r - n) [r >= n]  (f_x; nothing for f_u)  +  sum_{k >= n} f_B(k - n, r) y(k)
  // so a descriptor names the stored column (a), the run of y it meets (b), the top element of y and its scale (1 or dt).
  constexpr int NB = nd * nd + nd * nu;    // phase B: VF | VU
  constexpr int NC = NTEN + nd + nu;       // phase C: Q_xx | Q_ux | Q_uu | Q_x | Q_u
  constexpr int NF = nd * nd + nd;         // phase F: V_xx | V_x
  // PRE: descriptors live in registers for the whole kernel (small models, where the register budget of a resident
  // wave allows it); otherwise they are re-derived at every use.
  constexpr bool PRE = NJ <= 5 || BT == 1;
  auto mkB = [&](int e, unsigned& d0, unsigned& d1) {  // a | b << 16 ; top | scale << 16  (b relative to the current head buffer)
    const int r = e % nd, c = e / nd;  // c runs over the columns of f_x, then of f_u
    int bo, top = S::c_zero, ts = S::c_one;
    if (c < nd) {
      bo = S::h_fx + c * n;
      if (c < n) top = S::VxxR + r * nd + c; else { top = S::VxxR + r * nd + (c - n); ts = S::c_dt; }
    } else {
      bo = S::h_fu + (c - nd) * n;
    }
    d0 = (unsigned)(S::VxxR + r * nd + n) | ((unsigned)bo << 16);
    d1 = (unsigned)top | ((unsigned)ts << 16);
  };
  // a|b , l|w , dst|ldst (bit 15 of ldst: diagonal), top|scale.  a relative to the current head buffer; bit 15 of l: l is
  // relative to the head buffer too (l_u), otherwise an absolute offset (constants)
  auto mkC = [&](int e, unsigned& d0, unsigned& d1, unsigned& d2, unsigned& d3) {
    int a, bo, l = S::c_zero, w = S::c_zero, dst, ldst = 0x7fff, diag = 0, top = S::c_zero, ts = S::c_one;
    if (e < nQxx) {
      const int r = e % nd, c = e / nd;
      a = S::h_fx + r * n; bo = S::VF + c * nd + n; w = S::Wf + e; dst = S::Qxx + e;
      if (r < n) top = S::VF + c * nd + r; else { top = S::VF + c * nd + (r - n); ts = S::c_dt; }
    } else if (e < nQxx + nQux) {
      const int ee = e - nQxx, c = ee / nu, r = ee % nu;
      a = S::h_fu + r * n; bo = S::VF + c * nd + n; w = S::Wf + e; dst = S::Qux + ee;
    } else if (e < NTEN) {
      const int ee = e - nQxx - nQux, r = ee % nu, c = ee / nu;
      a = S::h_fu + r * n; bo = S::VU + c * nd + n; w = S::Wf + e; dst = S::Quu + ee;
      l = r == c ? S::c_c : S::c_zero;  // l_uu = c I
      ldst = S::L + ee; diag = r == c;
    } else if (e < NTEN + nd) {
      const int i = e - NTEN;
      a = S::h_fx + i * n; bo = S::Vx + n; dst = S::Qx + i;
      if (i < n) top = S::Vx + i; else { top = S::Vx + (i - n); ts = S::c_dt; }
    } else {
      const int i = e - NTEN - nd;
      a = S::h_fu + i * n; bo = S::Vx + n; l = (S::h_lu + i) | 0x8000; dst = S::Qu + i;
    }
    d0 = (unsigned)a | ((unsigned)bo << 16);
    d1 = (unsigned)l | ((unsigned)w << 16);
    d2 = (unsigned)dst | ((unsigned)(ldst | (diag << 15)) << 16);
    d3 = (unsigned)top | ((unsigned)ts << 16);
  };
  auto mkF = [&](int e, unsigned& d0, unsigned& d1) {  // a|b , q|dst
    if (e < nd * nd) {
      const int c = e % nd, r = e / nd;  // row-major destination: consecutive threads -> consecutive c
      d0 = (unsigned)(S::Qux + r * nu) | ((unsigned)(S::KK + c * nu) << 16);
      d1 = (unsigned)(S::Qxx + r + c * nd) | ((unsigned)(S::VxxR + e) << 16);
    } else {
      const int i = e - nd * nd;
      d0 = (unsigned)(S::Qux + i * nu) | ((unsigned)S::kk << 16);
      d1 = (unsigned)(S::Qx + i) | ((unsigned)(S::VxxR + e) << 16);
    }
  };
  constexpr int PB = PRE ? (NB + BT - 1) / BT : 1, PC = PRE ? (NC + BT - 1) / BT : 1, PF = PRE ? (NF + BT - 1) / BT : 1;
  unsigned dB0[PB], dB1[PB], dC0[PC], dC1[PC], dC2[PC], dC3[PC], dF0[PF], dF1[PF];
  if (PRE) {
    FOR_ELEMS(e, NB) mkB(e, dB0[PRE ? u_ : 0], dB1[PRE ? u_ : 0]);
    FOR_ELEMS(e, NC) mkC(e, dC0[PRE ? u_ : 0], dC1[PRE ? u_ : 0], dC2[PRE ? u_ : 0], dC3[PRE ? u_ : 0]);
    FOR_ELEMS(e, NF) mkF(e, dF0[PRE ? u_ : 0], dF1[PRE ? u_ : 0]);
  }
  if (tid == 0) { sm[S::c_zero] = 0.0; sm[S::c_c] = bd.c; sm[S::c_one] = 1.0; sm[S::c_dt] = bd.dt; }
  // tensor terms that do not exist in FD-of-Jacobians mode (f_ux slices i >= n, f_uu) stay exact zeros
  FOR_ELEMS(e, NTEN) sm[S::Wf + e] = 0.0;
  unsigned tma_phase = 0;     // parity of the staging mbarrier's current phase
  bool tma_pending = false;   // a prefetch has been issued and not yet waited for
#if defined(__CUDA_ARCH__) && DDPB_BWD_TMA
  if (tid == 0) { mbar_init(sm + S::mbar, 1); fence_proxy_async(); }
#endif
  (void)tma_phase;
  while (!success) {
    // V_xx = lfxx = 0, V_x = lfx = 0  (problem.hpp:958-959)
    FOR_ELEMS(e, nd * nd + nd) sm[S::VxxR + e] = 0.0;
    bsync();
    bool failed = false;
    // software pipeline: what knot t needs first -- the head of its record and round 0 of every warp's tensor share
    // (warp-private staging buffer) -- is fetched while knot t+1 is being processed
    int hcur = 0;  // which head buffer holds the knot being processed
#if defined(__CUDA_ARCH__)
    constexpr int NW = BT >= 32 ? BT / 32 : 1;
    using TD = typename S::template Ten<true, NW>;
    using TC = typename S::template Ten<false, NW>;
    constexpr int EW = dense ? TD::EW : TC::EW, RS = dense ? TD::RSM : TC::RSM, STG = dense ? TD::STG : TC::STG;
    const int warp = tid >> 5, lane = tid & 31;
    double* stg = sm + S::stage + warp * STG;
    const int e_beg = warp * EW < NEL ? warp * EW : NEL, e_end = (warp + 1) * EW < NEL ? (warp + 1) * EW : NEL;
#endif
    // fetch what knot tp needs first into the buffers that are idle while the current knot is processed: the head of
    // its record -> the other head buffer, round 0 of every warp's tensor share -> the warp's staging buffer.
    // TMA path: thread 0 issues one bulk copy per buffer, all completing on the CTA's mbarrier; it is called after a
    // barrier that follows the last read of the target buffers (the staging buffers were also written with ordinary
    // stores, hence the proxy fence).  Copies are whole 16-byte units: a round with an odd number of doubles takes the
    // next double of the (even-sized) record along.
    auto prefetch = [&](int tp, int hnext) {
      const RecLayout Lp = rec_layout(n, bd.ne[tp], bd.hess_mode);
      const double* recp = slab_rec(bd, tp, b, Lp.size);
      double* hdst = sm + (hnext ? S::head2 : S::head);
#if defined(__CUDA_ARCH__)
      constexpr int RL = dense ? nd : n;
#if DDPB_BWD_TMA
      if (tid == 0) {
        unsigned bytes = S::NHEAD * 8;
#pragma unroll
        for (int wq = 0; wq < NW; ++wq) {
          const int eb = wq * EW, ee = (wq + 1) * EW < NEL ? (wq + 1) * EW : NEL;
          const int nr = ee - eb < RS ? ee - eb : RS;
          if (nr > 0) bytes += (unsigned)(((nr * RL + 1) & ~1) * 8);
        }
        fence_proxy_async();
        mbar_expect_tx(sm + S::mbar, bytes);
        tma_bulk_g2s(hdst, recp + Lp.fx, S::NHEAD * 8, sm + S::mbar);
#pragma unroll
        for (int wq = 0; wq < NW; ++wq) {
          const int eb = wq * EW, ee = (wq + 1) * EW < NEL ? (wq + 1) * EW : NEL;
          const int nr = ee - eb < RS ? ee - eb : RS;
          if (nr > 0) tma_bulk_g2s(sm + S::stage + wq * STG, recp + Lp.fxx + (size_t)eb * RL, (unsigned)(((nr * RL + 1) & ~1) * 8), sm + S::mbar);
        }
      }
#else
      FOR_ELEMS(c, S::NHEAD / 2) cp_async16(hdst + 2 * c, recp + Lp.fx + 2 * c);
      const int ne_round = e_end - e_beg < RS ? e_end - e_beg : RS;
      const int nch = (ne_round * RL + 1) / 2;
      const double2* src = reinterpret_cast<const double2*>(recp + Lp.fxx + (size_t)e_beg * RL);
      for (int c = lane; c < nch; c += 32) cp_async16(reinterpret_cast<double2*>(stg) + c, src + c);
#endif
#else
      for (int e = 0; e < S::NHEAD; ++e) hdst[e] = recp[Lp.fx + e];
#endif
    };
    // wait for the staged data of the knot about to be processed
    auto staged_wait = [&]() {
#if defined(__CUDA_ARCH__)
#if DDPB_BWD_TMA
      mbar_wait(sm + S::mbar, tma_phase);
      tma_phase ^= 1u;
#else
      cp_async_wait_all();
      ex.sync_warp();
#endif
#endif
    };
    if (tma_pending) { staged_wait(); tma_pending = false; }  // a restarted pass still has the aborted knot's prefetch in flight
    bsync();
    prefetch(T - 1, hcur);
    tma_pending = true;
    for (int t = T - 1; t >= 0; --t) {
      DDPB_PHASE_START(bd, pc_);
      const int ne = bd.ne[t], eo = bd.eq_off[t];
      const RecLayout L = rec_layout(n, ne, bd.hess_mode);
      const double* rec = slab_rec(bd, t, b, L.size);
      // ---- phase A: head of the record (prefetched) ; tensors -> V_x contraction ----------------------------------------
      const double* hd = sm + (hcur ? S::head2 : S::head);
      {
        // W(e) = sum_k V_x[r0 + k] T_e[k] for every stored tensor element e (a run of RL doubles at L.fxx + e RL)
        auto contract = [&](auto rl_c, auto rs_c, auto ew_c) {
          constexpr int RL = decltype(rl_c)::value;
          constexpr int RS = decltype(rs_c)::value;              // elements per round (shadows the runtime value)
          constexpr bool MULTI = RS < decltype(ew_c)::value;     // more than one round per warp and knot
          (void)RS; (void)MULTI;
          const double* vx = sm + S::Vx + (RL == nd ? 0 : n);
#if defined(__CUDA_ARCH__)
          // Each warp owns a contiguous range of the element list (= a contiguous piece of the adjacent tensor blocks) and
          // streams it in rounds of RS elements: fully coalesced 16-byte loads (consecutive lanes -> consecutive chunks)
          // into the warp's staging buffer, then every lane contracts the k-run of its own element with V_x.  Only
          // __syncwarp is needed: the buffer is private to the warp.  Round 0 was fetched asynchronously during the
          // previous knot; the loads of round r+1 are in flight while round r is contracted.
          constexpr int LPR = (RS * RL / 2 + 31) / 32;  // 16-byte loads per lane and round
          constexpr bool RAL = (RL % 2) == 0;
          double2 buf[LPR];
          auto load_round = [&](int e0) {
            const int ne_round = e_end - e0 < RS ? e_end - e0 : RS;
            const int nch = (ne_round * RL + 1) / 2;
            const double2* src = reinterpret_cast<const double2*>(rec + L.fxx + (size_t)e0 * RL);
#pragma unroll
            for (int i = 0; i < LPR; ++i) {
              const int c = lane + 32 * i;
              if (c < nch) buf[i] = src[c];
            }
          };
          if (MULTI && e_beg + RS < e_end) load_round(e_beg + RS);
          DDPB_PHASE_MARK(bd, 13, pc_);
          staged_wait();
          tma_pending = false;
          DDPB_PHASE_MARK(bd, 14, pc_);
          for (int e0 = e_beg; e0 < e_end; e0 += RS) {
            const int ne_round = e_end - e0 < RS ? e_end - e0 : RS;
            for (int l = lane; l < ne_round; l += 32) sm[S::Wf + e0 + l] = contract_run<RL, RAL>(vx, stg + l * RL);  // a round may hold more than 32 elements
            __syncwarp();
            if (MULTI && e0 + RS < e_end) {
              const int nn = e_end - (e0 + RS) < RS ? e_end - (e0 + RS) : RS;
              const int nch = (nn * RL + 1) / 2;
#pragma unroll
              for (int i = 0; i < LPR; ++i) {
                const int c = lane + 32 * i;
                if (c < nch) reinterpret_cast<double2*>(stg)[c] = buf[i];
              }
              if (e0 + 2 * RS < e_end) load_round(e0 + 2 * RS);
              __syncwarp();
            }
          }
#else
          for (int e = 0; e < NEL; ++e) sm[S::Wf + e] = contract_run<RL, false>(vx, rec + L.fxx + (size_t)e * RL);
#endif
        };
#if defined(__CUDA_ARCH__)
        contract(IntC<(dense ? nd : n)>{}, IntC<RS>{}, IntC<EW>{});
#else
        contract(IntC<(dense ? nd : n)>{}, IntC<0>{}, IntC<0>{});
#endif
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
        DDPB_PHASE_MARK(bd, 15, pc_);
#if !(defined(__CUDA_ARCH__) && DDPB_BWD_TMA)
        if (t > 0) { prefetch(t - 1, hcur ^ 1); tma_pending = true; }  // the staging buffer and the other head buffer are free
#endif
        hcur ^= 1;
      }
      bsync();
      DDPB_PHASE_MARK(bd, 8, pc_);
#if defined(__CUDA_ARCH__) && DDPB_BWD_TMA
      // every warp is past its last read of the staging buffers: stage the next knot (hcur now names the idle head buffer)
      if (t > 0) { prefetch(t - 1, hcur); tma_pending = true; }
#endif
      // ---- phase B: tmp, tmp2, VF = V_xx f_x, VU = V_xx f_u ----------------------------------------------------
      for (int e = tid; e < ne; e += BT) sm[S::tmp + e] = sm[S::pe + e] + mu * sm[S::eqv + e];
      for (int e = tid; e < ne * nd; e += BT) sm[S::tmp2 + e] = sm[S::pex + e] + mu * sm[S::eqx + e];
      {
        // results are kept in registers until every chain of the thread is done: a shared-memory store between two
        // chains would order them (possible aliasing) and serialise the FMA latencies
        double acc[(NB + BT - 1) / BT];
        FOR_ELEMS(e, NB) {
          unsigned d0, d1;
          if (PRE) { d0 = dB0[PRE ? u_ : 0]; d1 = dB1[PRE ? u_ : 0]; } else mkB(e, d0, d1);
          const double run = dot_run<n, AL>(sm + (d0 & 0xffffu), hd + (d0 >> 16));
          acc[u_] = sm[d1 & 0xffffu] * sm[d1 >> 16] + run;
        }
        FOR_ELEMS(e, NB) sm[S::VF + e] = acc[u_];  // VF | VU are adjacent
      }
      bsync();
      DDPB_PHASE_MARK(bd, 9, pc_);
      const double* eqv = sm + S::eqv;
      const double* eqx = sm + S::eqx;
      const double* equ = sm + S::equ;
      const double* pex = sm + S::pex;
      const double* tmp = sm + S::tmp;
      const double* tmp2 = sm + S::tmp2;
      // ---- phase C: all Q elements as one task list: Q_xx | Q_ux | Q_uu | Q_x | Q_u ------------------------------
      // Q = l + f^T (V f) [+ equality terms] + tensor term, in the reference's order of additions
      {
        double qacc[(NC + BT - 1) / BT];
        unsigned qdst[(NC + BT - 1) / BT];
        FOR_ELEMS(e, NC) {
          unsigned c0, c1, c2, c3;
          if (PRE) { c0 = dC0[PRE ? u_ : 0]; c1 = dC1[PRE ? u_ : 0]; c2 = dC2[PRE ? u_ : 0]; c3 = dC3[PRE ? u_ : 0]; } else mkC(e, c0, c1, c2, c3);
          qdst[u_] = c2;
          const unsigned lo = c1 & 0xffffu;
          double q = (lo & 0x8000u) ? hd[lo & 0x7fffu] : sm[lo];
          const double run = dot_run<n, AL>(hd + (c0 & 0xffffu), sm + (c0 >> 16));
          q += sm[c3 & 0xffffu] * sm[c3 >> 16] + run;
          if (ne > 0) {
            // knots with equality rows (rare): terms by block
            double a = 0.0;
            if (e < nQxx) {
              const int r = e % nd, c = e / nd;
              for (int k = 0; k < ne; ++k) a += eqx[r * ne + k] * tmp2[c * ne + k];
              q += a;
              a = 0.0;
              for (int k = 0; k < ne; ++k) a += pex[r * ne + k] * eqx[c * ne + k];
              q += a;
              a = 0.0;
              const double* t3 = rec + L.eq_xx + r * ne + (size_t)c * ne * nd;
              for (int k = 0; k < ne; ++k) a += tmp[k] * t3[k];
              q += a;
            } else if (e < nQxx + nQux) {
              const int ee = e - nQxx, r = ee % nu, c = ee / nu;
              for (int k = 0; k < ne; ++k) a += equ[r * ne + k] * tmp2[c * ne + k];
              q += a;
              a = 0.0;
              const double* t3 = rec + L.eq_ux + r * ne + (size_t)c * ne * nu;
              for (int k = 0; k < ne; ++k) a += tmp[k] * t3[k];
              q += a;
            } else if (e < NTEN) {
              const int ee = e - nQxx - nQux, r = ee % nu, c = ee / nu;
              for (int k = 0; k < ne; ++k) a += equ[r * ne + k] * equ[c * ne + k];
              q += a * mu;
              a = 0.0;
              const double* t3 = rec + L.eq_uu + r * ne + (size_t)c * ne * nu;
              for (int k = 0; k < ne; ++k) a += tmp[k] * t3[k];
              q += a;
            } else if (e < NTEN + nd) {
              const int i = e - NTEN;
              for (int k = 0; k < ne; ++k) a += eqx[i * ne + k] * tmp[k];
              q += a;
              a = 0.0;
              for (int k = 0; k < ne; ++k) a += pex[i * ne + k] * eqv[k];
              q += a;
            } else {
              const int i = e - NTEN - nd;
              for (int k = 0; k < ne; ++k) a += equ[i * ne + k] * tmp[k];
              q += a;
            }
          }
          q += sm[c1 >> 16];
          qacc[u_] = q;
        }
        // the matrix the factorisation works on, Q_uu + reg I, is written by the owners of the Q_uu elements
        FOR_ELEMS(e, NC) {
          sm[qdst[u_] & 0xffffu] = qacc[u_];
          const unsigned ld = qdst[u_] >> 16;
          if ((ld & 0x7fffu) != 0x7fffu) sm[ld & 0x7fffu] = qacc[u_] + ((ld & 0x8000u) ? reg : 0.0);
        }
      }
      bsync();
      DDPB_PHASE_MARK(bd, 10, pc_);
      // ---- phase D (first warp): Cholesky of Q_uu + reg I, then k = -(LL^T)^-1 Q_u, K = -(LL^T)^-1 Q_ux ---------------------
#if defined(__CUDA_ARCH__)
      // One lane per column of the augmented matrix [Q_uu + reg I | -Q_ux | -Q_u] (nu + nd + 1 <= 32 columns), its nu
      // entries in registers.  Right-looking factorisation U^T U (U = L^T): at step k the pivot lane broadcasts its diagonal,
      // every lane scales its row-k entry by 1/sqrt(pivot) and subtracts U(k, i) U(k, c) from its entries i > k -- the
      // right-hand-side lanes thereby do the forward substitution in the same sweep (same order of subtractions as
      // Eigen::LLT's left-looking loop and a row-oriented solve; fails on a pivot <= 0 like Eigen::LLT).  The back
      // substitution is column-oriented on the right-hand-side lanes with L broadcast once.  ~12 dependent steps of one
      // multiply-add each instead of ~70 shared-memory round trips (this phase was a third of the knot, ncu).
#ifdef DDPB_BWD_SKIP_D  // development: what does this phase cost? (results are wrong)
      if (tid < nu * nd) sm[S::KK + tid] = 0.0;
      if (tid < nu) sm[S::kk + tid] = 0.0;
      if (tid == 0) sm[S::flags] = 1.0;
      if (false) {
#else
      if (tid < 32) {
#endif
        const unsigned FULL = 0xffffffffu;
        const int ln = tid;
        constexpr int NCOL = nu + nd + 1;
        static_assert(NCOL <= 32, "one lane per column of the augmented matrix");
        double a[nu], dinv_[nu];
        {
          // Eigen::LLT reads the LOWER triangle (Q_uu is not symmetric to the last bit: V_xx is not symmetrised):
          // lane c < nu takes row c of the matrix, U(k, c) = L(c, k)
          const double* src = ln < nu ? sm + S::L + ln : (ln < nu + nd ? sm + S::Qux + (ln - nu) * nu : sm + S::Qu);
          const int stride = ln < nu ? nu : 1;
          const double sgn = ln < nu ? 1.0 : -1.0;
#pragma unroll
          for (int i = 0; i < nu; ++i) a[i] = ln < NCOL ? sgn * src[i * stride] : 0.0;
        }
        bool ok = true;
#pragma unroll
        for (int k = 0; k < nu; ++k) {
          const double piv = __shfl_sync(FULL, a[k], k);
          if (piv <= 0.0) { ok = false; break; }
          const double linv = rsqrt(piv);  // column scaling by 1/sqrt(pivot): one special-function latency per step
          dinv_[k] = linv;
          a[k] = ln == k ? piv * linv : a[k] * linv;
#pragma unroll
          for (int i = 0; i < nu; ++i)
            if (i > k) {
              const double uki = __shfl_sync(FULL, a[k], i);  // U(k, i), final once row k is scaled
              a[i] -= uki * a[k];
            }
        }
        if (tid == 0) sm[S::flags] = ok ? 1.0 : 0.0;
        if (ok) {
          // x_k = y_k / L(k, k), y_i -= L(k, i) x_k  (k = nu-1 .. 0); L(k, i) = U(i, k) is lane k's a[i], fetched where it
          // is used (the factor lanes no longer change; holding all of L would cost 2 nu (nu - 1) / 2 registers)
#pragma unroll
          for (int k = nu - 1; k >= 0; --k) {
            a[k] *= dinv_[k];
#pragma unroll
            for (int i = 0; i < nu; ++i)
              if (i < k) {
                const double lki = __shfl_sync(FULL, a[i], k);
                if (ln >= nu) a[i] -= lki * a[k];
              }
          }
          if (ln >= nu && ln < nu + nd) {
            const int c = ln - nu;
#pragma unroll
            for (int i = 0; i < nu; ++i) {
              sm[S::KK + c * nu + i] = a[i];
              bd.fb_K[((size_t)t * nu * nd + c * nu + i) * B + b] = a[i];
            }
          } else if (ln == nu + nd) {
#pragma unroll
            for (int i = 0; i < nu; ++i) {
              sm[S::kk + i] = a[i];
              bd.fb_k[((size_t)t * nu + i) * B + b] = a[i];
            }
          }
        }
      }
#else
      {
        const bool ok = chol_host<nu>(sm + S::L, sm + S::dinv);
        sm[S::flags] = ok ? 1.0 : 0.0;
        if (ok) {
          for (int c = 0; c <= nd; ++c) {
            const double* A = sm + S::L;
            const double* di = sm + S::dinv;
            double rhs[nu];
            for (int i = 0; i < nu; ++i) rhs[i] = -(c == nd ? sm[S::Qu + i] : sm[S::Qux + c * nu + i]);
            for (int i = 0; i < nu; ++i) {
              double sacc = rhs[i];
              for (int j = 0; j < i; ++j) sacc -= A[j * nu + i] * rhs[j];
              rhs[i] = sacc * di[i];
            }
            for (int k = nu - 1; k >= 0; --k) {   // column-oriented, as on the device
              rhs[k] *= di[k];
              for (int i = 0; i < k; ++i) rhs[i] -= A[i * nu + k] * rhs[k];
            }
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
        }
      }
#endif
      bsync();
      DDPB_PHASE_MARK(bd, 11, pc_);
      if (sm[S::flags] == 0.0) {
        if (reg < mu) reg = mu;
        mu *= 2.0;
        reg *= 2.0;
        ++restarts;
        failed = true;
        bsync();
        break;
      }
      // ---- phase F: V_x = Q_x + Q_ux^T k ; V_xx = Q_xx + Q_ux^T K  (V_xx stored row-major) ----------------------------
      {
        double vacc[(NF + BT - 1) / BT];
        unsigned vdst[(NF + BT - 1) / BT];
        FOR_ELEMS(e, NF) {
          unsigned f0, f1;
          if (PRE) { f0 = dF0[PRE ? u_ : 0]; f1 = dF1[PRE ? u_ : 0]; } else mkF(e, f0, f1);
          vdst[u_] = f1 >> 16;
          const double* qa = sm + (f0 & 0xffffu);
          const double* kb = sm + (f0 >> 16);
          double a = 0.0;
#pragma unroll
          for (int k = 0; k < nu; ++k) a += qa[k] * kb[k];
          vacc[u_] = sm[f1 & 0xffffu] + a;
        }
        FOR_ELEMS(e, NF) sm[vdst[u_]] = vacc[u_];  // V_xx (row-major) | V_x are adjacent
      }
      bsync();
      DDPB_PHASE_MARK(bd, 12, pc_);
    }
    if (!failed) success = true;
    else if (bd.max_restarts > 0 && (restarts > bd.max_restarts || !(mu < 1.79e308))) {
      // guarded mode: this problem's Q_uu cannot be made positive definite by the reference's mu/reg schedule.
      // The copy staged for the knot after the failing one must land before the CTA (and its shared memory) goes away.
      if (tma_pending) { staged_wait(); tma_pending = false; }
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
