// kernels_nj.cu -- the hot-path kernels for one model size and joint-type variant; compile with -DDDPB_NJ=<0..7>
// [-DDDPB_PRISMATIC=1] (variant.cuh; the pendulum, NJ = 0, has no prismatic variant).
#ifndef DDPB_NJ
#error "compile with -DDDPB_NJ=<n>"
#endif
#include "backward.cuh"
#include "backward_lg.cuh"
#include "forward.cuh"
#include "launch.cuh"
#include "linearize.cuh"
#include "linearize_pt.cuh"
#include "model_eval.cuh"
#include "outer_warp.cuh"

DDPB_VARIANT_BEGIN

constexpr int KNJ = DDPB_NJ;

// Every kernel covers the problems [bd.b0, bd.b0 + bd.nb) of the batch: the solver splits a batch into sub-batches
// that run the iteration sequence on their own streams, so the latency-bound sweeps of one sub-batch overlap the
// throughput-bound linearization of another and a straggling problem delays only its own sub-batch.
#if DDPB_NJ > 0
// grid: (knots t0..t1-1) x (groups of `kpb` consecutive problems); all knots of a launch have the same ne
// FD-of-Jacobians kernels are compiled for 512 resident threads per SM (128 registers): CTAs of <= 128 threads for small
// models (2R chain: 16 knots each, 3 per SM by shared memory), 2 CTAs of 256 threads from 5 joints on (6R arm: three knots
// each, so the narrow serial sub-phases serve three knots and an SM holds six).  The value-based mode has
// 1 + n + n(n-1)/2 evaluation points per knot and is compiled for 255 registers; it is LAUNCHED with 128 threads from 5
// joints on, so two CTAs share an SM and overlap each other's narrow phases (7R: 26.0 -> 16.1 ms).  At 128 registers its
// value-only ABA tasks spill ~1.6 KB per thread and ncu shows 6x the algorithmic DRAM traffic (DDPB_LIN_LB_BLOCKS_FDV=2).
// The launch geometry (threads, knots per CTA) is chosen in ddpb_batch_create (ddpb_capi.cu).
#ifndef DDPB_LIN_LB_THREADS
#define DDPB_LIN_LB_THREADS (DDPB_NJ >= 5 ? 256 : 128)
#define DDPB_LIN_LB_BLOCKS (DDPB_NJ >= 5 ? 2 : 4)
#endif
#ifndef DDPB_LIN_LB_BLOCKS_FDV
#define DDPB_LIN_LB_BLOCKS_FDV 1
#endif
template <int NJ, int MODE> __global__ void __launch_bounds__(MODE == HESS_FD_JACOBIANS ? DDPB_LIN_LB_THREADS : 256, MODE == HESS_FD_JACOBIANS ? DDPB_LIN_LB_BLOCKS : DDPB_LIN_LB_BLOCKS_FDV) linearize_kernel(const __grid_constant__ ModelDev m, const __grid_constant__ BatchDev bd,
                                                                          int t0, int kpb, int ne_sm) {
  extern __shared__ __align__(16) double sm[];
  const int groups = (bd.nb + kpb - 1) / kpb;
  const int t = t0 + blockIdx.x / groups, b0 = bd.b0 + (blockIdx.x % groups) * kpb;
  const int nk = min(kpb, bd.b0 + bd.nb - b0);
  bool any = false;
  for (int k = 0; k < nk; ++k) any = any || bd.scal[(size_t)S_STATUS * bd.B + b0 + k] == 0.0;
  if (!any) return;  // every problem of the group is frozen (converged / diverged)
  linearize_group<NJ, MODE>(m, bd, t, b0, nk, kpb, ne_sm, sm, DevExec{});
}
template <int NJ> __global__ void fb_origin_kernel(const __grid_constant__ BatchDev bd) {  // (template: one copy per translation unit)
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (size_t)bd.T * bd.nx * bd.nb) fb_origin_knot(bd, i);
}
#if DDPB_NJ <= 2
// one thread per evaluation point (linearize_pt.cuh): 4 independent warps per CTA, 255 registers per thread
constexpr int PT_WARPS = 4;
#ifndef DDPB_PT_BLOCKS
#define DDPB_PT_BLOCKS 2   // CTAs per SM the register allocation aims at (2: 255 registers, no spill)
#endif
template <int NJ> __global__ void __launch_bounds__(32 * PT_WARPS, DDPB_PT_BLOCKS) linearize_pt_kernel(const __grid_constant__ ModelDev m, const __grid_constant__ BatchDev bd, int t0, int nt) {
  __shared__ __align__(16) double stage[PT_WARPS][PtGeom<NJ>::KPW * PtGeom<NJ>::SIZE];
  linearize_pt_warp<NJ>(m, bd, t0, nt, stage[threadIdx.x >> 5]);
}
// chains of <= 2 joints: one lane group of ndx^2 lanes per problem, registers and shuffles only (backward_lg.cuh)
constexpr int LG_WARPS = 4;  // warps per CTA (independent of each other)
template <int NJ, bool DENSE> __global__ void __launch_bounds__(32 * LG_WARPS) backward_lg_kernel(const __grid_constant__ BatchDev bd) {
  extern __shared__ __align__(16) double sm[];
  backward_lg_warp<NJ, DENSE>(bd, sm + (size_t)(threadIdx.x >> 5) * LgRec<NJ, DENSE>::smem_per_warp);
}
#else
// one CTA of BwdGeom<NJ>::BT threads per problem; enough CTAs per SM that a whole batch is resident in one wave
template <int NJ, bool DENSE> __global__ void __launch_bounds__(BwdGeom<NJ>::BT, NJ <= 3 ? 16 : (NJ <= 5 ? 8 : 7)) backward_kernel(const __grid_constant__ BatchDev bd) {
  extern __shared__ __align__(16) double sm[];
  const int b = bd.b0 + blockIdx.x;
  if (bd.scal[(size_t)S_STATUS * bd.B + b] != 0.0) return;
  backward_problem<NJ, BwdGeom<NJ>::BT, DENSE>(bd, b, sm, DevExec{});
}
#endif
#else
template <int NJ> __global__ void linearize_kernel(const __grid_constant__ ModelDev m, const __grid_constant__ BatchDev bd, int t0, int t1) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)bd.nb * (t1 - t0)) return;
  const int b = bd.b0 + (int)(i % bd.nb), t = t0 + (int)(i / bd.nb);
  if (bd.scal[(size_t)S_STATUS * bd.B + b] != 0.0) return;
  linearize_pendulum(m, bd, b, t);
}
#endif

// the Model concept's own functions for n independent samples (model_eval.cuh): one thread per sample, workspaces in local memory
template <int NJ> __global__ void __launch_bounds__(64) model_eval_kernel(const __grid_constant__ ModelDev m, const __grid_constant__ ModelEvalArgs A) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= A.n) return;
#if DDPB_NJ > 0
  double qrec[QRec<NJ>::SIZE], vrec[VRec<NJ>::SIZE];
  model_eval_sample<NJ>(m, A, s, qrec, vrec);
#else
  model_eval_pendulum(m, A, s);
#endif
}

// line search: cost_seq_aug of the current trajectory, then all trials at once (outer_warp.cuh)
// per-thread scratch of the fused forward dynamics in dynamic shared memory: element e of thread i at dyn_scr[e * blockDim.x + i]
extern __shared__ __align__(16) double dyn_scr[];
__device__ __forceinline__ Scratch thread_scratch() { return Scratch{dyn_scr + threadIdx.x, (int)blockDim.x}; }
// cost_seq_aug of the current trajectory (ddp.hpp:699-735): the terms of different knots are independent, so one thread
// per (knot, problem) instead of a T-step loop per problem -- problem index fastest => coalesced SoA accesses
template <int NJ> __global__ void cost_old_kernel(const __grid_constant__ ModelDev m, const __grid_constant__ BatchDev bd, double* costs_old) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)bd.nb * (bd.T + 1)) return;
  const int b = bd.b0 + (int)(i % bd.nb), t = (int)(i / bd.nb);
  if (bd.scal[(size_t)S_STATUS * bd.B + b] != 0.0) return;
  cost_term_knot<NJ>(m, bd, b, t, costs_old, bd.scal[(size_t)S_MU * bd.B + b], thread_scratch());
}
template <int NJ> __global__ void __launch_bounds__(32 * LsGeom<NJ>::WPB) ls_trials_kernel(const __grid_constant__ ModelDev m, const __grid_constant__ BatchDev bd,
                                                                                         const double* costs_old, double* xs_c, double* us_c, double* res,
                                                                                         const int* win, const int* wide, int c0, int c1) {
  ls_trials_block<NJ>(m, bd, costs_old, xs_c, us_c, res, win, wide, c0, c1, thread_scratch());
}
template <int NJ> __global__ void ls_select_kernel(const __grid_constant__ BatchDev bd, const double* xs_c, const double* us_c, const double* res, int* win,
                                                   const int* wide, int* wide_next, int c0, int c1, int last) {
  ls_select_warp<NJ>(bd, xs_c, us_c, res, win, wide, wide_next, c0, c1, last != 0);
}
template <int NJ> __global__ void rollout_kernel(const __grid_constant__ ModelDev m, const __grid_constant__ BatchDev bd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= bd.nb) return;
  rollout_problem<NJ>(m, bd, bd.b0 + i, thread_scratch());
}
template <int NJ> __global__ void cost_kernel(const __grid_constant__ ModelDev m, const __grid_constant__ BatchDev bd, double* costs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, b = bd.b0 + i;
  if (i >= bd.nb) return;
  const double plain = cost_seq_problem<NJ>(m, bd, b, costs, bd.scal[(size_t)S_MU * bd.B + b], thread_scratch());
  bd.scal[(size_t)S_COST * bd.B + b] = plain;
}
// outer-loop sweeps (outer_warp.cuh): per-(knot, problem) kernels and one-warp-per-problem adjoint recursion
template <int NJ> __global__ void origin_update_kernel(const __grid_constant__ BatchDev bd, int fb_all) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)bd.nb * bd.T) return;
  origin_update_knot<NJ>(bd, bd.b0 + (int)(i % bd.nb), (int)(i / bd.nb), fb_all != 0);
}
template <int NJ> __global__ void mult_update_kernel(const __grid_constant__ BatchDev bd) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)bd.nb * bd.T) return;
  mult_update_knot<NJ>(bd, bd.b0 + (int)(i % bd.nb), (int)(i / bd.nb));
}
template <int NJ> __global__ void al_decide_kernel(const __grid_constant__ BatchDev bd, double thr, int phase) {
  const int i = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) / AdjGeom<NJ>::G);
  const bool live = i < bd.nb;
  al_decide_warp<NJ>(bd, bd.b0 + (live ? i : 0), live, thr, phase);
}
template <int NJ> __global__ void optimality_kernel(const __grid_constant__ BatchDev bd) {
  const int i = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) / AdjGeom<NJ>::G);
  const bool live = i < bd.nb;
  const int b = bd.b0 + (live ? i : 0);
  const double oo = optimality_obj_warp<NJ>(bd, b, bd.scal[(size_t)S_MU * bd.B + b]);
  if ((threadIdx.x & (AdjGeom<NJ>::G - 1)) != 0 || !live) return;
  bd.scal[(size_t)S_OPT_OBJ * bd.B + b] = oo;
  bd.scal[(size_t)S_OPT_CONSTR * bd.B + b] = optimality_constr_problem<NJ>(bd, b);
}

constexpr int TPP = 32;  // threads per block of the thread-per-problem kernels
static size_t scr_bytes(int threads) { return (size_t)DynOps<KNJ>::SCR * threads * sizeof(double); }

DDPB_VARIANT_END

namespace ddpb {

template <> size_t Launch<KNJ, HAS_PRISMATIC>::lin_smem_bytes(int hess_mode, int ne_max, int eq_kind) {
#if DDPB_NJ > 0
  LinPlan<KNJ> pl;
  pl.init(hess_mode, ne_max, ne_max, eq_kind);
  return (size_t)pl.total * sizeof(double);
#else
  (void)hess_mode; (void)ne_max; (void)eq_kind;
  return 0;
#endif
}
template <> size_t Launch<KNJ, HAS_PRISMATIC>::bwd_smem_bytes(int hess_mode) {
#if DDPB_NJ > 2
  return (size_t)BwdSmem<KNJ>::size_for(BwdGeom<KNJ>::BT, hess_mode == HESS_FD_VALUES) * sizeof(double);
#elif DDPB_NJ > 0
  return (size_t)LG_WARPS * sizeof(double) * (hess_mode == HESS_FD_VALUES ? LgRec<KNJ, true>::smem_per_warp : LgRec<KNJ, false>::smem_per_warp);
#else
  (void)hess_mode;
  return 0;
#endif
}
template <> cudaError_t Launch<KNJ, HAS_PRISMATIC>::prepare(size_t lin_smem) {
  {
    cudaError_t e = cudaFuncSetAttribute(ls_trials_kernel<KNJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scr_bytes(32 * LsGeom<KNJ>::WPB));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(cost_old_kernel<KNJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scr_bytes(TPP));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(cost_kernel<KNJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scr_bytes(TPP));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rollout_kernel<KNJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scr_bytes(TPP));
    if (e != cudaSuccess) return e;
  }
#if DDPB_NJ > 0
  cudaError_t e = cudaFuncSetAttribute(linearize_kernel<KNJ, HESS_FD_JACOBIANS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lin_smem);
  if (e != cudaSuccess) return e;
#if DDPB_NJ <= 2
  // the record rings of the lane-group backward pass (8 knots per problem) exceed the default 48 KB in value-based mode
  e = cudaFuncSetAttribute(backward_lg_kernel<KNJ, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_smem_bytes(HESS_FD_VALUES));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(backward_lg_kernel<KNJ, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_smem_bytes(HESS_FD_JACOBIANS));
  if (e != cudaSuccess) return e;
#endif
#if DDPB_NJ <= 2
#endif
  return cudaFuncSetAttribute(linearize_kernel<KNJ, HESS_FD_VALUES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lin_smem);
#else
  (void)lin_smem;
  return cudaSuccess;
#endif
}
template <> void Launch<KNJ, HAS_PRISMATIC>::linearize(const ModelDev& m, const BatchDev& bd, int threads, int kpb, int t0, int t1, int ne_sm, cudaStream_t s) {
#if DDPB_NJ > 0
  const size_t smem = lin_smem_bytes(bd.hess_mode, ne_sm, bd.eq_kind) * kpb;
  const int groups = (bd.nb + kpb - 1) / kpb;
  const unsigned grid = (unsigned)((size_t)groups * (t1 - t0));
  if (bd.hess_mode == HESS_FD_JACOBIANS) linearize_kernel<KNJ, HESS_FD_JACOBIANS><<<grid, threads, smem, s>>>(m, bd, t0, kpb, ne_sm);
  else linearize_kernel<KNJ, HESS_FD_VALUES><<<grid, threads, smem, s>>>(m, bd, t0, kpb, ne_sm);
#else
  (void)threads; (void)kpb; (void)ne_sm;
  const size_t n = (size_t)bd.nb * (t1 - t0);
  linearize_kernel<KNJ><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(m, bd, t0, t1);
#endif
}
template <> bool Launch<KNJ, HAS_PRISMATIC>::linearize_pt(const ModelDev& m, const BatchDev& bd, int t0, int t1, cudaStream_t s) {
#if DDPB_NJ > 0 && DDPB_NJ <= 2
  if (bd.hess_mode != HESS_FD_JACOBIANS) return false;
  const long long knots = (long long)(t1 - t0) * bd.nb;
  const long long warps = (knots + PtGeom<KNJ>::KPW - 1) / PtGeom<KNJ>::KPW;
  linearize_pt_kernel<KNJ><<<(unsigned)((warps + PT_WARPS - 1) / PT_WARPS), 32 * PT_WARPS, 0, s>>>(m, bd, t0, t1 - t0);
  return true;
#else
  (void)m; (void)bd; (void)t0; (void)t1; (void)s;
  return false;
#endif
}
template <> void Launch<KNJ, HAS_PRISMATIC>::backward(const BatchDev& bd, int threads, size_t smem, cudaStream_t s) {
#if DDPB_NJ > 0
  {
    const size_t ne_ = (size_t)bd.T * bd.nx * bd.nb;
    fb_origin_kernel<KNJ><<<(unsigned)((ne_ + 255) / 256), 256, 0, s>>>(bd);
  }
#endif
#if DDPB_NJ > 2
  (void)threads;
  if (bd.hess_mode == HESS_FD_VALUES) backward_kernel<KNJ, true><<<bd.nb, BwdGeom<KNJ>::BT, smem, s>>>(bd);
  else backward_kernel<KNJ, false><<<bd.nb, BwdGeom<KNJ>::BT, smem, s>>>(bd);
#elif DDPB_NJ > 0
  (void)threads;
  if (bd.hess_mode == HESS_FD_VALUES) {
    constexpr int ppc = LG_WARPS * LgRec<KNJ, true>::GPW;  // problems per CTA
    backward_lg_kernel<KNJ, true><<<(bd.nb + ppc - 1) / ppc, 32 * LG_WARPS, smem, s>>>(bd);
  } else {
    constexpr int ppc = LG_WARPS * LgRec<KNJ, false>::GPW;
    backward_lg_kernel<KNJ, false><<<(bd.nb + ppc - 1) / ppc, 32 * LG_WARPS, smem, s>>>(bd);
  }
#else
  (void)bd; (void)threads; (void)smem; (void)s;
#endif
}
template <> void Launch<KNJ, HAS_PRISMATIC>::cost_old(const ModelDev& m, const BatchDev& bd, double* costs_old, cudaStream_t s) {
  const size_t nkb = (size_t)bd.nb * (bd.T + 1);
  cost_old_kernel<KNJ><<<(unsigned)((nkb + TPP - 1) / TPP), TPP, scr_bytes(TPP), s>>>(m, bd, costs_old);
}
template <> void Launch<KNJ, HAS_PRISMATIC>::forward(const ModelDev& m, const BatchDev& bd, double* costs_old, double* xs_c, double* us_c, double* ls_res, int* ls_win,
                                      const int* wide, int* wide_next, cudaStream_t s) {
  using G = LsGeom<KNJ>;
  const unsigned pb = (unsigned)((bd.nb + 31) / 32), sb = (unsigned)(((size_t)bd.nb * 32 + 127) / 128);
  // this pass's verdicts start from "narrow" (sub-batches start on multiples of 32 problems)
  cudaMemsetAsync(wide_next + (bd.b0 >> 5), 0, pb * sizeof(int), s);
  // first wave: the grid covers all 34 trials, the blocks of the trials past NC1 leave at once unless their problems are flagged
  ls_trials_kernel<KNJ><<<dim3(pb, (unsigned)((LS_MAX_TRIALS + G::WPB - 1) / G::WPB)), 32 * G::WPB, scr_bytes(32 * G::WPB), s>>>(
      m, bd, costs_old, xs_c, us_c, ls_res, ls_win, wide, 0, LS_MAX_TRIALS);
  ls_select_kernel<KNJ><<<sb, 128, 0, s>>>(bd, xs_c, us_c, ls_res, ls_win, wide, wide_next, 0, G::NC1, 0);
  // second wave: the remaining trials of the problems still unresolved
  ls_trials_kernel<KNJ><<<dim3(pb, (unsigned)((LS_MAX_TRIALS - G::NC1 + G::WPB - 1) / G::WPB)), 32 * G::WPB, scr_bytes(32 * G::WPB), s>>>(
      m, bd, costs_old, xs_c, us_c, ls_res, ls_win, wide, G::NC1, LS_MAX_TRIALS);
  ls_select_kernel<KNJ><<<sb, 128, 0, s>>>(bd, xs_c, us_c, ls_res, ls_win, wide, wide_next, G::NC1, LS_MAX_TRIALS, 1);
}
template <> int Launch<KNJ, HAS_PRISMATIC>::linesearch_slots() { return ls_slots<KNJ>(); }
template <> void Launch<KNJ, HAS_PRISMATIC>::rollout(const ModelDev& m, const BatchDev& bd, cudaStream_t s) {
  rollout_kernel<KNJ><<<(bd.nb + TPP - 1) / TPP, TPP, scr_bytes(TPP), s>>>(m, bd);
}
template <> void Launch<KNJ, HAS_PRISMATIC>::cost(const ModelDev& m, const BatchDev& bd, double* costs, cudaStream_t s) {
  cost_kernel<KNJ><<<(bd.nb + TPP - 1) / TPP, TPP, scr_bytes(TPP), s>>>(m, bd, costs);
}
template <> void Launch<KNJ, HAS_PRISMATIC>::al_update(const BatchDev& bd, double thr, cudaStream_t s) {
  const size_t nkb = (size_t)bd.nb * bd.T;
  const unsigned wblocks = (unsigned)(((size_t)bd.nb * AdjGeom<KNJ>::G + 127) / 128);
  origin_update_kernel<KNJ><<<(unsigned)((nkb + 127) / 128), 128, 0, s>>>(bd, thr > 0.0 ? 1 : 0);
  al_decide_kernel<KNJ><<<wblocks, 128, 0, s>>>(bd, thr, 0);
  mult_update_kernel<KNJ><<<(unsigned)((nkb + 127) / 128), 128, 0, s>>>(bd);
  al_decide_kernel<KNJ><<<wblocks, 128, 0, s>>>(bd, thr, 1);
}
template <> void Launch<KNJ, HAS_PRISMATIC>::model_eval(const ModelDev& m, const ModelEvalArgs& A, cudaStream_t s) {
  if (A.n > 0) model_eval_kernel<KNJ><<<(unsigned)((A.n + 63) / 64), 64, 0, s>>>(m, A);
}
template <> void Launch<KNJ, HAS_PRISMATIC>::optimality(const BatchDev& bd, cudaStream_t s) {
  optimality_kernel<KNJ><<<(unsigned)(((size_t)bd.nb * AdjGeom<KNJ>::G + 127) / 128), 128, 0, s>>>(bd);
}

}  // namespace ddpb
