// launch.cuh -- per-model-size kernel launchers.  Each instantiation Launch<NJ, PR> lives in its own
// translation unit (kernels_nj.cu compiled with -DDDPB_NJ=<n> -DDDPB_PRISMATIC=<0|1>, variant.cuh) so the build
// parallelises.  PR = false: every joint revolute; PR = true: the device code with the joint-type tests.
// NJ = 0 is the analytic pendulum (thread-per-problem kernels only; its backward pass is Launch<1>'s).
#pragma once
#include <cuda_runtime.h>

#include "batch.cuh"
#include "model_eval.cuh"

namespace ddpb {

template <int NJ, bool PR = false> struct Launch {
  static size_t lin_smem_bytes(int hess_mode, int ne_max, int eq_kind);
  static size_t bwd_smem_bytes(int hess_mode);
  static cudaError_t prepare(size_t lin_smem);
  // knots [t0, t1) (all with the same number of equality rows <= ne_sm), kpb problems per CTA
  static void linearize(const ModelDev& m, const BatchDev& bd, int threads, int kpb, int t0, int t1, int ne_sm, cudaStream_t s);
  // chains of <= 2 joints, FD-of-Jacobians Hessians, knots [t0, t1) without equality rows: one thread per evaluation point
  // (linearize_pt.cuh); false when this model size / mode has no such kernel
  static bool linearize_pt(const ModelDev& m, const BatchDev& bd, int t0, int t1, cudaStream_t s);
  static void backward(const BatchDev& bd, int threads, size_t smem, cudaStream_t s);
  // xs_c / us_c: scratch trajectories of the concurrent line-search trials, linesearch_slots() x the trajectory size
  // cost_seq_aug of the current trajectory (independent of the backward pass), then the concurrent line-search trials
  static void cost_old(const ModelDev& m, const BatchDev& bd, double* costs_old, cudaStream_t s);
  static void forward(const ModelDev& m, const BatchDev& bd, double* costs_old, double* xs_c, double* us_c, double* ls_res, int* ls_win,
                      const int* wide, int* wide_next, cudaStream_t s);
  static int linesearch_slots();
  static void rollout(const ModelDev& m, const BatchDev& bd, cudaStream_t s);
  static void cost(const ModelDev& m, const BatchDev& bd, double* costs, cudaStream_t s);
  static void al_update(const BatchDev& bd, double thr, cudaStream_t s);
  static void optimality(const BatchDev& bd, cudaStream_t s);
  // the Model concept's own functions for A.n independent samples (model_eval.cuh)
  static void model_eval(const ModelDev& m, const ModelEvalArgs& A, cudaStream_t s);
};

}  // namespace ddpb
