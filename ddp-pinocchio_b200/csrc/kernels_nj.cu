// kernels_nj.cu -- the hot-path kernels for one model size; compile with -DDDPB_NJ=<0..7>.
#ifndef DDPB_NJ
#error "compile with -DDDPB_NJ=<n>"
#endif
#include "backward.cuh"
#include "forward.cuh"
#include "launch.cuh"
#include "linearize.cuh"

namespace ddpb {

constexpr int KNJ = DDPB_NJ;

#if DDPB_NJ > 0
template <int NJ> __global__ void __launch_bounds__(256) linearize_kernel(const __grid_constant__ ModelDev m, const __grid_constant__ BatchDev bd) {
  extern __shared__ __align__(16) double sm[];
  const int b = blockIdx.x % bd.B, t = blockIdx.x / bd.B;
  if (bd.scal[(size_t)S_STATUS * bd.B + b] != 0.0) return;
  linearize_knot<NJ>(m, bd, b, t, sm, DevExec{});
}
template <int NJ> __global__ void __launch_bounds__(256) backward_kernel(const __grid_constant__ BatchDev bd) {
  extern __shared__ __align__(16) double sm[];
  const int b = blockIdx.x;
  if (bd.scal[(size_t)S_STATUS * bd.B + b] != 0.0) return;
  backward_problem<NJ>(bd, b, sm, DevExec{});
}
#else
template <int NJ> __global__ void linearize_kernel(const __grid_constant__ ModelDev m, const __grid_constant__ BatchDev bd) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)bd.B * bd.T) return;
  const int b = (int)(i % bd.B), t = (int)(i / bd.B);
  if (bd.scal[(size_t)S_STATUS * bd.B + b] != 0.0) return;
  linearize_pendulum(m, bd, b, t);
}
#endif

template <int NJ> __global__ void forward_kernel(const __grid_constant__ ModelDev m, const __grid_constant__ BatchDev bd, double* costs_old) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= bd.B) return;
  forward_problem<NJ>(m, bd, b, costs_old);
}
template <int NJ> __global__ void rollout_kernel(const __grid_constant__ ModelDev m, const __grid_constant__ BatchDev bd) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= bd.B) return;
  rollout_problem<NJ>(m, bd, b);
}
template <int NJ> __global__ void cost_kernel(const __grid_constant__ ModelDev m, const __grid_constant__ BatchDev bd, double* costs) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= bd.B) return;
  const double plain = cost_seq_problem<NJ>(m, bd, b, costs, bd.scal[(size_t)S_MU * bd.B + b]);
  bd.scal[(size_t)S_COST * bd.B + b] = plain;
}
template <int NJ> __global__ void al_update_kernel(const __grid_constant__ BatchDev bd, double thr) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= bd.B) return;
  al_update_problem<NJ>(bd, b, thr);
}
template <int NJ> __global__ void optimality_kernel(const __grid_constant__ BatchDev bd) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= bd.B) return;
  bd.scal[(size_t)S_OPT_OBJ * bd.B + b] = optimality_obj_problem<NJ>(bd, b, bd.scal[(size_t)S_MU * bd.B + b]);
  bd.scal[(size_t)S_OPT_CONSTR * bd.B + b] = optimality_constr_problem<NJ>(bd, b);
}

constexpr int TPP = 32;  // threads per block of the thread-per-problem kernels

template <> size_t Launch<KNJ>::lin_smem_bytes(int hess_mode, int ne_max) {
#if DDPB_NJ > 0
  LinPlan<KNJ> pl;
  pl.init(hess_mode, ne_max, ne_max);
  return (size_t)pl.total * sizeof(double);
#else
  (void)hess_mode; (void)ne_max;
  return 0;
#endif
}
template <> size_t Launch<KNJ>::bwd_smem_bytes() {
#if DDPB_NJ > 0
  return (size_t)BwdSmem<KNJ>::SIZE * sizeof(double);
#else
  return 0;
#endif
}
template <> cudaError_t Launch<KNJ>::prepare(size_t lin_smem) {
#if DDPB_NJ > 0
  return cudaFuncSetAttribute(linearize_kernel<KNJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lin_smem);
#else
  (void)lin_smem;
  return cudaSuccess;
#endif
}
template <> void Launch<KNJ>::linearize(const ModelDev& m, const BatchDev& bd, int threads, size_t smem, cudaStream_t s) {
#if DDPB_NJ > 0
  linearize_kernel<KNJ><<<(unsigned)((size_t)bd.B * bd.T), threads, smem, s>>>(m, bd);
#else
  (void)threads; (void)smem;
  const size_t n = (size_t)bd.B * bd.T;
  linearize_kernel<KNJ><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(m, bd);
#endif
}
template <> void Launch<KNJ>::backward(const BatchDev& bd, int threads, size_t smem, cudaStream_t s) {
#if DDPB_NJ > 0
  backward_kernel<KNJ><<<bd.B, threads, smem, s>>>(bd);
#else
  (void)bd; (void)threads; (void)smem; (void)s;
#endif
}
template <> void Launch<KNJ>::forward(const ModelDev& m, const BatchDev& bd, double* costs_old, cudaStream_t s) {
  forward_kernel<KNJ><<<(bd.B + TPP - 1) / TPP, TPP, 0, s>>>(m, bd, costs_old);
}
template <> void Launch<KNJ>::rollout(const ModelDev& m, const BatchDev& bd, cudaStream_t s) {
  rollout_kernel<KNJ><<<(bd.B + TPP - 1) / TPP, TPP, 0, s>>>(m, bd);
}
template <> void Launch<KNJ>::cost(const ModelDev& m, const BatchDev& bd, double* costs, cudaStream_t s) {
  cost_kernel<KNJ><<<(bd.B + TPP - 1) / TPP, TPP, 0, s>>>(m, bd, costs);
}
template <> void Launch<KNJ>::al_update(const BatchDev& bd, double thr, cudaStream_t s) {
  al_update_kernel<KNJ><<<(bd.B + TPP - 1) / TPP, TPP, 0, s>>>(bd, thr);
}
template <> void Launch<KNJ>::optimality(const BatchDev& bd, cudaStream_t s) {
  optimality_kernel<KNJ><<<(bd.B + TPP - 1) / TPP, TPP, 0, s>>>(bd);
}

}  // namespace ddpb
