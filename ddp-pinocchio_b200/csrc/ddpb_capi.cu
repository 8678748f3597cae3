// ddpb_capi.cu -- kernels and the extern "C" layer declared in include/ddpb.h.
// Built for sm_100a only (see __graft_entry__.build()); there is no CPU fallback: every entry point
// that needs the device returns DDPB_ERR_CUDA when none is usable.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ddpb.h"
#include "forward.cuh"
#include "launch.cuh"
#include "model_host.hpp"
#include "urdf_host.hpp"

using namespace ddpb;

static_assert((int)NSCALARS == (int)DDPB_NSCALARS, "scalar table mismatch");
static_assert(sizeof(ModelDev) < 3500, "model must fit the kernel parameter space");

// -------------------------------------------------------------------------------------------------
// kernels
// -------------------------------------------------------------------------------------------------
__global__ void post_forward_kernel(const __grid_constant__ BatchDev bd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= bd.nb) return;
  post_forward_problem(bd, bd.b0 + i);
}
// 16 independent FMA chains per thread: saturates the FP64 pipe
__global__ void fp64_peak_kernel(double* out, int iters) {
  double a[16];
  const double x = 1.0 + 1e-9 * threadIdx.x, y = 1e-9 * blockIdx.x;
#pragma unroll
  for (int k = 0; k < 16; ++k) a[k] = k * 0.5;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = fma(a[k], x, y);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += a[k];
  if (s == 12345.678) out[0] = s;  // keep the chains alive
}
__global__ void scalar_copy_kernel(double* scal, int B, int b0, int nb, int dst, int src) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nb) scal[(size_t)dst * B + b0 + i] = scal[(size_t)src * B + b0 + i];
}
__global__ void fill_kernel(double* p, size_t n, double v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
// dst[c * R + r] = src[r * C + c]   (host layout [R = problems][C] <-> SoA-by-batch [C][R])
__global__ void transpose_kernel(double* __restrict__ dst, const double* __restrict__ src, int R, int C) {
  __shared__ double tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < R && c < C) tile[i][threadIdx.x] = src[(size_t)r * C + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < C) dst[(size_t)c * R + r] = tile[threadIdx.x][i];
  }
}
// copy x of the current trajectory into an origin sequence: dst[T][nx][B] <- xs[T+1][nx][B] (first T knots)
__global__ void copy_kernel(double* dst, const double* src, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
// packed host sequence of one derivative field (the reference's dense layout) <-> stored slab records.  per_row:
// doubles per equality row (eq fields, ragged) or 0 for fields with a fixed per-knot size `per_knot`.  Entries the slab
// does not store (batch.cuh: constants of the reference's cost / Euler scheme) are produced from their constants on
// the way out and CHECKED on the way in: *mismatch counts the dense entries that differ from their structural value.
__global__ void slab_field_kernel(const __grid_constant__ BatchDev bd, double* packed, size_t field_total, int field, int per_knot,
                                  int per_row, int to_slab, unsigned long long* mismatch) {
  const int n = bd.n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < field_total * bd.B; i += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / field_total);
    const size_t pos = i % field_total;
    int t, e;
    if (per_row == 0) {
      t = (int)(pos / per_knot); e = (int)(pos % per_knot);
    } else {
      int lo = 0, hi = bd.T;  // largest t with eq_off[t] * per_row <= pos
      while (hi - lo > 1) {
        const int mid = (lo + hi) / 2;
        if ((size_t)bd.eq_off[mid] * per_row <= pos) lo = mid; else hi = mid;
      }
      t = lo; e = (int)(pos - (size_t)bd.eq_off[t] * per_row);
    }
    const RecLayout L = rec_layout(n, bd.ne[t], bd.hess_mode);
    double sval;
    const int off = stored_offset(L, n, field, e, bd.dt, bd.c, &sval);
    if (off >= 0) {
      const RecView r = rec_view(bd, t, b, L.size);
      if (to_slab) r[off] = packed[i]; else packed[i] = r[off];
    } else if (to_slab) {
      if (!(packed[i] == sval)) atomicAdd(mismatch, 1ull);
    } else {
      packed[i] = sval;
    }
  }
}

// -------------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) return fail(DDPB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

struct ddpb_model {
  ModelDev dev;
  int nq;
  urdf::Model src;  // the description the model was built from (arrays as in ddpb_model_desc; names for URDF models)
};

struct PhaseEv { cudaEvent_t a, b; int phase; };
struct SolveGraph {
  int iters, nsub; double thr, mu, reg, w, n; double* xs; cudaStream_t stream;
  cudaGraphExec_t exec = nullptr; unsigned long long launches = 0;
};

struct ddpb_batch {
  ModelDev model;
  BatchDev bd;
  int device = 0;
  int mt = 0;        // model tag of DISPATCH: 0 = pendulum, else number of joints (+ 8 for chains with prismatic joints)
  int nj = 1;        // joints (1 for the pendulum)
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::vector<int> ne, eq_off;
  std::vector<long long> slab_off;
  std::vector<void*> allocs;
  double* costs_old = nullptr;
  double *ls_xs = nullptr, *ls_us = nullptr;  // scratch trajectories of the concurrent line-search trials
  double* ls_res = nullptr;                   // [34][3][B] (sum, merit, plain) of every trial
  int* ls_win = nullptr;                      // [B] line-search state per problem
  int* ls_wide[2] = {nullptr, nullptr};       // [B/32] per block of 32 problems: run all trials in the first wave (ping-pong between passes)
  int ls_pass = 0;                            // parity of the line-search pass being enqueued
  double* stage = nullptr;
  size_t stage_bytes = 0;
  double* h_stage = nullptr;  // pinned bounce buffer for small scalar traffic
  double* snap = nullptr;     // device-resident snapshot: xs | us | m_origin | m_val | m_jac
  unsigned long long launches = 0;
  bool timing = false;
  std::vector<PhaseEv> events;
  float phase_ms[4] = {0, 0, 0, 0};
  size_t lin_smem = 0;
  int lin_threads = 128, bwd_threads = 128, lin_kpb = 1, lin_kpb_pref = 1;  // knots per CTA: for the most constrained knot / preferred
  bool lin_pt = true;  // use the thread-per-point linearization where it exists (DDPB_LIN_PT=0: development switch)
  // ddpb_solve runs the batch as `nsub` sub-batches, each on its own stream: the latency-bound sweeps (backward pass,
  // rollouts) of one sub-batch overlap the throughput-bound linearization of another, and a problem that needs
  // several backward restarts only delays its own sub-batch
  int nsub = 1;
  std::vector<cudaStream_t> sub_streams;     // high priority: backward pass, rollouts, outer-loop sweeps (latency-bound)
  std::vector<cudaStream_t> sub_streams_lo;  // low priority: linearization (throughput-bound, fills whatever is free)
  std::vector<cudaEvent_t> sub_done, sub_lin, sub_hi;
  cudaEvent_t fork_ev = nullptr;
  std::vector<SolveGraph> graphs;  // captured solves (ddpb_solve_async)
  size_t bwd_smem = 0;
};

template <class T> static int dev_alloc(ddpb_batch* h, T** p, size_t count) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T));
  if (e != cudaSuccess) return fail(DDPB_ERR_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  h->allocs.push_back(q);
  *p = (T*)q;
  return DDPB_OK;
}
static int ensure_stage(ddpb_batch* h, size_t bytes) {
  if (bytes <= h->stage_bytes) return DDPB_OK;
  if (h->stage) cudaFree(h->stage);
  h->stage = nullptr; h->stage_bytes = 0;
  cudaError_t e = cudaMalloc((void**)&h->stage, bytes);
  if (e != cudaSuccess) return fail(DDPB_ERR_NOMEM, std::string("cudaMalloc(stage): ") + cudaGetErrorString(e));
  h->stage_bytes = bytes;
  return DDPB_OK;
}

struct PhaseScope {
  ddpb_batch* h; int idx = -1; cudaStream_t s;
  PhaseScope(ddpb_batch* hh, int phase, cudaStream_t ss) : h(hh), s(ss) {
    if (!h->timing) return;
    PhaseEv ev; ev.phase = phase;
    cudaEventCreate(&ev.a); cudaEventCreate(&ev.b);
    cudaEventRecord(ev.a, s);
    h->events.push_back(ev); idx = (int)h->events.size() - 1;
  }
  ~PhaseScope() { if (idx >= 0) cudaEventRecord(h->events[idx].b, s); }
};
// the launch helpers take the (sub-)batch descriptor and the stream explicitly; LaunchCtx{h} = whole batch on h->stream
struct LaunchCtx {
  ddpb_batch* h; BatchDev bd; cudaStream_t s;
  explicit LaunchCtx(ddpb_batch* hh) : h(hh), bd(hh->bd), s(hh->stream) {}
  LaunchCtx(ddpb_batch* hh, int b0, int nb, cudaStream_t ss) : h(hh), bd(hh->bd), s(ss) { bd.b0 = b0; bd.nb = nb; }
};

// run STMT with L = Launch<k, PR> for the model tag: 0 = pendulum, 1..7 = joints of an all-revolute chain, 8 + (1..7) = joints
// of a chain with prismatic joints (the device code compiled with the joint-type tests, variant.cuh)
#define DISPATCH(tag, STMT)                                     \
  switch (tag) {                                                \
    case 9: { using L = Launch<1, true>; STMT; } break;         \
    case 10: { using L = Launch<2, true>; STMT; } break;        \
    case 11: { using L = Launch<3, true>; STMT; } break;        \
    case 12: { using L = Launch<4, true>; STMT; } break;        \
    case 13: { using L = Launch<5, true>; STMT; } break;        \
    case 14: { using L = Launch<6, true>; STMT; } break;        \
    case 15: { using L = Launch<7, true>; STMT; } break;        \
    case 0: { using L = Launch<0>; STMT; } break;               \
    case 1: { using L = Launch<1>; STMT; } break;               \
    case 2: { using L = Launch<2>; STMT; } break;               \
    case 3: { using L = Launch<3>; STMT; } break;               \
    case 4: { using L = Launch<4>; STMT; } break;               \
    case 5: { using L = Launch<5>; STMT; } break;               \
    case 6: { using L = Launch<6>; STMT; } break;               \
    case 7: { using L = Launch<7>; STMT; } break;               \
    default: return fail(DDPB_ERR_UNSUPPORTED, "n_joints must be 1..7"); \
  }

static int check_launch(ddpb_batch* h, const char* what) {
  ++h->launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(DDPB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  return DDPB_OK;
}

static int launch_linearize(const LaunchCtx& c) {
  ddpb_batch* h = c.h;
  PhaseScope ps(h, 0, c.s);
  // one launch per maximal run of knots with the same number of equality rows: shared memory (and with it the
  // number of resident CTAs) is sized for the run, not for the most constrained knot of the horizon
  const int T = h->bd.T;
  int t0 = 0;
  while (t0 < T) {
    int t1 = t0 + 1;
    while (t1 < T && h->ne[t1] == h->ne[t0]) ++t1;
    // knots without equality rows of a short chain go to the thread-per-point kernel where one exists
    bool done = false;
    if (h->ne[t0] == 0 && h->lin_pt) DISPATCH(h->mt, done = L::linearize_pt(h->model, c.bd, t0, t1, c.s));
    if (!done) {
      // knots per CTA for THIS run of knots: its workspace depends on its row count (a constrained knot of the horizon must
      // not shrink the groups of all the others)
      size_t knot_smem = 0;
      DISPATCH(h->mt, knot_smem = L::lin_smem_bytes(c.bd.hess_mode, h->ne[t0], c.bd.eq_kind));
      int kpb = h->lin_kpb_pref;
      while (kpb > 1 && knot_smem * kpb > h->lin_smem) --kpb;
      DISPATCH(h->mt, L::linearize(h->model, c.bd, h->lin_threads, kpb, t0, t1, h->ne[t0], c.s));
    }
    int rc = check_launch(h, "linearize");
    if (rc) return rc;
    t0 = t1;
  }
  return DDPB_OK;
}
static int launch_backward(const LaunchCtx& c) {
  ddpb_batch* h = c.h;
  PhaseScope ps(h, 1, c.s);
  DISPATCH(h->nj, L::backward(c.bd, h->bwd_threads, h->bwd_smem, c.s));
  h->launches += 1;  // the origin copy that precedes the recursion
  return check_launch(h, "backward");
}
static int launch_cost_old(const LaunchCtx& c) {
  ddpb_batch* h = c.h;
  PhaseScope ps(h, 2, c.s);
  DISPATCH(h->mt, L::cost_old(h->model, c.bd, h->costs_old, c.s));
  return check_launch(h, "cost_old");
}
// with_cost_old = false: cost_seq_aug of the current trajectory was already launched (on another stream) and joined
static int launch_forward(const LaunchCtx& c, bool with_cost_old = true) {
  ddpb_batch* h = c.h;
  if (with_cost_old) { int rc = launch_cost_old(c); if (rc) return rc; }
  PhaseScope ps(h, 2, c.s);
  DISPATCH(h->mt, L::forward(h->model, c.bd, h->costs_old, h->ls_xs, h->ls_us, h->ls_res, h->ls_win, h->ls_wide[h->ls_pass & 1],
                             h->ls_wide[(h->ls_pass + 1) & 1], c.s));
  h->launches += 3;  // two waves of trials, each with its selection kernel
  return check_launch(h, "forward");
}
static int launch_al_update(const LaunchCtx& c, double thr) {
  ddpb_batch* h = c.h;
  PhaseScope ps(h, 3, c.s);
  DISPATCH(h->mt, L::al_update(c.bd, thr, c.s));
  h->launches += 3;  // origin update, decision sweep, multiplier update, second sweep
  return check_launch(h, "al_update");
}
static int launch_post_forward(const LaunchCtx& c) {
  ddpb_batch* h = c.h;
  PhaseScope ps(h, 3, c.s);
  const int th = 128;
  post_forward_kernel<<<(c.bd.nb + th - 1) / th, th, 0, c.s>>>(c.bd);
  return check_launch(h, "post_forward");
}
static int scalar_copy(const LaunchCtx& c, int dst, int src) {
  scalar_copy_kernel<<<(c.bd.nb + 127) / 128, 128, 0, c.s>>>(c.bd.scal, c.bd.B, c.bd.b0, c.bd.nb, dst, src);
  return check_launch(c.h, "scalar_copy");
}
static int fill(ddpb_batch* h, double* p, size_t n, double v) {
  if (n == 0) return DDPB_OK;
  fill_kernel<<<(unsigned)std::min<size_t>((n + 255) / 256, 4096), 256, 0, h->stream>>>(p, n, v);
  return check_launch(h, "fill");
}
static int dcopy(ddpb_batch* h, double* dst, const double* src, size_t n) {
  if (n == 0) return DDPB_OK;
  copy_kernel<<<(unsigned)std::min<size_t>((n + 255) / 256, 4096), 256, 0, h->stream>>>(dst, src, n);
  return check_launch(h, "copy");
}
// host [B][C] -> device [C][B]
static int upload_T(ddpb_batch* h, double* dev, const double* host, size_t C) {
  if (C == 0) return DDPB_OK;
  const int B = h->bd.B;
  int rc = ensure_stage(h, (size_t)B * C * sizeof(double));
  if (rc) return rc;
  CU(cudaMemcpyAsync(h->stage, host, (size_t)B * C * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  dim3 grid((unsigned)((C + 31) / 32), (unsigned)((B + 31) / 32)), block(32, 8);
  transpose_kernel<<<grid, block, 0, h->stream>>>(dev, h->stage, B, (int)C);
  return check_launch(h, "transpose_in");
}
// device [C][B] -> host [B][C]
static int download_T(ddpb_batch* h, double* host, const double* dev, size_t C) {
  if (C == 0) return DDPB_OK;
  const int B = h->bd.B;
  int rc = ensure_stage(h, (size_t)B * C * sizeof(double));
  if (rc) return rc;
  dim3 grid((unsigned)((B + 31) / 32), (unsigned)((C + 31) / 32)), block(32, 8);
  transpose_kernel<<<grid, block, 0, h->stream>>>(h->stage, dev, (int)C, B);
  rc = check_launch(h, "transpose_out");
  if (rc) return rc;
  CU(cudaMemcpyAsync(host, h->stage, (size_t)B * C * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return DDPB_OK;
}

// everything of ddpb_batch_create that can fail after the handle exists; the caller destroys the handle on failure
static int batch_init(ddpb_batch* h, const ddpb_model* m, int B, int T, const int* ne_per_knot, const ddpb_params* p, int device, const cudaDeviceProp& prop) {
  const int n = m->nq;
  h->model = m->dev;
  h->device = device;
  h->mt = m->dev.kind == DDPB_MODEL_PENDULUM ? 0 : n + (m->dev.any_prismatic ? 8 : 0);
  h->nj = n;
  BatchDev& bd = h->bd;
  std::memset(&bd, 0, sizeof(bd));
  bd.B = B; bd.T = T; bd.n = n; bd.nx = 2 * n; bd.nu = n;
  bd.b0 = 0; bd.nb = B;
  bd.hess_mode = p->hess_mode; bd.eq_kind = p->eq_kind; bd.frame_id = p->frame_id; bd.jac_mode = p->frame_jac_mode;
  bd.dt = p->dt; bd.c = p->c;
  bd.max_restarts = p->max_backward_restarts < 0 ? 0 : p->max_backward_restarts;
  h->ne.assign(T, 0); h->eq_off.assign(T + 1, 0); h->slab_off.assign(T + 1, 0);
  const int row_cap = p->eq_kind == DDPB_EQ_SPATIAL ? 3 : n;
  for (int t = 0; t < T; ++t) {
    const int ne = (p->eq_kind == DDPB_EQ_NONE || !ne_per_knot) ? 0 : ne_per_knot[t];
    if (ne < 0 || ne > row_cap || (p->eq_kind == DDPB_EQ_SPATIAL && ne != 0 && ne != 3)) return fail(DDPB_ERR_INVALID, "bad equality row count");
    h->ne[t] = ne;
    h->eq_off[t + 1] = h->eq_off[t] + ne;
    bd.ne_max = std::max(bd.ne_max, ne);
    h->slab_off[t + 1] = h->slab_off[t] + rec_layout(n, ne, p->hess_mode).size;
  }
  bd.ne_total = h->eq_off[T];
  CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  h->own_stream = true;
  const size_t nx = bd.nx, nu = bd.nu, nd = bd.nx, net = bd.ne_total;
  int rc = 0;
  int *d_ne = nullptr, *d_eo = nullptr;
  long long* d_so = nullptr;
#define A(ptr, count) if (!rc) rc = dev_alloc(h, &(ptr), (size_t)(count))
  A(bd.xs, (T + 1) * nx * B); A(bd.us, T * nu * B); A(bd.xs_new, (T + 1) * nx * B); A(bd.us_new, T * nu * B);
  A(bd.targets, net * B); A(bd.m_origin, T * nx * B); A(bd.m_val, net * B); A(bd.m_jac, net * nd * B);
  A(bd.fb_origin, T * nx * B); A(bd.fb_k, T * nu * B); A(bd.fb_K, T * nu * nd * B);
  A(bd.slab, (size_t)h->slab_off[T] * B); A(bd.scal, (size_t)NSCALARS * B + 2); A(h->costs_old, (size_t)(T + 1) * B);
  A(d_ne, T); A(d_eo, T + 1); A(d_so, T + 1);
  {
    int slots = 1;
    DISPATCH(h->mt, { slots = L::linesearch_slots(); });
    A(h->ls_xs, (size_t)slots * (T + 1) * nx * B); A(h->ls_us, (size_t)slots * T * nu * B);
    A(h->ls_res, (size_t)34 * 3 * B); A(h->ls_win, (size_t)B);
    A(h->ls_wide[0], (size_t)(B + 31) / 32 + 1); A(h->ls_wide[1], (size_t)(B + 31) / 32 + 1);
  }
#undef A
  if (rc) return rc;
  bd.ne = d_ne; bd.eq_off = d_eo; bd.slab_off = d_so;
  CU(cudaMemcpy(d_ne, h->ne.data(), T * sizeof(int), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(d_eo, h->eq_off.data(), (T + 1) * sizeof(int), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(d_so, h->slab_off.data(), (T + 1) * sizeof(long long), cudaMemcpyHostToDevice));
  CU(cudaMemsetAsync(bd.scal, 0, (size_t)NSCALARS * B * sizeof(double), h->stream));
  CU(cudaMemsetAsync(bd.slab, 0, (size_t)h->slab_off[T] * B * sizeof(double), h->stream));
  CU(cudaMemsetAsync(bd.m_val, 0, std::max<size_t>(net * B, 1) * sizeof(double), h->stream));
  CU(cudaMemsetAsync(bd.m_jac, 0, std::max<size_t>(net * nd * B, 1) * sizeof(double), h->stream));
  CU(cudaMemsetAsync(bd.targets, 0, std::max<size_t>(net * B, 1) * sizeof(double), h->stream));
  CU(cudaMemsetAsync(bd.fb_k, 0, T * nu * B * sizeof(double), h->stream));
  CU(cudaMemsetAsync(h->ls_wide[0], 0, ((size_t)(B + 31) / 32 + 1) * sizeof(int), h->stream));
  CU(cudaMemsetAsync(h->ls_wide[1], 0, ((size_t)(B + 31) / 32 + 1) * sizeof(int), h->stream));
  CU(cudaMemsetAsync(bd.fb_K, 0, T * nu * nd * B * sizeof(double), h->stream));
  CU(cudaHostAlloc((void**)&h->h_stage, (size_t)NSCALARS * B * sizeof(double), cudaHostAllocDefault));
  // launch geometry
  const bool fdv = p->hess_mode == DDPB_HESS_FD_VALUES;
  // measured on B200 (tools/lin_geom_sweep.sh): small CTAs keep more knots / problems resident per SM.
  // Value-based Hessians: one knot per CTA; from 5 joints on 128 threads, so that two CTAs
  // (255 registers per thread) share an SM and the narrow phases of one knot overlap the wide phases of another.
  h->lin_threads = fdv ? (n <= 2 ? 32 : 128) : (n >= 5 ? 256 : 128);
  h->bwd_threads = 32;  // informational: the backward kernel fixes its own CTA size (BwdGeom in backward.cuh)
  // development knobs (launch geometry sweeps); not part of the interface
  if (const char* e = getenv("DDPB_LIN_PT")) h->lin_pt = atoi(e) != 0;
  if (const char* e = getenv("DDPB_LIN_THREADS")) { const int v = atoi(e); if (v >= 32 && v <= 256 && v % 32 == 0) h->lin_threads = v; }

  // problems per CTA of the linearization: the serial sub-phases of a knot occupy a handful of lanes, so a CTA owns as
  // many knots as shared memory allows while two (256-thread) or three (128-thread) CTAs still fit an SM
  size_t knot_smem = 0, knot_smem_max = 0;  // workspace of an unconstrained knot / of the knot with the most rows
  DISPATCH(h->mt, { knot_smem = L::lin_smem_bytes(p->hess_mode, 0, p->eq_kind); knot_smem_max = L::lin_smem_bytes(p->hess_mode, bd.ne_max, p->eq_kind); });
  h->lin_kpb = fdv ? 1 : (n <= 1 ? 16 : (n == 2 ? 16 : 2));
  if (!fdv && n >= 5 && knot_smem > 0) h->lin_kpb = (int)std::min<size_t>(3, std::max<size_t>(1, ((size_t)113 << 10) / knot_smem));
  if (const char* e = getenv("DDPB_LIN_KPB")) { const int v = atoi(e); if (v >= 1 && v <= 32) h->lin_kpb = v; }
  h->lin_kpb_pref = h->lin_kpb;
  while (h->lin_kpb > 1 && knot_smem_max * h->lin_kpb > (size_t)prop.sharedMemPerBlockOptin) --h->lin_kpb;
  // dynamic shared memory the kernels may ask for: the larger of (most constrained knot x its group size) and
  // (unconstrained knot x preferred group size); every launch sizes its own request (launch_linearize)
  h->lin_smem = std::max(knot_smem_max * h->lin_kpb, std::min<size_t>(knot_smem * h->lin_kpb_pref, (size_t)prop.sharedMemPerBlockOptin));
  if (h->lin_smem > (size_t)prop.sharedMemPerBlockOptin) return fail(DDPB_ERR_UNSUPPORTED, "linearization workspace exceeds shared memory");
  DISPATCH(h->mt, { CU(L::prepare(h->lin_smem)); });
  DISPATCH(h->nj, { h->bwd_smem = L::bwd_smem_bytes(p->hess_mode); });
  // sub-batches of ddpb_solve: enough problems per sub-batch to keep whole warps / CTAs busy
  {
    int nsub = B >= 512 ? 4 : (B >= 128 ? 2 : 1);
    if (const char* e = getenv("DDPB_SUBBATCHES")) { const int v = atoi(e); if (v >= 1 && v <= 32) nsub = v; }
    const int rc2 = ddpb_set_subbatches(h, nsub);
    if (rc2) return rc2;
  }
  CU(cudaStreamSynchronize(h->stream));
  return DDPB_OK;
}

extern "C" {

const char* ddpb_last_error(void) { return g_err.c_str(); }

int ddpb_model_create(const ddpb_model_desc* d, ddpb_model** out) {
  if (!d || !out) return fail(DDPB_ERR_INVALID, "null argument");
  ddpb_model* m = new ddpb_model();
  std::string err;
  const int rc = model_from_desc(d, m->dev, err);
  if (rc) { delete m; return fail(rc, err); }
  m->nq = m->dev.n;
  if (d->kind == DDPB_MODEL_REVOLUTE_CHAIN) {
    const int n = d->n_joints, f = d->n_frames;
    urdf::Model& s = m->src;
    s.parent.assign(d->parent, d->parent + n);
    s.placement_rot.assign(d->placement_rot, d->placement_rot + 9 * n); s.placement_trans.assign(d->placement_trans, d->placement_trans + 3 * n);
    s.axis.assign(d->axis, d->axis + 3 * n); s.mass.assign(d->mass, d->mass + n); s.com.assign(d->com, d->com + 3 * n);
    s.inertia.assign(d->inertia, d->inertia + 6 * n);
    s.joint_type.assign(m->dev.jtype, m->dev.jtype + n);
    if (f > 0) { s.frame_parent.assign(d->frame_parent, d->frame_parent + f); s.frame_trans.assign(d->frame_trans, d->frame_trans + 3 * f); }
  }
  *out = m;
  return DDPB_OK;
}
int ddpb_model_create_from_urdf(const char* xml, unsigned long long length, ddpb_model** out) {
  if (!xml || !out) return fail(DDPB_ERR_INVALID, "null argument");
  urdf::Model src;
  std::string err;
  int rc = urdf::model_from_urdf(xml, (size_t)length, src, err);
  if (rc) return fail(rc, "URDF: " + err);
  const ddpb_model_desc d = src.desc();
  ddpb_model* m = new ddpb_model();
  rc = model_from_desc(&d, m->dev, err);
  if (rc) { delete m; return fail(rc, "URDF: " + err); }
  m->nq = m->dev.n;
  m->src = std::move(src);
  *out = m;
  return DDPB_OK;
}
int ddpb_model_dims(const ddpb_model* m, int* dims2) {
  if (!m || !dims2) return fail(DDPB_ERR_INVALID, "null argument");
  dims2[0] = m->dev.n; dims2[1] = m->dev.nframes;
  return DDPB_OK;
}
int ddpb_model_get(const ddpb_model* m, double* placement_rot, double* placement_trans, double* axis, double* mass, double* com,
                   double* inertia, int* frame_parent, double* frame_trans) {
  if (!m) return fail(DDPB_ERR_INVALID, "null model");
  const urdf::Model& s = m->src;
  auto put = [](auto* dst, const auto& v) { if (dst) std::copy(v.begin(), v.end(), dst); };
  put(placement_rot, s.placement_rot); put(placement_trans, s.placement_trans); put(axis, s.axis); put(mass, s.mass); put(com, s.com);
  put(inertia, s.inertia); put(frame_parent, s.frame_parent); put(frame_trans, s.frame_trans);
  return DDPB_OK;
}
int ddpb_model_joint_types(const ddpb_model* m, int* joint_type) {
  if (!m || !joint_type) return fail(DDPB_ERR_INVALID, "null argument");
  for (int i = 0; i < m->dev.n; ++i) joint_type[i] = m->dev.jtype[i];
  return DDPB_OK;
}
const char* ddpb_model_frame_name(const ddpb_model* m, int frame) {
  if (!m || frame < 0 || frame >= (int)m->src.frame_names.size()) return "";
  return m->src.frame_names[frame].c_str();
}
void ddpb_model_destroy(ddpb_model* m) { delete m; }

// ---- the Model concept's own functions for n independent samples (csrc/model_eval.cuh) ---------------------------------------
namespace {
struct DevBuf {  // device copy of a host array, freed on scope exit
  double* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t in(const double* h, size_t count) {
    cudaError_t e = cudaMalloc((void**)&p, std::max<size_t>(count, 1) * sizeof(double));
    return e != cudaSuccess ? e : cudaMemcpy(p, h, count * sizeof(double), cudaMemcpyHostToDevice);
  }
  cudaError_t out(size_t count) { return cudaMalloc((void**)&p, std::max<size_t>(count, 1) * sizeof(double)); }
};
int model_eval_host(const ddpb_model* m, int device, int n, int frame, int jac_mode, const double* q, const double* v, const double* tau,
                    const double* acc, double* a, double* a_dq, double* a_dv, double* a_dtau, double* tau_out, double* p3, double* J) {
  if (!m || !q) return fail(DDPB_ERR_INVALID, "null argument");
  if (n < 0) return fail(DDPB_ERR_INVALID, "negative sample count");
  const bool pend = m->dev.kind == DDPB_MODEL_PENDULUM;
  const bool fwd = a || a_dq || a_dv;
  if (fwd && (!v || !tau)) return fail(DDPB_ERR_INVALID, "forward dynamics needs v and tau");
  if (tau_out && (!v || !acc)) return fail(DDPB_ERR_INVALID, "inverse dynamics needs v and a");
  if (pend && (tau_out || p3 || J)) return fail(DDPB_ERR_UNSUPPORTED, "the pendulum model has no inverse dynamics and no frames (pendulum_model.hpp)");
  if ((p3 || J) && (frame < 0 || frame >= m->dev.nframes)) return fail(DDPB_ERR_INVALID, "frame index out of range");
  if (J && jac_mode != DDPB_FJ_WORLD && jac_mode != DDPB_FJ_LOCAL_WORLD_ALIGNED) return fail(DDPB_ERR_INVALID, "unknown frame Jacobian convention");
  if (n == 0) return DDPB_OK;
  CU(cudaSetDevice(device));
  const size_t nq = (size_t)m->nq, N = (size_t)n;
  DevBuf dq, dv, dtau, dacc, oa, oadq, oadv, oadtau, otau, op3, oJ;
  CU(dq.in(q, N * nq));
  if (v) CU(dv.in(v, N * nq));
  if (tau) CU(dtau.in(tau, N * nq));
  if (acc) CU(dacc.in(acc, N * nq));
  if (a) CU(oa.out(N * nq));
  if (a_dq) CU(oadq.out(N * nq * nq));
  if (a_dv) CU(oadv.out(N * nq * nq));
  if (a_dtau) CU(oadtau.out(N * nq * nq));
  if (tau_out) CU(otau.out(N * nq));
  if (p3) CU(op3.out(N * 3));
  if (J) CU(oJ.out(N * 3 * nq));
  ModelEvalArgs A{};
  A.n = n; A.frame = frame; A.jac_mode = jac_mode;
  A.q = dq.p; A.v = dv.p; A.tau = dtau.p; A.acc = dacc.p;
  A.a = oa.p; A.a_dq = oadq.p; A.a_dv = oadv.p; A.a_dtau = oadtau.p; A.tau_out = otau.p; A.p3 = op3.p; A.J = oJ.p;
  const int tag = pend ? 0 : m->dev.n + (m->dev.any_prismatic ? 8 : 0);
  DISPATCH(tag, L::model_eval(m->dev, A, 0));
  CU(cudaGetLastError());
  auto back = [&](double* h, const DevBuf& d, size_t count) { return h ? cudaMemcpy(h, d.p, count * sizeof(double), cudaMemcpyDeviceToHost) : cudaSuccess; };
  CU(back(a, oa, N * nq)); CU(back(a_dq, oadq, N * nq * nq)); CU(back(a_dv, oadv, N * nq * nq)); CU(back(a_dtau, oadtau, N * nq * nq));
  CU(back(tau_out, otau, N * nq)); CU(back(p3, op3, N * 3)); CU(back(J, oJ, N * 3 * nq));
  return DDPB_OK;
}
}  // namespace
int ddpb_model_dynamics_aba(const ddpb_model* m, int device, int n, const double* q, const double* v, const double* tau, double* a) {
  if (!a) return fail(DDPB_ERR_INVALID, "null output");
  return model_eval_host(m, device, n, 0, 0, q, v, tau, nullptr, a, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
}
int ddpb_model_d_dynamics_aba(const ddpb_model* m, int device, int n, const double* q, const double* v, const double* tau, double* a_dq,
                              double* a_dv, double* a_dtau) {
  if (!a_dq && !a_dv && !a_dtau) return fail(DDPB_ERR_INVALID, "null outputs");
  return model_eval_host(m, device, n, 0, 0, q, v, tau, nullptr, nullptr, a_dq, a_dv, a_dtau, nullptr, nullptr, nullptr);
}
int ddpb_model_inverse_dynamics_rnea(const ddpb_model* m, int device, int n, const double* q, const double* v, const double* a, double* tau) {
  if (!tau) return fail(DDPB_ERR_INVALID, "null output");
  return model_eval_host(m, device, n, 0, 0, q, v, nullptr, a, nullptr, nullptr, nullptr, nullptr, tau, nullptr, nullptr);
}
int ddpb_model_frame_coordinates(const ddpb_model* m, int device, int n, int frame, const double* q, double* p) {
  if (!p) return fail(DDPB_ERR_INVALID, "null output");
  return model_eval_host(m, device, n, frame, DDPB_FJ_WORLD, q, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, p, nullptr);
}
int ddpb_model_d_frame_coordinates(const ddpb_model* m, int device, int n, int frame, int frame_jac_mode, const double* q, double* J) {
  if (!J) return fail(DDPB_ERR_INVALID, "null output");
  return model_eval_host(m, device, n, frame, frame_jac_mode, q, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, J);
}

int ddpb_batch_create(const ddpb_model* m, int B, int T, const int* ne_per_knot, const ddpb_params* p, int device, ddpb_batch** out) {
  if (!m || !p || !out || B < 1 || T < 1) return fail(DDPB_ERR_INVALID, "bad arguments");
  if (p->hess_mode != DDPB_HESS_FD_JACOBIANS && p->hess_mode != DDPB_HESS_FD_VALUES) return fail(DDPB_ERR_INVALID, "bad hess_mode");
  if (p->eq_kind < 0 || p->eq_kind > 2) return fail(DDPB_ERR_INVALID, "bad eq_kind");
  const int n = m->nq;
  if (p->eq_kind == DDPB_EQ_SPATIAL && (m->dev.kind == DDPB_MODEL_PENDULUM || p->frame_id < 0 || p->frame_id >= m->dev.nframes))
    return fail(DDPB_ERR_INVALID, "frame index must be in bounds");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(DDPB_ERR_CUDA, "no CUDA device available (there is no CPU fallback)"); }
  if (device < 0 || device >= ndev) return fail(DDPB_ERR_INVALID, "device index out of range");
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) return fail(DDPB_ERR_CUDA, "this library is built for sm_100a (B200) only");
  // every failure past this point goes through one cleanup path (batch_init never frees, the handle owns all it got)
  ddpb_batch* h = new ddpb_batch();
  const int rc = batch_init(h, m, B, T, ne_per_knot, p, device, prop);
  if (rc != DDPB_OK) { ddpb_batch_destroy(h); return rc; }
  *out = h;
  return DDPB_OK;
}

void ddpb_batch_destroy(ddpb_batch* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (void* p : h->allocs) cudaFree(p);
  if (h->stage) cudaFree(h->stage);
  if (h->h_stage) cudaFreeHost(h->h_stage);
  for (auto& e : h->events) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
  for (auto& g : h->graphs) cudaGraphExecDestroy(g.exec);
  for (auto st : h->sub_streams) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
  for (auto st : h->sub_streams_lo) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
  for (auto ev : h->sub_done) cudaEventDestroy(ev);
  for (auto ev : h->sub_lin) cudaEventDestroy(ev);
  for (auto ev : h->sub_hi) cudaEventDestroy(ev);
  if (h->fork_ev) cudaEventDestroy(h->fork_ev);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int ddpb_batch_set_stream(ddpb_batch* h, void* s) {
  if (!h) return fail(DDPB_ERR_INVALID, "null batch");
  if (h->own_stream && h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
  h->stream = (cudaStream_t)s;
  h->own_stream = false;
  return DDPB_OK;
}
int ddpb_set_subbatches(ddpb_batch* h, int n) {
  if (!h || n < 1 || n > 32) return fail(DDPB_ERR_INVALID, "sub-batch count must be in 1..32");
  CU(cudaSetDevice(h->device));
  if ((int)h->sub_streams.size() < n && n > 1) {
    int prio_lo = 0, prio_hi = 0;
    CU(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    for (int i = (int)h->sub_streams.size(); i < n; ++i) {
      cudaStream_t hi = nullptr, lo = nullptr;
      cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
      CU(cudaStreamCreateWithPriority(&hi, cudaStreamNonBlocking, prio_hi));
      CU(cudaStreamCreateWithPriority(&lo, cudaStreamNonBlocking, prio_lo));
      CU(cudaEventCreateWithFlags(&e0, cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
      h->sub_streams.push_back(hi); h->sub_streams_lo.push_back(lo);
      h->sub_done.push_back(e0); h->sub_lin.push_back(e1); h->sub_hi.push_back(e2);
    }
    if (!h->fork_ev) CU(cudaEventCreateWithFlags(&h->fork_ev, cudaEventDisableTiming));
  }
  h->nsub = n;
  return DDPB_OK;
}
int ddpb_batch_sync(ddpb_batch* h) {
  if (!h) return fail(DDPB_ERR_INVALID, "null batch");
  CU(cudaStreamSynchronize(h->stream));
  return DDPB_OK;
}
int ddpb_host_alloc(void** ptr, unsigned long long bytes) {
  CU(cudaHostAlloc(ptr, (size_t)bytes, cudaHostAllocDefault));
  return DDPB_OK;
}
int ddpb_host_free(void* ptr) {
  CU(cudaFreeHost(ptr));
  return DDPB_OK;
}

int ddpb_batch_dims(const ddpb_batch* h, int* o) {
  if (!h || !o) return fail(DDPB_ERR_INVALID, "null argument");
  o[0] = h->bd.n; o[1] = h->bd.nx; o[2] = h->bd.nx; o[3] = h->bd.nu; o[4] = h->bd.ne_total; o[5] = h->bd.T; o[6] = h->bd.B;
  return DDPB_OK;
}
unsigned long long ddpb_launch_count(const ddpb_batch* h) { return h ? h->launches : 0; }
int ddpb_enable_phase_timing(ddpb_batch* h, int on) {
  if (!h) return fail(DDPB_ERR_INVALID, "null batch");
  h->timing = on != 0;
  return DDPB_OK;
}
int ddpb_phase_times(ddpb_batch* h, float* ms4) {
  if (!h || !ms4) return fail(DDPB_ERR_INVALID, "null argument");
  CU(cudaSetDevice(h->device));
  CU(cudaStreamSynchronize(h->stream));
  for (int k = 0; k < 4; ++k) ms4[k] = 0.f;
  for (auto& e : h->events) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e.a, e.b);
    ms4[e.phase] += ms;
    cudaEventDestroy(e.a); cudaEventDestroy(e.b);
  }
  h->events.clear();
  return DDPB_OK;
}

#define ENTER(h)                                              \
  if (!(h)) return fail(DDPB_ERR_INVALID, "null batch");      \
  CU(cudaSetDevice((h)->device));

int ddpb_set_trajectory(ddpb_batch* h, const double* xs, const double* us) {
  ENTER(h);
  const BatchDev& bd = h->bd;
  int rc = DDPB_OK;
  if (xs) rc = upload_T(h, bd.xs, xs, (size_t)(bd.T + 1) * bd.nx);
  if (!rc && xs) CU(cudaStreamSynchronize(h->stream));  // the staging buffer is reused below
  if (!rc && us) rc = upload_T(h, bd.us, us, (size_t)bd.T * bd.nu);
  if (rc) return rc;
  // the new-trajectory storage starts as a clone (ddp.hpp:752); x_0 of it is never rewritten
  rc = dcopy(h, bd.xs_new, bd.xs, (size_t)(bd.T + 1) * bd.nx * bd.B);
  if (!rc) rc = dcopy(h, bd.us_new, bd.us, (size_t)bd.T * bd.nu * bd.B);
  CU(cudaStreamSynchronize(h->stream));
  return rc;
}
int ddpb_get_trajectory(ddpb_batch* h, double* xs, double* us) {
  ENTER(h);
  int rc = DDPB_OK;
  if (xs) rc = download_T(h, xs, h->bd.xs, (size_t)(h->bd.T + 1) * h->bd.nx);
  if (!rc && us) rc = download_T(h, us, h->bd.us, (size_t)h->bd.T * h->bd.nu);
  return rc;
}
int ddpb_get_new_trajectory(ddpb_batch* h, double* xs, double* us) {
  ENTER(h);
  int rc = DDPB_OK;
  if (xs) rc = download_T(h, xs, h->bd.xs_new, (size_t)(h->bd.T + 1) * h->bd.nx);
  if (!rc && us) rc = download_T(h, us, h->bd.us_new, (size_t)h->bd.T * h->bd.nu);
  return rc;
}
int ddpb_make_trajectory(ddpb_batch* h, const double* x0, const double* us) {
  ENTER(h);
  const BatchDev& bd = h->bd;
  if (!x0) return fail(DDPB_ERR_INVALID, "x0 is required");
  int rc = upload_T(h, bd.xs, x0, bd.nx);  // knot 0 occupies the first nx rows of xs
  if (rc) return rc;
  CU(cudaStreamSynchronize(h->stream));
  if (us) { rc = upload_T(h, bd.us, us, (size_t)bd.T * bd.nu); if (rc) return rc; }
  else CU(cudaMemsetAsync(bd.us, 0, (size_t)bd.T * bd.nu * bd.B * sizeof(double), h->stream));
  DISPATCH(h->mt, L::rollout(h->model, bd, h->stream));
  rc = check_launch(h, "rollout");
  if (!rc) rc = dcopy(h, bd.xs_new, bd.xs, (size_t)(bd.T + 1) * bd.nx * bd.B);
  if (!rc) rc = dcopy(h, bd.us_new, bd.us, (size_t)bd.T * bd.nu * bd.B);
  CU(cudaStreamSynchronize(h->stream));
  return rc;
}
int ddpb_set_targets(ddpb_batch* h, const double* targets) {
  ENTER(h);
  if (h->bd.ne_total == 0) return DDPB_OK;
  if (!targets) return fail(DDPB_ERR_INVALID, "targets is null");
  int rc = upload_T(h, h->bd.targets, targets, h->bd.ne_total);
  CU(cudaStreamSynchronize(h->stream));
  return rc;
}
int ddpb_set_multipliers(ddpb_batch* h, const double* val, const double* jac) {
  ENTER(h);
  const BatchDev& bd = h->bd;
  const size_t net = bd.ne_total;
  int rc = DDPB_OK;
  if (net) {
    if (val) { rc = upload_T(h, bd.m_val, val, net); CU(cudaStreamSynchronize(h->stream)); }
    else CU(cudaMemsetAsync(bd.m_val, 0, net * bd.B * sizeof(double), h->stream));
    if (rc) return rc;
    if (jac) { rc = upload_T(h, bd.m_jac, jac, net * bd.nx); CU(cudaStreamSynchronize(h->stream)); }
    else CU(cudaMemsetAsync(bd.m_jac, 0, net * bd.nx * bd.B * sizeof(double), h->stream));
    if (rc) return rc;
  }
  rc = dcopy(h, bd.m_origin, bd.xs, (size_t)bd.T * bd.nx * bd.B);  // eq.origin() = xu.x()  (ddp.hpp:763)
  CU(cudaStreamSynchronize(h->stream));
  return rc;
}
int ddpb_get_multipliers(ddpb_batch* h, double* origin, double* val, double* jac) {
  ENTER(h);
  const BatchDev& bd = h->bd;
  int rc = DDPB_OK;
  if (origin) rc = download_T(h, origin, bd.m_origin, (size_t)bd.T * bd.nx);
  if (!rc && val && bd.ne_total) rc = download_T(h, val, bd.m_val, bd.ne_total);
  if (!rc && jac && bd.ne_total) rc = download_T(h, jac, bd.m_jac, (size_t)bd.ne_total * bd.nx);
  return rc;
}
int ddpb_set_feedback(ddpb_batch* h, const double* k, const double* K) {
  ENTER(h);
  const BatchDev& bd = h->bd;
  int rc = DDPB_OK;
  if (k) { rc = upload_T(h, bd.fb_k, k, (size_t)bd.T * bd.nu); CU(cudaStreamSynchronize(h->stream)); }
  if (!rc && K) { rc = upload_T(h, bd.fb_K, K, (size_t)bd.T * bd.nu * bd.nx); CU(cudaStreamSynchronize(h->stream)); }
  if (!rc) rc = dcopy(h, bd.fb_origin, bd.xs, (size_t)bd.T * bd.nx * bd.B);
  CU(cudaStreamSynchronize(h->stream));
  return rc;
}
int ddpb_get_feedback(ddpb_batch* h, double* origin, double* k, double* K) {
  ENTER(h);
  const BatchDev& bd = h->bd;
  int rc = DDPB_OK;
  if (origin) rc = download_T(h, origin, bd.fb_origin, (size_t)bd.T * bd.nx);
  if (!rc && k) rc = download_T(h, k, bd.fb_k, (size_t)bd.T * bd.nu);
  if (!rc && K) rc = download_T(h, K, bd.fb_K, (size_t)bd.T * bd.nu * bd.nx);
  return rc;
}

static int derivs_xfer(ddpb_batch* h, const ddpb_derivs* d, int to_slab) {
  const BatchDev& bd = h->bd;
  const int nd = bd.nx, nu = bd.nu, nx = bd.nx, T = bd.T;
  double* ptrs[17] = {d->lx, d->lu, d->lxx, d->lux, d->luu, d->f_val, d->fx, d->fu, d->fxx, d->fux, d->fuu,
                      d->eq_val, d->eq_x, d->eq_u, d->eq_xx, d->eq_ux, d->eq_uu};
  const int per_knot[11] = {nd, nu, nd * nd, nu * nd, nu * nu, nx, nd * nd, nd * nu, nd * nd * nd, nd * nu * nd, nd * nu * nu};
  const int per_row[6] = {1, nd, nu, nd * nd, nu * nd, nu * nu};
  unsigned long long* mism = (unsigned long long*)(h->bd.scal + (size_t)NSCALARS * bd.B);  // one spare word behind the scalars
  if (to_slab) CU(cudaMemsetAsync(mism, 0, sizeof(unsigned long long), h->stream));
  for (int f = 0; f < 17; ++f) {
    if (!ptrs[f]) continue;
    const size_t total = f < 11 ? (size_t)T * per_knot[f] : (size_t)bd.ne_total * per_row[f - 11];
    if (total == 0) continue;
    const size_t bytes = total * bd.B * sizeof(double);
    int rc = ensure_stage(h, bytes);
    if (rc) return rc;
    if (to_slab) CU(cudaMemcpyAsync(h->stage, ptrs[f], bytes, cudaMemcpyHostToDevice, h->stream));
    slab_field_kernel<<<(unsigned)std::min<size_t>((total * bd.B + 255) / 256, 8192), 256, 0, h->stream>>>(
        bd, h->stage, total, f, f < 11 ? per_knot[f] : 0, f < 11 ? 0 : per_row[f - 11], to_slab, mism);
    rc = check_launch(h, "slab_field");
    if (rc) return rc;
    if (!to_slab) CU(cudaMemcpyAsync(ptrs[f], h->stage, bytes, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
  }
  if (to_slab) {
    unsigned long long bad = 0;
    CU(cudaMemcpyAsync(&bad, mism, sizeof(bad), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (bad)
      return fail(DDPB_ERR_UNSUPPORTED, "derivatives differ from the structure of the reference's problem_t (l = c/2 |u|^2, Euler scheme on a vector-space "
                                        "configuration): " + std::to_string(bad) + " entries that the slab keeps as constants");
  }
  if (!to_slab) {
    if (d->lfx) std::memset(d->lfx, 0, (size_t)bd.B * nd * sizeof(double));        // lfx = 0  (problem.hpp:958)
    if (d->lfxx) std::memset(d->lfxx, 0, (size_t)bd.B * nd * nd * sizeof(double)); // lfxx = 0 (problem.hpp:959)
  }
  return DDPB_OK;
}
int ddpb_set_derivatives(ddpb_batch* h, const ddpb_derivs* in) {
  ENTER(h);
  if (!in) return fail(DDPB_ERR_INVALID, "null derivs");
  return derivs_xfer(h, in, 1);
}
int ddpb_get_derivatives(ddpb_batch* h, const ddpb_derivs* out) {
  ENTER(h);
  if (!out) return fail(DDPB_ERR_INVALID, "null derivs");
  return derivs_xfer(h, out, 0);
}
int ddpb_set_scalar(ddpb_batch* h, int field, const double* values) {
  ENTER(h);
  if (field < 0 || field >= NSCALARS || !values) return fail(DDPB_ERR_INVALID, "bad scalar field");
  CU(cudaMemcpyAsync(h->bd.scal + (size_t)field * h->bd.B, values, h->bd.B * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return DDPB_OK;
}
int ddpb_get_scalars(ddpb_batch* h, double* out) {
  ENTER(h);
  if (!out) return fail(DDPB_ERR_INVALID, "null output");
  CU(cudaMemcpyAsync(out, h->bd.scal, (size_t)NSCALARS * h->bd.B * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return DDPB_OK;
}
int ddpb_scalars_device_ptr(ddpb_batch* h, void** ptr) {
  if (!h || !ptr) return fail(DDPB_ERR_INVALID, "null argument");
  *ptr = h->bd.scal;
  return DDPB_OK;
}

int ddpb_compute_derivatives(ddpb_batch* h) {
  ENTER(h);
  int rc = launch_linearize(LaunchCtx(h));
  if (rc) return rc;
  CU(cudaStreamSynchronize(h->stream));
  return DDPB_OK;
}
int ddpb_backward_pass(ddpb_batch* h, double* mu, double* reg) {
  ENTER(h);
  if (!mu || !reg) return fail(DDPB_ERR_INVALID, "mu / reg are required");
  const int B = h->bd.B;
  CU(cudaMemcpyAsync(h->bd.scal + (size_t)S_MU * B, mu, B * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->bd.scal + (size_t)S_REG * B, reg, B * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  int rc = launch_backward(LaunchCtx(h));
  if (rc) return rc;
  CU(cudaMemcpyAsync(mu, h->bd.scal + (size_t)S_MU * B, B * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(reg, h->bd.scal + (size_t)S_REG * B, B * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return DDPB_OK;
}
int ddpb_forward_pass(ddpb_batch* h, double* step_out) {
  ENTER(h);
  int rc = launch_forward(LaunchCtx(h));
  if (rc) return rc;
  if (step_out) CU(cudaMemcpyAsync(step_out, h->bd.scal + (size_t)S_STEP * h->bd.B, h->bd.B * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return DDPB_OK;
}
int ddpb_cost_seq_aug(ddpb_batch* h, double* costs) {
  ENTER(h);
  if (!costs) return fail(DDPB_ERR_INVALID, "null output");
  DISPATCH(h->mt, L::cost(h->model, h->bd, h->costs_old, h->stream));
  int rc = check_launch(h, "cost_seq_aug");
  if (rc) return rc;
  return download_T(h, costs, h->costs_old, (size_t)h->bd.T + 1);
}
int ddpb_optimality(ddpb_batch* h, double* opt_obj, double* opt_constr) {
  ENTER(h);
  const int B = h->bd.B;
  DISPATCH(h->mt, L::optimality(h->bd, h->stream));
  int rc = check_launch(h, "optimality");
  if (rc) return rc;
  if (opt_obj) CU(cudaMemcpyAsync(opt_obj, h->bd.scal + (size_t)S_OPT_OBJ * B, B * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (opt_constr) CU(cudaMemcpyAsync(opt_constr, h->bd.scal + (size_t)S_OPT_CONSTR * B, B * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return DDPB_OK;
}

static size_t snap_layout(const BatchDev& bd, size_t* off) {
  const size_t B = bd.B, T = bd.T, nx = bd.nx, nu = bd.nu, net = bd.ne_total;
  off[0] = 0; off[1] = off[0] + (T + 1) * nx * B; off[2] = off[1] + T * nu * B; off[3] = off[2] + T * nx * B;
  off[4] = off[3] + net * B; off[5] = off[4] + net * nx * B;
  return off[5];
}
int ddpb_snapshot_save(ddpb_batch* h) {
  ENTER(h);
  const BatchDev& bd = h->bd;
  size_t off[6];
  const size_t total = snap_layout(bd, off);
  if (!h->snap) { int rc = dev_alloc(h, &h->snap, total); if (rc) return rc; }
  double* src[5] = {bd.xs, bd.us, bd.m_origin, bd.m_val, bd.m_jac};
  for (int k = 0; k < 5; ++k)
    if (off[k + 1] > off[k]) CU(cudaMemcpyAsync(h->snap + off[k], src[k], (off[k + 1] - off[k]) * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return DDPB_OK;
}
int ddpb_snapshot_restore(ddpb_batch* h) {
  ENTER(h);
  const BatchDev& bd = h->bd;
  if (!h->snap) return fail(DDPB_ERR_INVALID, "no snapshot saved");
  size_t off[6];
  snap_layout(bd, off);
  double* dst[5] = {bd.xs, bd.us, bd.m_origin, bd.m_val, bd.m_jac};
  for (int k = 0; k < 5; ++k)
    if (off[k + 1] > off[k]) CU(cudaMemcpyAsync(dst[k], h->snap + off[k], (off[k + 1] - off[k]) * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  // the new-trajectory storage is a clone of the trajectory (ddp.hpp:752)
  CU(cudaMemcpyAsync(bd.xs_new, h->snap + off[0], (off[1] - off[0]) * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  CU(cudaMemcpyAsync(bd.us_new, h->snap + off[1], (off[2] - off[1]) * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  return DDPB_OK;
}
// development aid (not declared in ddpb.h): the (phase, start ms, end ms) of every launch group recorded while phase timing
// was on, relative to the first one; consumes the recorded events.  Returns the number of entries written.
int ddpb_debug_timeline(ddpb_batch* h, float* out3, int max_entries) {
  if (!h || !out3) return 0;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  int n = 0;
  for (auto& e : h->events) {
    if (n < max_entries) {
      float t0 = 0.f, t1 = 0.f;
      cudaEventElapsedTime(&t0, h->events[0].a, e.a);
      cudaEventElapsedTime(&t1, h->events[0].a, e.b);
      out3[3 * n] = (float)e.phase; out3[3 * n + 1] = t0; out3[3 * n + 2] = t1;
      ++n;
    }
  }
  for (auto& e : h->events) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
  h->events.clear();
  return n;
}
// development aid (not declared in ddpb.h): per-phase cycle counters of the linearization kernel
int ddpb_debug_phase_cycles(ddpb_batch* h, int enable, unsigned long long* out16) {
  ENTER(h);
  if (!DDPB_PHASE_CYCLES) return fail(DDPB_ERR_UNSUPPORTED, "the per-phase cycle counters are compiled out: build with DDPB_NVCC_EXTRA=-DDDPB_PHASE_CYCLES=1");
  if (enable && !h->bd.dbg) {
    int rc = dev_alloc(h, &h->bd.dbg, 16);
    if (rc) return rc;
    CU(cudaMemsetAsync(h->bd.dbg, 0, 16 * sizeof(unsigned long long), h->stream));
  }
  if (out16 && h->bd.dbg) {
    CU(cudaMemcpyAsync(out16, h->bd.dbg, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaMemsetAsync(h->bd.dbg, 0, 16 * sizeof(unsigned long long), h->stream));
  }
  return DDPB_OK;
}
int ddpb_measure_fp64_peak(int device, double* tflops) {
  if (!tflops) return fail(DDPB_ERR_INVALID, "null output");
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  double* d = nullptr;
  CU(cudaMalloc((void**)&d, sizeof(double)));
  cudaEvent_t a, b;
  CU(cudaEventCreate(&a)); CU(cudaEventCreate(&b));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 20000;
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    CU(cudaEventRecord(a));
    fp64_peak_kernel<<<blocks, threads>>>(d, iters);
    CU(cudaEventRecord(b));
    CU(cudaEventSynchronize(b));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, a, b));
    const double tf = 2.0 * 16.0 * iters * (double)blocks * threads / (ms * 1e-3) * 1e-12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(d);
  *tflops = best;
  return DDPB_OK;
}

// enqueue the whole fixed-iteration solve on h->stream and the sub-batch streams (forked from / joined into h->stream)
static int enqueue_solve(ddpb_batch* h, const ddpb_solver_params* p) {
  BatchDev& bd = h->bd;
  const size_t B = bd.B;
  int rc = DDPB_OK;
#define R(x) do { rc = (x); if (rc) return rc; } while (0)
  CU(cudaMemsetAsync(bd.scal, 0, (size_t)NSCALARS * B * sizeof(double), h->stream));
  R(fill(h, bd.scal + S_MU * B, B, p->mu));
  R(fill(h, bd.scal + S_REG * B, B, p->reg));
  R(fill(h, bd.scal + S_W * B, B, p->w));
  R(fill(h, bd.scal + S_N * B, B, p->n));
  // sub-batches: contiguous ranges of problems, multiples of 32 (one warp of the thread-per-problem kernels)
  const int nsub = h->nsub;
  std::vector<LaunchCtx> sub;
  if (nsub <= 1) {
    sub.emplace_back(h);
  } else {
    const int per = (((int)B + nsub - 1) / nsub + 31) / 32 * 32;
    for (int i = 0, b0 = 0; i < nsub && b0 < (int)B; ++i, b0 += per) sub.emplace_back(h, b0, std::min(per, (int)B - b0), h->sub_streams[i]);
    CU(cudaEventRecord(h->fork_ev, h->stream));
    for (auto& c : sub) CU(cudaStreamWaitEvent(c.s, h->fork_ev, 0));
  }
  // warm-up derivative / backward / forward (ddp.hpp:768-773): mu is taken from it, reg is not,
  // and the trial trajectory is not swapped in
  // linearization of sub-batch i runs on its low-priority stream, everything else on the high-priority one
  auto linearize_sub = [&](size_t i) -> int {
    LaunchCtx& c = sub[i];
    if (nsub <= 1) return launch_linearize(c);
    LaunchCtx lo = c;
    lo.s = h->sub_streams_lo[i];
    if (cudaEventRecord(h->sub_hi[i], c.s) != cudaSuccess || cudaStreamWaitEvent(lo.s, h->sub_hi[i], 0) != cudaSuccess) return fail(DDPB_ERR_CUDA, "stream fork");
    int rc_ = launch_linearize(lo);
    if (rc_) return rc_;
    if (cudaEventRecord(h->sub_lin[i], lo.s) != cudaSuccess || cudaStreamWaitEvent(c.s, h->sub_lin[i], 0) != cudaSuccess) return fail(DDPB_ERR_CUDA, "stream join");
    return DDPB_OK;
  };
  h->ls_pass = 0;  // the line-search passes of a solve alternate the two flag buffers from a fixed start (graph replay)
  for (size_t i = 0; i < sub.size(); ++i) {
    LaunchCtx& c = sub[i];
    R(scalar_copy(c, S_SPARE, S_REG));
    R(linearize_sub(i));
    R(launch_backward(c));
    R(scalar_copy(c, S_REG, S_SPARE));
    R(launch_forward(c));
    CU(cudaMemsetAsync(bd.scal + S_ROLLOUTS * B + c.bd.b0, 0, c.bd.nb * sizeof(double), c.s));
    CU(cudaMemsetAsync(bd.scal + S_RESTARTS * B + c.bd.b0, 0, c.bd.nb * sizeof(double), c.s));
  }
  for (int iter = 0; iter < p->max_iterations; ++iter) {
    ++h->ls_pass;
    for (size_t i = 0; i < sub.size(); ++i) {
      LaunchCtx& c = sub[i];
      c.bd.xs = bd.xs; c.bd.us = bd.us; c.bd.xs_new = bd.xs_new; c.bd.us_new = bd.us_new;
      R(linearize_sub(i));
      R(launch_al_update(c, p->optimality_stopping_threshold));
      R(launch_backward(c));
      R(launch_forward(c));
      R(launch_post_forward(c));
    }
    std::swap(bd.xs, bd.xs_new);   // swap(traj, new_traj)  (ddp.hpp:826)
    std::swap(bd.us, bd.us_new);
  }
  if (nsub > 1) {
    for (size_t i = 0; i < sub.size(); ++i) {
      CU(cudaEventRecord(h->sub_done[i], sub[i].s));
      CU(cudaStreamWaitEvent(h->stream, h->sub_done[i], 0));
    }
  }
#undef R
  return DDPB_OK;
}
// The solve is a fixed sequence of ~10^3 launches and stream dependencies: it is captured once into a CUDA graph per
// (parameters, sub-batch count, trajectory-buffer roles) and replayed, so the host cost of a solve is one graph launch.
// Per-launch timing (ddpb_enable_phase_timing) and DDPB_NO_GRAPH=1 use the direct path.
int ddpb_solve_async(ddpb_batch* h, const ddpb_solver_params* p) {
  ENTER(h);
  if (!p || p->max_iterations < 0) return fail(DDPB_ERR_INVALID, "bad solver parameters");
  static const bool no_graph = getenv("DDPB_NO_GRAPH") != nullptr;
  if (h->timing || no_graph) return enqueue_solve(h, p);
  for (auto& g : h->graphs) {
    if (g.iters == p->max_iterations && g.thr == p->optimality_stopping_threshold && g.mu == p->mu && g.reg == p->reg && g.w == p->w &&
        g.n == p->n && g.nsub == h->nsub && g.xs == h->bd.xs && g.stream == h->stream) {
      CU(cudaGraphLaunch(g.exec, h->stream));
      h->launches += g.launches;
      if (p->max_iterations & 1) { std::swap(h->bd.xs, h->bd.xs_new); std::swap(h->bd.us, h->bd.us_new); }
      return DDPB_OK;
    }
  }
  SolveGraph g;
  g.iters = p->max_iterations; g.thr = p->optimality_stopping_threshold; g.mu = p->mu; g.reg = p->reg; g.w = p->w; g.n = p->n;
  g.nsub = h->nsub; g.xs = h->bd.xs; g.stream = h->stream;
  const unsigned long long l0 = h->launches;
  double* const xs0 = h->bd.xs; double* const xsn0 = h->bd.xs_new; double* const us0 = h->bd.us; double* const usn0 = h->bd.us_new;
  CU(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
  const int rc = enqueue_solve(h, p);
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
  if (rc != DDPB_OK || ce != cudaSuccess || !graph) {
    // capture failed: restore the host-side state and run the sequence directly
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    h->bd.xs = xs0; h->bd.xs_new = xsn0; h->bd.us = us0; h->bd.us_new = usn0;
    h->launches = l0;
    return enqueue_solve(h, p);
  }
  g.launches = h->launches - l0;
  const cudaError_t ie = cudaGraphInstantiate(&g.exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ie != cudaSuccess) {
    cudaGetLastError();
    h->bd.xs = xs0; h->bd.xs_new = xsn0; h->bd.us = us0; h->bd.us_new = usn0;
    h->launches = l0;
    return enqueue_solve(h, p);
  }
  if (h->graphs.size() >= 8) { cudaGraphExecDestroy(h->graphs.front().exec); h->graphs.erase(h->graphs.begin()); }
  h->graphs.push_back(g);
  CU(cudaGraphLaunch(g.exec, h->stream));  // the capture itself executed nothing
  return DDPB_OK;
}
int ddpb_solve(ddpb_batch* h, const ddpb_solver_params* p, double* scalars_out) {
  int rc = ddpb_solve_async(h, p);
  if (rc) return rc;
  if (scalars_out) return ddpb_get_scalars(h, scalars_out);
  CU(cudaStreamSynchronize(h->stream));
  return DDPB_OK;
}

}  // extern "C"
