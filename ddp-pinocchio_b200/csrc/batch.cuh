// batch.cuh -- device-side description of a batch and the HBM layouts (DESIGN.md "Data layout").
//
//  * small per-knot vectors (x, u, k, K, multipliers, targets) and per-problem scalars are
//    SoA-by-batch: element e of knot t of problem b lives at ((t * E + e) * B + b), problem index
//    fastest, so thread-per-problem kernels (rollouts, adjoint sweeps) are fully coalesced;
//  * derivative slabs (>= 90% of all bytes) hold, per (knot, problem), the STATE-DEPENDENT entries of the 17 sequences
//    of the reference's derivative_storage_t (/root/reference/include/ddp/ddp.hpp:84-105).  The reference's
//    problem_t fixes the cost (l = c/2 |u|^2, problem.hpp:932-942 => lx = lxx = lux = 0, luu = c I) and the Euler
//    scheme on a vector-space configuration (problem.hpp:441-503 => rows k < n of f_x, f_u are the constants
//    [I, dt I], 0, and with FD-of-Jacobians Hessians rows k < n of f_xx / f_ux, the v-slices of f_ux and all of f_uu are
//    exact zeros): those entries are never stored -- every consumer uses the constants -- which cuts a record of the
//    6R arm from 3522 to 1206 doubles.  What is stored, each block in the reference's element order restricted to the
//    stored rows (column-major, row index fastest):
//        f[nx] | fx_B[n x ndx] | fu_B[n x nu] | lu[nu] | tensors | eq | eq_x | eq_u | eq_xx | eq_ux | eq_uu
//    tensors, FD-of-Jacobians mode: fxx_B[(k-n) + n (j + ndx i)] | fux_B[(k-n) + n (j + nu i)], i < n
//    tensors, value-based mode    : fxx | fux | fuu dense (their "structural" entries carry the reference's FD
//                                   rounding noise there, so they are data)
//    The slab is record-major: one contiguous record per (knot, problem), ordered [knot][problem][element], which the
//    backward pass streams with bulk asynchronous copies (TMA / cp.async).  The host boundary (ddpb_get/set_derivatives)
//    always speaks the reference's dense packed layout.
#pragma once
#include "rbd.cuh"

namespace ddpb {

enum { EQ_NONE = 0, EQ_CONFIG = 1, EQ_SPATIAL = 2 };
enum { HESS_FD_JACOBIANS = 0, HESS_FD_VALUES = 1 };
enum {
  S_MU = 0, S_REG, S_W, S_N, S_STEP, S_OPT_OBJ, S_OPT_CONSTR, S_COST, S_MERIT, S_ITERATIONS, S_STATUS, S_ROLLOUTS, S_RESTARTS,
  S_MULT_UPDATES, S_FLAG /* internal: multiplier-update decision of the current iteration */, S_SPARE, NSCALARS = 16
};

// offsets (doubles) of the stored blocks inside one slab record of an n-joint model for a knot with `ne` equality
// rows; `rt` = stored rows per tensor column (n: bottom rows only, or ndx: dense), `ri` = stored x-perturbation slices
// of f_ux (n or ndx), `n_fuu` = stored doubles of f_uu (0 or dense)
struct RecLayout {
  int f, fx, fu, lu, fxx, fux, fuu, eq, eq_x, eq_u, eq_xx, eq_ux, eq_uu, size;
  int rt, ri, r0;  // tensor rows stored per column, f_ux right-index count, first stored tensor row (n or 0)
};
DDPB_HD RecLayout rec_layout(int n, int ne, int hess_mode) {
  const int nd = 2 * n, nu = n, nx = 2 * n;
  const bool dense = hess_mode == HESS_FD_VALUES;
  RecLayout L;
  L.rt = dense ? nd : n; L.ri = dense ? nd : n; L.r0 = dense ? 0 : n;
  int o = 0;
  L.f = o; o += nx;
  L.fx = o; o += n * nd;
  L.fu = o; o += n * nu;
  L.lu = o; o += nu;                 // n (n + 1) is even: the tensors start on a 16-byte boundary
  L.fxx = o; o += L.rt * nd * nd;
  L.fux = o; o += L.rt * nu * L.ri;
  L.fuu = o; o += dense ? nd * nu * nu : 0;
  L.eq = o; o += ne;
  L.eq_x = o; o += ne * nd;
  L.eq_u = o; o += ne * nu;
  L.eq_xx = o; o += ne * nd * nd;
  L.eq_ux = o; o += ne * nu * nd;
  L.eq_uu = o; o += ne * nu * nu;
  L.size = (o + 1) & ~1;  // 16-byte multiple: records stay aligned for vector / bulk copies
  return L;
}

// Dense element e (reference element order) of derivative field `field` (0 lx, 1 lu, 2 lxx, 3 lux, 4 luu, 5 f, 6 fx, 7 fu,
// 8 fxx, 9 fux, 10 fuu, 11 eq, 12 eq_x, 13 eq_u, 14 eq_xx, 15 eq_ux, 16 eq_uu) -> offset inside the stored record, or -1
// when the entry is structural; *sval is its constant value then.  Used by the host-boundary converters only.
DDPB_HD int stored_offset(const RecLayout& L, int n, int field, int e, double dt, double c, double* sval) {
  const int nd = 2 * n, nu = n;
  *sval = 0.0;
  switch (field) {
    case 0: case 2: case 3: return -1;
    case 1: return L.lu + e;
    case 4: *sval = (e / nu == e % nu) ? c : 0.0; return -1;
    case 5: return L.f + e;
    case 6: {
      const int k = e % nd, col = e / nd;
      if (k >= n) return L.fx + (k - n) + n * col;
      *sval = col < n ? (k == col ? 1.0 : 0.0) : (k == col - n ? dt : 0.0);
      return -1;
    }
    case 7: {
      const int k = e % nd, col = e / nd;
      return k >= n ? L.fu + (k - n) + n * col : -1;
    }
    case 8: {
      if (L.rt == nd) return L.fxx + e;
      const int k = e % nd, ji = e / nd;
      return k >= n ? L.fxx + (k - n) + n * ji : -1;
    }
    case 9: {
      if (L.rt == nd) return L.fux + e;
      const int k = e % nd, ji = e / nd, j = ji % nu, i = ji / nu;
      return (k >= n && i < n) ? L.fux + (k - n) + n * (j + nu * i) : -1;
    }
    case 10: return L.rt == nd ? L.fuu + e : -1;
    case 11: return L.eq + e;
    case 12: return L.eq_x + e;
    case 13: return L.eq_u + e;
    case 14: return L.eq_xx + e;
    case 15: return L.eq_ux + e;
    default: return L.eq_uu + e;
  }
}

struct BatchDev {
  int B, T, n, nx, nu;         // nx = ndx = 2n, nu = n
  int hess_mode, eq_kind, frame_id, jac_mode;
  int ne_max, ne_total;
  int max_restarts;                   // 0 = unbounded restarts (reference behaviour)
  int pad_;
  int b0, nb;                         // problems [b0, b0 + nb) covered by a launch (sub-batches run on their own streams)
  double dt, c;
  double *xs, *us, *xs_new, *us_new;  // [(T+1)][nx][B], [T][nu][B]
  double* targets;                    // [ne_total][B]
  double *m_origin, *m_val, *m_jac;   // [T][nx][B], [ne_total][B], [ne_total*nd][B]
  double *fb_origin, *fb_k, *fb_K;    // [T][nx][B], [T][nu][B], [T][nu*nd][B]
  double* slab;                       // records of knot t start at slab_off[t]*B (see rec_view)
  double* scal;                       // [NSCALARS][B]
  const int* ne;                      // [T]
  const int* eq_off;                  // [T+1]
  const long long* slab_off;          // [T+1] prefix sums of the per-knot record sizes
  unsigned long long* dbg;            // optional [16] per-phase cycle counters (development; nullptr = off)
};

// the contiguous record of (knot t, problem b)
DDPB_HD double* slab_rec(const BatchDev& bd, int t, int b, int rec_size) {
  return bd.slab + (size_t)bd.slab_off[t] * bd.B + (size_t)b * rec_size;
}
// the same without knowing the knot's row count: the record size of knot t is the step of the prefix table (two cached
// loads instead of evaluating rec_layout -- the offsets of the state-dependent blocks do not depend on the row count)
DDPB_HD double* slab_rec_t(const BatchDev& bd, int t, int b) {
  const long long o0 = bd.slab_off[t], o1 = bd.slab_off[t + 1];
  return bd.slab + (size_t)o0 * bd.B + (size_t)b * (size_t)(o1 - o0);
}
using RecView = double*;
DDPB_HD RecView rec_view(const BatchDev& bd, int t, int b, int rec_size) { return slab_rec(bd, t, b, rec_size); }

// ---- execution context: the same CTA body runs on the device (threads of one block) and, for the
// CPU unit tests of tests/hostemu, as a plain sequential loop over the tasks of every phase -------
// development aid: accumulate the cycles thread 0 of a CTA spent since *t0 into bd.dbg[slot].  Compiled in only with
// -DDDPB_PHASE_CYCLES=1 (tools/phase_cycles.py builds such a library): the ~30 guarded clock reads of the linearization and
// backward kernels were measured to cost the product 2 % of a C3 solve although never taken (instruction fetch, see variant.cuh).
#ifndef DDPB_PHASE_CYCLES
#define DDPB_PHASE_CYCLES 0
#endif
#if defined(__CUDA_ARCH__) && DDPB_PHASE_CYCLES
#define DDPB_PHASE_MARK(bd, slot, t0)                                                            \
  do {                                                                                           \
    if ((bd).dbg && threadIdx.x == 0) {                                                          \
      const long long now_ = clock64();                                                          \
      atomicAdd((bd).dbg + (slot), (unsigned long long)(now_ - (t0)));                           \
      (t0) = now_;                                                                               \
    }                                                                                            \
  } while (0)
#define DDPB_PHASE_START(bd, t0) long long t0 = ((bd).dbg && threadIdx.x == 0) ? clock64() : 0
#else
#define DDPB_PHASE_MARK(bd, slot, t0) do { } while (0)
#define DDPB_PHASE_START(bd, t0) long long t0 = 0; (void)t0
#endif

struct DevExec {
  DDPB_HD int tid() const {
#if defined(__CUDA_ARCH__)
    return threadIdx.x;
#else
    return 0;
#endif
  }
  DDPB_HD int nthr() const {
#if defined(__CUDA_ARCH__)
    return blockDim.x;
#else
    return 1;
#endif
  }
  DDPB_HD void sync() const {
#if defined(__CUDA_ARCH__)
    __syncthreads();
#endif
  }
  DDPB_HD void sync_warp() const {
#if defined(__CUDA_ARCH__)
    __syncwarp();
#endif
  }
};

}  // namespace ddpb
