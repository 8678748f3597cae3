// batch.cuh -- device-side description of a batch and the HBM layouts (DESIGN.md "Data layout").
//
//  * small per-knot vectors (x, u, k, K, multipliers, targets) and per-problem scalars are
//    SoA-by-batch: element e of knot t of problem b lives at ((t * E + e) * B + b), problem index
//    fastest, so thread-per-problem kernels (rollouts, adjoint sweeps) are fully coalesced;
//  * derivative slabs (>= 90% of all bytes) are one contiguous record per (knot, problem), ordered
//    [knot][problem], holding the 17 sequences of the reference's derivative_storage_t
//    (/root/reference/include/ddp/ddp.hpp:84-105) back to back, each block in the reference's element
//    order (column-major matrices, tensors k + i*O + j*O*L).
#pragma once
#include "rbd.cuh"

namespace ddpb {

enum { EQ_NONE = 0, EQ_CONFIG = 1, EQ_SPATIAL = 2 };
enum { HESS_FD_JACOBIANS = 0, HESS_FD_VALUES = 1 };
enum {
  S_MU = 0, S_REG, S_W, S_N, S_STEP, S_OPT_OBJ, S_OPT_CONSTR, S_COST, S_MERIT, S_ITERATIONS, S_STATUS, S_ROLLOUTS, S_RESTARTS,
  S_MULT_UPDATES, S_FLAG /* internal: multiplier-update decision of the current iteration */, S_SPARE, NSCALARS = 16
};

// offsets (doubles) of the 17 blocks inside one slab record for a knot with `ne` equality rows
struct RecLayout {
  int lx, lu, lxx, lux, luu, f, fx, fu, fxx, fux, fuu, eq, eq_x, eq_u, eq_xx, eq_ux, eq_uu, size;
};
DDPB_HD RecLayout rec_layout(int nd, int nu, int nx, int ne) {
  RecLayout L;
  int o = 0;
  L.lx = o; o += nd;
  L.lu = o; o += nu;
  L.lxx = o; o += nd * nd;
  L.lux = o; o += nu * nd;
  L.luu = o; o += nu * nu;
  L.f = o; o += nx;
  L.fx = o; o += nd * nd;
  L.fu = o; o += nd * nu;
  L.fxx = o; o += nd * nd * nd;
  L.fux = o; o += nd * nu * nd;
  L.fuu = o; o += nd * nu * nu;
  L.eq = o; o += ne;
  L.eq_x = o; o += ne * nd;
  L.eq_u = o; o += ne * nu;
  L.eq_xx = o; o += ne * nd * nd;
  L.eq_ux = o; o += ne * nu * nd;
  L.eq_uu = o; o += ne * nu * nu;
  L.size = (o + 1) & ~1;  // 16-byte multiple: records stay aligned for vector / bulk copies
  return L;
}

struct BatchDev {
  int B, T, n, nx, nu;         // nx = ndx = 2n, nu = n
  int hess_mode, eq_kind, frame_id, jac_mode;
  int ne_max, ne_total;
  int max_restarts, pad_;             // 0 = unbounded restarts (reference behaviour)
  int b0, nb;                         // problems [b0, b0 + nb) covered by a launch (sub-batches run on their own streams)
  double dt, c;
  double *xs, *us, *xs_new, *us_new;  // [(T+1)][nx][B], [T][nu][B]
  double* targets;                    // [ne_total][B]
  double *m_origin, *m_val, *m_jac;   // [T][nx][B], [ne_total][B], [ne_total*nd][B]
  double *fb_origin, *fb_k, *fb_K;    // [T][nx][B], [T][nu][B], [T][nu*nd][B]
  double* slab;                       // records, knot t of problem b at slab_off[t]*B + b*rec_size(t)
  double* scal;                       // [NSCALARS][B]
  const int* ne;                      // [T]
  const int* eq_off;                  // [T+1]
  const long long* slab_off;          // [T+1] prefix sums of the per-knot record sizes
  unsigned long long* dbg;            // optional [16] per-phase cycle counters (development; nullptr = off)
};

DDPB_HD double* slab_rec(const BatchDev& bd, int t, int b, int rec_size) {
  return bd.slab + (size_t)bd.slab_off[t] * bd.B + (size_t)b * rec_size;
}

// ---- execution context: the same CTA body runs on the device (threads of one block) and, for the
// CPU unit tests of tests/hostemu, as a plain sequential loop over the tasks of every phase -------
// development aid: accumulate the cycles thread 0 of a CTA spent since *t0 into bd.dbg[slot]
#if defined(__CUDA_ARCH__)
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
