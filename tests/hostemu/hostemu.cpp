// hostemu.cpp -- TEST INFRASTRUCTURE: runs the product's CTA bodies (csrc/*.cuh) on the CPU, one
// "thread" at a time, so the kernel arithmetic can be checked against the oracle in this GPU-less
// container (`pytest -m "not gpu"`).  It is never loaded by the product: libddpb.so has no CPU path and
// fails with DDPB_ERR_CUDA without a device.  Arrays use the DEVICE layouts (SoA-by-batch, slab records);
// tests/hostemu_lib.py does the host<->device layout conversion with numpy.
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../ddp-pinocchio_b200/csrc/backward.cuh"
#include "../../ddp-pinocchio_b200/csrc/backward_lg.cuh"
#include "../../ddp-pinocchio_b200/csrc/forward.cuh"
#include "../../ddp-pinocchio_b200/csrc/linearize.cuh"
#include "../../ddp-pinocchio_b200/csrc/linearize_pt.cuh"
#include "../../ddp-pinocchio_b200/csrc/model_eval.cuh"
#include "../../ddp-pinocchio_b200/csrc/model_host.hpp"

using namespace ddpb;

struct Emu {
  ModelDev model;
  BatchDev bd;
  int mt, nj;
  std::vector<int> ne, eq_off;
  std::vector<long long> slab_off;
  std::vector<std::vector<double>> bufs;
  std::vector<double> costs_old;
  std::string err;
};

#define EMU_NJ(nj, STMT)                           \
  switch (nj) {                                    \
    case 1: { constexpr int NJ = 1; STMT; } break; \
    case 2: { constexpr int NJ = 2; STMT; } break; \
    case 3: { constexpr int NJ = 3; STMT; } break; \
    case 4: { constexpr int NJ = 4; STMT; } break; \
    case 5: { constexpr int NJ = 5; STMT; } break; \
    case 6: { constexpr int NJ = 6; STMT; } break; \
    case 7: { constexpr int NJ = 7; STMT; } break; \
    default: break;                                \
  }
#define EMU_MT(mt, STMT)                           \
  switch (mt) {                                    \
    case 0: { constexpr int MT = 0; STMT; } break; \
    case 1: { constexpr int MT = 1; STMT; } break; \
    case 2: { constexpr int MT = 2; STMT; } break; \
    case 3: { constexpr int MT = 3; STMT; } break; \
    case 4: { constexpr int MT = 4; STMT; } break; \
    case 5: { constexpr int MT = 5; STMT; } break; \
    case 6: { constexpr int MT = 6; STMT; } break; \
    case 7: { constexpr int MT = 7; STMT; } break; \
    default: break;                                \
  }

template <int NJ> static void linearize_host(const ModelDev& m, const BatchDev& bd, int t, int b0, int nk, int kpb, double* sm) {
  if (bd.hess_mode == HESS_FD_JACOBIANS) linearize_group<NJ, HESS_FD_JACOBIANS>(m, bd, t, b0, nk, kpb, bd.ne[t], sm, DevExec{});
  else linearize_group<NJ, HESS_FD_VALUES>(m, bd, t, b0, nk, kpb, bd.ne[t], sm, DevExec{});
}
// chains of <= 2 joints run the lane-group body (backward_lg.cuh), longer ones the CTA body -- as on the device
template <int NJ> static void backward_host(const BatchDev& bd, int b, double* sm) {
  if (NJ <= 2) {
    if (bd.hess_mode == HESS_FD_VALUES) backward_lg_plain<(NJ <= 2 ? NJ : 1), true>(bd, b);
    else backward_lg_plain<(NJ <= 2 ? NJ : 1), false>(bd, b);
  } else {
    if (bd.hess_mode == HESS_FD_VALUES) backward_problem<NJ, 1, true>(bd, b, sm, DevExec{});
    else backward_problem<NJ, 1, false>(bd, b, sm, DevExec{});
  }
}

extern "C" {

void* emu_create(const ddpb_model_desc* d, int B, int T, const int* ne_per_knot, const ddpb_params* p) {
  Emu* h = new Emu();
  if (model_from_desc(d, h->model, h->err) != DDPB_OK) { delete h; return nullptr; }
  const int n = h->model.n;
  h->mt = h->model.kind == DDPB_MODEL_PENDULUM ? 0 : n;
  h->nj = n;
  BatchDev& bd = h->bd;
  std::memset(&bd, 0, sizeof(bd));
  bd.B = B; bd.T = T; bd.n = n; bd.nx = 2 * n; bd.nu = n;
  bd.b0 = 0; bd.nb = B;
  bd.hess_mode = p->hess_mode; bd.eq_kind = p->eq_kind; bd.frame_id = p->frame_id; bd.jac_mode = p->frame_jac_mode;
  bd.dt = p->dt; bd.c = p->c;
  bd.max_restarts = p->max_backward_restarts < 0 ? 0 : p->max_backward_restarts;
  h->ne.assign(T, 0); h->eq_off.assign(T + 1, 0); h->slab_off.assign(T + 1, 0);
  for (int t = 0; t < T; ++t) {
    const int ne = (p->eq_kind == DDPB_EQ_NONE || !ne_per_knot) ? 0 : ne_per_knot[t];
    h->ne[t] = ne;
    h->eq_off[t + 1] = h->eq_off[t] + ne;
    bd.ne_max = bd.ne_max > ne ? bd.ne_max : ne;
    h->slab_off[t + 1] = h->slab_off[t] + rec_layout(n, ne, p->hess_mode).size;
  }
  bd.ne_total = h->eq_off[T];
  const size_t nx = bd.nx, nu = bd.nu, nd = bd.nx, net = bd.ne_total;
  auto A = [&](double*& ptr, size_t count) {
    h->bufs.emplace_back(count ? count : 1, 0.0);
    ptr = h->bufs.back().data();
  };
  A(bd.xs, (T + 1) * nx * B); A(bd.us, T * nu * B); A(bd.xs_new, (T + 1) * nx * B); A(bd.us_new, T * nu * B);
  A(bd.targets, net * B); A(bd.m_origin, T * nx * B); A(bd.m_val, net * B); A(bd.m_jac, net * nd * B);
  A(bd.fb_origin, T * nx * B); A(bd.fb_k, T * nu * B); A(bd.fb_K, T * nu * nd * B);
  A(bd.slab, (size_t)h->slab_off[T] * B); A(bd.scal, (size_t)NSCALARS * B);
  h->costs_old.assign((size_t)(T + 1) * B, 0.0);
  bd.ne = h->ne.data(); bd.eq_off = h->eq_off.data(); bd.slab_off = h->slab_off.data();
  return h;
}
void emu_destroy(void* hv) { delete (Emu*)hv; }

// which: 0 xs, 1 us, 2 xs_new, 3 us_new, 4 targets, 5 m_origin, 6 m_val, 7 m_jac, 8 fb_origin, 9 fb_k, 10 fb_K, 11 scal, 12 costs_old
double* emu_ptr(void* hv, int which) {
  Emu* h = (Emu*)hv;
  BatchDev& bd = h->bd;
  double* p[13] = {bd.xs, bd.us, bd.xs_new, bd.us_new, bd.targets, bd.m_origin, bd.m_val, bd.m_jac, bd.fb_origin, bd.fb_k, bd.fb_K, bd.scal, h->costs_old.data()};
  return p[which];
}

void emu_linearize(void* hv) {
  Emu* h = (Emu*)hv;
  const BatchDev& bd = h->bd;
  if (h->mt == 0) {
    for (int t = 0; t < bd.T; ++t)
      for (int b = 0; b < bd.B; ++b)
        if (bd.scal[(size_t)S_STATUS * bd.B + b] == 0.0) linearize_pendulum(h->model, bd, b, t);
    return;
  }
  EMU_NJ(h->nj, {
    const int kpb = 3;  // exercise the several-knots-per-CTA path, including a ragged last group
    for (int t = 0; t < bd.T; ++t) {
      if (NJ <= 2 && bd.hess_mode == HESS_FD_JACOBIANS && bd.ne[t] == 0) {   // as the launcher does (ddpb_capi.cu launch_linearize)
        for (int b = 0; b < bd.B; ++b)
          if (bd.scal[(size_t)S_STATUS * bd.B + b] == 0.0) linearize_pt_knot_plain<(NJ <= 2 ? NJ : 1)>(h->model, bd, t, b);
        continue;
      }
      LinPlan<NJ> pl;
      pl.init(bd.hess_mode, bd.ne[t], bd.ne[t], bd.eq_kind);   // shared memory sized for this knot's row count, as the launcher does
      std::vector<double> sm((size_t)pl.total * kpb);
      for (int b0 = 0; b0 < bd.B; b0 += kpb) {
        const int nk = bd.B - b0 < kpb ? bd.B - b0 : kpb;
        bool any = false;
        for (int k = 0; k < nk; ++k) any = any || bd.scal[(size_t)S_STATUS * bd.B + b0 + k] == 0.0;
        if (any) linearize_host<NJ>(h->model, bd, t, b0, nk, kpb, sm.data());
      }
    }
  });
}
void emu_backward(void* hv) {
  Emu* h = (Emu*)hv;
  const BatchDev& bd = h->bd;
  for (size_t i = 0; i < (size_t)bd.T * bd.nx * bd.B; ++i) fb_origin_knot(bd, i);   // bd.b0 = 0, bd.nb = B in the emulation
  EMU_NJ(h->nj, {
    std::vector<double> sm(BwdSmem<NJ>::SIZE);
    for (int b = 0; b < bd.B; ++b)
      if (bd.scal[(size_t)S_STATUS * bd.B + b] == 0.0) backward_host<NJ>(bd, b, sm.data());
  });
}
void emu_forward(void* hv) {
  Emu* h = (Emu*)hv;
  EMU_MT(h->mt, { for (int b = 0; b < h->bd.B; ++b) forward_problem<MT>(h->model, h->bd, b, h->costs_old.data()); });
}
void emu_rollout(void* hv) {
  Emu* h = (Emu*)hv;
  EMU_MT(h->mt, { for (int b = 0; b < h->bd.B; ++b) { double scr_[DynOps<MT>::SCR]; rollout_problem<MT>(h->model, h->bd, b, Scratch{scr_, 1}); } });
}
void emu_cost(void* hv) {
  Emu* h = (Emu*)hv;
  EMU_MT(h->mt, {
    for (int b = 0; b < h->bd.B; ++b)
      h->bd.scal[(size_t)S_COST * h->bd.B + b] = [&] { double scr_[DynOps<MT>::SCR]; return cost_seq_problem<MT>(h->model, h->bd, b, h->costs_old.data(), h->bd.scal[(size_t)S_MU * h->bd.B + b], Scratch{scr_, 1}); }();
  });
}
void emu_al_update(void* hv, double thr) {
  Emu* h = (Emu*)hv;
  EMU_MT(h->mt, { for (int b = 0; b < h->bd.B; ++b) al_update_problem<MT>(h->bd, b, thr); });
}
void emu_optimality(void* hv) {
  Emu* h = (Emu*)hv;
  const BatchDev& bd = h->bd;
  EMU_MT(h->mt, {
    for (int b = 0; b < bd.B; ++b) {
      bd.scal[(size_t)S_OPT_OBJ * bd.B + b] = optimality_obj_problem<MT>(bd, b, bd.scal[(size_t)S_MU * bd.B + b]);
      bd.scal[(size_t)S_OPT_CONSTR * bd.B + b] = optimality_constr_problem<MT>(bd, b);
    }
  });
}
void emu_post_forward(void* hv) {
  Emu* h = (Emu*)hv;
  for (int b = 0; b < h->bd.B; ++b) post_forward_problem(h->bd, b);
}
void emu_swap(void* hv) {
  Emu* h = (Emu*)hv;
  std::swap(h->bd.xs, h->bd.xs_new);
  std::swap(h->bd.us, h->bd.us_new);
}
// the launch sequence of ddpb_solve_async (ddpb_capi.cu), minus streams
void emu_solve(void* hv, const ddpb_solver_params* p) {
  Emu* h = (Emu*)hv;
  BatchDev& bd = h->bd;
  const size_t B = bd.B;
  std::memset(bd.scal, 0, sizeof(double) * NSCALARS * B);
  for (size_t b = 0; b < B; ++b) {
    bd.scal[S_MU * B + b] = p->mu; bd.scal[S_REG * B + b] = p->reg; bd.scal[S_W * B + b] = p->w; bd.scal[S_N * B + b] = p->n;
    bd.scal[S_SPARE * B + b] = p->reg;
  }
  emu_linearize(h);
  emu_backward(h);
  for (size_t b = 0; b < B; ++b) bd.scal[S_REG * B + b] = bd.scal[S_SPARE * B + b];
  emu_forward(h);
  for (size_t b = 0; b < B; ++b) { bd.scal[S_ROLLOUTS * B + b] = 0.0; bd.scal[S_RESTARTS * B + b] = 0.0; }
  for (int iter = 0; iter < p->max_iterations; ++iter) {
    emu_linearize(h);
    emu_al_update(h, p->optimality_stopping_threshold);
    emu_backward(h);
    emu_forward(h);
    emu_post_forward(h);
    emu_swap(h);
  }
}
// packed dense field [B][total] <-> stored slab (host version of slab_field_kernel); returns the number of dense entries
// that differ from their structural constant (to_slab only)
int emu_field(void* hv, double* packed, int field, int to_slab) {
  Emu* h = (Emu*)hv;
  const BatchDev& bd = h->bd;
  const int n = bd.n, nd = bd.nx, nu = bd.nu, nx = bd.nx, T = bd.T;
  const int per_knot[11] = {nd, nu, nd * nd, nu * nd, nu * nu, nx, nd * nd, nd * nu, nd * nd * nd, nd * nu * nd, nd * nu * nu};
  const int per_row[6] = {1, nd, nu, nd * nd, nu * nd, nu * nu};
  const size_t total = field < 11 ? (size_t)T * per_knot[field] : (size_t)bd.ne_total * per_row[field - 11];
  int bad = 0;
  for (int b = 0; b < bd.B; ++b)
    for (int t = 0; t < T; ++t) {
      const RecLayout L = rec_layout(n, bd.ne[t], bd.hess_mode);
      double* r = slab_rec(bd, t, b, L.size);
      const size_t cnt = field < 11 ? per_knot[field] : (size_t)bd.ne[t] * per_row[field - 11];
      const size_t pos = field < 11 ? (size_t)t * per_knot[field] : (size_t)bd.eq_off[t] * per_row[field - 11];
      for (size_t e = 0; e < cnt; ++e) {
        double sval;
        const int off = stored_offset(L, n, field, (int)e, bd.dt, bd.c, &sval);
        double& pk = packed[(size_t)b * total + pos + e];
        if (off >= 0) { if (to_slab) r[off] = pk; else pk = r[off]; }
        else if (to_slab) { if (!(pk == sval)) ++bad; }
        else pk = sval;
      }
    }
  return bad;
}

// the Model concept's own functions for A->n samples: the body of model_eval_kernel (kernels_nj.cu) one sample at a time
int emu_model_eval(const ddpb_model_desc* d, const ModelEvalArgs* A) {
  ModelDev m;
  std::string err;
  if (model_from_desc(d, m, err) != DDPB_OK) return -1;
  if (m.kind == DDPB_MODEL_PENDULUM) {
    for (int s = 0; s < A->n; ++s) model_eval_pendulum(m, *A, s);
    return 0;
  }
  EMU_NJ(m.n, {
    std::vector<double> qrec(QRec<NJ>::SIZE);
    std::vector<double> vrec(VRec<NJ>::SIZE);
    for (int s = 0; s < A->n; ++s) model_eval_sample<NJ>(m, *A, s, qrec.data(), vrec.data());
  });
  return 0;
}

}  // extern "C"
