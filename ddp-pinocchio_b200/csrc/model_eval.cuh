// model_eval.cuh -- the Model concept's own functions, batched: one thread evaluates one sample.
//
// The reference's dynamics / constraints call these per knot on one shared pinocchio::Data
// (/root/reference/include/ddp/pinocchio_model.hpp:147-184, pinocchio_model.ipp:336-462:
//  dynamics_aba -> pinocchio::aba, d_dynamics_aba -> computeABADerivatives, inverse_dynamics_rnea -> rnea,
//  frame_coordinates -> framesForwardKinematics, d_frame_coordinates -> getFrameJacobian(WORLD);
//  pendulum_model.hpp:105-130).  The hot path never goes through them one sample at a time (linearize.cuh evaluates
// whole stencils per knot), but a user of the reference's Model concept finds them here, on the device, for n
// independent samples per call -- the same rigid-body tasks (rbd.cuh) the linearization is made of, so these entry
// points are also the direct parity check of that layer against the oracle's ABA / RNEA / derivatives.
#pragma once
#include "batch.cuh"

namespace ddpb {

// device (or, in tests/hostemu, host) pointers, sample-major; nullptr = not requested
struct ModelEvalArgs {
  int n;                 // samples
  int frame, jac_mode;   // frame functions: frame index, DDPB_FJ_WORLD (the reference's convention) / DDPB_FJ_LOCAL_WORLD_ALIGNED
  const double* q;       // [n][nq]
  const double* v;       // [n][nv]
  const double* tau;     // [n][nv]   torques (forward dynamics)
  const double* acc;     // [n][nv]   accelerations (inverse dynamics)
  double* a;             // [n][nv]        dynamics_aba
  double* a_dq;          // [n][nv * nv]   d_dynamics_aba, column-major per sample
  double* a_dv;          // [n][nv * nv]
  double* a_dtau;        // [n][nv * nv]   = M^-1
  double* tau_out;       // [n][nv]        inverse_dynamics_rnea
  double* p3;            // [n][3]         frame_coordinates
  double* J;             // [n][3 * nv]    d_frame_coordinates, column-major per sample
};

}  // namespace ddpb

DDPB_VARIANT_BEGIN

// tau = ID(q, qd, qdd) on the q-data of q_task (world-frame recursive Newton-Euler: the forward sweep of v_task with the
// accelerations given, then the subtree forces projected on the joint axes)
template <int NJ> DDPB_HD void rnea_task(const ModelDev& m, const double* qrec, const double* qd, const double* qdd, double* tau) {
  using Q = QRec<NJ>;
  double f[NJ][6];
  double v[6] = {0, 0, 0, 0, 0, 0};
  double a[6] = {0.0, 0.0, 0.0, -m.grav[0], -m.grav[1], -m.grav[2]};
#pragma unroll
  for (int i = 0; i < NJ; ++i) {
    const double* S = qrec + Q::S + 6 * i;
    double sq[6], c[6], h[6], Ya[6], vh[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) { sq[k] = S[k] * qd[i]; v[k] += sq[k]; }
    crm(v, sq, c);
#pragma unroll
    for (int k = 0; k < 6; ++k) { a[k] += c[k]; a[k] += S[k] * qdd[i]; }
    y_mul(qrec + Q::Y + 10 * i, v, h);
    y_mul(qrec + Q::Y + 10 * i, a, Ya);
    crf(v, h, vh);
#pragma unroll
    for (int k = 0; k < 6; ++k) f[i][k] = Ya[k] + vh[k];
  }
  double F[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int i = NJ - 1; i >= 0; --i) {
#pragma unroll
    for (int k = 0; k < 6; ++k) F[k] += f[i][k];
    tau[i] = dot6(qrec + Q::S + 6 * i, F);
  }
}

// sample s of a chain of NJ joints; qrec / vrec: per-thread workspaces of QRec<NJ>::SIZE / VRec<NJ>::SIZE doubles
template <int NJ> DDPB_HD void model_eval_sample(const ModelDev& m, const ModelEvalArgs& A, int s, double* qrec, double* vrec) {
  const double* q = A.q + (size_t)s * NJ;
  const bool fwd = A.a || A.a_dq || A.a_dv;
  if (fwd || A.a_dtau || A.tau_out) q_task<NJ>(m, q, qrec);
  if (A.tau_out) {
    double tau[NJ];
    rnea_task<NJ>(m, qrec, A.v + (size_t)s * NJ, A.acc + (size_t)s * NJ, tau);
#pragma unroll
    for (int k = 0; k < NJ; ++k) A.tau_out[(size_t)s * NJ + k] = tau[k];
  }
  double Mi[NJ * NJ];
  if (A.a_dq || A.a_dv || A.a_dtau) {
#pragma unroll
    for (int c = 0; c < NJ; ++c) minv_col<NJ>(qrec, c, Mi + c * NJ);
    if (A.a_dtau) {
#pragma unroll
      for (int k = 0; k < NJ * NJ; ++k) A.a_dtau[(size_t)s * NJ * NJ + k] = Mi[k];
    }
  }
  if (fwd) {
    double qdd[NJ];
    v_task<NJ, true>(m, qrec, A.v + (size_t)s * NJ, A.tau + (size_t)s * NJ, qdd, vrec);
    if (A.a) {
#pragma unroll
      for (int k = 0; k < NJ; ++k) A.a[(size_t)s * NJ + k] = qdd[k];
    }
    // d a / d q = -M^-1 d tau / d q at the ABA solution, likewise for qdot (what phases 3-4 of the linearization do)
#pragma unroll 1
    for (int wrt = 0; wrt < 2; ++wrt) {
      double* out = wrt ? A.a_dv : A.a_dq;
      if (!out) continue;
#pragma unroll 1
      for (int j = 0; j < NJ; ++j) {
        double col[NJ];
        col_task<NJ>(qrec, vrec, j, wrt, col);
#pragma unroll
        for (int r = 0; r < NJ; ++r) {
          double acc = 0.0;
#pragma unroll
          for (int k = 0; k < NJ; ++k) acc += Mi[k * NJ + r] * col[k];
          out[(size_t)s * NJ * NJ + j * NJ + r] = -acc;
        }
      }
    }
  }
  if (A.p3 || A.J) {
    double p3[3], J[3 * NJ];
    frame_task<NJ>(m, q, A.frame, A.jac_mode, p3, A.J ? J : nullptr);
    if (A.p3) { A.p3[(size_t)s * 3] = p3[0]; A.p3[(size_t)s * 3 + 1] = p3[1]; A.p3[(size_t)s * 3 + 2] = p3[2]; }
    if (A.J) {
#pragma unroll
      for (int k = 0; k < 3 * NJ; ++k) A.J[(size_t)s * 3 * NJ + k] = J[k];
    }
  }
}

// the analytic pendulum (pendulum_model.hpp:105-130): nq = nv = 1, no frames, no inverse dynamics in its concept
DDPB_HD void model_eval_pendulum(const ModelDev& m, const ModelEvalArgs& A, int s) {
  const double q = A.q[s];
  if (A.a) A.a[s] = pend_acc(m, q, A.tau[s]);
  if (A.a_dq) A.a_dq[s] = pend_dacc_dq(m, q);
  if (A.a_dv) A.a_dv[s] = 0.0;
  if (A.a_dtau) A.a_dtau[s] = 1.0 / m.pend_m;
}

DDPB_VARIANT_END
