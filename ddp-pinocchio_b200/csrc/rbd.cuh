// rbd.cuh -- rigid-body dynamics of a serial revolute chain, written as small independent
// per-thread tasks whose inputs/outputs live in shared-memory records.
//
// Replaces the third-party Pinocchio calls behind the reference's Model concept
// (/root/reference/include/ddp/pinocchio_model.ipp:336-400 aba / computeABADerivatives,
// :417-462 frame position / WORLD frame Jacobian) and pendulum_model.hpp:105-130.
//
// Work split (one CTA = one knot, see linearize.cuh):
//   q_task        : everything that depends on q only -- joint rotations, articulated-body
//                   inertias IA_i, U_i = IA_i S_i, 1/D_i.  Shared by all evaluation points with
//                   the same q (velocity / control perturbations do not change it).
//   v_task        : the velocity/torque dependent sweeps of the articulated-body algorithm ->
//                   joint accelerations (+ per-body v, X a_parent, subtree force F for tangents).
//   tangent_col   : one column of d tau / d q or d tau / d qdot of inverse dynamics at the ABA
//                   solution (forward-mode tangent of RNEA restricted to bodies i >= j).
//   minv_col      : one column of M^-1(q) (unit-torque ABA sweeps on the cached U, 1/D).
// Spatial vectors are [angular; linear] in body coordinates; X_i = (E_i, r_i) maps parent to
// child coordinates with E_i = (Rp_i * Rodrigues(axis_i, q_i))^T and r_i = placement translation.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define DDPB_HD __host__ __device__ __forceinline__
#else
#define DDPB_HD inline
#endif

namespace ddpb {

constexpr int MAXJ = 7;
constexpr int MAXF = 16;
constexpr int MAXE = 7;

struct ModelDev {
  int kind, n, nframes, pad_;
  double pend_m, pend_l;
  double grav[3];
  double Rp[MAXJ][9];
  double pp[MAXJ][3];
  double axis[MAXJ][3];
  double Io[MAXJ][6];   // rotational inertia about the body-frame origin: xx,xy,xz,yy,yz,zz
  double h[MAXJ][3];    // mass * com
  double mass[MAXJ];
  int fparent[MAXF];
  double fp[MAXF][3];
};

// ---- small vector helpers -------------------------------------------------------------------
DDPB_HD void cross3(const double* a, const double* b, double* o) {
  double x = a[1] * b[2] - a[2] * b[1];
  double y = a[2] * b[0] - a[0] * b[2];
  double z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
DDPB_HD void cross3_add(const double* a, const double* b, double* o) {
  o[0] += a[1] * b[2] - a[2] * b[1];
  o[1] += a[2] * b[0] - a[0] * b[2];
  o[2] += a[0] * b[1] - a[1] * b[0];
}
DDPB_HD void m3v(const double* M, const double* v, double* o) {  // row-major M * v
  double x = M[0] * v[0] + M[1] * v[1] + M[2] * v[2];
  double y = M[3] * v[0] + M[4] * v[1] + M[5] * v[2];
  double z = M[6] * v[0] + M[7] * v[1] + M[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
DDPB_HD void m3tv(const double* M, const double* v, double* o) {  // M^T * v
  double x = M[0] * v[0] + M[3] * v[1] + M[6] * v[2];
  double y = M[1] * v[0] + M[4] * v[1] + M[7] * v[2];
  double z = M[2] * v[0] + M[5] * v[1] + M[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
DDPB_HD void sym3v(const double* S, const double* v, double* o) {  // S = xx,xy,xz,yy,yz,zz
  double x = S[0] * v[0] + S[1] * v[1] + S[2] * v[2];
  double y = S[1] * v[0] + S[3] * v[1] + S[4] * v[2];
  double z = S[2] * v[0] + S[4] * v[1] + S[5] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
// motion transform parent -> child:  [E w ; E (v - r x w)]
DDPB_HD void xf_motion(const double* E, const double* r, const double* m, double* o) {
  double t[3], u[3];
  cross3(r, m, t);
  u[0] = m[3] - t[0]; u[1] = m[4] - t[1]; u[2] = m[5] - t[2];
  m3v(E, m, o);
  m3v(E, u, o + 3);
}
// force transform child -> parent (transpose):  [E^T n + r x E^T f ; E^T f], accumulated into o
DDPB_HD void xf_force_T_add(const double* E, const double* r, const double* f, double* o) {
  double n[3], l[3];
  m3tv(E, f, n);
  m3tv(E, f + 3, l);
  cross3_add(r, l, n);
  o[0] += n[0]; o[1] += n[1]; o[2] += n[2];
  o[3] += l[0]; o[4] += l[1]; o[5] += l[2];
}
// motion cross a x b, force cross a x* f
DDPB_HD void crm(const double* a, const double* b, double* o) {
  double t0[3], t1[3];
  cross3(a, b, t0);
  cross3(a, b + 3, t1);
  cross3_add(a + 3, b, t1);
  o[0] = t0[0]; o[1] = t0[1]; o[2] = t0[2]; o[3] = t1[0]; o[4] = t1[1]; o[5] = t1[2];
}
DDPB_HD void crf(const double* a, const double* f, double* o) {
  double t0[3], t1[3];
  cross3(a, f, t0);
  cross3_add(a + 3, f + 3, t0);
  cross3(a, f + 3, t1);
  o[0] = t0[0]; o[1] = t0[1]; o[2] = t0[2]; o[3] = t1[0]; o[4] = t1[1]; o[5] = t1[2];
}
// motion cross with a pure-rotation second operand b = [ax*s; 0]
DDPB_HD void crm_axis(const double* a, const double* ax, double s, double* o) {
  double b[3] = {ax[0] * s, ax[1] * s, ax[2] * s};
  cross3(a, b, o);
  cross3(a + 3, b, o + 3);
}
// rigid-body spatial inertia times motion vector: [Io w + h x v ; m v - h x w]
DDPB_HD void inertia_mul(const ModelDev& m, int i, const double* v, double* o) {
  double a[3], l[3];
  sym3v(m.Io[i], v, a);
  cross3_add(m.h[i], v + 3, a);
  cross3(m.h[i], v, l);
  o[0] = a[0]; o[1] = a[1]; o[2] = a[2];
  o[3] = m.mass[i] * v[3] - l[0]; o[4] = m.mass[i] * v[4] - l[1]; o[5] = m.mass[i] * v[5] - l[2];
}

DDPB_HD void sincos_(double q, double* s, double* c) {
#if defined(__CUDA_ARCH__)
  sincos(q, s, c);
#else
  *s = sin(q); *c = cos(q);
#endif
}

// E_i(q) = (Rp_i * R_axis(q))^T, row-major
DDPB_HD void joint_E(const ModelDev& m, int i, double q, double* E) {
  double s, c;
  sincos_(q, &s, &c);
  const double v = 1.0 - c;
  const double* a = m.axis[i];
  const double RJ[9] = {c + v * a[0] * a[0],        v * a[0] * a[1] - s * a[2], v * a[0] * a[2] + s * a[1],
                        v * a[1] * a[0] + s * a[2], c + v * a[1] * a[1],        v * a[1] * a[2] - s * a[0],
                        v * a[2] * a[0] - s * a[1], v * a[2] * a[1] + s * a[0], c + v * a[2] * a[2]};
  const double* Rp = m.Rp[i];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) E[cc * 3 + r] = Rp[r * 3 + 0] * RJ[0 * 3 + cc] + Rp[r * 3 + 1] * RJ[1 * 3 + cc] + Rp[r * 3 + 2] * RJ[2 * 3 + cc];
}

// ---- record layouts (offsets in doubles; sizes padded to an odd count so that consecutive
// tasks' records start in different shared-memory banks) -------------------------------------
template <int NJ> struct QRec {
  static constexpr int E = 0;                 // [NJ][9]
  static constexpr int IA = E + 9 * NJ;       // [NJ][21]: A(6 sym) | B(9 row-major) | C(6 sym)
  static constexpr int U = IA + 21 * NJ;      // [NJ][6]
  static constexpr int DINV = U + 6 * NJ;     // [NJ]
  static constexpr int MINV = DINV + NJ;      // [NJ][NJ] column-major
  static constexpr int RAW = MINV + NJ * NJ;
  static constexpr int SIZE = RAW | 1;
};
template <int NJ> struct VRec {
  static constexpr int V = 0;                 // [NJ][6] body velocities
  static constexpr int XA = V + 6 * NJ;       // [NJ][6] X_i a_parent
  static constexpr int F = XA + 6 * NJ;       // [NJ][6] subtree forces of RNEA at the ABA solution
  static constexpr int QD = F + 6 * NJ;       // [NJ]
  static constexpr int QDD = QD + NJ;         // [NJ]
  static constexpr int DQ = QDD + NJ;         // [NJ][NJ] d tau/d q, then da/dq (column-major)
  static constexpr int DV = DQ + NJ * NJ;     // [NJ][NJ] d tau/d qdot, then da/dqdot
  static constexpr int RAW = DV + NJ * NJ;
  static constexpr int SIZE = RAW | 1;
};

// ---- q_task ------------------------------------------------------------------------------------
template <int NJ> DDPB_HD void q_task(const ModelDev& m, const double* q, double* rec) {
  using R = QRec<NJ>;
#pragma unroll
  for (int i = 0; i < NJ; ++i) {
    double E[9];
    joint_E(m, i, q[i], E);
#pragma unroll
    for (int k = 0; k < 9; ++k) rec[R::E + 9 * i + k] = E[k];
  }
  // articulated inertia of the current body: A (ang-ang, sym), B (ang-lin), C (lin-lin, sym)
  double A[6], B[9], C[6];
#pragma unroll
  for (int i = NJ - 1; i >= 0; --i) {
    const double* hh = m.h[i];
    if (i == NJ - 1) {
#pragma unroll
      for (int k = 0; k < 6; ++k) { A[k] = m.Io[i][k]; C[k] = 0.0; }
#pragma unroll
      for (int k = 0; k < 9; ++k) B[k] = 0.0;
    } else {
#pragma unroll
      for (int k = 0; k < 6; ++k) A[k] += m.Io[i][k];
    }
    // B += skew(h), C += mass * I
    B[1] -= hh[2]; B[2] += hh[1]; B[3] += hh[2]; B[5] -= hh[0]; B[6] -= hh[1]; B[7] += hh[0];
    C[0] += m.mass[i]; C[3] += m.mass[i]; C[5] += m.mass[i];
    double* ia = rec + R::IA + 21 * i;
#pragma unroll
    for (int k = 0; k < 6; ++k) { ia[k] = A[k]; ia[15 + k] = C[k]; }
#pragma unroll
    for (int k = 0; k < 9; ++k) ia[6 + k] = B[k];
    const double* ax = m.axis[i];
    double Ua[3], Ul[3];
    sym3v(A, ax, Ua);
    m3tv(B, ax, Ul);
    const double D = ax[0] * Ua[0] + ax[1] * Ua[1] + ax[2] * Ua[2];
    const double Dinv = 1.0 / D;
    rec[R::U + 6 * i + 0] = Ua[0]; rec[R::U + 6 * i + 1] = Ua[1]; rec[R::U + 6 * i + 2] = Ua[2];
    rec[R::U + 6 * i + 3] = Ul[0]; rec[R::U + 6 * i + 4] = Ul[1]; rec[R::U + 6 * i + 5] = Ul[2];
    rec[R::DINV + i] = Dinv;
    if (i > 0) {
      // Ia = IA - U D^-1 U^T
      const double sa[3] = {Ua[0] * Dinv, Ua[1] * Dinv, Ua[2] * Dinv};
      const double sl[3] = {Ul[0] * Dinv, Ul[1] * Dinv, Ul[2] * Dinv};
      A[0] -= sa[0] * Ua[0]; A[1] -= sa[0] * Ua[1]; A[2] -= sa[0] * Ua[2];
      A[3] -= sa[1] * Ua[1]; A[4] -= sa[1] * Ua[2]; A[5] -= sa[2] * Ua[2];
      C[0] -= sl[0] * Ul[0]; C[1] -= sl[0] * Ul[1]; C[2] -= sl[0] * Ul[2];
      C[3] -= sl[1] * Ul[1]; C[4] -= sl[1] * Ul[2]; C[5] -= sl[2] * Ul[2];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) B[r * 3 + c] -= sa[r] * Ul[c];
      // rotate the three blocks into the parent orientation:  M' = E^T M E
      const double* E = rec + R::E + 9 * i;
      double Af[9] = {A[0], A[1], A[2], A[1], A[3], A[4], A[2], A[4], A[5]};
      double Cf[9] = {C[0], C[1], C[2], C[1], C[3], C[4], C[2], C[4], C[5]};
      double T[9], Ab[9], Bb[9], Cb[9];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) T[r * 3 + c] = Af[r * 3] * E[c] + Af[r * 3 + 1] * E[3 + c] + Af[r * 3 + 2] * E[6 + c];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) Ab[r * 3 + c] = E[r] * T[c] + E[3 + r] * T[3 + c] + E[6 + r] * T[6 + c];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) T[r * 3 + c] = B[r * 3] * E[c] + B[r * 3 + 1] * E[3 + c] + B[r * 3 + 2] * E[6 + c];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) Bb[r * 3 + c] = E[r] * T[c] + E[3 + r] * T[3 + c] + E[6 + r] * T[6 + c];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) T[r * 3 + c] = Cf[r * 3] * E[c] + Cf[r * 3 + 1] * E[3 + c] + Cf[r * 3 + 2] * E[6 + c];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) Cb[r * 3 + c] = E[r] * T[c] + E[3 + r] * T[3 + c] + E[6 + r] * T[6 + c];
      // shift to the parent origin (r = placement translation):
      //   C' = Cb ; B' = Bb + rx Cb ; A' = Ab - Bb rx + rx B'^T
      const double* rr = m.pp[i];
      double Bn[9];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        Bn[0 * 3 + c] = Bb[0 * 3 + c] + (-rr[2] * Cb[3 + c] + rr[1] * Cb[6 + c]);
        Bn[1 * 3 + c] = Bb[1 * 3 + c] + (rr[2] * Cb[c] - rr[0] * Cb[6 + c]);
        Bn[2 * 3 + c] = Bb[2 * 3 + c] + (-rr[1] * Cb[c] + rr[0] * Cb[3 + c]);
      }
      double An[9];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        // (Bb rx)[r][:]
        const double b0 = Bb[r * 3 + 1] * rr[2] - Bb[r * 3 + 2] * rr[1];
        const double b1 = -Bb[r * 3 + 0] * rr[2] + Bb[r * 3 + 2] * rr[0];
        const double b2 = Bb[r * 3 + 0] * rr[1] - Bb[r * 3 + 1] * rr[0];
        An[r * 3 + 0] = Ab[r * 3 + 0] - b0;
        An[r * 3 + 1] = Ab[r * 3 + 1] - b1;
        An[r * 3 + 2] = Ab[r * 3 + 2] - b2;
      }
      // + rx B'^T : row r of rx times column c of B'^T (= row c of B')
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        An[0 * 3 + c] += -rr[2] * Bn[c * 3 + 1] + rr[1] * Bn[c * 3 + 2];
        An[1 * 3 + c] += rr[2] * Bn[c * 3 + 0] - rr[0] * Bn[c * 3 + 2];
        An[2 * 3 + c] += -rr[1] * Bn[c * 3 + 0] + rr[0] * Bn[c * 3 + 1];
      }
      A[0] = An[0]; A[1] = An[1]; A[2] = An[2]; A[3] = An[4]; A[4] = An[5]; A[5] = An[8];
      C[0] = Cb[0]; C[1] = Cb[1]; C[2] = Cb[2]; C[3] = Cb[4]; C[4] = Cb[5]; C[5] = Cb[8];
#pragma unroll
      for (int k = 0; k < 9; ++k) B[k] = Bn[k];
    }
  }
}

// IA_i * c  with IA stored as A | B | C
DDPB_HD void ia_mul(const double* ia, const double* c, double* o) {
  double a[3], l[3], t[3];
  sym3v(ia, c, a);
  m3v(ia + 6, c + 3, t);
  a[0] += t[0]; a[1] += t[1]; a[2] += t[2];
  m3tv(ia + 6, c, l);
  sym3v(ia + 15, c + 3, t);
  o[0] = a[0]; o[1] = a[1]; o[2] = a[2];
  o[3] = l[0] + t[0]; o[4] = l[1] + t[1]; o[5] = l[2] + t[2];
}

// ---- v_task --------------------------------------------------------------------------------------
// qdd_out[NJ] always; when vrec != nullptr also the per-body data the tangent columns need.
template <int NJ> DDPB_HD void v_task(const ModelDev& m, const double* qrec, const double* qd, const double* tau, double* qdd_out, double* vrec) {
  using Q = QRec<NJ>;
  using V = VRec<NJ>;
  double v[NJ][6], pA[NJ][6], u[NJ];
#pragma unroll
  for (int i = 0; i < NJ; ++i) {
    const double* ax = m.axis[i];
    if (i == 0) {
      v[0][0] = ax[0] * qd[0]; v[0][1] = ax[1] * qd[0]; v[0][2] = ax[2] * qd[0];
      v[0][3] = 0.0; v[0][4] = 0.0; v[0][5] = 0.0;
    } else {
      xf_motion(qrec + Q::E + 9 * i, m.pp[i], v[i - 1], v[i]);
      v[i][0] += ax[0] * qd[i]; v[i][1] += ax[1] * qd[i]; v[i][2] += ax[2] * qd[i];
    }
    double hv[6];
    inertia_mul(m, i, v[i], hv);
    crf(v[i], hv, pA[i]);
  }
#pragma unroll
  for (int i = NJ - 1; i >= 0; --i) {
    const double* ax = m.axis[i];
    u[i] = tau[i] - (ax[0] * pA[i][0] + ax[1] * pA[i][1] + ax[2] * pA[i][2]);
    if (i > 0) {
      double c[6], Iac[6], pa[6];
      crm_axis(v[i], ax, qd[i], c);
      const double* U = qrec + Q::U + 6 * i;
      const double Uc = U[0] * c[0] + U[1] * c[1] + U[2] * c[2] + U[3] * c[3] + U[4] * c[4] + U[5] * c[5];
      const double t = qrec[Q::DINV + i] * (u[i] - Uc);
      ia_mul(qrec + Q::IA + 21 * i, c, Iac);
#pragma unroll
      for (int k = 0; k < 6; ++k) pa[k] = pA[i][k] + Iac[k] + U[k] * t;
      xf_force_T_add(qrec + Q::E + 9 * i, m.pp[i], pa, pA[i - 1]);
    }
  }
  double a[6] = {0.0, 0.0, 0.0, -m.grav[0], -m.grav[1], -m.grav[2]};
  double F[NJ][6];
#pragma unroll
  for (int i = 0; i < NJ; ++i) {
    const double* ax = m.axis[i];
    double Xa[6], c[6];
    xf_motion(qrec + Q::E + 9 * i, m.pp[i], a, Xa);
    crm_axis(v[i], ax, qd[i], c);
#pragma unroll
    for (int k = 0; k < 6; ++k) a[k] = Xa[k] + c[k];
    const double* U = qrec + Q::U + 6 * i;
    const double Ua = U[0] * a[0] + U[1] * a[1] + U[2] * a[2] + U[3] * a[3] + U[4] * a[4] + U[5] * a[5];
    const double qdd = qrec[Q::DINV + i] * (u[i] - Ua);
    a[0] += ax[0] * qdd; a[1] += ax[1] * qdd; a[2] += ax[2] * qdd;
    qdd_out[i] = qdd;
    if (vrec) {
#pragma unroll
      for (int k = 0; k < 6; ++k) { vrec[V::V + 6 * i + k] = v[i][k]; vrec[V::XA + 6 * i + k] = Xa[k]; }
      vrec[V::QD + i] = qd[i];
      vrec[V::QDD + i] = qdd;
      double Ia[6], hv[6], vf[6];
      inertia_mul(m, i, a, Ia);
      inertia_mul(m, i, v[i], hv);
      crf(v[i], hv, vf);
#pragma unroll
      for (int k = 0; k < 6; ++k) F[i][k] = Ia[k] + vf[k];
    }
  }
  if (vrec) {
#pragma unroll
    for (int i = NJ - 1; i >= 0; --i) {
      if (i > 0) xf_force_T_add(qrec + Q::E + 9 * i, m.pp[i], F[i], F[i - 1]);
#pragma unroll
      for (int k = 0; k < 6; ++k) vrec[V::F + 6 * i + k] = F[i][k];
    }
  }
}

// ---- tangent_col -----------------------------------------------------------------------------------
// column j of d tau/d q (wrt_v = 0) or d tau/d qdot (wrt_v = 1) of RNEA(q, qdot, qddot = ABA result)
template <int NJ> DDPB_HD void tangent_col(const ModelDev& m, const double* qrec, const double* vrec, int j, int wrt_v, double* out /*[NJ]*/) {
  using Q = QRec<NJ>;
  using V = VRec<NJ>;
  double dF[NJ][6];
  double dv[6] = {0, 0, 0, 0, 0, 0}, da[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < NJ; ++i) {
    if (i < j) {
#pragma unroll
      for (int k = 0; k < 6; ++k) dF[i][k] = 0.0;
      continue;
    }
    const double* ax = m.axis[i];
    const double* vi = vrec + V::V + 6 * i;
    const double qdi = vrec[V::QD + i];
    if (i == j) {
      if (!wrt_v) {
        // d(X_j m)/dq_j = (X_j m) x S_j  with  X_j v_parent = v_j - S_j qd_j,  X_j a_parent = XA_j
        double Xv[6] = {vi[0] - ax[0] * qdi, vi[1] - ax[1] * qdi, vi[2] - ax[2] * qdi, vi[3], vi[4], vi[5]};
        crm_axis(Xv, ax, 1.0, dv);
        crm_axis(vrec + V::XA + 6 * i, ax, 1.0, da);
      } else {
        dv[0] = ax[0]; dv[1] = ax[1]; dv[2] = ax[2]; dv[3] = 0.0; dv[4] = 0.0; dv[5] = 0.0;
        crm_axis(vi, ax, 1.0, da);
      }
    } else {
      double t[6];
      xf_motion(qrec + Q::E + 9 * i, m.pp[i], dv, t);
#pragma unroll
      for (int k = 0; k < 6; ++k) dv[k] = t[k];
      xf_motion(qrec + Q::E + 9 * i, m.pp[i], da, t);
#pragma unroll
      for (int k = 0; k < 6; ++k) da[k] = t[k];
    }
    {
      double t[6];
      crm_axis(dv, ax, qdi, t);
#pragma unroll
      for (int k = 0; k < 6; ++k) da[k] += t[k];
    }
    double Ida[6], Idv[6], hv[6], t1[6], t2[6];
    inertia_mul(m, i, da, Ida);
    inertia_mul(m, i, dv, Idv);
    inertia_mul(m, i, vi, hv);
    crf(dv, hv, t1);
    crf(vi, Idv, t2);
#pragma unroll
    for (int k = 0; k < 6; ++k) dF[i][k] = Ida[k] + t1[k] + t2[k];
  }
#pragma unroll
  for (int i = NJ - 1; i >= 0; --i) {
    const double* ax = m.axis[i];
    out[i] = ax[0] * dF[i][0] + ax[1] * dF[i][1] + ax[2] * dF[i][2];
    if (i > 0 && i >= j) {
      double tot[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) tot[k] = dF[i][k];
      if (i == j && !wrt_v) {
        // + S_j x* F_j
        const double* Fj = vrec + V::F + 6 * i;
        cross3_add(ax, Fj, tot);
        cross3_add(ax, Fj + 3, tot + 3);
      }
      xf_force_T_add(qrec + Q::E + 9 * i, m.pp[i], tot, dF[i - 1]);
    } else if (i > 0) {
      // bodies above j only pass the accumulated force on
      xf_force_T_add(qrec + Q::E + 9 * i, m.pp[i], dF[i], dF[i - 1]);
    }
  }
}

// ---- minv_col ---------------------------------------------------------------------------------------
template <int NJ> DDPB_HD void minv_col(const ModelDev& m, const double* qrec, int col, double* out /*[NJ]*/) {
  using Q = QRec<NJ>;
  double pA[6] = {0, 0, 0, 0, 0, 0};
  double u[NJ];
#pragma unroll
  for (int i = NJ - 1; i >= 0; --i) {
    const double* ax = m.axis[i];
    if (i > col) { u[i] = 0.0; continue; }
    u[i] = (i == col ? 1.0 : 0.0) - (ax[0] * pA[0] + ax[1] * pA[1] + ax[2] * pA[2]);
    if (i > 0) {
      const double* U = qrec + Q::U + 6 * i;
      const double t = qrec[Q::DINV + i] * u[i];
      double pa[6], np[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int k = 0; k < 6; ++k) pa[k] = pA[k] + U[k] * t;
      xf_force_T_add(qrec + Q::E + 9 * i, m.pp[i], pa, np);
#pragma unroll
      for (int k = 0; k < 6; ++k) pA[k] = np[k];
    }
  }
  double a[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < NJ; ++i) {
    const double* ax = m.axis[i];
    if (i > 0) {
      double t[6];
      xf_motion(qrec + Q::E + 9 * i, m.pp[i], a, t);
#pragma unroll
      for (int k = 0; k < 6; ++k) a[k] = t[k];
    }
    const double* U = qrec + Q::U + 6 * i;
    const double Ua = U[0] * a[0] + U[1] * a[1] + U[2] * a[2] + U[3] * a[3] + U[4] * a[4] + U[5] * a[5];
    const double qdd = qrec[Q::DINV + i] * (u[i] - Ua);
    a[0] += ax[0] * qdd; a[1] += ax[1] * qdd; a[2] += ax[2] * qdd;
    out[i] = qdd;
  }
}

// ---- frame position and its "Jacobian" (pinocchio_model.ipp:417-462) ----------------------------
// J is 3 x NJ column-major.  mode 0 = WORLD convention (p_j x w_j, the reference's quirk),
// mode 1 = LOCAL_WORLD_ALIGNED (w_j x (p_f - p_j), the true dp/dq).
template <int NJ> DDPB_HD void frame_task(const ModelDev& m, const double* q, int frame, int mode, double* p3, double* J) {
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, p[3] = {0, 0, 0};
  double w[NJ][3], pj[NJ][3];
  const int jf = m.fparent[frame];
#pragma unroll
  for (int i = 0; i < NJ; ++i) {
    if (i > jf) continue;
    double E[9], t[3], Rn[9];
    joint_E(m, i, q[i], E);
    m3v(R, m.pp[i], t);
    p[0] += t[0]; p[1] += t[1]; p[2] += t[2];
    // R_world_i = R_world_parent * E^T
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) Rn[r * 3 + c] = R[r * 3] * E[c * 3] + R[r * 3 + 1] * E[c * 3 + 1] + R[r * 3 + 2] * E[c * 3 + 2];
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = Rn[k];
    m3v(R, m.axis[i], w[i]);
    pj[i][0] = p[0]; pj[i][1] = p[1]; pj[i][2] = p[2];
  }
  double t[3];
  m3v(R, m.fp[frame], t);
  p3[0] = p[0] + t[0]; p3[1] = p[1] + t[1]; p3[2] = p[2] + t[2];
  if (J) {
#pragma unroll
    for (int i = 0; i < NJ; ++i) {
      double lin[3] = {0, 0, 0};
      if (i <= jf) {
        if (mode == 0) {
          cross3(pj[i], w[i], lin);
        } else {
          double d[3] = {p3[0] - pj[i][0], p3[1] - pj[i][1], p3[2] - pj[i][2]};
          cross3(w[i], d, lin);
        }
      }
      J[3 * i] = lin[0]; J[3 * i + 1] = lin[1]; J[3 * i + 2] = lin[2];
    }
  }
}

// ---- pendulum (pendulum_model.hpp:105-130) ------------------------------------------------------
DDPB_HD double pend_acc(const ModelDev& m, double q, double tau) { return -9.81 / m.pend_l * sin(q) + tau / m.pend_m; }
DDPB_HD double pend_dacc_dq(const ModelDev& m, double q) { return -9.81 / m.pend_l * cos(q); }

}  // namespace ddpb
