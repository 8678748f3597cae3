// rbd.cuh -- rigid-body dynamics of a serial chain of revolute / prismatic joints, written as small independent
// per-thread tasks whose inputs/outputs live in shared-memory records.
//
// Replaces the third-party Pinocchio calls behind the reference's Model concept
// (/root/reference/include/ddp/pinocchio_model.ipp:336-400 aba / computeABADerivatives,
// :417-462 frame position / WORLD frame Jacobian) and pendulum_model.hpp:105-130.
//
// Work split (one CTA = one knot, see linearize.cuh):
//   q_task        : everything that depends on q only -- world-frame joint axes S_i and body
//                   inertias Y_i, articulated-body inertias IA_i, U_i = IA_i S_i, 1/D_i.  Shared by
//                   all evaluation points with the same q (velocity / control perturbations do
//                   not change it).
//   v_task        : the velocity/torque dependent sweeps of the articulated-body algorithm ->
//                   joint accelerations (+ per-body v, a, momentum h, subtree force F for the columns).
//   col_task      : one column of d tau / d q or d tau / d qdot of inverse dynamics at the ABA
//                   solution (closed form per (body, joint) in world coordinates).
//   minv_col      : one column of M^-1(q) (unit-torque ABA sweeps on the cached S, U, 1/D).
// Spatial vectors are [angular; linear] Pluecker coordinates in the WORLD frame (linear part taken at
// the world origin), so quantities of different bodies add without transforms.
#pragma once
#include <math.h>

#include "variant.cuh"

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
  int kind, n, nframes;
  int any_prismatic;    // some jtype[i] != 0 (host: picks the kernels of the HAS_PRISMATIC variant)
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
  // joint types (what pinocchio::urdf::buildModel makes of URDF revolute / prismatic joints): 0 = revolute about `axis`,
  // 1 = prismatic along it.  Only the q-dependent placement (joint_local_*) and the motion subspace S_i (body_world,
  // frame_task) depend on the type; everything downstream is written on S_i and holds for both.  The tests on jtype are
  // compiled only into the HAS_PRISMATIC variant of the device code (variant.cuh); any_prismatic selects that variant.
  int jtype[MAXJ + 1];
};

}  // namespace ddpb

DDPB_VARIANT_BEGIN

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
DDPB_HD void sincos_(double q, double* s, double* c) {
#if defined(__CUDA_ARCH__)
  sincos(q, s, c);
#else
  *s = sin(q); *c = cos(q);
#endif
}

// E_i(q) = (Rp_i * R_axis(q))^T, row-major (used by frame_task only: frame positions follow the body-frame recursion
// of the reference's forward kinematics)
DDPB_HD void joint_E(const ModelDev& m, int i, double q, double* E) {
  if (HAS_PRISMATIC && m.jtype[i] != 0) {  // prismatic: no rotation, E = Rp^T
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) E[cc * 3 + r] = m.Rp[i][r * 3 + cc];
    return;
  }
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

// ---- world-frame spatial algebra -----------------------------------------------------------------
// All spatial vectors of the tasks below are Pluecker coordinates in the WORLD frame, [angular; linear at the
// world origin].  A chain in world coordinates needs no per-joint transforms: parent/child quantities simply add,
// the articulated-body recursion is IA_parent += IA - U U^T / D, and the derivative of inverse dynamics with respect
// to q_j / qdot_j has a closed form per (body k, joint j) (see col_task).
//
// rigid-body inertia Y = [mass, h (3) = mass * com, I (6) = rotational inertia about the world origin: xx,xy,xz,yy,yz,zz]
DDPB_HD void y_mul(const double* Y, const double* v, double* o) {  // [I w + h x vl ; m vl - h x w]
  double a[3], l[3];
  sym3v(Y + 4, v, a);
  cross3_add(Y + 1, v + 3, a);
  cross3(Y + 1, v, l);
  o[0] = a[0]; o[1] = a[1]; o[2] = a[2];
  o[3] = Y[0] * v[3] - l[0]; o[4] = Y[0] * v[4] - l[1]; o[5] = Y[0] * v[5] - l[2];
}
// packed upper triangle of a symmetric 6x6: (r, c), r <= c  ->  r*6 - r(r-1)/2 + (c - r)
DDPB_HD constexpr int s6(int r, int c) { return r <= c ? r * 6 - (r * (r - 1)) / 2 + (c - r) : c * 6 - (c * (c - 1)) / 2 + (r - c); }
DDPB_HD void sym6_mul(const double* A, const double* x, double* o) {
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    double a = 0.0;
#pragma unroll
    for (int c = 0; c < 6; ++c) a += A[s6(r, c)] * x[c];
    o[r] = a;
  }
}
DDPB_HD double dot6(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5]; }

// ---- record layouts (offsets in doubles; sizes padded to an odd count so that consecutive
// tasks' records start in different shared-memory banks) -------------------------------------
template <int NJ> struct QRec {
  static constexpr int S = 0;                 // [NJ][6] joint axes (world)
  static constexpr int Y = S + 6 * NJ;        // [NJ][10] rigid-body inertias (world)
  static constexpr int U = Y + 10 * NJ;       // [NJ][6]  IA_i S_i
  static constexpr int DINV = U + 6 * NJ;     // [NJ]     1 / (S_i . U_i)
  // [NJ-1][21] articulated inertias (packed symmetric) of bodies 1 .. NJ-1, body i at IA + 21 (i - 1): the v-sweeps
  // need IA_i only to hand body i's bias force to its parent, so the root's is never stored
  static constexpr int IA = DINV + NJ;
  // [NJ][NJ] column-major, ALIASED onto the IA block: M^-1 is computed (minv_col: S, U, 1/D only) in the phase after
  // the last reader of IA, so the two never live at the same time -- 8 NJ^2 bytes of shared memory less per record
  static constexpr int MINV = IA;
  // scratch of the q sub-phases (QScratch, 24 NJ doubles): over U | DINV | IA, all written by q_articulate afterwards
  static constexpr int SCR = U;
  static constexpr int END_IA = IA + 21 * (NJ - 1), END_MINV = MINV + NJ * NJ, END_SCR = SCR + 24 * NJ;
  static constexpr int RAW = END_IA > END_MINV ? (END_IA > END_SCR ? END_IA : END_SCR) : (END_MINV > END_SCR ? END_MINV : END_SCR);
  static constexpr int SIZE = RAW | 1;
};
template <int NJ> struct VRec {
  static constexpr int V = 0;                 // [NJ][6] body velocities
  static constexpr int A = V + 6 * NJ;        // [NJ][6] body accelerations at the ABA solution (gravity folded in)
  static constexpr int H = A + 6 * NJ;        // [NJ][6] momenta Y_i v_i
  static constexpr int F = H + 6 * NJ;        // [NJ][6] subtree forces of inverse dynamics at the ABA solution
  static constexpr int DQ = F + 6 * NJ;       // [NJ][NJ] d tau/d q, then dt * da/dq (column-major); v_bias scratch before that
  static constexpr int DV = DQ + NJ * NJ;     // [NJ][NJ] d tau/d qdot, then I + dt * da/dqdot
  static constexpr int RAW = DV + NJ * NJ;
  static constexpr int SIZE = RAW | 1;
};

// local rotation of joint i:  Rl = Rp_i Rot(axis_i, q_i)   (row-major 3x3)
DDPB_HD void joint_local_rot(const ModelDev& m, int i, double q, double* Rl) {
  if (HAS_PRISMATIC && m.jtype[i] != 0) {  // prismatic: Rl = Rp_i
#pragma unroll
    for (int k = 0; k < 9; ++k) Rl[k] = m.Rp[i][k];
    return;
  }
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
    for (int cc = 0; cc < 3; ++cc) Rl[r * 3 + cc] = Rp[r * 3 + 0] * RJ[0 * 3 + cc] + Rp[r * 3 + 1] * RJ[1 * 3 + cc] + Rp[r * 3 + 2] * RJ[2 * 3 + cc];
}
// origin of joint frame i in parent coordinates:  pl = pp_i (revolute) ,  pp_i + Rp_i (axis_i q) (prismatic)
DDPB_HD void joint_local_trans(const ModelDev& m, int i, double q, double* pl) {
  pl[0] = m.pp[i][0]; pl[1] = m.pp[i][1]; pl[2] = m.pp[i][2];
  if (HAS_PRISMATIC && m.jtype[i] != 0) {
    const double d[3] = {m.axis[i][0] * q, m.axis[i][1] * q, m.axis[i][2] * q};
    double t[3];
    m3v(m.Rp[i], d, t);
    pl[0] += t[0]; pl[1] += t[1]; pl[2] += t[2];
  }
}
// world rotation / origin of joint frame i from those of its parent:  p <- p + R_parent pl_i ,  R <- R Rl_i
DDPB_HD void joint_compose(const double* pl, const double* Rl, double* R, double* p) {
  double t[3];
  m3v(R, pl, t);
  p[0] += t[0]; p[1] += t[1]; p[2] += t[2];
  double Rn[9];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) Rn[r * 3 + cc] = R[r * 3 + 0] * Rl[0 * 3 + cc] + R[r * 3 + 1] * Rl[1 * 3 + cc] + R[r * 3 + 2] * Rl[2 * 3 + cc];
#pragma unroll
  for (int k = 0; k < 9; ++k) R[k] = Rn[k];
}
DDPB_HD void joint_advance(const ModelDev& m, int i, double q, double* R, double* p) {
  double Rl[9], pl[3];
  joint_local_rot(m, i, q, Rl);
  joint_local_trans(m, i, q, pl);
  joint_compose(pl, Rl, R, p);
}
// joint axis S = [w ; p x w] (revolute) or [0 ; w] (prismatic) and rigid-body inertia Y of body i in world coordinates
// about the world origin.
// With h_p = R h_i (first moment about the joint origin), g = h_p + (m/2) p:
//   h_O = h_p + m p ,  I_O = R Io_i R^T + 2 (g . p) Id - (g p^T + p g^T)      (parallel-axis shift, no division by m)
DDPB_HD void body_world(const ModelDev& m, int i, const double* R, const double* p, double* S, double* Y) {
  double w[3], mo[3];
  m3v(R, m.axis[i], w);
  if (HAS_PRISMATIC && m.jtype[i] != 0) {
    S[0] = 0.0; S[1] = 0.0; S[2] = 0.0; S[3] = w[0]; S[4] = w[1]; S[5] = w[2];
  } else {
    cross3(p, w, mo);
    S[0] = w[0]; S[1] = w[1]; S[2] = w[2]; S[3] = mo[0]; S[4] = mo[1]; S[5] = mo[2];
  }
  const double mass = m.mass[i];
  double hp[3];
  m3v(R, m.h[i], hp);
  const double g[3] = {hp[0] + 0.5 * mass * p[0], hp[1] + 0.5 * mass * p[1], hp[2] + 0.5 * mass * p[2]};
  const double* Io = m.Io[i];
  const double If[9] = {Io[0], Io[1], Io[2], Io[1], Io[3], Io[4], Io[2], Io[4], Io[5]};
  double T[9];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) T[r * 3 + c] = R[r * 3 + 0] * If[0 * 3 + c] + R[r * 3 + 1] * If[1 * 3 + c] + R[r * 3 + 2] * If[2 * 3 + c];
  const double gp2 = 2.0 * (g[0] * p[0] + g[1] * p[1] + g[2] * p[2]);
  Y[0] = mass;
  Y[1] = hp[0] + mass * p[0]; Y[2] = hp[1] + mass * p[1]; Y[3] = hp[2] + mass * p[2];
  Y[4] = (T[0] * R[0] + T[1] * R[1] + T[2] * R[2]) + gp2 - 2.0 * g[0] * p[0];
  Y[5] = (T[0] * R[3] + T[1] * R[4] + T[2] * R[5]) - (g[0] * p[1] + p[0] * g[1]);
  Y[6] = (T[0] * R[6] + T[1] * R[7] + T[2] * R[8]) - (g[0] * p[2] + p[0] * g[2]);
  Y[7] = (T[3] * R[3] + T[4] * R[4] + T[5] * R[5]) + gp2 - 2.0 * g[1] * p[1];
  Y[8] = (T[3] * R[6] + T[4] * R[7] + T[5] * R[8]) - (g[1] * p[2] + p[1] * g[2]);
  Y[9] = (T[6] * R[6] + T[7] * R[7] + T[8] * R[8]) + gp2 - 2.0 * g[2] * p[2];
}
// IA (packed symmetric 6x6) += rigid-body inertia Y:  [[I, skew(h)], [skew(h)^T, m Id]]
DDPB_HD void ia_add_body(double* IA, const double* Y) {
  IA[s6(0, 0)] += Y[4]; IA[s6(0, 1)] += Y[5]; IA[s6(0, 2)] += Y[6]; IA[s6(1, 1)] += Y[7]; IA[s6(1, 2)] += Y[8]; IA[s6(2, 2)] += Y[9];
  IA[s6(0, 4)] -= Y[3]; IA[s6(0, 5)] += Y[2]; IA[s6(1, 3)] += Y[3]; IA[s6(1, 5)] -= Y[1]; IA[s6(2, 3)] -= Y[2]; IA[s6(2, 4)] += Y[1];
  IA[s6(3, 3)] += Y[0]; IA[s6(4, 4)] += Y[0]; IA[s6(5, 5)] += Y[0];
}

// ---- q_task ------------------------------------------------------------------------------------
// everything that depends on q only: world-frame joint axes and body inertias, articulated inertias IA_i, U_i, 1/D_i.
// The linearization kernel runs it as four sub-phases so that only the inherently serial parts are narrow:
//   joint_local_rot  per (configuration, joint)   -> Rl_i            (parallel; scratch in the IA block of the record)
//   q_chain          per configuration            -> R_i, p_i        (serial over joints, 2 small products per joint)
//   body_world       per (configuration, body)    -> S_i, Y_i        (parallel)
//   q_articulate     per configuration            -> IA_i, U_i, 1/D_i (serial, leaf to root)
// q_task is the same arithmetic in one call (rollouts of small models, tests).
template <int NJ> struct QScratch {  // scratch at QRec::SCR (24 * NJ doubles) until q_articulate overwrites it
  static constexpr int RL = 0;            // [NJ][12]: local Rl_i (9) | pl_i (3)
  static constexpr int RP = 12 * NJ;      // [NJ][12]: world R_i (9) | p_i (3)
};
template <int NJ> DDPB_HD void q_chain(const ModelDev& m, double* rec) {
  using Q = QRec<NJ>;
  using X = QScratch<NJ>;
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, p[3] = {0, 0, 0};
#pragma unroll
  for (int i = 0; i < NJ; ++i) {
    double Rl[9], pl[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) Rl[k] = rec[Q::SCR + X::RL + 12 * i + k];
#pragma unroll
    for (int k = 0; k < 3; ++k) pl[k] = HAS_PRISMATIC ? rec[Q::SCR + X::RL + 12 * i + 9 + k] : m.pp[i][k];  // (constant when every joint is revolute)
    joint_compose(pl, Rl, R, p);
#pragma unroll
    for (int k = 0; k < 9; ++k) rec[Q::SCR + X::RP + 12 * i + k] = R[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) rec[Q::SCR + X::RP + 12 * i + 9 + k] = p[k];
  }
}
template <int NJ> DDPB_HD void q_body(const ModelDev& m, int i, double* rec) {
  using Q = QRec<NJ>;
  using X = QScratch<NJ>;
  double R[9], p[3], S[6], Y[10];
#pragma unroll
  for (int k = 0; k < 9; ++k) R[k] = rec[Q::SCR + X::RP + 12 * i + k];
#pragma unroll
  for (int k = 0; k < 3; ++k) p[k] = rec[Q::SCR + X::RP + 12 * i + 9 + k];
  body_world(m, i, R, p, S, Y);
#pragma unroll
  for (int k = 0; k < 6; ++k) rec[Q::S + 6 * i + k] = S[k];
#pragma unroll
  for (int k = 0; k < 10; ++k) rec[Q::Y + 10 * i + k] = Y[k];
}
template <int NJ> DDPB_HD void q_articulate(double* rec) {
  using Q = QRec<NJ>;
  double IA[21];
#pragma unroll
  for (int k = 0; k < 21; ++k) IA[k] = 0.0;
#pragma unroll
  for (int i = NJ - 1; i >= 0; --i) {
    double Y[10], S[6], U[6];
#pragma unroll
    for (int k = 0; k < 10; ++k) Y[k] = rec[Q::Y + 10 * i + k];
#pragma unroll
    for (int k = 0; k < 6; ++k) S[k] = rec[Q::S + 6 * i + k];
    ia_add_body(IA, Y);
    if (i > 0) {
#pragma unroll
      for (int k = 0; k < 21; ++k) rec[Q::IA + 21 * (i - 1) + k] = IA[k];
    }
    sym6_mul(IA, S, U);
    const double Dinv = 1.0 / dot6(S, U);
#pragma unroll
    for (int k = 0; k < 6; ++k) rec[Q::U + 6 * i + k] = U[k];
    rec[Q::DINV + i] = Dinv;
    if (i > 0) {
      // what the parent sees of this subtree: IA - U D^-1 U^T
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        const double sr = U[r] * Dinv;
#pragma unroll
        for (int c = r; c < 6; ++c) IA[s6(r, c)] -= sr * U[c];
      }
    }
  }
}
template <int NJ> DDPB_HD void q_task(const ModelDev& m, const double* q, double* rec) {
  using Q = QRec<NJ>;
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, p[3] = {0, 0, 0};
#pragma unroll
  for (int i = 0; i < NJ; ++i) {
    joint_advance(m, i, q[i], R, p);
    double S[6], Y[10];
    body_world(m, i, R, p, S, Y);
#pragma unroll
    for (int k = 0; k < 6; ++k) rec[Q::S + 6 * i + k] = S[k];
#pragma unroll
    for (int k = 0; k < 10; ++k) rec[Q::Y + 10 * i + k] = Y[k];
  }
  q_articulate<NJ>(rec);
}

// ---- v_task --------------------------------------------------------------------------------------
// velocity / torque dependent sweeps of the articulated-body algorithm -> joint accelerations qdd_out[NJ].
// REC = true: the per-body intermediates live in the point's shared-memory record `vrec`, which afterwards holds
// what the derivative columns need (V, A, H, F).  REC = false: value only, intermediates in registers.
template <int NJ, bool REC> DDPB_HD void v_task(const ModelDev& m, const double* qrec, const double* qd, const double* tau, double* qdd_out, double* vrec) {
  using Q = QRec<NJ>;
  using V = VRec<NJ>;
  double cl[REC ? 1 : NJ][6], pl[REC ? 1 : NJ][6];  // bias accelerations / bias forces when there is no record
  double u[NJ];
  {
    double v[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < NJ; ++i) {
      const double* S = qrec + Q::S + 6 * i;
      double sq[6], c[6], h[6], pa[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) { sq[k] = S[k] * qd[i]; v[k] += sq[k]; }
      crm(v, sq, c);                      // c_i = v_i x (S_i qd_i)
      y_mul(qrec + Q::Y + 10 * i, v, h);  // h_i = Y_i v_i
      crf(v, h, pa);                      // p_i = v_i x* h_i
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        if (REC) { vrec[V::V + 6 * i + k] = v[k]; vrec[V::H + 6 * i + k] = h[k]; vrec[V::A + 6 * i + k] = c[k]; vrec[V::F + 6 * i + k] = pa[k]; }
        else { cl[REC ? 0 : i][k] = c[k]; pl[REC ? 0 : i][k] = pa[k]; }
      }
    }
  }
  double carry[6] = {0, 0, 0, 0, 0, 0};  // child -> parent bias force, in registers (see v_sweeps)
#pragma unroll
  for (int i = NJ - 1; i >= 0; --i) {
    const double* S = qrec + Q::S + 6 * i;
    double pa[6], c[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      pa[k] = REC ? vrec[V::F + 6 * i + k] : pl[REC ? 0 : i][k];
      if (i < NJ - 1) pa[k] += carry[k];
      c[k] = REC ? vrec[V::A + 6 * i + k] : cl[REC ? 0 : i][k];
    }
    u[i] = tau[i] - dot6(S, pa);
    if (i > 0) {
      const double* U = qrec + Q::U + 6 * i;
      double Iac[6];
      const double t = qrec[Q::DINV + i] * (u[i] - dot6(U, c));
      sym6_mul(qrec + Q::IA + 21 * (i - 1), c, Iac);
#pragma unroll
      for (int k = 0; k < 6; ++k) carry[k] = pa[k] + Iac[k] + U[k] * t;
    }
  }
  double a[6] = {0.0, 0.0, 0.0, -m.grav[0], -m.grav[1], -m.grav[2]};
#pragma unroll
  for (int i = 0; i < NJ; ++i) {
    const double* S = qrec + Q::S + 6 * i;
    const double* U = qrec + Q::U + 6 * i;
#pragma unroll
    for (int k = 0; k < 6; ++k) a[k] += REC ? vrec[V::A + 6 * i + k] : cl[REC ? 0 : i][k];
    const double qdd = qrec[Q::DINV + i] * (u[i] - dot6(U, a));
#pragma unroll
    for (int k = 0; k < 6; ++k) a[k] += S[k] * qdd;
    qdd_out[i] = qdd;
    if (REC) {
#pragma unroll
      for (int k = 0; k < 6; ++k) vrec[V::A + 6 * i + k] = a[k];
    }
  }
  if (REC) {
    // subtree forces of inverse dynamics at this solution: F_i = sum_{k >= i} (Y_k a_k + v_k x* h_k)
    double F[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int i = NJ - 1; i >= 0; --i) {
      double ak[6], vk[6], hk[6], Ya[6], vh[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) { ak[k] = vrec[V::A + 6 * i + k]; vk[k] = vrec[V::V + 6 * i + k]; hk[k] = vrec[V::H + 6 * i + k]; }
      y_mul(qrec + Q::Y + 10 * i, ak, Ya);
      crf(vk, hk, vh);
#pragma unroll
      for (int k = 0; k < 6; ++k) { F[k] += Ya[k] + vh[k]; vrec[V::F + 6 * i + k] = F[k]; }
    }
  }
}

// ---- v_task<NJ, true> as sub-phases (same arithmetic; only the sweeps that couple bodies are serial) ------------------
//   v_prefix  per point         : body velocities v_i = sum_{l <= i} S_l qd_l
//   v_bias    per (point, body) : c_i = v_i x (S_i qd_i), h_i = Y_i v_i, p_i = v_i x* h_i   (into the A, H, F slots)
//   v_sweeps  per point         : leaf-to-root bias-force sweep and root-to-leaf acceleration sweep -> qdd, a_i
//   v_force   per (point, body) : f_i = Y_i a_i + v_i x* h_i                                 (into the F slot)
//   v_suffix  per point         : F_i = sum_{k >= i} f_k
template <int NJ> DDPB_HD void v_prefix(const double* qrec, const double* qd, double* vrec) {
  using Q = QRec<NJ>;
  using V = VRec<NJ>;
  double v[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < NJ; ++i) {
    const double* S = qrec + Q::S + 6 * i;
#pragma unroll
    for (int k = 0; k < 6; ++k) { const double sq = S[k] * qd[i]; v[k] += sq; vrec[V::V + 6 * i + k] = v[k]; }
  }
}
template <int NJ> DDPB_HD void v_bias(const double* qrec, int i, double qdi, double* vrec) {
  using Q = QRec<NJ>;
  using V = VRec<NJ>;
  const double* S = qrec + Q::S + 6 * i;
  double v[6], sq[6], c[6], h[6], pa[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) { v[k] = vrec[V::V + 6 * i + k]; sq[k] = S[k] * qdi; }
  crm(v, sq, c);
  y_mul(qrec + Q::Y + 10 * i, v, h);
  crf(v, h, pa);
#pragma unroll
  for (int k = 0; k < 6; ++k) { vrec[V::H + 6 * i + k] = h[k]; vrec[V::A + 6 * i + k] = c[k]; vrec[V::F + 6 * i + k] = pa[k]; }
  if (i > 0) {
    // the part of the leaf-to-root sweep that does not depend on the recursion: IA_i c_i and U_i . c_i, parked in the
    // DQ block (7 doubles per body) until the derivative columns overwrite it
    double Iac[6];
    sym6_mul(qrec + Q::IA + 21 * (i - 1), c, Iac);
#pragma unroll
    for (int k = 0; k < 6; ++k) vrec[V::DQ + 7 * i + k] = Iac[k];
    vrec[V::DQ + 7 * i + 6] = dot6(qrec + Q::U + 6 * i, c);
  }
}
template <int NJ> DDPB_HD void v_sweeps(const ModelDev& m, const double* qrec, const double* tau, double* qdd_out, double* vrec) {
  using Q = QRec<NJ>;
  using V = VRec<NJ>;
  double u[NJ];
  // what the child hands to its parent is carried in registers (parent bias force = its own + carry: the same single
  // addition as accumulating into the parent's slot, without the shared-memory round trip on the dependent path)
  double carry[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int i = NJ - 1; i >= 0; --i) {
    const double* S = qrec + Q::S + 6 * i;
    double pa[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) { pa[k] = vrec[V::F + 6 * i + k]; if (i < NJ - 1) pa[k] += carry[k]; }
    u[i] = tau[i] - dot6(S, pa);
    if (i > 0) {
      const double* U = qrec + Q::U + 6 * i;
      const double t = qrec[Q::DINV + i] * (u[i] - vrec[V::DQ + 7 * i + 6]);
#pragma unroll
      for (int k = 0; k < 6; ++k) carry[k] = pa[k] + vrec[V::DQ + 7 * i + k] + U[k] * t;
    }
  }
  double a[6] = {0.0, 0.0, 0.0, -m.grav[0], -m.grav[1], -m.grav[2]};
#pragma unroll
  for (int i = 0; i < NJ; ++i) {
    const double* S = qrec + Q::S + 6 * i;
    const double* U = qrec + Q::U + 6 * i;
#pragma unroll
    for (int k = 0; k < 6; ++k) a[k] += vrec[V::A + 6 * i + k];
    const double qdd = qrec[Q::DINV + i] * (u[i] - dot6(U, a));
#pragma unroll
    for (int k = 0; k < 6; ++k) { a[k] += S[k] * qdd; vrec[V::A + 6 * i + k] = a[k]; }
    qdd_out[i] = qdd;
  }
}
template <int NJ> DDPB_HD void v_force(const double* qrec, int i, double* vrec) {
  using Q = QRec<NJ>;
  using V = VRec<NJ>;
  double ak[6], vk[6], hk[6], Ya[6], vh[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) { ak[k] = vrec[V::A + 6 * i + k]; vk[k] = vrec[V::V + 6 * i + k]; hk[k] = vrec[V::H + 6 * i + k]; }
  y_mul(qrec + Q::Y + 10 * i, ak, Ya);
  crf(vk, hk, vh);
#pragma unroll
  for (int k = 0; k < 6; ++k) vrec[V::F + 6 * i + k] = Ya[k] + vh[k];
}
template <int NJ> DDPB_HD void v_suffix(double* vrec) {
  using V = VRec<NJ>;
  double F[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int i = NJ - 1; i >= 0; --i) {
#pragma unroll
    for (int k = 0; k < 6; ++k) { F[k] += vrec[V::F + 6 * i + k]; vrec[V::F + 6 * i + k] = F[k]; }
  }
}

// ---- aba_fused ------------------------------------------------------------------------------------
// Forward dynamics qdd = ABA(q, qd, tau) for ONE evaluation (the rollouts): the same arithmetic as q_task followed by
// v_task<NJ, false> -- bit for bit, the derivative records' f(x, u) must equal the rollout -- but with the three
// sweeps fused so that only 28 doubles per body survive between them.  They live in `scr`, a per-thread strided
// scratch (element e at scr[e * stride]): shared memory with stride = threads per block on the device (conflict-free,
// no local-memory traffic), a plain array on the host.
constexpr int ABA_SCR_PER_BODY = 28;  // S(6) | c(6) | pA(6) | Y(10, reused as U(6), 1/D, u after the inertia is consumed)
template <int NJ> DDPB_HD void aba_fused(const ModelDev& m, const double* q, const double* qd, const double* tau, double* qdd_out, double* scr, int stride) {
#define SCR(i, e) scr[((i) * ABA_SCR_PER_BODY + (e)) * stride]
  {
    double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, p[3] = {0, 0, 0};
    double v[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < NJ; ++i) {
      joint_advance(m, i, q[i], R, p);
      double S[6], Y[10], sq[6], c[6], h[6], pa[6];
      body_world(m, i, R, p, S, Y);
#pragma unroll
      for (int k = 0; k < 6; ++k) { sq[k] = S[k] * qd[i]; v[k] += sq[k]; }
      crm(v, sq, c);
      y_mul(Y, v, h);
      crf(v, h, pa);
#pragma unroll
      for (int k = 0; k < 6; ++k) { SCR(i, k) = S[k]; SCR(i, 6 + k) = c[k]; SCR(i, 12 + k) = pa[k]; }
#pragma unroll
      for (int k = 0; k < 10; ++k) SCR(i, 18 + k) = Y[k];
    }
  }
  {
    double IA[21];
#pragma unroll
    for (int k = 0; k < 21; ++k) IA[k] = 0.0;
    double carry[6] = {0, 0, 0, 0, 0, 0};  // child -> parent bias force, in registers (see v_sweeps)
#pragma unroll
    for (int i = NJ - 1; i >= 0; --i) {
      double Y[10], S[6], U[6], pa[6], c[6];
#pragma unroll
      for (int k = 0; k < 10; ++k) Y[k] = SCR(i, 18 + k);
#pragma unroll
      for (int k = 0; k < 6; ++k) { S[k] = SCR(i, k); c[k] = SCR(i, 6 + k); pa[k] = SCR(i, 12 + k); if (i < NJ - 1) pa[k] += carry[k]; }
      ia_add_body(IA, Y);
      sym6_mul(IA, S, U);
      const double Dinv = 1.0 / dot6(S, U);
      const double u = tau[i] - dot6(S, pa);
#pragma unroll
      for (int k = 0; k < 6; ++k) SCR(i, 18 + k) = U[k];
      SCR(i, 24) = Dinv;
      SCR(i, 25) = u;
      if (i > 0) {
        double Iac[6];
        const double t = Dinv * (u - dot6(U, c));
        sym6_mul(IA, c, Iac);
#pragma unroll
        for (int k = 0; k < 6; ++k) carry[k] = pa[k] + Iac[k] + U[k] * t;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          const double sr = U[r] * Dinv;
#pragma unroll
          for (int cc = r; cc < 6; ++cc) IA[s6(r, cc)] -= sr * U[cc];
        }
      }
    }
  }
  double a[6] = {0.0, 0.0, 0.0, -m.grav[0], -m.grav[1], -m.grav[2]};
#pragma unroll
  for (int i = 0; i < NJ; ++i) {
    double S[6], U[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) { S[k] = SCR(i, k); U[k] = SCR(i, 18 + k); a[k] += SCR(i, 6 + k); }
    const double qdd = SCR(i, 24) * (SCR(i, 25) - dot6(U, a));
#pragma unroll
    for (int k = 0; k < 6; ++k) a[k] += S[k] * qdd;
    qdd_out[i] = qdd;
  }
#undef SCR
}

// ---- col_task -----------------------------------------------------------------------------------
// column j of d tau/d q (wrt_v = 0) or d tau/d qdot (wrt_v = 1) of inverse dynamics at (q, qdot, qddot = ABA result).
// World-frame closed form (bodies k >= j move rigidly with joint j; S_j x is the spatial motion cross product):
//   d = v_j x S_j ,   g_k = Y_k (alpha + beta x (v_k - v_j)) + gamma x* h_k + v_k x* (Y_k gamma)
//     wrt q_j    : alpha = a_j x S_j, beta = gamma = d ;   d f_k = S_j x* f_k + g_k
//     wrt qdot_j : alpha = d, beta = gamma = S_j       ;   d f_k = g_k
//   out[i] = S_i . sum_{k >= max(i,j)} g_k   ( + S_i . (S_j x* F_j) for i < j, wrt q: for i >= j that term cancels
//   against the derivative of S_i itself )
template <int NJ> DDPB_HD void col_task(const double* qrec, const double* vrec, int j, int wrt_v, double* out /*[NJ]*/) {
  using Q = QRec<NJ>;
  using V = VRec<NJ>;
  double Sj[6], vj[6], d[6], al[6], be[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) { Sj[k] = qrec[Q::S + 6 * j + k]; vj[k] = vrec[V::V + 6 * j + k]; }
  crm(vj, Sj, d);
  if (!wrt_v) {
    double aj[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) aj[k] = vrec[V::A + 6 * j + k];
    crm(aj, Sj, al);
#pragma unroll
    for (int k = 0; k < 6; ++k) be[k] = d[k];
  } else {
#pragma unroll
    for (int k = 0; k < 6; ++k) { al[k] = d[k]; be[k] = Sj[k]; }
  }
  double G[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int k = NJ - 1; k >= 0; --k) {
    if (k < j) continue;
    const double* Yk = qrec + Q::Y + 10 * k;
    double vk[6], hk[6], dv[6], t[6], g1[6], g2[6], Yg[6], g3[6];
#pragma unroll
    for (int e = 0; e < 6; ++e) { vk[e] = vrec[V::V + 6 * k + e]; hk[e] = vrec[V::H + 6 * k + e]; dv[e] = vk[e] - vj[e]; }
    crm(be, dv, t);
#pragma unroll
    for (int e = 0; e < 6; ++e) t[e] += al[e];
    y_mul(Yk, t, g1);
    crf(be, hk, g2);
    y_mul(Yk, be, Yg);
    crf(vk, Yg, g3);
#pragma unroll
    for (int e = 0; e < 6; ++e) G[e] += g1[e] + g2[e] + g3[e];
    out[k] = dot6(qrec + Q::S + 6 * k, G);
  }
  if (!wrt_v) {
    double Fj[6], x[6];
#pragma unroll
    for (int e = 0; e < 6; ++e) Fj[e] = vrec[V::F + 6 * j + e];
    crf(Sj, Fj, x);
#pragma unroll
    for (int e = 0; e < 6; ++e) G[e] += x[e];
  }
#pragma unroll
  for (int i = 0; i < NJ; ++i)
    if (i < j) out[i] = dot6(qrec + Q::S + 6 * i, G);
}

// Both derivative kinds of column j in one task: d tau/d q_j -> out_q, d tau/d qdot_j -> out_v.  The same arithmetic in the
// same order as two col_task calls (bit-identical results); the body data (v_k, h_k, Y_k, S_k) is loaded once for both.
template <int NJ> DDPB_HD void col_task2(const double* qrec, const double* vrec, int j, double* out_q /*[NJ]*/, double* out_v /*[NJ]*/) {
  using Q = QRec<NJ>;
  using V = VRec<NJ>;
  double Sj[6], vj[6], d[6], al[6];
  {
    double aj[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) { Sj[k] = qrec[Q::S + 6 * j + k]; vj[k] = vrec[V::V + 6 * j + k]; aj[k] = vrec[V::A + 6 * j + k]; }
    crm(vj, Sj, d);
    crm(aj, Sj, al);
  }
  double Gq[6] = {0, 0, 0, 0, 0, 0}, Gv[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int k = NJ - 1; k >= 0; --k) {
    if (k < j) continue;
    double Yk[10], Sk[6], vk[6], hk[6], dv[6];
#pragma unroll
    for (int e = 0; e < 10; ++e) Yk[e] = qrec[Q::Y + 10 * k + e];
#pragma unroll
    for (int e = 0; e < 6; ++e) { Sk[e] = qrec[Q::S + 6 * k + e]; vk[e] = vrec[V::V + 6 * k + e]; hk[e] = vrec[V::H + 6 * k + e]; dv[e] = vk[e] - vj[e]; }
    {  // wrt q_j: alpha = a_j x S_j, beta = gamma = d
      double t[6], g1[6], g2[6], Yg[6], g3[6];
      crm(d, dv, t);
#pragma unroll
      for (int e = 0; e < 6; ++e) t[e] += al[e];
      y_mul(Yk, t, g1);
      crf(d, hk, g2);
      y_mul(Yk, d, Yg);
      crf(vk, Yg, g3);
#pragma unroll
      for (int e = 0; e < 6; ++e) Gq[e] += g1[e] + g2[e] + g3[e];
      out_q[k] = dot6(Sk, Gq);
    }
    {  // wrt qdot_j: alpha = d, beta = gamma = S_j
      double t[6], g1[6], g2[6], Yg[6], g3[6];
      crm(Sj, dv, t);
#pragma unroll
      for (int e = 0; e < 6; ++e) t[e] += d[e];
      y_mul(Yk, t, g1);
      crf(Sj, hk, g2);
      y_mul(Yk, Sj, Yg);
      crf(vk, Yg, g3);
#pragma unroll
      for (int e = 0; e < 6; ++e) Gv[e] += g1[e] + g2[e] + g3[e];
      out_v[k] = dot6(Sk, Gv);
    }
  }
  {
    double Fj[6], x[6];
#pragma unroll
    for (int e = 0; e < 6; ++e) Fj[e] = vrec[V::F + 6 * j + e];
    crf(Sj, Fj, x);
#pragma unroll
    for (int e = 0; e < 6; ++e) Gq[e] += x[e];
  }
#pragma unroll
  for (int i = 0; i < NJ; ++i)
    if (i < j) {
      double Si[6];
#pragma unroll
      for (int e = 0; e < 6; ++e) Si[e] = qrec[Q::S + 6 * i + e];
      out_q[i] = dot6(Si, Gq);
      out_v[i] = dot6(Si, Gv);
    }
}

// col_task split by body for the narrow case (one point per knot needs columns, value-based Hessians): col_term is the
// contribution g_k of one body to column (j, wrt), col_reduce the suffix sums and projections -- the same arithmetic
// in the same order as col_task, with 2 * NJ(NJ+1)/2 terms in parallel instead of 2 * NJ columns.
template <int NJ> DDPB_HD void col_term(const double* qrec, const double* vrec, int j, int k, int wrt_v, double* g /*[6]*/) {
  using Q = QRec<NJ>;
  using V = VRec<NJ>;
  double Sj[6], vj[6], d[6], al[6], be[6];
#pragma unroll
  for (int e = 0; e < 6; ++e) { Sj[e] = qrec[Q::S + 6 * j + e]; vj[e] = vrec[V::V + 6 * j + e]; }
  crm(vj, Sj, d);
  if (!wrt_v) {
    double aj[6];
#pragma unroll
    for (int e = 0; e < 6; ++e) aj[e] = vrec[V::A + 6 * j + e];
    crm(aj, Sj, al);
#pragma unroll
    for (int e = 0; e < 6; ++e) be[e] = d[e];
  } else {
#pragma unroll
    for (int e = 0; e < 6; ++e) { al[e] = d[e]; be[e] = Sj[e]; }
  }
  const double* Yk = qrec + Q::Y + 10 * k;
  double vk[6], hk[6], dv[6], t[6], g1[6], g2[6], Yg[6], g3[6];
#pragma unroll
  for (int e = 0; e < 6; ++e) { vk[e] = vrec[V::V + 6 * k + e]; hk[e] = vrec[V::H + 6 * k + e]; dv[e] = vk[e] - vj[e]; }
  crm(be, dv, t);
#pragma unroll
  for (int e = 0; e < 6; ++e) t[e] += al[e];
  y_mul(Yk, t, g1);
  crf(be, hk, g2);
  y_mul(Yk, be, Yg);
  crf(vk, Yg, g3);
#pragma unroll
  for (int e = 0; e < 6; ++e) g[e] = g1[e] + g2[e] + g3[e];
}
// terms: g_k of this column at terms[6 * k] for k >= j
template <int NJ> DDPB_HD void col_reduce(const double* qrec, const double* vrec, const double* terms, int j, int wrt_v, double* out /*[NJ]*/) {
  using Q = QRec<NJ>;
  using V = VRec<NJ>;
  double G[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int k = NJ - 1; k >= 0; --k) {
    if (k < j) continue;
#pragma unroll
    for (int e = 0; e < 6; ++e) G[e] += terms[6 * k + e];
    out[k] = dot6(qrec + Q::S + 6 * k, G);
  }
  if (!wrt_v) {
    double Sj[6], Fj[6], x[6];
#pragma unroll
    for (int e = 0; e < 6; ++e) { Sj[e] = qrec[Q::S + 6 * j + e]; Fj[e] = vrec[V::F + 6 * j + e]; }
    crf(Sj, Fj, x);
#pragma unroll
    for (int e = 0; e < 6; ++e) G[e] += x[e];
  }
#pragma unroll
  for (int i = 0; i < NJ; ++i)
    if (i < j) out[i] = dot6(qrec + Q::S + 6 * i, G);
}

// ---- minv_col ---------------------------------------------------------------------------------------
// column `col` of M^-1(q): unit-torque articulated-body sweeps on the cached S, U, 1/D
template <int NJ> DDPB_HD void minv_col(const double* qrec, int col, double* out /*[NJ]*/) {
  using Q = QRec<NJ>;
  double pA[6] = {0, 0, 0, 0, 0, 0};
  double u[NJ];
#pragma unroll
  for (int i = NJ - 1; i >= 0; --i) {
    if (i > col) { u[i] = 0.0; continue; }
    u[i] = (i == col ? 1.0 : 0.0) - dot6(qrec + Q::S + 6 * i, pA);
    if (i > 0) {
      const double* U = qrec + Q::U + 6 * i;
      const double t = qrec[Q::DINV + i] * u[i];
#pragma unroll
      for (int k = 0; k < 6; ++k) pA[k] += U[k] * t;
    }
  }
  double a[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < NJ; ++i) {
    const double* S = qrec + Q::S + 6 * i;
    const double qdd = qrec[Q::DINV + i] * (u[i] - dot6(qrec + Q::U + 6 * i, a));
#pragma unroll
    for (int k = 0; k < 6; ++k) a[k] += S[k] * qdd;
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
    double E[9], t[3], Rn[9], pl[3];
    joint_E(m, i, q[i], E);
    joint_local_trans(m, i, q[i], pl);
    m3v(R, pl, t);
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
        if (HAS_PRISMATIC && m.jtype[i] != 0) {  // prismatic: the linear part of [0 ; w_i], the same in both conventions
          lin[0] = w[i][0]; lin[1] = w[i][1]; lin[2] = w[i][2];
        } else if (mode == 0) {
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

DDPB_VARIANT_END
