// ddp_oracle.hpp -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
//
// A serial, single-problem restatement of the reference hot path of
// s-elkazdadi/ddp-pinocchio, templated on the scalar type (double, __float128,
// flop-counting double).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference leg may build, link or call this.
//
// PARITY STATUS: "parity unpinned" for the rigid-body arithmetic.  The reference
// delegates ABA / ABA-derivatives / frame Jacobians to the third-party Pinocchio
// library (pkg-config module `pinocchio`, version unpinned, CMakeLists.txt:81-86;
// call sites include/ddp/pinocchio_model.ipp:354,391,428,458-460), which is absent
// from /root/reference and from this image, and the reference's own tests hold no
// numeric golden vector for this path.  The rigid-body part
// therefore restates the published algorithms (Featherstone ABA; forward-mode
// tangent of RNEA; M^-1 by the ABA-style unit-torque passes) and is validated by
// identities (M(q) a + nle = tau with an independent RNEA, __float128 central
// differences).  Everything above the Model concept follows the reference line by
// line; each function cites the file:line it restates.
//
// Conventions: spatial vectors are [angular(3); linear(3)] in body coordinates.
// Matrices at the API boundary are column-major; tensors use the reference index
// k + i*O + j*O*L (include/ddp/detail/tensor.hpp:141-147).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>
#ifdef ORC_HAVE_QUAD
#include <quadmath.h>
#endif

namespace orc {

// ----------------------------------------------------------------------------------
// scalar helpers
// ----------------------------------------------------------------------------------
struct FlopCounter {
  static uint64_t& n() {
    static thread_local uint64_t c = 0;
    return c;
  }
};
// double that counts arithmetic operations (used to pin the algorithmic flop figure
// of one first-order evaluation).
struct Counted {
  double v;
  Counted() : v(0) {}
  Counted(double x) : v(x) {}
  Counted(long double x) : v((double)x) {}
  Counted(int x) : v(x) {}
  explicit operator double() const { return v; }
};
inline Counted operator+(Counted a, Counted b) { ++FlopCounter::n(); return {a.v + b.v}; }
inline Counted operator-(Counted a, Counted b) { ++FlopCounter::n(); return {a.v - b.v}; }
inline Counted operator*(Counted a, Counted b) { ++FlopCounter::n(); return {a.v * b.v}; }
inline Counted operator/(Counted a, Counted b) { ++FlopCounter::n(); return {a.v / b.v}; }
inline Counted operator-(Counted a) { return {-a.v}; }
inline Counted& operator+=(Counted& a, Counted b) { a = a + b; return a; }
inline Counted& operator-=(Counted& a, Counted b) { a = a - b; return a; }
inline Counted& operator*=(Counted& a, Counted b) { a = a * b; return a; }
inline Counted& operator/=(Counted& a, Counted b) { a = a / b; return a; }
inline bool operator<(Counted a, Counted b) { return a.v < b.v; }
inline bool operator<=(Counted a, Counted b) { return a.v <= b.v; }
inline bool operator>(Counted a, Counted b) { return a.v > b.v; }
inline bool operator>=(Counted a, Counted b) { return a.v >= b.v; }

inline double Sin(double x) { return std::sin(x); }
inline double Cos(double x) { return std::cos(x); }
inline double Sqrt(double x) { return std::sqrt(x); }
inline double Pow(double x, double y) { return std::pow(x, y); }
inline double ToDouble(double x) { return x; }
inline Counted Sin(Counted x) { FlopCounter::n() += 8; return {std::sin(x.v)}; }
inline Counted Cos(Counted x) { FlopCounter::n() += 8; return {std::cos(x.v)}; }
inline Counted Sqrt(Counted x) { FlopCounter::n() += 1; return {std::sqrt(x.v)}; }
inline Counted Pow(Counted x, Counted y) { FlopCounter::n() += 16; return {std::pow(x.v, y.v)}; }
inline double ToDouble(Counted x) { return x.v; }
#ifdef ORC_HAVE_QUAD
inline __float128 Sin(__float128 x) { return sinq(x); }
inline __float128 Cos(__float128 x) { return cosq(x); }
inline __float128 Sqrt(__float128 x) { return sqrtq(x); }
inline __float128 Pow(__float128 x, __float128 y) { return powq(x, y); }
inline double ToDouble(__float128 x) { return (double)x; }
#endif

template <class S> struct Eps { static S value() { return S(std::numeric_limits<double>::epsilon()); } };
#ifdef ORC_HAVE_QUAD
template <> struct Eps<__float128> { static __float128 value() { return scalbnq((__float128)1, -112); } };
#endif

constexpr int MAXJ = 8;   // joints
constexpr int MAXF = 16;  // operational frames
constexpr int MAXE = 8;   // equality rows per knot

enum ModelKind { KIND_PENDULUM = 0, KIND_REVOLUTE_TREE = 1 };
enum JointType { JOINT_REVOLUTE = 0, JOINT_PRISMATIC = 1 };  // what pinocchio::urdf::buildModel makes of URDF revolute / prismatic joints
enum ConstraintKind { EQ_NONE = 0, EQ_CONFIG = 1, EQ_SPATIAL = 2 };
enum FrameJacMode { FJ_WORLD = 0, FJ_LOCAL_WORLD_ALIGNED = 1 };

// ----------------------------------------------------------------------------------
// 3-vector / spatial helpers
// ----------------------------------------------------------------------------------
template <class S> inline void cross3(const S* a, const S* b, S* o) {
  S x = a[1] * b[2] - a[2] * b[1];
  S y = a[2] * b[0] - a[0] * b[2];
  S z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
template <class S> inline void mat3_mul_vec(const S* M, const S* v, S* o) {  // row-major M
  S x = M[0] * v[0] + M[1] * v[1] + M[2] * v[2];
  S y = M[3] * v[0] + M[4] * v[1] + M[5] * v[2];
  S z = M[6] * v[0] + M[7] * v[1] + M[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
template <class S> inline void mat3T_mul_vec(const S* M, const S* v, S* o) {
  S x = M[0] * v[0] + M[3] * v[1] + M[6] * v[2];
  S y = M[1] * v[0] + M[4] * v[1] + M[7] * v[2];
  S z = M[2] * v[0] + M[5] * v[1] + M[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
// motion cross product  a x b   (a, b motion vectors)
template <class S> inline void crm(const S* a, const S* b, S* o) {
  S t0[3], t1[3], t2[3];
  cross3(a, b, t0);          // w x bw
  cross3(a, b + 3, t1);      // w x bv
  cross3(a + 3, b, t2);      // v x bw
  for (int k = 0; k < 3; ++k) { o[k] = t0[k]; o[3 + k] = t1[k] + t2[k]; }
}
// force cross product  a x* f   (a motion, f force)
template <class S> inline void crf(const S* a, const S* f, S* o) {
  S t0[3], t1[3], t2[3];
  cross3(a, f, t0);          // w x n
  cross3(a + 3, f + 3, t1);  // v x f
  cross3(a, f + 3, t2);      // w x f
  for (int k = 0; k < 3; ++k) { o[k] = t0[k] + t1[k]; o[3 + k] = t2[k]; }
}

// Pluecker transform parent->child: E (row-major rotation taking parent coordinates
// to child coordinates) and r (child origin in parent coordinates).
template <class S> struct Xform { S E[9]; S r[3]; };
template <class S> inline void xf_motion(const Xform<S>& X, const S* m, S* o) {
  S t[3], u[3];
  cross3(X.r, m, t);  // r x w
  for (int k = 0; k < 3; ++k) u[k] = m[3 + k] - t[k];
  mat3_mul_vec(X.E, m, o);
  mat3_mul_vec(X.E, u, o + 3);
}
// transpose action: child force -> parent force
template <class S> inline void xf_force_T(const Xform<S>& X, const S* f, S* o) {
  S n[3], l[3], t[3];
  mat3T_mul_vec(X.E, f, n);
  mat3T_mul_vec(X.E, f + 3, l);
  cross3(X.r, l, t);
  for (int k = 0; k < 3; ++k) { o[k] = n[k] + t[k]; o[3 + k] = l[k]; }
}
// dense 6x6 (row-major) form of X, for the congruence X^T IA X
template <class S> inline void xf_dense(const Xform<S>& X, S* M) {
  // [ E        0 ]
  // [ -E rx    E ]
  S rx[9] = {S(0), -X.r[2], X.r[1], X.r[2], S(0), -X.r[0], -X.r[1], X.r[0], S(0)};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      M[i * 6 + j] = X.E[i * 3 + j];
      M[i * 6 + 3 + j] = S(0);
      M[(3 + i) * 6 + 3 + j] = X.E[i * 3 + j];
      S acc = S(0);
      for (int k = 0; k < 3; ++k) acc += X.E[i * 3 + k] * rx[k * 3 + j];
      M[(3 + i) * 6 + j] = -acc;
    }
}

// ----------------------------------------------------------------------------------
// Model (replaces pinocchio::ModelTpl + the pendulum model)
// ----------------------------------------------------------------------------------
template <class S> struct Model {
  int kind = KIND_REVOLUTE_TREE;
  int n = 0;                 // number of one-degree-of-freedom joints (nq = nv = n)
  int jtype[MAXJ] = {0, 0, 0, 0, 0, 0, 0, 0};  // JointType: the motion subspace is [axis; 0] (revolute) or [0; axis] (prismatic)
  int parent[MAXJ];          // -1 = fixed base
  S Rp[MAXJ][9];             // joint placement rotation (joint frame -> parent coordinates), row-major
  S pp[MAXJ][3];             // joint placement translation in parent coordinates
  S axis[MAXJ][3];           // unit joint axis in the joint frame
  int so(int i) const { return jtype[i] == JOINT_PRISMATIC ? 3 : 0; }  // offset of the axis inside the 6-vector S_i
  S mass[MAXJ];
  S com[MAXJ][3];            // centre of mass in body frame
  S Ic[MAXJ][6];             // rotational inertia about the COM: xx,xy,xz,yy,yz,zz
  S grav[3];                 // gravity field (0,0,-9.81)
  int nframes = 0;
  int fparent[MAXF];         // joint each frame is attached to
  S fR[MAXF][9];
  S fp[MAXF][3];
  S pend_m = S(1), pend_l = S(1);  // pendulum_model.hpp:23-24
  int nq() const { return kind == KIND_PENDULUM ? 1 : n; }
  int nv() const { return nq(); }
};

// spatial inertia (6x6 row-major) of body i in its own frame
template <class S> inline void body_inertia(const Model<S>& m, int i, S* I6) {
  const S* c = m.com[i];
  S ms = m.mass[i];
  S Ic[9] = {m.Ic[i][0], m.Ic[i][1], m.Ic[i][2], m.Ic[i][1], m.Ic[i][3], m.Ic[i][4], m.Ic[i][2], m.Ic[i][4], m.Ic[i][5]};
  S cx[9] = {S(0), -c[2], c[1], c[2], S(0), -c[0], -c[1], c[0], S(0)};
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) {
      S cc = S(0);
      for (int k = 0; k < 3; ++k) cc += cx[a * 3 + k] * cx[k * 3 + b];
      I6[a * 6 + b] = Ic[a * 3 + b] - ms * cc;       // I_o = I_c - m cx cx
      I6[a * 6 + 3 + b] = ms * cx[a * 3 + b];        // m cx
      I6[(3 + a) * 6 + b] = -(ms * cx[a * 3 + b]);   // -m cx
      I6[(3 + a) * 6 + 3 + b] = (a == b) ? ms : S(0);
    }
}
template <class S> inline void mat6_mul_vec(const S* M, const S* v, S* o) {
  S t[6];
  for (int i = 0; i < 6; ++i) {
    S acc = S(0);
    for (int j = 0; j < 6; ++j) acc += M[i * 6 + j] * v[j];
    t[i] = acc;
  }
  for (int i = 0; i < 6; ++i) o[i] = t[i];
}

// joint transforms X_i(q_i) = X_J(q_i) X_T(i)   (Rodrigues rotation about the joint axis)
template <class S> inline void joint_xform(const Model<S>& m, int i, S q, Xform<S>& X) {
  const S* a = m.axis[i];
  if (m.jtype[i] == JOINT_PRISMATIC) {
    // translation by q along the axis: E = Rp^T, child origin = pp + Rp (axis q)
    S d[3] = {a[0] * q, a[1] * q, a[2] * q}, t[3];
    mat3_mul_vec(m.Rp[i], d, t);
    for (int r = 0; r < 3; ++r)
      for (int cidx = 0; cidx < 3; ++cidx) X.E[cidx * 3 + r] = m.Rp[i][r * 3 + cidx];
    for (int k = 0; k < 3; ++k) X.r[k] = m.pp[i][k] + t[k];
    return;
  }
  S s = Sin(q), c = Cos(q), v = S(1) - c;
  S RJ[9] = {c + v * a[0] * a[0],        v * a[0] * a[1] - s * a[2], v * a[0] * a[2] + s * a[1],
             v * a[1] * a[0] + s * a[2], c + v * a[1] * a[1],        v * a[1] * a[2] - s * a[0],
             v * a[2] * a[0] - s * a[1], v * a[2] * a[1] + s * a[0], c + v * a[2] * a[2]};
  // R = Rp * RJ (child -> parent);  E = R^T
  for (int r = 0; r < 3; ++r)
    for (int cidx = 0; cidx < 3; ++cidx) {
      S acc = S(0);
      for (int k = 0; k < 3; ++k) acc += m.Rp[i][r * 3 + k] * RJ[k * 3 + cidx];
      X.E[cidx * 3 + r] = acc;
    }
  for (int k = 0; k < 3; ++k) X.r[k] = m.pp[i][k];
}

// workspace of one ABA evaluation, kept for the derivative passes
template <class S> struct AbaData {
  Xform<S> X[MAXJ];
  S I6[MAXJ][36];
  S v[MAXJ][6], c[MAXJ][6], a[MAXJ][6];
  S pA[MAXJ][6], IA[MAXJ][36], U[MAXJ][6], Dinv[MAXJ], u[MAXJ];
  S qdd[MAXJ];
};

// Featherstone articulated-body algorithm (what pinocchio::aba computes,
// pinocchio_model.ipp:354): gravity enters as base acceleration -g.
template <class S>
inline void aba(const Model<S>& m, const S* q, const S* qd, const S* tau, AbaData<S>& d) {
  const int n = m.n;
  for (int i = 0; i < n; ++i) {
    joint_xform(m, i, q[i], d.X[i]);
    body_inertia(m, i, d.I6[i]);
    const int o = m.so(i);
    S Sq[6] = {S(0), S(0), S(0), S(0), S(0), S(0)};
    for (int k = 0; k < 3; ++k) Sq[o + k] = m.axis[i][k] * qd[i];
    if (m.parent[i] >= 0) {
      xf_motion(d.X[i], d.v[m.parent[i]], d.v[i]);
      for (int k = 0; k < 6; ++k) d.v[i][k] += Sq[k];
    } else {
      for (int k = 0; k < 6; ++k) d.v[i][k] = Sq[k];
    }
    crm(d.v[i], Sq, d.c[i]);
    S h[6];
    mat6_mul_vec(d.I6[i], d.v[i], h);
    crf(d.v[i], h, d.pA[i]);
    for (int k = 0; k < 36; ++k) d.IA[i][k] = d.I6[i][k];
  }
  for (int i = n - 1; i >= 0; --i) {
    const int o = m.so(i);
    S Sv[6] = {S(0), S(0), S(0), S(0), S(0), S(0)};
    for (int k = 0; k < 3; ++k) Sv[o + k] = m.axis[i][k];
    mat6_mul_vec(d.IA[i], Sv, d.U[i]);
    S D = S(0), SpA = S(0);
    for (int k = 0; k < 3; ++k) { D += Sv[o + k] * d.U[i][o + k]; SpA += Sv[o + k] * d.pA[i][o + k]; }
    d.Dinv[i] = S(1) / D;
    d.u[i] = tau[i] - SpA;
    if (m.parent[i] >= 0) {
      S Ia[36], pa[6], Iac[6];
      for (int r = 0; r < 6; ++r)
        for (int cidx = 0; cidx < 6; ++cidx) Ia[r * 6 + cidx] = d.IA[i][r * 6 + cidx] - d.U[i][r] * d.Dinv[i] * d.U[i][cidx];
      mat6_mul_vec(Ia, d.c[i], Iac);
      for (int k = 0; k < 6; ++k) pa[k] = d.pA[i][k] + Iac[k] + d.U[i][k] * (d.Dinv[i] * d.u[i]);
      S Xd[36], T[36];
      xf_dense(d.X[i], Xd);
      // T = Ia * X ; IA_parent += X^T * T
      for (int r = 0; r < 6; ++r)
        for (int cidx = 0; cidx < 6; ++cidx) {
          S acc = S(0);
          for (int k = 0; k < 6; ++k) acc += Ia[r * 6 + k] * Xd[k * 6 + cidx];
          T[r * 6 + cidx] = acc;
        }
      S* IAp = d.IA[m.parent[i]];
      for (int r = 0; r < 6; ++r)
        for (int cidx = 0; cidx < 6; ++cidx) {
          S acc = S(0);
          for (int k = 0; k < 6; ++k) acc += Xd[k * 6 + r] * T[k * 6 + cidx];
          IAp[r * 6 + cidx] += acc;
        }
      S pf[6];
      xf_force_T(d.X[i], pa, pf);
      for (int k = 0; k < 6; ++k) d.pA[m.parent[i]][k] += pf[k];
    }
  }
  S a0[6] = {S(0), S(0), S(0), -m.grav[0], -m.grav[1], -m.grav[2]};
  for (int i = 0; i < n; ++i) {
    S ap[6];
    xf_motion(d.X[i], m.parent[i] >= 0 ? d.a[m.parent[i]] : a0, ap);
    for (int k = 0; k < 6; ++k) ap[k] += d.c[i][k];
    S Ua = S(0);
    for (int k = 0; k < 6; ++k) Ua += d.U[i][k] * ap[k];
    d.qdd[i] = d.Dinv[i] * (d.u[i] - Ua);
    for (int k = 0; k < 6; ++k) d.a[i][k] = ap[k];
    for (int k = 0; k < 3; ++k) d.a[i][m.so(i) + k] += m.axis[i][k] * d.qdd[i];
  }
}

// independent recursive Newton-Euler (only used to validate ABA in the tests)
template <class S>
inline void rnea(const Model<S>& m, const S* q, const S* qd, const S* qdd, S* tau) {
  const int n = m.n;
  Xform<S> X[MAXJ];
  S v[MAXJ][6], a[MAXJ][6], f[MAXJ][6];
  S a0[6] = {S(0), S(0), S(0), -m.grav[0], -m.grav[1], -m.grav[2]};
  for (int i = 0; i < n; ++i) {
    joint_xform(m, i, q[i], X[i]);
    S I6[36];
    body_inertia(m, i, I6);
    S Sq[6] = {S(0), S(0), S(0), S(0), S(0), S(0)};
    for (int k = 0; k < 3; ++k) Sq[m.so(i) + k] = m.axis[i][k] * qd[i];
    S vp[6] = {S(0), S(0), S(0), S(0), S(0), S(0)};
    if (m.parent[i] >= 0) xf_motion(X[i], v[m.parent[i]], vp);
    for (int k = 0; k < 6; ++k) v[i][k] = vp[k] + Sq[k];
    S ap[6], cr[6];
    xf_motion(X[i], m.parent[i] >= 0 ? a[m.parent[i]] : a0, ap);
    crm(v[i], Sq, cr);
    for (int k = 0; k < 6; ++k) a[i][k] = ap[k] + cr[k];
    for (int k = 0; k < 3; ++k) a[i][m.so(i) + k] += m.axis[i][k] * qdd[i];
    S Ia[6], Iv[6], vf[6];
    mat6_mul_vec(I6, a[i], Ia);
    mat6_mul_vec(I6, v[i], Iv);
    crf(v[i], Iv, vf);
    for (int k = 0; k < 6; ++k) f[i][k] = Ia[k] + vf[k];
  }
  for (int i = n - 1; i >= 0; --i) {
    S t = S(0);
    for (int k = 0; k < 3; ++k) t += m.axis[i][k] * f[i][m.so(i) + k];
    tau[i] = t;
    if (m.parent[i] >= 0) {
      S fp[6];
      xf_force_T(X[i], f[i], fp);
      for (int k = 0; k < 6; ++k) f[m.parent[i]][k] += fp[k];
    }
  }
}

// M^-1(q) column by column: ABA with zero velocity, zero gravity and unit torques,
// reusing U, D^-1 of the articulated inertias (third output of
// pinocchio::computeABADerivatives, pinocchio_model.ipp:391-399; full symmetric matrix).
template <class S>
inline void minv_from_aba(const Model<S>& m, const AbaData<S>& d, S* Minv /* n x n col-major */) {
  const int n = m.n;
  for (int col = 0; col < n; ++col) {
    S pA[MAXJ][6], u[MAXJ], a[MAXJ][6];
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < 6; ++k) pA[i][k] = S(0);
    for (int i = n - 1; i >= 0; --i) {
      S SpA = S(0);
      for (int k = 0; k < 3; ++k) SpA += m.axis[i][k] * pA[i][m.so(i) + k];
      u[i] = (i == col ? S(1) : S(0)) - SpA;
      if (m.parent[i] >= 0) {
        S pa[6], pf[6];
        for (int k = 0; k < 6; ++k) pa[k] = pA[i][k] + d.U[i][k] * (d.Dinv[i] * u[i]);
        xf_force_T(d.X[i], pa, pf);
        for (int k = 0; k < 6; ++k) pA[m.parent[i]][k] += pf[k];
      }
    }
    for (int i = 0; i < n; ++i) {
      S ap[6] = {S(0), S(0), S(0), S(0), S(0), S(0)};
      if (m.parent[i] >= 0) xf_motion(d.X[i], a[m.parent[i]], ap);
      S Ua = S(0);
      for (int k = 0; k < 6; ++k) Ua += d.U[i][k] * ap[k];
      S qdd = d.Dinv[i] * (u[i] - Ua);
      for (int k = 0; k < 6; ++k) a[i][k] = ap[k];
      for (int k = 0; k < 3; ++k) a[i][m.so(i) + k] += m.axis[i][k] * qdd;
      Minv[col * n + i] = qdd;
    }
  }
}

// d tau / d q_j and d tau / d qd_j of inverse dynamics at (q, qd, qdd = aba result):
// forward-mode tangent of RNEA in body coordinates.
template <class S>
inline void rnea_tangents(const Model<S>& m, const S* qd, const AbaData<S>& d, S* dtau_dq, S* dtau_dv /* n x n col-major */) {
  const int n = m.n;
  // RNEA forces at the ABA solution, accumulated over subtrees
  S F[MAXJ][6], h[MAXJ][6];
  for (int i = 0; i < n; ++i) {
    S Ia[6], vf[6];
    mat6_mul_vec(d.I6[i], d.a[i], Ia);
    mat6_mul_vec(d.I6[i], d.v[i], h[i]);
    crf(d.v[i], h[i], vf);
    for (int k = 0; k < 6; ++k) F[i][k] = Ia[k] + vf[k];
  }
  for (int i = n - 1; i >= 0; --i)
    if (m.parent[i] >= 0) {
      S fp[6];
      xf_force_T(d.X[i], F[i], fp);
      for (int k = 0; k < 6; ++k) F[m.parent[i]][k] += fp[k];
    }
  for (int wrt_v = 0; wrt_v < 2; ++wrt_v)
    for (int j = 0; j < n; ++j) {
      S dv[MAXJ][6], da[MAXJ][6], dF[MAXJ][6];
      for (int i = 0; i < n; ++i) {
        S Sj[6] = {S(0), S(0), S(0), S(0), S(0), S(0)};
        for (int k = 0; k < 3; ++k) Sj[m.so(i) + k] = m.axis[i][k];
        S Sq[6];
        for (int k = 0; k < 6; ++k) Sq[k] = Sj[k] * qd[i];
        for (int k = 0; k < 6; ++k) { dv[i][k] = S(0); da[i][k] = S(0); }
        if (m.parent[i] >= 0) {
          xf_motion(d.X[i], dv[m.parent[i]], dv[i]);
          xf_motion(d.X[i], da[m.parent[i]], da[i]);
        }
        if (i == j) {
          if (!wrt_v) {
            // d(X_j m)/dq_j = (X_j m) x S_j
            S Xv[6], Xa[6], t[6];
            for (int k = 0; k < 6; ++k) {
              Xv[k] = d.v[i][k] - Sq[k];
              Xa[k] = d.a[i][k] - d.c[i][k] - Sj[k] * d.qdd[i];
            }
            crm(Xv, Sj, t);
            for (int k = 0; k < 6; ++k) dv[i][k] += t[k];
            crm(Xa, Sj, t);
            for (int k = 0; k < 6; ++k) da[i][k] += t[k];
          } else {
            S t[6];
            for (int k = 0; k < 6; ++k) dv[i][k] += Sj[k];
            crm(d.v[i], Sj, t);
            for (int k = 0; k < 6; ++k) da[i][k] += t[k];
          }
        }
        S t[6];
        crm(dv[i], Sq, t);
        for (int k = 0; k < 6; ++k) da[i][k] += t[k];
        S Ida[6], Idv[6], t1[6], t2[6];
        mat6_mul_vec(d.I6[i], da[i], Ida);
        mat6_mul_vec(d.I6[i], dv[i], Idv);
        crf(dv[i], h[i], t1);
        crf(d.v[i], Idv, t2);
        for (int k = 0; k < 6; ++k) dF[i][k] = Ida[k] + t1[k] + t2[k];
      }
      S* out = wrt_v ? dtau_dv : dtau_dq;
      for (int i = n - 1; i >= 0; --i) {
        S t = S(0);
        for (int k = 0; k < 3; ++k) t += m.axis[i][k] * dF[i][m.so(i) + k];
        out[j * n + i] = t;
        if (m.parent[i] >= 0) {
          S tot[6], fp[6];
          for (int k = 0; k < 6; ++k) tot[k] = dF[i][k];
          if (i == j && !wrt_v) {
            S Sj[6] = {S(0), S(0), S(0), S(0), S(0), S(0)};
            for (int k = 0; k < 3; ++k) Sj[m.so(i) + k] = m.axis[i][k];
            S t3[6];
            crf(Sj, F[i], t3);
            for (int k = 0; k < 6; ++k) tot[k] += t3[k];
          }
          xf_force_T(d.X[i], tot, fp);
          for (int k = 0; k < 6; ++k) dF[m.parent[i]][k] += fp[k];
        }
      }
    }
}

// ----------------------------------------------------------------------------------
// Model concept (pinocchio_model.hpp:82-185, pendulum_model.hpp:35-132)
// ----------------------------------------------------------------------------------
// pendulum_model.hpp:105-130.  `g` is a `long double` constant there (:26), so with
// scalar_t = double the expression -g / l * sin(q) + tau / m is evaluated in extended
// precision and rounded once on assignment; the double instantiation below does the same.
template <class S> inline S pend_acc(const Model<S>& m, S q, S tau) {
  S g = S(9.81L);
  return -g / m.pend_l * Sin(q) + tau / m.pend_m;
}
template <class S> inline S pend_dacc_dq(const Model<S>& m, S q) {
  S g = S(9.81L);
  return -g / m.pend_l * Cos(q);
}
template <> inline double pend_acc<double>(const Model<double>& m, double q, double tau) {
  const long double g = 9.81L;
  return (double)(-g / m.pend_l * std::sin(q) + tau / m.pend_m);
}
template <> inline double pend_dacc_dq<double>(const Model<double>& m, double q) {
  const long double g = 9.81L;
  return (double)(-g / m.pend_l * std::cos(q));
}

template <class S> inline void dynamics_aba(const Model<S>& m, const S* q, const S* v, const S* tau, S* out_a) {
  if (m.kind == KIND_PENDULUM) {
    out_a[0] = pend_acc(m, q[0], tau[0]);
    return;
  }
  AbaData<S> d;
  aba(m, q, v, tau, d);
  for (int i = 0; i < m.n; ++i) out_a[i] = d.qdd[i];
}
// d_dynamics_aba: da/dq, da/dv, da/dtau = M^-1 (nv x nv col-major, fully overwritten)
template <class S>
inline void d_dynamics_aba(const Model<S>& m, const S* q, const S* v, const S* tau, S* dq, S* dv, S* dtau) {
  if (m.kind == KIND_PENDULUM) {
    dq[0] = pend_dacc_dq(m, q[0]);
    dv[0] = S(0);
    dtau[0] = S(1) / m.pend_m;
    return;
  }
  const int n = m.n;
  AbaData<S> d;
  aba(m, q, v, tau, d);
  minv_from_aba(m, d, dtau);
  S tq[MAXJ * MAXJ], tv[MAXJ * MAXJ];
  rnea_tangents(m, v, d, tq, tv);
  for (int c = 0; c < n; ++c)
    for (int r = 0; r < n; ++r) {
      S aq = S(0), av = S(0);
      for (int k = 0; k < n; ++k) {
        aq += dtau[k * n + r] * tq[c * n + k];
        av += dtau[k * n + r] * tv[c * n + k];
      }
      dq[c * n + r] = -aq;
      dv[c * n + r] = -av;
    }
}

// world placement of every joint frame (for frame_coordinates / d_frame_coordinates)
template <class S> struct WorldKin { S R[MAXJ][9]; S p[MAXJ][3]; };
template <class S> inline void forward_kinematics(const Model<S>& m, const S* q, WorldKin<S>& w) {
  for (int i = 0; i < m.n; ++i) {
    Xform<S> X;
    joint_xform(m, i, q[i], X);
    // local rotation child->parent is E^T
    S Rl[9];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) Rl[r * 3 + c] = X.E[c * 3 + r];
    if (m.parent[i] < 0) {
      for (int k = 0; k < 9; ++k) w.R[i][k] = Rl[k];
      for (int k = 0; k < 3; ++k) w.p[i][k] = X.r[k];
    } else {
      const S* Rpar = w.R[m.parent[i]];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
          S acc = S(0);
          for (int k = 0; k < 3; ++k) acc += Rpar[r * 3 + k] * Rl[k * 3 + c];
          w.R[i][r * 3 + c] = acc;
        }
      S t[3];
      mat3_mul_vec(Rpar, X.r, t);
      for (int k = 0; k < 3; ++k) w.p[i][k] = w.p[m.parent[i]][k] + t[k];
    }
  }
}
// frame_coordinates: translation of oMf (pinocchio_model.ipp:417-430)
template <class S> inline void frame_coordinates(const Model<S>& m, int frame, const S* q, S* out3) {
  WorldKin<S> w;
  forward_kinematics(m, q, w);
  int j = m.fparent[frame];
  S t[3];
  mat3_mul_vec(w.R[j], m.fp[frame], t);
  for (int k = 0; k < 3; ++k) out3[k] = w.p[j][k] + t[k];
}
// d_frame_coordinates: top 3 (linear) rows of getFrameJacobian(..., WORLD)
// (pinocchio_model.ipp:433-462).  WORLD convention = spatial velocity of the
// supporting joint axes taken at the WORLD ORIGIN: p_j x w_j, which equals dp/dq
// only where the frame sits at the origin.  mode
// FJ_LOCAL_WORLD_ALIGNED gives the true dp/dq = w_j x (p_f - p_j) instead.
template <class S>
inline void d_frame_coordinates(const Model<S>& m, int frame, const S* q, int mode, S* out /* 3 x nv col-major */) {
  WorldKin<S> w;
  forward_kinematics(m, q, w);
  const int n = m.n;
  for (int k = 0; k < 3 * n; ++k) out[k] = S(0);
  S pf[3], t[3];
  int jf = m.fparent[frame];
  mat3_mul_vec(w.R[jf], m.fp[frame], t);
  for (int k = 0; k < 3; ++k) pf[k] = w.p[jf][k] + t[k];
  for (int j = jf; j >= 0; j = m.parent[j]) {
    S wj[3], lin[3];
    mat3_mul_vec(w.R[j], m.axis[j], wj);
    if (m.jtype[j] == JOINT_PRISMATIC) {
      for (int k = 0; k < 3; ++k) lin[k] = wj[k];  // linear part of [0; w_j]: the same in both conventions
    } else if (mode == FJ_WORLD) {
      cross3(w.p[j], wj, lin);
    } else {
      S dlt[3] = {pf[0] - w.p[j][0], pf[1] - w.p[j][1], pf[2] - w.p[j][2]};
      cross3(wj, dlt, lin);
    }
    for (int k = 0; k < 3; ++k) out[j * 3 + k] = lin[k];
  }
}

// ----------------------------------------------------------------------------------
// dynamics_t (problem.hpp:343-525): explicit Euler on ABA
// ----------------------------------------------------------------------------------
template <class S> struct Dynamics {
  const Model<S>* model;
  S dt;
  bool second_order_finite_diff;  // true: value-based FD (second_order_deriv_2); false: FD of Jacobians (_1)
  int nq() const { return model->nq(); }
  int nv() const { return model->nv(); }
  int nx() const { return nq() + nv(); }
  int ndx() const { return 2 * nv(); }
  int nu() const { return nv(); }

  // problem.hpp:441-461
  void eval_to(S* x_out, const S* x, const S* u) const {
    const int n = nv();
    const S* q = x;
    const S* v = x + n;
    S acc[MAXJ];
    for (int i = 0; i < n; ++i) x_out[n + i] = dt * v[i];          // v_out = dt * v
    for (int i = 0; i < n; ++i) x_out[i] = q[i] + x_out[n + i];    // integrate (vector space: q + v)
    dynamics_aba(*model, q, v, u, acc);                             // v_out = acc
    for (int i = 0; i < n; ++i) x_out[n + i] = v[i] + acc[i] * dt;  // v_out = v + v_out * dt
  }
  // problem.hpp:463-503   fx: ndx x ndx col-major, fu: ndx x nu col-major
  void first_order_deriv(S* fx, S* fu, S* f, const S* x, const S* u) const {
    eval_to(f, x, u);
    const int n = nv(), nd = 2 * n;
    S dq[MAXJ * MAXJ], dv[MAXJ * MAXJ], dtau[MAXJ * MAXJ];
    d_dynamics_aba(*model, x, x + n, u, dq, dv, dtau);
    for (int c = 0; c < n; ++c)
      for (int r = 0; r < n; ++r) {
        S id = (r == c) ? S(1) : S(0);
        fx[c * nd + r] = id;                               // d_integrate_dq = I
        fx[(n + c) * nd + r] = id * dt;                    // d_integrate_dv * dt
        fx[c * nd + n + r] = dq[c * n + r] * dt;           // fx_bot_left *= dt
        fx[(n + c) * nd + n + r] = dv[c * n + r] * dt + id;  // fx_bot_right = dt*dv + I
        fu[c * nd + r] = S(0);                             // fu_top = 0
        fu[c * nd + n + r] = dtau[c * n + r] * dt;         // fu_bot *= dt
      }
  }
};

// ----------------------------------------------------------------------------------
// "Fn" concept for the finite-difference Hessian driver (dynamics or a constraint)
// ----------------------------------------------------------------------------------
template <class S> struct Fn {
  virtual ~Fn() {}
  virtual int out_dim(int t) const = 0;
  virtual void eval_to(S* out, int t, const S* x, const S* u) const = 0;
  virtual void first_order_deriv(S* out_x, S* out_u, S* out, int t, const S* x, const S* u) const = 0;
  virtual const Dynamics<S>& dyn() const = 0;
  // integrate_x / integrate_u / difference_out: vector-space (revolute, pendulum)
};

template <class S> struct DynamicsFn : Fn<S> {
  Dynamics<S> d;
  explicit DynamicsFn(const Dynamics<S>& dd) : d(dd) {}
  int out_dim(int) const override { return d.nx(); }
  void eval_to(S* out, int, const S* x, const S* u) const override { d.eval_to(out, x, u); }
  void first_order_deriv(S* ox, S* ou, S* o, int, const S* x, const S* u) const override { d.first_order_deriv(ox, ou, o, x, u); }
  const Dynamics<S>& dyn() const override { return d; }
};

// per-knot equality targets: ne[t] rows for BASE index t (the reference's base indexer lives
// on [2, T+2), test/pendulum_ddp.cpp:37-47); target(t) has ne[t] entries.
template <class S> struct Targets {
  int t_begin = 0, t_end = 0;      // base index range
  std::vector<int> ne;             // rows per base index
  std::vector<int> off;
  std::vector<S> data;
  int rows(int t) const { return (t < t_begin || t >= t_end) ? 0 : ne[t - t_begin]; }
  const S* at(int t) const { return data.data() + off[t - t_begin]; }
};

// config_constraint_t (problem.hpp:744-864): eq = q - target, eq_x = [I 0], eq_u = 0
template <class S> struct ConfigConstraint : Fn<S> {
  Dynamics<S> d;
  const Targets<S>* tg;
  ConfigConstraint(const Dynamics<S>& dd, const Targets<S>* t) : d(dd), tg(t) {}
  int out_dim(int t) const override { return tg->rows(t); }
  void eval_to(S* out, int t, const S* x, const S*) const override {
    int ne = tg->rows(t);
    if (ne == 0) return;
    const S* target = tg->at(t);
    for (int i = 0; i < ne; ++i) out[i] = x[i] - target[i];  // difference_out(target_q, x_q) = finish - start
  }
  void first_order_deriv(S* ox, S* ou, S* o, int t, const S* x, const S* u) const override {
    int ne = tg->rows(t);
    if (ne == 0) return;
    eval_to(o, t, x, u);
    const int nd = d.ndx(), nu = d.nu();
    for (int c = 0; c < nd; ++c)
      for (int r = 0; r < ne; ++r) ox[c * ne + r] = (c == r) ? S(1) : S(0);  // d_difference_dq_finish = I
    for (int k = 0; k < ne * nu; ++k) ou[k] = S(0);
  }
  const Dynamics<S>& dyn() const override { return d; }
};

// spatial_constraint_t (problem.hpp:631-742): eq = p_frame(q) - target (3 rows)
template <class S> struct SpatialConstraint : Fn<S> {
  Dynamics<S> d;
  const Targets<S>* tg;
  int frame;
  int jac_mode;
  SpatialConstraint(const Dynamics<S>& dd, const Targets<S>* t, int f, int jm) : d(dd), tg(t), frame(f), jac_mode(jm) {}
  int out_dim(int t) const override { return tg->rows(t); }
  void eval_to(S* out, int t, const S* x, const S*) const override {
    if (tg->rows(t) == 0) return;
    S p[3];
    frame_coordinates(*d.model, frame, x, p);
    const S* target = tg->at(t);
    for (int i = 0; i < 3; ++i) out[i] = p[i] - target[i];
  }
  void first_order_deriv(S* ox, S* ou, S* o, int t, const S* x, const S* u) const override {
    if (tg->rows(t) == 0) return;
    eval_to(o, t, x, u);
    const int nv = d.nv(), nd = d.ndx(), nu = d.nu();
    S J[3 * MAXJ];
    d_frame_coordinates(*d.model, frame, x, jac_mode, J);
    for (int c = 0; c < nd; ++c)
      for (int r = 0; r < 3; ++r) ox[c * 3 + r] = (c < nv) ? J[c * 3 + r] : S(0);
    for (int k = 0; k < 3 * nu; ++k) ou[k] = S(0);
  }
  const Dynamics<S>& dyn() const override { return d; }
};

// constraint_advance_time_t (problem.hpp:527-624): eq(t,x,u) = c(t+1, f(x,u), u)
template <class S> struct AdvanceTime : Fn<S> {
  const Fn<S>* inner;
  explicit AdvanceTime(const Fn<S>* c) : inner(c) {}
  int out_dim(int t) const override { return inner->out_dim(t + 1); }
  void eval_to(S* out, int t, const S* x, const S* u) const override {
    S xn[2 * MAXJ];
    inner->dyn().eval_to(xn, x, u);
    inner->eval_to(out, t + 1, xn, u);
  }
  void first_order_deriv(S* ox, S* ou, S* o, int t, const S* x, const S* u) const override {
    const int ne = out_dim(t);
    if (ne == 0) return;  // problem.hpp:578-580
    const Dynamics<S>& d = inner->dyn();
    const int nd = d.ndx(), nu = d.nu();
    S xn[2 * MAXJ], fx[4 * MAXJ * MAXJ], fu[2 * MAXJ * MAXJ];
    d.first_order_deriv(fx, fu, xn, x, u);
    S ex[MAXE * 2 * MAXJ], eu[MAXE * MAXJ];
    inner->first_order_deriv(ex, eu, o, t + 1, xn, u);
    // out_x = eq_n_x * fx_n ; out_u = eq_n_x * fu_n   (problem.hpp:603-604)
    for (int c = 0; c < nd; ++c)
      for (int r = 0; r < ne; ++r) {
        S acc = S(0);
        for (int k = 0; k < nd; ++k) acc += ex[k * ne + r] * fx[c * nd + k];
        ox[c * ne + r] = acc;
      }
    for (int c = 0; c < nu; ++c)
      for (int r = 0; r < ne; ++r) {
        S acc = S(0);
        for (int k = 0; k < nd; ++k) acc += ex[k * ne + r] * fu[c * nd + k];
        ou[c * ne + r] = acc;
      }
  }
  const Dynamics<S>& dyn() const override { return inner->dyn(); }
};

// ----------------------------------------------------------------------------------
// finite_diff_hessian_compute (problem.hpp:50-341)
// tensors: T(k,i,j) at k + i*O + j*O*L ; xx: O x nx x nx, ux: O x nu x nx, uu: O x nu x nu
// ----------------------------------------------------------------------------------
template <class S>
inline void second_order_deriv_1(const Fn<S>& fn, S* xx, S* ux, S* uu, S* ox, S* ou, S* o, int t, const S* x, const S* u) {
  // problem.hpp:67-150
  const Dynamics<S>& d = fn.dyn();
  const int no = fn.out_dim(t), nx = d.ndx(), nu = d.nu(), nxs = d.nx();
  fn.first_order_deriv(ox, ou, o, t, x, u);
  std::vector<S> fx_(ox, ox + no * nx), fu_(ou, ou + no * nu), out_(o, o + no);
  S x_[2 * MAXJ], u_[MAXJ], dx[2 * MAXJ], du[MAXJ];
  for (int i = 0; i < nx; ++i) dx[i] = S(0);
  for (int i = 0; i < nu; ++i) du[i] = S(0);
  S eps = Sqrt(Eps<S>::value());
  for (int i = 0; i < nx + nu; ++i) {
    bool at_x = i < nx;
    int idx = at_x ? i : i - nx;
    S& in_var = at_x ? dx[idx] : du[idx];
    in_var = eps;
    for (int k = 0; k < nxs; ++k) x_[k] = x[k] + dx[k];  // integrate_x (vector space)
    for (int k = 0; k < nu; ++k) u_[k] = u[k] + du[k];
    fn.first_order_deriv(fx_.data(), fu_.data(), out_.data(), t, x_, u_);
    if (at_x) {
      for (int k = 0; k < no; ++k) {
        for (int j = 0; j < nx; ++j) xx[k + j * no + idx * no * nx] = (fx_[j * no + k] - ox[j * no + k]) / eps;
        for (int j = 0; j < nu; ++j) ux[k + j * no + idx * no * nu] = (fu_[j * no + k] - ou[j * no + k]) / eps;
      }
    } else {
      for (int k = 0; k < no; ++k)
        for (int j = 0; j < nu; ++j) uu[k + j * no + idx * no * nu] = (fu_[j * no + k] - ou[j * no + k]) / eps;
    }
    in_var = S(0);
  }
}

template <class S>
inline void second_order_deriv_2(const Fn<S>& fn, S* xx, S* ux, S* uu, S* ox, S* ou, S* o, int t, const S* x, const S* u) {
  // problem.hpp:152-298
  const Dynamics<S>& d = fn.dyn();
  const int no = fn.out_dim(t), nx = d.ndx(), nu = d.nu(), nxs = d.nx();
  fn.first_order_deriv(ox, ou, o, t, x, u);
  std::vector<S> f1(no > 0 ? no : 1, S(0)), df(no > 0 ? no : 1, S(0));
  S x1[2 * MAXJ], u1[MAXJ], dx[2 * MAXJ], du[MAXJ];
  for (int i = 0; i < nx; ++i) dx[i] = S(0);
  for (int i = 0; i < nu; ++i) du[i] = S(0);
  S eps = Sqrt(Sqrt(Eps<S>::value()));
  S eps2 = eps * eps;
  auto tensor_of = [&](bool ax1, bool ax2) -> S* { return ax1 ? (ax2 ? xx : ux) : uu; };
  auto tdiml = [&](bool ax1, bool ax2) -> int { return ax1 ? (ax2 ? nx : nu) : nu; };
  // diagonal (problem.hpp:192-222)
  for (int i = 0; i < nx + nu; ++i) {
    bool at_x = i < nx;
    int idx = at_x ? i : i - nx;
    S& in_var = at_x ? dx[idx] : du[idx];
    const S* f_col = at_x ? ox + idx * no : ou + idx * no;
    S* tensor = at_x ? xx : uu;
    int L = at_x ? nx : nu;
    in_var = eps;
    for (int k = 0; k < nxs; ++k) x1[k] = x[k] + dx[k];
    for (int k = 0; k < nu; ++k) u1[k] = u[k] + du[k];
    fn.eval_to(f1.data(), t, x1, u1);
    for (int k = 0; k < no; ++k) {
      df[k] = f1[k] - o[k];            // difference_out(f0, f1)
      df[k] -= eps * f_col[k];
      df[k] *= S(2);
      tensor[k + idx * no + idx * no * L] = df[k] / eps2;
    }
    in_var = S(0);
  }
  // off-diagonal (problem.hpp:226-296)
  for (int i = 0; i < nx + nu; ++i) {
    bool ax1 = i < nx;
    int i1 = ax1 ? i : i - nx;
    S& v1 = ax1 ? dx[i1] : du[i1];
    const S* col1 = ax1 ? ox + i1 * no : ou + i1 * no;
    S* t1 = ax1 ? xx : uu;
    int L1 = ax1 ? nx : nu;
    v1 = eps;
    for (int j = i + 1; j < nx + nu; ++j) {
      bool ax2 = j < nx;
      int i2 = ax2 ? j : j - nx;
      S& v2 = ax2 ? dx[i2] : du[i2];
      const S* col2 = ax2 ? ox + i2 * no : ou + i2 * no;
      S* t2 = ax2 ? xx : uu;
      int L2 = ax2 ? nx : nu;
      // at_x_1 && at_x_2 -> xx ; at_x_1 && !at_x_2 -> ux ; both u -> uu
      S* tensor = tensor_of(ax1, ax2);
      int L = tdiml(ax1, ax2);
      v2 = eps;
      for (int k = 0; k < nxs; ++k) x1[k] = x[k] + dx[k];
      for (int k = 0; k < nu; ++k) u1[k] = u[k] + du[k];
      fn.eval_to(f1.data(), t, x1, u1);
      for (int k = 0; k < no; ++k) {
        df[k] = f1[k] - o[k];
        df[k] -= eps * col1[k];
        df[k] -= eps * col2[k];
        df[k] *= S(2);
        S val = S(0.5) * (df[k] / eps2 - t1[k + i1 * no + i1 * no * L1] - t2[k + i2 * no + i2 * no * L2]);
        tensor[k + i2 * no + i1 * no * L] = val;
        if (ax1 == ax2) tensor[k + i1 * no + i2 * no * L] = val;
      }
      v2 = S(0);
    }
    v1 = S(0);
  }
}

template <class S>
inline void second_order_deriv(const Fn<S>& fn, S* xx, S* ux, S* uu, S* ox, S* ou, S* o, int t, const S* x, const S* u) {
  if (fn.dyn().second_order_finite_diff) second_order_deriv_2(fn, xx, ux, uu, ox, ou, o, t, x, u);
  else second_order_deriv_1(fn, xx, ux, uu, ox, ou, o, t, x, u);
}

// ----------------------------------------------------------------------------------
// packed per-problem sequences (mat_seq_t / tensor_seq_t layout, detail/mat_seq.hpp:10-128,
// detail/tensor.hpp:346-452): one flat array per sequence, knot t at offset sum_{t'<t} stride(t')
// ----------------------------------------------------------------------------------
template <class S> struct Problem {
  Model<S> model;
  Dynamics<S> dyn;
  int T = 0;
  S c = S(1);                      // cost weight, l = c/2 |u|^2 (problem.hpp:936-942)
  int eq_kind = EQ_NONE;
  int frame = 0, jac_mode = FJ_WORLD;
  int max_restarts = 0;            // guarded mode of the product (ddpb_params.max_backward_restarts); 0 = reference behaviour
  Targets<S> targets;              // base indices [2, T+2)
  // constraint chain advance(advance(inner))
  Fn<S>* inner = nullptr;
  AdvanceTime<S>* adv1 = nullptr;
  AdvanceTime<S>* adv2 = nullptr;
  std::vector<int> ne, eq_off;     // per knot rows (knot t <-> base index t+2) and prefix sums

  int nx() const { return dyn.nx(); }
  int ndx() const { return dyn.ndx(); }
  int nu() const { return dyn.nu(); }
  int ne_total() const { return eq_off.empty() ? 0 : eq_off.back(); }
  ~Problem() { delete adv2; delete adv1; delete inner; }
  Problem() = default;
  Problem(const Problem&) = delete;
  Problem& operator=(const Problem&) = delete;

  void finalize() {
    dyn.model = &model;
    ne.assign(T, 0);
    eq_off.assign(T + 1, 0);
    for (int t = 0; t < T; ++t) {
      ne[t] = (eq_kind == EQ_NONE) ? 0 : targets.rows(t + 2);
      eq_off[t + 1] = eq_off[t] + ne[t];
    }
    delete adv2; delete adv1; delete inner;
    if (eq_kind == EQ_SPATIAL) inner = new SpatialConstraint<S>(dyn, &targets, frame, jac_mode);
    else inner = new ConfigConstraint<S>(dyn, &targets);
    adv1 = new AdvanceTime<S>(inner);
    adv2 = new AdvanceTime<S>(adv1);
  }
  // eval_eq_to (problem.hpp:950): doubly advanced constraint; evaluated even when ne(t)==0
  void eval_eq_to(S* out, int t, const S* x, const S* u) const { adv2->eval_to(out, t, x, u); }
  S l(const S* u) const {
    S s = S(0);
    for (int i = 0; i < nu(); ++i) s += u[i] * u[i];
    return S(0.5) * c * s;
  }
};

template <class S> struct Derivs {
  // 17 sequences of derivative_storage_t (ddp.hpp:84-105) + lfx, lfxx
  std::vector<S> lfx, lfxx, lx, lu, lxx, lux, luu, f_val, fx, fu, fxx, fux, fuu, eq_val, eq_x, eq_u, eq_xx, eq_ux, eq_uu;
  void resize(const Problem<S>& p) {
    const int T = p.T, nd = p.ndx(), nu = p.nu(), nx = p.nx(), ne = p.ne_total();
    lfx.assign(nd, S(0)); lfxx.assign(nd * nd, S(0));
    lx.assign(T * nd, S(0)); lu.assign(T * nu, S(0)); lxx.assign(T * nd * nd, S(0));
    lux.assign(T * nu * nd, S(0)); luu.assign(T * nu * nu, S(0));
    f_val.assign(T * nx, S(0)); fx.assign(T * nd * nd, S(0)); fu.assign(T * nd * nu, S(0));
    fxx.assign(T * nd * nd * nd, S(0)); fux.assign(T * nd * nu * nd, S(0)); fuu.assign(T * nd * nu * nu, S(0));
    eq_val.assign(ne, S(0)); eq_x.assign(ne * nd, S(0)); eq_u.assign(ne * nu, S(0));
    eq_xx.assign(ne * nd * nd, S(0)); eq_ux.assign(ne * nu * nd, S(0)); eq_uu.assign(ne * nu * nu, S(0));
  }
};

// problem_t::compute_derivatives (problem.hpp:956-997); the MPFR-only Taylor self-check
// (:999-1139) and the per-knot chronometer file writes (:989-997) are not restated.
template <class S>
inline void compute_derivatives(const Problem<S>& p, Derivs<S>& d, const S* xs, const S* us, int t_begin = 0, int t_end = -1) {
  const int T = p.T, nd = p.ndx(), nu = p.nu(), nx = p.nx();
  if (t_end < 0) t_end = T;
  for (int k = 0; k < nd; ++k) d.lfx[k] = S(0);
  for (int k = 0; k < nd * nd; ++k) d.lfxx[k] = S(0);
  DynamicsFn<S> dynfn(p.dyn);
  for (int t = t_begin; t < t_end; ++t) {
    const S* x = xs + t * nx;
    const S* u = us + t * nu;
    for (int k = 0; k < nd; ++k) d.lx[t * nd + k] = S(0);
    for (int k = 0; k < nd * nd; ++k) d.lxx[t * nd * nd + k] = S(0);
    for (int k = 0; k < nu * nd; ++k) d.lux[t * nu * nd + k] = S(0);
    for (int k = 0; k < nu; ++k) d.lu[t * nu + k] = p.c * u[k];
    for (int a = 0; a < nu; ++a)
      for (int b = 0; b < nu; ++b) d.luu[t * nu * nu + a * nu + b] = (a == b ? S(1) : S(0)) * p.c;
    second_order_deriv<S>(dynfn, &d.fxx[t * nd * nd * nd], &d.fux[t * nd * nu * nd], &d.fuu[t * nd * nu * nu],
                          &d.fx[t * nd * nd], &d.fu[t * nd * nu], &d.f_val[t * nx], t, x, u);
    const int o = p.eq_off[t];
    S dummy[1];
    second_order_deriv<S>(*p.adv2, p.ne[t] ? &d.eq_xx[o * nd * nd] : dummy, p.ne[t] ? &d.eq_ux[o * nu * nd] : dummy,
                          p.ne[t] ? &d.eq_uu[o * nu * nu] : dummy, p.ne[t] ? &d.eq_x[o * nd] : dummy,
                          p.ne[t] ? &d.eq_u[o * nu] : dummy, p.ne[t] ? &d.eq_val[o] : dummy, t, x, u);
  }
}

// affine_vector_function_seq_t (detail/mat_seq_common.hpp:12-177): out(t)(x) = val + jac (x - origin)
template <class S> struct AffineSeq {
  std::vector<int> rows, off;  // per knot
  int nd = 0, nx = 0, T = 0;
  std::vector<S> origin, val, jac;
  void init(int T_, const std::vector<int>& r, int nx_, int nd_) {
    T = T_; rows = r; nx = nx_; nd = nd_;
    off.assign(T + 1, 0);
    for (int t = 0; t < T; ++t) off[t + 1] = off[t] + rows[t];
    origin.assign(T * nx, S(0)); val.assign(off[T], S(0)); jac.assign(off[T] * nd, S(0));
  }
  S* val_at(int t) { return val.data() + off[t]; }
  S* jac_at(int t) { return jac.data() + off[t] * nd; }  // rows[t] x nd col-major
  const S* val_at(int t) const { return val.data() + off[t]; }
  const S* jac_at(int t) const { return jac.data() + off[t] * nd; }
  // proxy_t::operator() (mat_seq_common.hpp:105-115)
  void eval(int t, const S* x, S* out) const {
    const int r = rows[t];
    const S* J = jac_at(t);
    for (int i = 0; i < r; ++i) out[i] = val_at(t)[i];
    for (int i = 0; i < r; ++i) {
      S acc = S(0);
      for (int k = 0; k < nd; ++k) acc += J[k * r + i] * (x[k] - origin[t * nx + k]);
      out[i] += acc;
    }
  }
  // update_origin (mat_seq_common.hpp:62-89); d_difference_dfinish = I for vector spaces
  void update_origin(const S* xs_new) {
    for (int t = 0; t < T; ++t) {
      const int r = rows[t];
      S* J = jac_at(t);
      for (int i = 0; i < r; ++i) {
        S acc = S(0);
        for (int k = 0; k < nd; ++k) acc += J[k * r + i] * (xs_new[t * nx + k] - origin[t * nx + k]);
        val_at(t)[i] += acc;
      }
      for (int k = 0; k < nx; ++k) origin[t * nx + k] = xs_new[t * nx + k];
    }
  }
};

// ----------------------------------------------------------------------------------
// backward_pass (ddp_bwd.ipp:9-155)
// ----------------------------------------------------------------------------------
template <class S> struct BwdResult { S mu, reg; int restarts; bool diverged; };

// in-place lower Cholesky (Eigen::LLT unblocked kernel semantics: fails when a pivot <= 0)
template <class S> inline bool llt_inplace(S* A, int n) {  // col-major, lower
  for (int k = 0; k < n; ++k) {
    S x = A[k * n + k];
    for (int j = 0; j < k; ++j) x -= A[j * n + k] * A[j * n + k];
    if (x <= S(0)) return false;
    x = Sqrt(x);
    A[k * n + k] = x;
    for (int i = k + 1; i < n; ++i) {
      S s = A[k * n + i];
      for (int j = 0; j < k; ++j) s -= A[j * n + i] * A[j * n + k];
      A[k * n + i] = s / x;
    }
  }
  return true;
}
template <class S> inline void llt_solve(const S* L, int n, S* b) {
  for (int i = 0; i < n; ++i) {
    S s = b[i];
    for (int j = 0; j < i; ++j) s -= L[j * n + i] * b[j];
    b[i] = s / L[i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    S s = b[i];
    for (int j = i + 1; j < n; ++j) s -= L[i * n + j] * b[j];
    b[i] = s / L[i * n + i];
  }
}

template <class S>
inline BwdResult<S> backward_pass(const Problem<S>& p, const Derivs<S>& d, const S* xs, const AffineSeq<S>& mults,
                                  AffineSeq<S>& fb, S reg, S mu) {
  const int T = p.T, nd = p.ndx(), nu = p.nu(), nx = p.nx();
  std::vector<S> Vxx(nd * nd), Vx(nd), Qx(nd), Qu(nu), Qxx(nd * nd), Quu(nu * nu), Qux(nu * nd), tmpM(nd * nd), A(nu * nu), rhs(nu);
  int restarts = 0;
  bool success = false;
  bool diverged = false;
  while (!success) {
    bool failed = false;
    for (int k = 0; k < nd * nd; ++k) Vxx[k] = d.lfxx[k];
    for (int k = 0; k < nd; ++k) Vx[k] = d.lfx[k];
    for (int t = T - 1; t >= 0; --t) {
      const int ne = p.ne[t], eo = p.eq_off[t];
      const S* pe = mults.val_at(t);
      const S* pex = mults.jac_at(t);
      const S* eqv = &d.eq_val[eo];
      const S* eqx = &d.eq_x[eo * nd];
      const S* equ = &d.eq_u[eo * nu];
      const S* fx = &d.fx[t * nd * nd];
      const S* fu = &d.fu[t * nd * nu];
      S tmp[MAXE], tmp2[MAXE * 2 * MAXJ];
      for (int i = 0; i < ne; ++i) tmp[i] = pe[i] + mu * eqv[i];
      for (int i = 0; i < ne * nd; ++i) tmp2[i] = pex[i] + mu * eqx[i];
      // Q_x
      for (int i = 0; i < nd; ++i) {
        S q = d.lx[t * nd + i];
        S a = S(0);
        for (int k = 0; k < nd; ++k) a += fx[i * nd + k] * Vx[k];
        q += a;
        a = S(0);
        for (int k = 0; k < ne; ++k) a += eqx[i * ne + k] * tmp[k];
        q += a;
        a = S(0);
        for (int k = 0; k < ne; ++k) a += pex[i * ne + k] * eqv[k];
        q += a;
        Qx[i] = q;
      }
      // Q_u
      for (int i = 0; i < nu; ++i) {
        S q = d.lu[t * nu + i];
        S a = S(0);
        for (int k = 0; k < nd; ++k) a += fu[i * nd + k] * Vx[k];
        q += a;
        a = S(0);
        for (int k = 0; k < ne; ++k) a += equ[i * ne + k] * tmp[k];
        q += a;
        Qu[i] = q;
      }
      // tmpM = Vxx * fx  (nd x nd), tmpU = Vxx * fu
      std::vector<S>& VF = tmpM;
      for (int c = 0; c < nd; ++c)
        for (int r = 0; r < nd; ++r) {
          S a = S(0);
          for (int k = 0; k < nd; ++k) a += Vxx[k * nd + r] * fx[c * nd + k];
          VF[c * nd + r] = a;
        }
      S VU[2 * MAXJ * MAXJ];
      for (int c = 0; c < nu; ++c)
        for (int r = 0; r < nd; ++r) {
          S a = S(0);
          for (int k = 0; k < nd; ++k) a += Vxx[k * nd + r] * fu[c * nd + k];
          VU[c * nd + r] = a;
        }
      // Q_xx
      for (int c = 0; c < nd; ++c)
        for (int r = 0; r < nd; ++r) {
          S q = d.lxx[t * nd * nd + c * nd + r];
          S a = S(0);
          for (int k = 0; k < nd; ++k) a += fx[r * nd + k] * VF[c * nd + k];
          q += a;
          a = S(0);
          for (int k = 0; k < ne; ++k) a += eqx[r * ne + k] * tmp2[c * ne + k];
          q += a;
          a = S(0);
          for (int k = 0; k < ne; ++k) a += pex[r * ne + k] * eqx[c * ne + k];
          q += a;
          a = S(0);
          for (int k = 0; k < ne; ++k) a += tmp[k] * d.eq_xx[eo * nd * nd + k + r * ne + c * ne * nd];
          q += a;
          a = S(0);
          for (int k = 0; k < nd; ++k) a += Vx[k] * d.fxx[t * nd * nd * nd + k + r * nd + c * nd * nd];
          q += a;
          Qxx[c * nd + r] = q;
        }
      // Q_uu
      for (int c = 0; c < nu; ++c)
        for (int r = 0; r < nu; ++r) {
          S q = d.luu[t * nu * nu + c * nu + r];
          S a = S(0);
          for (int k = 0; k < nd; ++k) a += fu[r * nd + k] * VU[c * nd + k];
          q += a;
          a = S(0);
          for (int k = 0; k < ne; ++k) a += equ[r * ne + k] * equ[c * ne + k];
          q += a * mu;
          a = S(0);
          for (int k = 0; k < ne; ++k) a += tmp[k] * d.eq_uu[eo * nu * nu + k + r * ne + c * ne * nu];
          q += a;
          a = S(0);
          for (int k = 0; k < nd; ++k) a += Vx[k] * d.fuu[t * nd * nu * nu + k + r * nd + c * nd * nu];
          q += a;
          Quu[c * nu + r] = q;
        }
      // Q_ux  (nu x nd)
      for (int c = 0; c < nd; ++c)
        for (int r = 0; r < nu; ++r) {
          S q = d.lux[t * nu * nd + c * nu + r];
          S a = S(0);
          for (int k = 0; k < nd; ++k) a += fu[r * nd + k] * VF[c * nd + k];
          q += a;
          a = S(0);
          for (int k = 0; k < ne; ++k) a += equ[r * ne + k] * tmp2[c * ne + k];
          q += a;
          a = S(0);
          for (int k = 0; k < ne; ++k) a += tmp[k] * d.eq_ux[eo * nu * nd + k + r * ne + c * ne * nu];
          q += a;
          a = S(0);
          for (int k = 0; k < nd; ++k) a += Vx[k] * d.fux[t * nd * nu * nd + k + r * nd + c * nd * nu];
          q += a;
          Qux[c * nu + r] = q;
        }
      // LLT of Q_uu + reg I (lower triangle)
      for (int k = 0; k < nu * nu; ++k) A[k] = Quu[k];
      for (int k = 0; k < nu; ++k) A[k * nu + k] += reg;
      if (!llt_inplace(A.data(), nu)) {
        // ddp_bwd.ipp:105-132
        if (reg < mu) reg = mu;
        mu *= S(2);
        reg *= S(2);
        ++restarts;
        failed = true;
        break;
      }
      for (int k = 0; k < nx; ++k) fb.origin[t * nx + k] = xs[t * nx + k];
      S* kk = fb.val_at(t);
      S* KK = fb.jac_at(t);
      for (int i = 0; i < nu; ++i) rhs[i] = -Qu[i];
      llt_solve(A.data(), nu, rhs.data());
      for (int i = 0; i < nu; ++i) kk[i] = rhs[i];
      for (int c = 0; c < nd; ++c) {
        for (int i = 0; i < nu; ++i) rhs[i] = -Qux[c * nu + i];
        llt_solve(A.data(), nu, rhs.data());
        for (int i = 0; i < nu; ++i) KK[c * nu + i] = rhs[i];
      }
      // V_x = Q_x + Q_ux^T k ; V_xx = Q_xx + Q_ux^T K
      for (int i = 0; i < nd; ++i) {
        S a = S(0);
        for (int k = 0; k < nu; ++k) a += Qux[i * nu + k] * kk[k];
        Vx[i] = Qx[i] + a;
      }
      for (int c = 0; c < nd; ++c)
        for (int r = 0; r < nd; ++r) {
          S a = S(0);
          for (int k = 0; k < nu; ++k) a += Qux[r * nu + k] * KK[c * nu + k];
          Vxx[c * nd + r] = Qxx[c * nd + r] + a;
        }
      if (t == 0) success = true;
    }
    if (failed && p.max_restarts > 0 && (restarts > p.max_restarts || !(mu < S(1.79e308)))) { diverged = true; break; }
  }
  return {mu, reg, restarts, diverged};
}

// ----------------------------------------------------------------------------------
// cost_seq_aug (ddp.hpp:699-735) and forward_pass (ddp_fwd.ipp:10-67)
// ----------------------------------------------------------------------------------
template <class S>
inline void cost_seq_aug(const Problem<S>& p, S* out /* T+1 */, const S* xs, const S* us, const AffineSeq<S>& mults, S mu) {
  const int T = p.T, nx = p.nx(), nu = p.nu();
  for (int t = 0; t < T; ++t) {
    const S* x = xs + t * nx;
    const S* u = us + t * nu;
    S ce[MAXE], pe[MAXE];
    S l = p.l(u);
    p.eval_eq_to(ce, t, x, u);
    mults.eval(t, x, pe);
    S dot = S(0), sq = S(0);
    for (int i = 0; i < p.ne[t]; ++i) { dot += pe[i] * ce[i]; sq += ce[i] * ce[i]; }
    out[t] = l + dot + (mu / S(2)) * sq;
  }
  out[T] = S(0);  // lf = 0 (problem.hpp:932-935)
}

template <class S> struct FwdResult { S step; int rollouts; S merit; };

template <class S>
inline FwdResult<S> forward_pass(const Problem<S>& p, S* xs_new, S* us_new, const S* xs, const S* us, const AffineSeq<S>& mults,
                                 const AffineSeq<S>& fb, S mu, bool do_linesearch = true) {
  const int T = p.T, nx = p.nx(), nu = p.nu(), nd = p.ndx();
  std::vector<S> c_old(T + 1), c_new(T + 1);
  if (do_linesearch) cost_seq_aug(p, c_old.data(), xs, us, mults, mu);
  S step = S(1);
  bool success = false;
  int rollouts = 0;
  S merit = S(0);
  while (!success) {
    if (step < S(1e-10)) break;
    ++rollouts;
    for (int t = 0; t < T; ++t) {
      S dx[2 * MAXJ];
      for (int k = 0; k < nd; ++k) dx[k] = xs_new[t * nx + k] - xs[t * nx + k];  // difference(x_old, x_new)
      const S* kk = fb.val_at(t);
      const S* KK = fb.jac_at(t);
      for (int i = 0; i < nu; ++i) {
        S a = S(0);
        for (int k = 0; k < nd; ++k) a += KK[k * nu + i] * dx[k];
        us_new[t * nu + i] = us[t * nu + i] + step * kk[i] + a;
      }
      p.dyn.eval_to(xs_new + (t + 1) * nx, xs_new + t * nx, us_new + t * nu);
    }
    if (do_linesearch) {
      cost_seq_aug(p, c_new.data(), xs_new, us_new, mults, mu);
      S sum = S(0);
      merit = S(0);
      for (int t = 0; t <= T; ++t) { sum += c_new[t] - c_old[t]; merit += c_new[t]; }
      if (sum <= S(0)) success = true;
      else step *= S(0.5);
    } else {
      success = true;
    }
  }
  return {step, rollouts, merit};
}

// ----------------------------------------------------------------------------------
// optimality measures (ddp.hpp:516-627) and update_derivatives (ddp.hpp:641-696)
// ----------------------------------------------------------------------------------
template <class S> inline S norm2(const S* v, int n) {
  S s = S(0);
  for (int i = 0; i < n; ++i) s += v[i] * v[i];
  return Sqrt(s);
}
template <class S> inline S optimality_constr(const Problem<S>& p, const Derivs<S>& d) {
  S r = S(0);
  for (int t = 0; t < p.T; ++t) {
    S nrm = norm2(&d.eq_val[p.eq_off[t]], p.ne[t]);
    if (r < nrm) r = nrm;
  }
  return r;
}
template <class S>
inline S optimality_obj(const Problem<S>& p, const S* xs, const AffineSeq<S>& mults, S mu, const Derivs<S>& d) {
  const int T = p.T, nd = p.ndx(), nu = p.nu(), nx = p.nx();
  S retval = S(0);
  std::vector<S> adj(d.lfx), nadj(nd);
  for (int t = T - 1; t >= 0; --t) {
    const int ne = p.ne[t], eo = p.eq_off[t];
    S pe[MAXE], lu[MAXJ];
    mults.eval(t, xs + t * nx, pe);
    const S* eqv = &d.eq_val[eo];
    const S* eqx = &d.eq_x[eo * nd];
    const S* equ = &d.eq_u[eo * nu];
    const S* fx = &d.fx[t * nd * nd];
    const S* fu = &d.fu[t * nd * nu];
    const S* pj = mults.jac_at(t);
    for (int i = 0; i < nu; ++i) {
      S v = d.lu[t * nu + i];
      S a = S(0);
      for (int k = 0; k < ne; ++k) a += pe[k] * equ[i * ne + k];
      v += a;
      a = S(0);
      for (int k = 0; k < ne; ++k) a += (mu * eqv[k]) * equ[i * ne + k];
      v += a;
      a = S(0);
      for (int k = 0; k < nd; ++k) a += adj[k] * fu[i * nd + k];
      v += a;
      lu[i] = v;
    }
    S nrm = norm2(lu, nu);
    if (retval < nrm) retval = nrm;
    for (int i = 0; i < nd; ++i) {
      S a = S(0);
      for (int k = 0; k < nd; ++k) a += adj[k] * fx[i * nd + k];
      S v = a;
      v += d.lx[t * nd + i];
      a = S(0);
      for (int k = 0; k < ne; ++k) a += (mu * eqv[k]) * eqx[i * ne + k];
      v += a;
      a = S(0);
      for (int k = 0; k < ne; ++k) a += pe[k] * eqx[i * ne + k];
      v += a;
      a = S(0);
      for (int k = 0; k < ne; ++k) a += eqv[k] * pj[i * ne + k];
      v += a;
      nadj[i] = v;
    }
    adj = nadj;
  }
  return retval;
}

enum MultUpdate { NO_UPDATE = 0, UPDATE_SUCCESS = 1, UPDATE_FAILURE = 2, OPTIMUM_ATTAINED = 3 };

template <class S>
inline int update_derivatives(const Problem<S>& p, Derivs<S>& d, AffineSeq<S>& fb, AffineSeq<S>& mults, const S* xs,
                              const S* us, S mu, S w, S n, S thr, S* opt_obj_out, S* opt_constr_out) {
  const int T = p.T, nd = p.ndx(), nu = p.nu();
  compute_derivatives(p, d, xs, us);
  mults.update_origin(xs);
  fb.update_origin(xs);
  S opt_obj = optimality_obj(p, xs, mults, mu, d);
  S opt_constr = optimality_constr(p, d);
  *opt_obj_out = opt_obj;
  *opt_constr_out = opt_constr;
  if (opt_constr < thr && opt_obj < thr) return OPTIMUM_ATTAINED;
  if (opt_obj < w) {
    if (opt_constr < n) {
      for (int t = 0; t < T; ++t) {
        const int ne = p.ne[t], eo = p.eq_off[t];
        const S* eqv = &d.eq_val[eo];
        const S* eqx = &d.eq_x[eo * nd];
        const S* equ = &d.eq_u[eo * nu];
        const S* kk = fb.val_at(t);
        const S* KK = fb.jac_at(t);
        S* pv = mults.val_at(t);
        S* pj = mults.jac_at(t);
        for (int i = 0; i < ne; ++i) {
          S a = S(0);
          for (int k = 0; k < nu; ++k) a += equ[k * ne + i] * kk[k];
          pv[i] += mu * (eqv[i] + a);
        }
        for (int c = 0; c < nd; ++c)
          for (int i = 0; i < ne; ++i) {
            S a = S(0);
            for (int k = 0; k < nu; ++k) a += equ[k * ne + i] * KK[c * nu + k];
            pj[c * ne + i] += mu * (eqx[c * ne + i] + a);
          }
      }
      return UPDATE_SUCCESS;
    }
    return UPDATE_FAILURE;
  }
  return NO_UPDATE;
}

// ----------------------------------------------------------------------------------
// solve (ddp.hpp:744-842), method::primal_dual_affine_multipliers
// ----------------------------------------------------------------------------------
template <class S> struct SolverParams { int max_iterations; S thr, mu, reg, w, n; };
template <class S> struct IterLog { S mu, reg, w, n, step, opt_obj, opt_constr; int rollouts, restarts, mult_update; };
template <class S> struct SolveResult {
  std::vector<S> xs, us;
  AffineSeq<S> fb, mults;
  std::vector<IterLog<S>> log;
  int iterations = 0;
  int status = 0;  // 0 = max iterations, 1 = optimum attained
};

// mult_jac_init replaces the reference's unseeded eq.jac().setRandom() (ddp.hpp:762)
template <class S>
inline void solve(const Problem<S>& p, const SolverParams<S>& sp, const S* xs0, const S* us0, const S* mult_jac_init, SolveResult<S>& R) {
  const int T = p.T, nx = p.nx(), nu = p.nu(), nd = p.ndx();
  Derivs<S> d;
  d.resize(p);
  std::vector<S> xs(xs0, xs0 + (T + 1) * nx), us(us0, us0 + T * nu);
  std::vector<S> xs_new(xs), us_new(us);
  S reg = sp.reg, mu = sp.mu, w = sp.w, n = sp.n;
  AffineSeq<S>& mults = R.mults;
  mults.init(T, p.ne, nx, nd);
  for (size_t k = 0; k < mults.jac.size(); ++k) mults.jac[k] = mult_jac_init ? mult_jac_init[k] : S(0);
  for (int t = 0; t < T; ++t)
    for (int k = 0; k < nx; ++k) mults.origin[t * nx + k] = xs[t * nx + k];
  AffineSeq<S>& fb = R.fb;
  fb.init(T, std::vector<int>(T, nu), nx, nd);
  compute_derivatives(p, d, xs.data(), us.data());
  BwdResult<S> b = backward_pass(p, d, xs.data(), mults, fb, reg, mu);
  mu = b.mu;  // reg from the warm-up pass is NOT taken (ddp.hpp:771)
  R.status = 0;
  if (b.diverged) R.status = 2;
  else forward_pass(p, xs_new.data(), us_new.data(), xs.data(), us.data(), mults, fb, mu, true);
  int iter = 0;
  for (; iter < sp.max_iterations && R.status != 2; ++iter) {
    IterLog<S> lg{};
    int rv = update_derivatives(p, d, fb, mults, xs.data(), us.data(), mu, w, n, sp.thr, &lg.opt_obj, &lg.opt_constr);
    lg.mult_update = rv;
    if (rv == UPDATE_FAILURE) {
      mu *= S(10);
    } else if (rv == UPDATE_SUCCESS) {
      S opt_obj = optimality_obj(p, xs.data(), mults, mu, d);
      n = opt_obj / Pow(mu, S(0.1L));
      w /= mu;
    } else if (rv == OPTIMUM_ATTAINED) {
      R.status = 1;
      R.log.push_back(lg);
      break;
    }
    b = backward_pass(p, d, xs.data(), mults, fb, reg, mu);
    mu = b.mu;
    reg = b.reg;
    lg.restarts = b.restarts;
    if (b.diverged) { R.status = 2; R.log.push_back(lg); break; }
    FwdResult<S> f = forward_pass(p, xs_new.data(), us_new.data(), xs.data(), us.data(), mults, fb, mu, true);
    if (p.max_restarts > 0 && (!(f.merit == f.merit) || !(f.merit < S(1.79e308)) || !(-f.merit < S(1.79e308)) || !(mu < S(1.79e308)))) {
      R.status = 2; R.log.push_back(lg); break;  // guarded mode: the trial trajectory is not swapped in
    }
    if (f.step >= S(0.5)) {
      reg /= S(2);
      if (reg < S(1e-5)) reg = S(0);
    }
    xs.swap(xs_new);
    us.swap(us_new);
    lg.mu = mu; lg.reg = reg; lg.w = w; lg.n = n; lg.step = f.step; lg.rollouts = f.rollouts;
    R.log.push_back(lg);
  }
  R.iterations = iter;
  R.xs = xs;
  R.us = us;
}

}  // namespace orc
