/*
 * wb_dynamics.hpp — CPU ORACLE (test infrastructure only): Pinocchio-free rigid-body algorithms for the
 * Mini-Cheetah whole-body model, numeric, templated on the scalar so that the same recursion also runs on
 * forward-mode dual numbers (that is how this oracle obtains what the reference gets from
 * pinocchio::computeRNEADerivatives / computeGeneralizedGravityDerivatives).
 *
 * Third-party dependency restated: Pinocchio 2.6.10 (README.md:8 of the reference; robotpkg install, not vendored).
 * Published algorithms used: recursive Newton-Euler (RNEA), unit-acceleration columns for the joint-space inertia
 * (equal to CRBA), frame Jacobian / velocity / classical acceleration in LOCAL_WORLD_ALIGNED.
 * Call sites followed: MHPC/MHPC-Trajopt/WBM.cpp:368-456 (KKT contact dynamics, impact), :459-543 (derivatives),
 * :260-364 (kinematics getters); tree: MHPC/MHPC-Trajopt/PinocchioInteface.cpp:17-56; inertial data and joint
 * placements: urdf/mini_cheetah_simple_correctedInertia.urdf:5-140 (FL leg; FR/HL/HR repeat the pattern).
 *
 * Pinned by: test/testKKTDynamics.cpp:95-121 (qdd / GRF known answers, reproduced to the printed 4 decimals when
 * the hip yaw is pi) and by the reference's CasADi kinematic partials (footVelPartialDq etc., machine precision).
 */
#pragma once
#include <cmath>
#include <cstring>

namespace oracle {
namespace wb {

/* ---- forward-mode dual number with N directions */
template <int N>
struct Dual {
  double v;
  double d[N];
  Dual() : v(0) { std::memset(d, 0, sizeof(d)); }
  Dual(double c) : v(c) { std::memset(d, 0, sizeof(d)); }
};
template <int N> inline Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v + b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
template <int N> inline Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v - b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
template <int N> inline Dual<N> operator-(const Dual<N>& a) { Dual<N> r; r.v = -a.v; for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r; }
template <int N> inline Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v * b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
template <int N> inline Dual<N> sin(const Dual<N>& a) { Dual<N> r; r.v = std::sin(a.v); double c = std::cos(a.v); for (int i = 0; i < N; ++i) r.d[i] = c * a.d[i]; return r; }
template <int N> inline Dual<N> cos(const Dual<N>& a) { Dual<N> r; r.v = std::cos(a.v); double s = -std::sin(a.v); for (int i = 0; i < N; ++i) r.d[i] = s * a.d[i]; return r; }
inline double sin(double a) { return std::sin(a); }
inline double cos(double a) { return std::cos(a); }

template <class S> struct V3 { S x, y, z; V3() : x(0.0), y(0.0), z(0.0) {} V3(S a, S b, S c) : x(a), y(b), z(c) {} };
template <class S> inline V3<S> operator+(const V3<S>& a, const V3<S>& b) { return V3<S>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <class S> inline V3<S> operator-(const V3<S>& a, const V3<S>& b) { return V3<S>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <class S> inline V3<S> operator*(const V3<S>& a, const S& s) { return V3<S>(a.x * s, a.y * s, a.z * s); }
template <class S> inline V3<S> cross(const V3<S>& a, const V3<S>& b) { return V3<S>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
template <class S> inline S dot(const V3<S>& a, const V3<S>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class S> struct M3 {
  S m[3][3];
  M3() { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m[i][j] = S(i == j ? 1.0 : 0.0); }
};
template <class S> inline M3<S> operator*(const M3<S>& A, const M3<S>& B) { M3<S> C; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) C.m[i][j] = A.m[i][0] * B.m[0][j] + A.m[i][1] * B.m[1][j] + A.m[i][2] * B.m[2][j]; return C; }
template <class S> inline V3<S> operator*(const M3<S>& A, const V3<S>& v) { return V3<S>(A.m[0][0] * v.x + A.m[0][1] * v.y + A.m[0][2] * v.z, A.m[1][0] * v.x + A.m[1][1] * v.y + A.m[1][2] * v.z, A.m[2][0] * v.x + A.m[2][1] * v.y + A.m[2][2] * v.z); }
template <class S> inline M3<S> transposed(const M3<S>& A) { M3<S> T; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) T.m[i][j] = A.m[j][i]; return T; }

/* ---- model data (urdf/mini_cheetah_simple_correctedInertia.urdf) */
struct Body { double m; double com[3]; double I[6]; /* ixx ixy ixz iyy iyz izz */ };
struct Leg { double abd_xyz[3], hip_xyz[3], knee_xyz[3], foot_xyz[3]; Body abd, thigh, shank; };
struct Params {
  double hip_yaw;
  Body body;
  Leg leg[4];  // FL FR HL HR
};
inline Params make_params(double hip_yaw) {
  Params P;
  P.hip_yaw = hip_yaw;
  P.body = Body{3.3, {0, 0, 0}, {0.011253, 0, 0, 0.036203, 0, 0.042673}};  // urdf:5-9
  const double sx[4] = {1, 1, -1, -1}, sy[4] = {1, -1, 1, -1};
  for (int l = 0; l < 4; ++l) {
    Leg& L = P.leg[l];
    L.abd_xyz[0] = sx[l] * 0.19; L.abd_xyz[1] = sy[l] * 0.049; L.abd_xyz[2] = 0;          // urdf:27
    L.hip_xyz[0] = 0; L.hip_xyz[1] = sy[l] * 0.062; L.hip_xyz[2] = 0;                      // urdf:79 (rpy 0 0 3.1415)
    L.knee_xyz[0] = 0; L.knee_xyz[1] = 0; L.knee_xyz[2] = -0.209;                          // urdf:110
    L.foot_xyz[0] = 0; L.foot_xyz[1] = 0; L.foot_xyz[2] = -0.195;                          // urdf:137
    L.abd = Body{0.54, {0, sy[l] * 0.036, 0}, {0.000381, sy[l] * 0.000058, 0.00000045, 0.000560, sy[l] * 0.00000095, 0.000444}};     // urdf:34-38
    L.thigh = Body{0.634, {0, sy[l] * 0.016, -0.02}, {0.001983, sy[l] * 0.000245, 0.000013, 0.002103, sy[l] * 0.0000015, 0.000408}};  // urdf:86-90
    L.shank = Body{0.064, {0, 0, -0.061}, {0.000245, 0, 0, 0.000248, 0, 0.000006}};                                                  // urdf:117-120
  }
  return P;
}

/* ---- kinematic tree: 18 one-dof joints. 0..2 prismatic x,y,z; 3 RZ; 4 RY; 5 RX (carries the trunk);
 * leg l: 6+3l abduction (RX), 7+3l hip (RY after a fixed Rz(hip_yaw)), 8+3l knee (RY). */
struct JointDesc { int parent; int type; /*0 prismatic 1 revolute*/ int axis; /*0 x 1 y 2 z*/ double p[3]; double yaw; const Body* body; };

inline void build_tree(const Params& P, JointDesc J[18]) {
  const int ax[6] = {0, 1, 2, 2, 1, 0};
  for (int i = 0; i < 6; ++i) J[i] = JointDesc{i - 1, i < 3 ? 0 : 1, ax[i], {0, 0, 0}, 0.0, i == 5 ? &P.body : nullptr};
  for (int l = 0; l < 4; ++l) {
    const Leg& L = P.leg[l];
    int b = 6 + 3 * l;
    J[b] = JointDesc{5, 1, 0, {L.abd_xyz[0], L.abd_xyz[1], L.abd_xyz[2]}, 0.0, &L.abd};
    J[b + 1] = JointDesc{b, 1, 1, {L.hip_xyz[0], L.hip_xyz[1], L.hip_xyz[2]}, P.hip_yaw, &L.thigh};
    J[b + 2] = JointDesc{b + 1, 1, 1, {L.knee_xyz[0], L.knee_xyz[1], L.knee_xyz[2]}, 0.0, &L.shank};
  }
}

template <class S> struct Kin { M3<S> R; V3<S> p, axw, w, v, al, a; };

template <class S>
inline M3<S> axis_rotation(int axis, const S& q) {
  M3<S> R;
  S c = cos(q), s = sin(q);
  if (axis == 0) { R.m[1][1] = c; R.m[1][2] = -s; R.m[2][1] = s; R.m[2][2] = c; }
  else if (axis == 1) { R.m[0][0] = c; R.m[0][2] = s; R.m[2][0] = -s; R.m[2][2] = c; }
  else { R.m[0][0] = c; R.m[0][1] = -s; R.m[1][0] = s; R.m[1][1] = c; }
  return R;
}

/* world-frame forward pass; linear acceleration is that of the joint-frame origin, gravity excluded */
template <class S>
inline void forward_kinematics(const JointDesc J[18], const S* q, const S* v, const S* a, Kin<S> K[18]) {
  for (int i = 0; i < 18; ++i) {
    const JointDesc& j = J[i];
    M3<S> Rp; V3<S> pp, wp, vp, alp, ap;
    if (j.parent >= 0) { const Kin<S>& P = K[j.parent]; Rp = P.R; pp = P.p; wp = P.w; vp = P.v; alp = P.al; ap = P.a; }
    M3<S> Ro = Rp;
    if (j.yaw != 0.0) Ro = Rp * axis_rotation<S>(2, S(j.yaw));
    V3<S> e(S(j.axis == 0 ? 1.0 : 0.0), S(j.axis == 1 ? 1.0 : 0.0), S(j.axis == 2 ? 1.0 : 0.0));
    V3<S> axw = Ro * e;
    V3<S> r = Rp * V3<S>(S(j.p[0]), S(j.p[1]), S(j.p[2]));
    S qd = v ? v[i] : S(0.0), qdd = a ? a[i] : S(0.0);
    Kin<S>& O = K[i];
    O.axw = axw;
    if (j.type == 1) {
      O.R = Ro * axis_rotation<S>(j.axis, q[i]);
      O.p = pp + r;
      O.w = wp + axw * qd;
      O.v = vp + cross(wp, r);
      O.al = alp + axw * qdd + cross(wp, axw * qd);
      O.a = ap + cross(alp, r) + cross(wp, cross(wp, r));
    } else {
      O.R = Ro;
      r = r + axw * q[i];
      O.p = pp + r;
      O.w = wp;
      O.v = vp + cross(wp, r) + axw * qd;
      O.al = alp;
      O.a = ap + cross(alp, r) + cross(wp, cross(wp, r)) + cross(wp, axw) * (qd * S(2.0)) + axw * qdd;
    }
  }
}

/* tau = M(q) a + C(q,v) v + g(q)   (gravity (0,0,-9.81) when with_gravity) */
template <class S>
inline void rnea(const JointDesc J[18], const S* q, const S* v, const S* a, bool with_gravity, S* tau) {
  Kin<S> K[18];
  forward_kinematics<S>(J, q, v, a, K);
  V3<S> F[18], N[18];
  for (int i = 0; i < 18; ++i) {
    const Body* b = J[i].body;
    if (!b) continue;
    const Kin<S>& k = K[i];
    V3<S> cw = k.R * V3<S>(S(b->com[0]), S(b->com[1]), S(b->com[2]));
    V3<S> ac = k.a + cross(k.al, cw) + cross(k.w, cross(k.w, cw));
    if (with_gravity) ac.z = ac.z + S(9.81);
    V3<S> f = ac * S(b->m);
    M3<S> Ib;
    Ib.m[0][0] = S(b->I[0]); Ib.m[0][1] = S(b->I[1]); Ib.m[0][2] = S(b->I[2]);
    Ib.m[1][0] = S(b->I[1]); Ib.m[1][1] = S(b->I[3]); Ib.m[1][2] = S(b->I[4]);
    Ib.m[2][0] = S(b->I[2]); Ib.m[2][1] = S(b->I[4]); Ib.m[2][2] = S(b->I[5]);
    M3<S> Iw = k.R * Ib * transposed(k.R);
    V3<S> n = Iw * k.al + cross(k.w, Iw * k.w);
    F[i] = f;
    N[i] = n + cross(cw, f);
  }
  for (int i = 17; i >= 0; --i) {
    const Kin<S>& k = K[i];
    tau[i] = J[i].type == 1 ? dot(k.axw, N[i]) : dot(k.axw, F[i]);
    int pa = J[i].parent;
    if (pa >= 0) { F[pa] = F[pa] + F[i]; N[pa] = N[pa] + N[i] + cross(k.p - K[pa].p, F[i]); }
  }
}

struct Feet { double p[4][3], v[4][3], acc[4][3], J[4][3][18]; };
/* foot position, velocity, d^2p/dt^2 for the given (v, a), and translational Jacobian (LOCAL_WORLD_ALIGNED) */
inline void feet_kinematics(const Params& P, const JointDesc J[18], const double* q, const double* v, const double* a, Feet& out) {
  Kin<double> K[18];
  forward_kinematics<double>(J, q, v, a, K);
  for (int f = 0; f < 4; ++f) {
    int jp = 8 + 3 * f;
    const Kin<double>& k = K[jp];
    V3<double> r = k.R * V3<double>(P.leg[f].foot_xyz[0], P.leg[f].foot_xyz[1], P.leg[f].foot_xyz[2]);
    V3<double> p = k.p + r, vel = k.v + cross(k.w, r), acc = k.a + cross(k.al, r) + cross(k.w, cross(k.w, r));
    out.p[f][0] = p.x; out.p[f][1] = p.y; out.p[f][2] = p.z;
    out.v[f][0] = vel.x; out.v[f][1] = vel.y; out.v[f][2] = vel.z;
    out.acc[f][0] = acc.x; out.acc[f][1] = acc.y; out.acc[f][2] = acc.z;
    for (int r_ = 0; r_ < 3; ++r_) for (int c = 0; c < 18; ++c) out.J[f][r_][c] = 0;
    for (int i = jp; i >= 0; i = J[i].parent) {
      V3<double> col = J[i].type == 1 ? cross(K[i].axw, p - K[i].p) : K[i].axw;
      out.J[f][0][i] = col.x; out.J[f][1][i] = col.y; out.J[f][2][i] = col.z;
    }
  }
}

}  // namespace wb
}  // namespace oracle
