// cafe_pinocchio_shim.hpp — TEST INFRASTRUCTURE ONLY (part of oracle/, never linked into the product).
//
// Stand-in for the part of the Pinocchio 2.6 API that the reference's whole-body model (MHPC/MHPC-Trajopt/WBM.cpp) and constraints call.
// Pinocchio (and urdfdom behind it) is not in this image. With this header on the include path as <pinocchio/...> the reference's
// WBM.cpp / MHPCProblem.cpp / MHPCCost.cpp / MHPCConstraint.cpp / MHPCReset.cpp / MHPCReference.cpp compile UNCHANGED (oracle/refbuild).
//
// What stands behind the names: the oracle's own rigid-body recursions for the Mini Cheetah (oracle/wb_dynamics.hpp: RNEA on doubles and on
// forward-mode dual numbers, unit-acceleration columns for the joint-space inertia, foot kinematics in LOCAL_WORLD_ALIGNED), wrapped in
// Pinocchio's documented semantics for each call (which quantity lands where in Data; the KKT formulas of contact-dynamics.hxx: forwardDynamics
// with inv_damping on J M^-1 J^T, impulseDynamics with restitution 0, the block inverse of [[M, J^T], [J, 0]]). So a run of the reference's
// whole-body code built this way pins the reference's OWN layers - contact / impact glue incl. its quirks, costs, constraints, reset maps,
// problem set-up, references, the MPC update - but NOT the rigid-body algorithms, which are common to this build and the oracle (those are
// pinned by test/testKKTDynamics.cpp's known answers and the reference's CasADi kinematic partials, tests/test_cpu_mhpc.py).
// Only T = double, the fixed 18-dof tree of PinocchioInteface.cpp:17-56 (PX PY PZ RZ RY RX + 4 legs) and the four foot frames (ids 11, 19, 27,
// 35, WBM.h:20) are supported.
#pragma once
#include <eigen3/Eigen/Dense>
#include <stdexcept>
#include <string>
#include <vector>
#include "wb_dynamics.hpp"

namespace pinocchio {

enum ReferenceFrame { WORLD = 0, LOCAL = 1, LOCAL_WORLD_ALIGNED = 2 };
enum FrameType { OP_FRAME = 1, JOINT = 2, FIXED_JOINT = 4, BODY = 8, SENSOR = 16 };
typedef std::size_t FrameIndex;
typedef std::size_t JointIndex;

template <class T> struct SE3Tpl {
  Eigen::Matrix<T, 3, 3> rot; Eigen::Matrix<T, 3, 1> trans;
  SE3Tpl() { rot.setIdentity(); trans.setZero(); }
  static SE3Tpl Identity() { return SE3Tpl(); }
  const Eigen::Matrix<T, 3, 1>& translation() const { return trans; }
  Eigen::Matrix<T, 3, 1>& translation() { return trans; }
  const Eigen::Matrix<T, 3, 3>& rotation() const { return rot; }
};
typedef SE3Tpl<double> SE3;
template <class T> struct MotionTpl {
  Eigen::Matrix<T, 3, 1> lin, ang;
  MotionTpl() { lin.setZero(); ang.setZero(); }
  const Eigen::Matrix<T, 3, 1>& linear() const { return lin; }
  const Eigen::Matrix<T, 3, 1>& angular() const { return ang; }
  Eigen::Matrix<T, 6, 1> toVector() const { Eigen::Matrix<T, 6, 1> v; v << lin, ang; return v; }
};
template <class T> using ForceTpl = MotionTpl<T>;
template <class T> struct FrameTpl {
  std::string name; JointIndex parent = 0; FrameIndex previousFrame = 0; SE3Tpl<T> placement; FrameType type = OP_FRAME;
  FrameTpl() {}
  FrameTpl(const std::string& n, JointIndex p, FrameIndex pf, const SE3Tpl<T>& pl, FrameType t) : name(n), parent(p), previousFrame(pf), placement(pl), type(t) {}
};

template <class T> struct ModelTpl {
  int nq = 18, nv = 18, njoints = 19, nbodies = 19, nframes = 40;
  std::string name = "mini_cheetah";
  std::vector<std::string> names;
  std::vector<FrameTpl<T>> frames;
  oracle::wb::Params P;
  oracle::wb::JointDesc J[18];
  ModelTpl() { set_hip_yaw(3.1415); }
  ModelTpl(const ModelTpl& o) : nq(o.nq), nv(o.nv), njoints(o.njoints), nbodies(o.nbodies), nframes(o.nframes), name(o.name), names(o.names), frames(o.frames) { set_hip_yaw(o.P.hip_yaw); }
  ModelTpl& operator=(const ModelTpl& o) { nq = o.nq; nv = o.nv; njoints = o.njoints; nbodies = o.nbodies; nframes = o.nframes; name = o.name; names = o.names; frames = o.frames; set_hip_yaw(o.P.hip_yaw); return *this; }
  // the tree holds pointers into P: rebuilt whenever the model is copied
  void set_hip_yaw(double yaw) {
    P = oracle::wb::make_params(yaw);
    oracle::wb::build_tree(P, J);
    static const char* jn[19] = {"universe", "PX", "PY", "PZ", "RZ", "RY", "RX", "abduct_fl", "thigh_fl", "shank_fl", "abduct_fr", "thigh_fr", "shank_fr",
                                 "abduct_hl", "thigh_hl", "shank_hl", "abduct_hr", "thigh_hr", "shank_hr"};
    names.assign(jn, jn + 19);
    frames.assign(nframes, FrameTpl<T>());
    for (int i = 0; i < nframes; ++i) frames[i].name = "frame_" + std::to_string(i);
    const char* fn[4] = {"toe_fl", "toe_fr", "toe_hl", "toe_hr"};
    for (int f = 0; f < 4; ++f) frames[11 + 8 * f].name = fn[f];
  }
  FrameIndex getFrameId(const std::string& n) const { for (std::size_t i = 0; i < frames.size(); ++i) if (frames[i].name == n) return i; return frames.size(); }
  bool existFrame(const std::string& n) const { return getFrameId(n) < frames.size(); }
};
typedef ModelTpl<double> Model;

inline int foot_of_frame(std::size_t id) {
  if (id < 11 || (id - 11) % 8 != 0 || (id - 11) / 8 > 3) throw std::invalid_argument("pinocchio stand-in: only the foot frames 11, 19, 27, 35 are known");
  return (int)((id - 11) / 8);
}

template <class T> struct DataTpl {
  typedef Eigen::Matrix<T, Eigen::Dynamic, Eigen::Dynamic> MatrixXs;
  typedef Eigen::Matrix<T, Eigen::Dynamic, 1> VectorXs;
  MatrixXs M, dtau_dq, dtau_dv, Minv, JMinvJt_inv, Jc;
  VectorXs nle, ddq, lambda_c, impulse_c, dq_after, tau, torque_residual;
  std::vector<SE3Tpl<T>> oMf;
  ForceTpl<T> hg, dhg;
  // state of the last kinematics pass
  double q[18], v[18], a[18];
  oracle::wb::Kin<double> K[18];
  oracle::wb::Feet feet;
  DataTpl() {}
  explicit DataTpl(const ModelTpl<T>& m) : oMf(m.nframes) {
    M.setZero(18, 18); dtau_dq.setZero(18, 18); dtau_dv.setZero(18, 18); Minv.setZero(18, 18);
    nle.setZero(18); ddq.setZero(18); dq_after.setZero(18); tau.setZero(18); torque_residual.setZero(18);
    for (int i = 0; i < 18; ++i) { q[i] = 0; v[i] = 0; a[i] = 0; }
  }
};
typedef DataTpl<double> Data;

namespace detail {
template <class T, class V> inline void load(const Eigen::DenseBase<V>& x, double* out) { for (int i = 0; i < 18; ++i) out[i] = (double)x[i]; }
template <class T> inline void kin_pass(const ModelTpl<T>& m, DataTpl<T>& d) {
  oracle::wb::forward_kinematics<double>(m.J, d.q, d.v, d.a, d.K);
  oracle::wb::feet_kinematics(m.P, m.J, d.q, d.v, d.a, d.feet);
}
// dense symmetric positive-definite solve (Cholesky), X = A^-1 B
template <class T> inline Eigen::Matrix<T, Eigen::Dynamic, Eigen::Dynamic> spd_solve(const Eigen::Matrix<T, Eigen::Dynamic, Eigen::Dynamic>& A,
                                                                                     const Eigen::Matrix<T, Eigen::Dynamic, Eigen::Dynamic>& B) {
  const Eigen::Index n = A.rows();
  Eigen::Matrix<T, Eigen::Dynamic, Eigen::Dynamic> L(n, n), X(B);
  L.setZero();
  for (Eigen::Index j = 0; j < n; ++j) {
    T s = A(j, j);
    for (Eigen::Index k = 0; k < j; ++k) s -= L(j, k) * L(j, k);
    if (!(s > T(0))) throw std::runtime_error("pinocchio stand-in: matrix not positive definite");
    L(j, j) = std::sqrt(s);
    for (Eigen::Index i = j + 1; i < n; ++i) { T t = A(i, j); for (Eigen::Index k = 0; k < j; ++k) t -= L(i, k) * L(j, k); L(i, j) = t / L(j, j); }
  }
  for (Eigen::Index c = 0; c < X.cols(); ++c) {
    for (Eigen::Index i = 0; i < n; ++i) { T s = X(i, c); for (Eigen::Index k = 0; k < i; ++k) s -= L(i, k) * X(k, c); X(i, c) = s / L(i, i); }
    for (Eigen::Index i = n - 1; i >= 0; --i) { T s = X(i, c); for (Eigen::Index k = i + 1; k < n; ++k) s -= L(k, i) * X(k, c); X(i, c) = s / L(i, i); }
  }
  return X;
}
template <class T> inline void mass_matrix(const ModelTpl<T>& m, DataTpl<T>& d, const double* q) {
  double zero[18] = {0}, e[18], col[18];
  for (int c = 0; c < 18; ++c) {
    for (int i = 0; i < 18; ++i) e[i] = (i == c) ? 1.0 : 0.0;
    oracle::wb::rnea<double>(m.J, q, zero, e, false, col);
    for (int r = 0; r < 18; ++r) d.M(r, c) = col[r];
  }
  for (int r = 0; r < 18; ++r) for (int c = r + 1; c < 18; ++c) { const double s = 0.5 * (d.M(r, c) + d.M(c, r)); d.M(r, c) = s; d.M(c, r) = s; }
  Eigen::Matrix<T, Eigen::Dynamic, Eigen::Dynamic> I(18, 18); I.setIdentity();
  d.Minv = spd_solve<T>(d.M, I);
}
template <class T, class JM> inline void contact_blocks(DataTpl<T>& d, const Eigen::DenseBase<JM>& J, T inv_damping) {
  d.Jc = J;
  const Eigen::Index nc = d.Jc.rows();
  if (nc == 0) { d.JMinvJt_inv.setZero(0, 0); return; }
  Eigen::Matrix<T, Eigen::Dynamic, Eigen::Dynamic> S = d.Jc * d.Minv * d.Jc.transpose(), I(nc, nc);
  for (Eigen::Index i = 0; i < nc; ++i) S(i, i) += inv_damping;
  I.setIdentity();
  d.JMinvJt_inv = spd_solve<T>(S, I);
}
}  // namespace detail

// ---- kinematics.hpp / frames.hpp / jacobian.hpp
template <class T, class Q> void forwardKinematics(const ModelTpl<T>& m, DataTpl<T>& d, const Eigen::DenseBase<Q>& q) {
  detail::load<T>(q, d.q); for (int i = 0; i < 18; ++i) { d.v[i] = 0; d.a[i] = 0; }
  detail::kin_pass(m, d);
}
template <class T, class Q, class V> void forwardKinematics(const ModelTpl<T>& m, DataTpl<T>& d, const Eigen::DenseBase<Q>& q, const Eigen::DenseBase<V>& v) {
  detail::load<T>(q, d.q); detail::load<T>(v, d.v); for (int i = 0; i < 18; ++i) d.a[i] = 0;
  detail::kin_pass(m, d);
}
template <class T, class Q, class V, class A> void forwardKinematics(const ModelTpl<T>& m, DataTpl<T>& d, const Eigen::DenseBase<Q>& q, const Eigen::DenseBase<V>& v, const Eigen::DenseBase<A>& a) {
  detail::load<T>(q, d.q); detail::load<T>(v, d.v); detail::load<T>(a, d.a);
  detail::kin_pass(m, d);
}
template <class T> const SE3Tpl<T>& updateFramePlacement(const ModelTpl<T>&, DataTpl<T>& d, std::size_t id) {
  const int f = foot_of_frame(id);
  for (int r = 0; r < 3; ++r) d.oMf[id].trans[r] = d.feet.p[f][r];
  const oracle::wb::Kin<double>& k = d.K[8 + 3 * f];
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) d.oMf[id].rot(r, c) = k.R.m[r][c];
  return d.oMf[id];
}
template <class T> void updateFramePlacements(const ModelTpl<T>& m, DataTpl<T>& d) { for (int f = 0; f < 4; ++f) updateFramePlacement(m, d, 11 + 8 * f); }
template <class T> void framesForwardKinematics(const ModelTpl<T>& m, DataTpl<T>& d) { updateFramePlacements(m, d); }
template <class T, class Q> void framesForwardKinematics(const ModelTpl<T>& m, DataTpl<T>& d, const Eigen::DenseBase<Q>& q) { forwardKinematics(m, d, q); updateFramePlacements(m, d); }
// LOCAL_WORLD_ALIGNED: velocity of the frame origin and angular velocity of its link, world axes
template <class T> MotionTpl<T> getFrameVelocity(const ModelTpl<T>&, const DataTpl<T>& d, std::size_t id, ReferenceFrame rf = LOCAL) {
  if (rf != LOCAL_WORLD_ALIGNED) throw std::invalid_argument("pinocchio stand-in: LOCAL_WORLD_ALIGNED only");
  const int f = foot_of_frame(id);
  const oracle::wb::Kin<double>& k = d.K[8 + 3 * f];
  MotionTpl<T> mt;
  for (int r = 0; r < 3; ++r) mt.lin[r] = d.feet.v[f][r];
  mt.ang[0] = k.w.x; mt.ang[1] = k.w.y; mt.ang[2] = k.w.z;
  return mt;
}
// spatial acceleration: its linear part is the classical acceleration of the frame origin minus w x v
template <class T> MotionTpl<T> getFrameAcceleration(const ModelTpl<T>& m, const DataTpl<T>& d, std::size_t id, ReferenceFrame rf = LOCAL) {
  const MotionTpl<T> vel = getFrameVelocity(m, d, id, rf);
  const int f = foot_of_frame(id);
  const oracle::wb::Kin<double>& k = d.K[8 + 3 * f];
  MotionTpl<T> mt;
  const Eigen::Matrix<T, 3, 1> wxv = vel.ang.cross(vel.lin);
  for (int r = 0; r < 3; ++r) mt.lin[r] = d.feet.acc[f][r] - wxv[r];
  mt.ang[0] = k.al.x; mt.ang[1] = k.al.y; mt.ang[2] = k.al.z;
  return mt;
}
template <class T, class Q> void computeJointJacobians(const ModelTpl<T>& m, DataTpl<T>& d, const Eigen::DenseBase<Q>& q) { forwardKinematics(m, d, q); }
template <class T> void computeJointJacobians(const ModelTpl<T>&, DataTpl<T>&) {}   // the last kinematics pass holds what the Jacobians need
template <class T, class JM> void getFrameJacobian(const ModelTpl<T>& m, const DataTpl<T>& d, std::size_t id, ReferenceFrame rf, const Eigen::DenseBase<JM>& J_) {
  if (rf != LOCAL_WORLD_ALIGNED) throw std::invalid_argument("pinocchio stand-in: LOCAL_WORLD_ALIGNED only");
  Eigen::DenseBase<JM>& J = const_cast<Eigen::DenseBase<JM>&>(J_);
  const int f = foot_of_frame(id);
  for (int r = 0; r < 6; ++r) for (int c = 0; c < 18; ++c) J(r, c) = 0;
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 18; ++c) J(r, c) = d.feet.J[f][r][c];
  for (int i = 8 + 3 * f; i >= 0; i = m.J[i].parent)
    if (m.J[i].type == 1) { J(3, i) = d.K[i].axw.x; J(4, i) = d.K[i].axw.y; J(5, i) = d.K[i].axw.z; }
}

// ---- crba.hpp / rnea.hpp / aba.hpp
template <class T, class Q> const typename DataTpl<T>::MatrixXs& crba(const ModelTpl<T>& m, DataTpl<T>& d, const Eigen::DenseBase<Q>& q) {
  double qq[18]; detail::load<T>(q, qq);
  detail::mass_matrix(m, d, qq);
  return d.M;
}
template <class T, class Q, class V> const typename DataTpl<T>::VectorXs& nonLinearEffects(const ModelTpl<T>& m, DataTpl<T>& d, const Eigen::DenseBase<Q>& q, const Eigen::DenseBase<V>& v) {
  double qq[18], vv[18], zero[18] = {0}, out[18];
  detail::load<T>(q, qq); detail::load<T>(v, vv);
  oracle::wb::rnea<double>(m.J, qq, vv, zero, true, out);
  for (int i = 0; i < 18; ++i) d.nle[i] = out[i];
  return d.nle;
}
template <class T, class Q, class V, class A> const typename DataTpl<T>::VectorXs& rnea(const ModelTpl<T>& m, DataTpl<T>& d, const Eigen::DenseBase<Q>& q, const Eigen::DenseBase<V>& v, const Eigen::DenseBase<A>& a) {
  double qq[18], vv[18], aa[18], out[18];
  detail::load<T>(q, qq); detail::load<T>(v, vv); detail::load<T>(a, aa);
  oracle::wb::rnea<double>(m.J, qq, vv, aa, true, out);
  for (int i = 0; i < 18; ++i) d.tau[i] = out[i];
  return d.tau;
}
template <class T, class Q, class V, class U> const typename DataTpl<T>::VectorXs& aba(const ModelTpl<T>& m, DataTpl<T>& d, const Eigen::DenseBase<Q>& q, const Eigen::DenseBase<V>& v, const Eigen::DenseBase<U>& tau) {
  crba(m, d, q); nonLinearEffects(m, d, q, v);
  typename DataTpl<T>::VectorXs b = tau - d.nle;
  d.ddq = d.Minv * b;
  return d.ddq;
}

// ---- contact-dynamics.hpp (contact-dynamics.hxx of Pinocchio 2.6)
// forwardDynamics: M and nle in Data are up to date (crba + nonLinearEffects called before, WBM.cpp:409-411)
template <class T, class U, class JM, class G>
const typename DataTpl<T>::VectorXs& forwardDynamics(const ModelTpl<T>&, DataTpl<T>& d, const Eigen::DenseBase<U>& tau, const Eigen::DenseBase<JM>& J,
                                                     const Eigen::DenseBase<G>& gamma, const T inv_damping = 0.) {
  detail::contact_blocks(d, J, inv_damping);
  typename DataTpl<T>::VectorXs b = tau - d.nle;
  d.torque_residual = d.Minv * b;                       // M^-1 (tau - nle)
  typename DataTpl<T>::VectorXs rhs = -(d.Jc * d.torque_residual);
  rhs -= gamma;
  d.lambda_c = d.JMinvJt_inv * rhs;
  typename DataTpl<T>::VectorXs jt = d.Jc.transpose() * d.lambda_c;
  d.ddq = d.Minv * jt;
  d.ddq += d.torque_residual;
  return d.ddq;
}
template <class T, class Q, class V, class JM>
const typename DataTpl<T>::VectorXs& impulseDynamics(const ModelTpl<T>& m, DataTpl<T>& d, const Eigen::DenseBase<Q>& q, const Eigen::DenseBase<V>& v_before,
                                                     const Eigen::DenseBase<JM>& J, const T r_coeff = 0., const T inv_damping = 0.) {
  crba(m, d, q);
  detail::contact_blocks(d, J, inv_damping);
  typename DataTpl<T>::VectorXs vb(v_before);
  typename DataTpl<T>::VectorXs rhs = -(1. + r_coeff) * (d.Jc * vb);
  d.impulse_c = d.JMinvJt_inv * rhs;
  typename DataTpl<T>::VectorXs jt = d.Jc.transpose() * d.impulse_c;
  d.dq_after = d.Minv * jt;
  d.dq_after += vb;
  return d.dq_after;
}
// [[M, J^T], [J, 0]]^-1 by blocks, from the decomposition of the last forwardDynamics / impulseDynamics call
template <class T, class JM, class KM>
void getKKTContactDynamicMatrixInverse(const ModelTpl<T>&, const DataTpl<T>& d, const Eigen::DenseBase<JM>& J, const Eigen::DenseBase<KM>& K_) {
  Eigen::DenseBase<KM>& K = const_cast<Eigen::DenseBase<KM>&>(K_);
  const Eigen::Index nc = J.rows();
  typename DataTpl<T>::MatrixXs Jc(J), MinvJt = d.Minv * Jc.transpose();
  typename DataTpl<T>::MatrixXs tr = MinvJt * d.JMinvJt_inv;         // 18 x nc
  typename DataTpl<T>::MatrixXs tl = d.Minv - tr * MinvJt.transpose();
  K.block(0, 0, 18, 18) = tl;
  if (nc > 0) {
    K.block(0, 18, 18, nc) = tr;
    K.block(18, 0, nc, 18) = tr.transpose();
    K.block(18, 18, nc, nc) = -d.JMinvJt_inv;
  }
}
template <class T, class Q, class JM, class KM>
void computeKKTContactDynamicMatrixInverse(const ModelTpl<T>& m, DataTpl<T>& d, const Eigen::DenseBase<Q>& q, const Eigen::DenseBase<JM>& J,
                                           const Eigen::DenseBase<KM>& K, const T inv_damping = 0.) {
  crba(m, d, q);
  detail::contact_blocks(d, J, inv_damping);
  getKKTContactDynamicMatrixInverse(m, d, J, K);
}

// ---- rnea-derivatives.hpp: forward-mode dual numbers through the same RNEA recursion
template <class T, class Q, class V, class A>
void computeRNEADerivatives(const ModelTpl<T>& m, DataTpl<T>& d, const Eigen::DenseBase<Q>& q, const Eigen::DenseBase<V>& v, const Eigen::DenseBase<A>& a) {
  typedef oracle::wb::Dual<36> D;
  D qd[18], vd[18], ad[18], tau[18];
  for (int i = 0; i < 18; ++i) { qd[i] = D((double)q[i]); qd[i].d[i] = 1.0; vd[i] = D((double)v[i]); vd[i].d[18 + i] = 1.0; ad[i] = D((double)a[i]); }
  oracle::wb::rnea<D>(m.J, qd, vd, ad, true, tau);
  for (int r = 0; r < 18; ++r) for (int c = 0; c < 18; ++c) { d.dtau_dq(r, c) = tau[r].d[c]; d.dtau_dv(r, c) = tau[r].d[18 + c]; }
  double qq[18]; detail::load<T>(q, qq);
  detail::mass_matrix(m, d, qq);     // dtau_da = M (Pinocchio fills its upper triangle; here the whole matrix)
}
template <class T, class Q, class GM>
void computeGeneralizedGravityDerivatives(const ModelTpl<T>& m, DataTpl<T>&, const Eigen::DenseBase<Q>& q, const Eigen::DenseBase<GM>& G_) {
  Eigen::DenseBase<GM>& G = const_cast<Eigen::DenseBase<GM>&>(G_);
  typedef oracle::wb::Dual<18> D;
  D qd[18], zd[18], tau[18];
  for (int i = 0; i < 18; ++i) { qd[i] = D((double)q[i]); qd[i].d[i] = 1.0; zd[i] = D(0.0); }
  oracle::wb::rnea<D>(m.J, qd, zd, zd, true, tau);
  for (int r = 0; r < 18; ++r) for (int c = 0; c < 18; ++c) G(r, c) = tau[r].d[c];
}

// ---- centroidal.hpp: only used by the reference's diagnostic getters (WBM.cpp:140-165) when its stand-alone programs publish a trajectory
//      for the visualiser, never on the solve path: NOT provided, hg / dhg stay zero
template <class T, class Q, class V> const ForceTpl<T>& computeCentroidalMomentum(const ModelTpl<T>&, DataTpl<T>& d, const Eigen::DenseBase<Q>&, const Eigen::DenseBase<V>&) { return d.hg; }
template <class T, class Q, class V, class A> const ForceTpl<T>& computeCentroidalMomentumTimeVariation(const ModelTpl<T>&, DataTpl<T>& d, const Eigen::DenseBase<Q>&, const Eigen::DenseBase<V>&, const Eigen::DenseBase<A>&) { return d.dhg; }

}  // namespace pinocchio
