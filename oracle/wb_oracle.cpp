/*
 * wb_oracle.cpp — CPU ORACLE (test infrastructure only): whole-body (WB) phase of the MHPC problem —
 * KKT contact dynamics and their derivatives, impact map, WB costs and constraints, WB->WB / WB->SRB reset.
 *
 *   WBM::Model::dynamics / dynamics_partial          MHPC/MHPC-Trajopt/WBM.cpp:17-139
 *   KKTContactDynamics / KKTContactDynamicsDerivatives   WBM.cpp:368-424, :459-505
 *   KKTImpact / KKTImpactDerivatives / impact(_partial)  WBM.cpp:178-254, :427-456, :508-543
 *   MHPCReset::reset_map(_partial)                   MHPC/MHPC-Trajopt/MHPCReset.cpp:4-53, MHPCReset.h:24-26
 *   WBTrackingCost, WBFootPlaceReg, SwingFootPosTracking, SwingFootVelTracking, TDVelocityPenalty
 *                                                    MHPC/MHPC-Trajopt/MHPCCost.h:8-205, MHPCCost.cpp:4-291
 *   TorqueLimit, JointLimit, WBMinimumHeight, WBGRF, WBTouchDown   MHPC/MHPC-Trajopt/MHPCConstraint.cpp:9-288
 *   problem wiring                                   MHPC/MHPC-Trajopt/MHPCProblem.cpp:403-601
 *
 * Pinocchio (absent here) is replaced by wb_dynamics.hpp; its contact-dynamics routines are restated from the
 * published 2.6.x algorithm (algorithm/contact-dynamics.hxx): forwardDynamics = Cholesky(M), JMinvJt + mu I, LLT;
 * lambda = -(JMinvJt)^-1 (J Minv (tau - nle) + gamma); qdd = Minv (tau - nle + J^T lambda);
 * impulseDynamics (r = 0): Lambda = -(JMinvJt)^-1 J v, v+ = v + Minv J^T Lambda;
 * computeKKTContactDynamicMatrixInverse = inverse of [[M, J^T],[J, 0]].
 * The reference's CasADi kinematic partials are called unchanged from oracle/_ref.
 */
#include "hsddp_oracle.hpp"
#include "casadi_ref.hpp"
#include "wb_dynamics.hpp"
#include <stdexcept>
#include <string>

namespace oracle {

/* dense helpers local to the WB model */
static Mat chol_lower(const Mat& A) {
  int n = A.r;
  Mat L(n, n);
  for (int j = 0; j < n; ++j) {
    double d = A(j, j);
    for (int k = 0; k < j; ++k) d -= L(j, k) * L(j, k);
    if (!(d > 0)) throw std::runtime_error("WB oracle: matrix not positive definite");
    L(j, j) = std::sqrt(d);
    for (int i = j + 1; i < n; ++i) {
      double s = A(i, j);
      for (int k = 0; k < j; ++k) s -= L(i, k) * L(j, k);
      L(i, j) = s / L(j, j);
    }
  }
  return L;
}
static Mat chol_solve(const Mat& L, const Mat& B) {  // (L L^T)^-1 B
  int n = L.r;
  Mat X = B;
  for (int c = 0; c < B.c; ++c) {
    for (int i = 0; i < n; ++i) { double s = X(i, c); for (int k = 0; k < i; ++k) s -= L(i, k) * X(k, c); X(i, c) = s / L(i, i); }
    for (int i = n - 1; i >= 0; --i) { double s = X(i, c); for (int k = i + 1; k < n; ++k) s -= L(k, i) * X(k, c); X(i, c) = s / L(i, i); }
  }
  return X;
}
static Mat col(const Vec& v) { Mat m((int)v.size(), 1); m.a = v; return m; }

class WBPhase : public Phase {
 public:
  double BG_alpha;
  wb::Params P;
  wb::JointDesc J[18];
  /* scratch of the last KKT evaluation */
  Mat M, Jc, Lm;  // M (18x18), active-contact Jacobian (3nc x 18), chol(M)
  Vec nle, qdd, GRF, gamma, lam;
  std::vector<int> feet_active;

  WBPhase(double bg, double hip_yaw) : BG_alpha(bg), P(wb::make_params(hip_yaw)) { wb::build_tree(P, J); }

  void build_model() override {
    /* MHPCProblem.cpp:436-481: Torque, Joint, MinHeight, then GRF when any foot is in contact */
    PathConstraint tq; tq.kind = 0; tq.create(24, h, n, m, p, ph->reb_torque); pcon.push_back(tq);
    /* BarrelRollTO.cpp:190-198: the joint-speed barrier sits between the torque and the joint-limit barrier */
    if (ph->joint_speed_limit) { PathConstraint jv; jv.kind = 4; jv.create(24, h, n, m, p, ph->reb_jointvel); pcon.push_back(jv); }
    /* LocoProblem.cpp:64-82 keeps Torque and GRF only */
    if (!ph->no_joint_limit) { PathConstraint jl; jl.kind = 1; jl.create(24, h, n, m, p, ph->reb_joint); pcon.push_back(jl); }
    if (!ph->no_min_height) { PathConstraint mh; mh.kind = 2; mh.create(1, h, n, m, p, ph->reb_minheight); pcon.push_back(mh); }
    int nc = 0;
    for (int l = 0; l < 4; ++l) nc += ph->contact[l] > 0;
    if (nc > 0) { PathConstraint g; g.kind = 3; g.create(5 * nc, h, n, m, p, ph->reb_grf); pcon.push_back(g); }
    if (ph->n_td > 0) { TermConstraint tc; tc.create(ph->n_td, n, ph->al_td); tcon.push_back(tc); }
  }

  void mass_and_bias(const double* q, const double* v) {
    M = Mat(18, 18);
    double zero[18] = {0}, e[18], colv[18];
    for (int c = 0; c < 18; ++c) {
      for (int i = 0; i < 18; ++i) e[i] = (i == c) ? 1.0 : 0.0;
      wb::rnea<double>(J, q, zero, e, false, colv);
      for (int r = 0; r < 18; ++r) M(r, c) = colv[r];
    }
    for (int r = 0; r < 18; ++r) for (int c = r + 1; c < 18; ++c) { double s = 0.5 * (M(r, c) + M(c, r)); M(r, c) = s; M(c, r) = s; }
    nle.assign(18, 0.0);
    wb::rnea<double>(J, q, v, zero, true, nle.data());
    Lm = chol_lower(M);
  }

  /* KKTContactDynamics (WBM.cpp:368-424) */
  void kkt_dynamics(const double* q, const double* v, const double* tau, const int* contact) {
    feet_active.clear();
    for (int l = 0; l < 4; ++l) if (contact[l] > 0) feet_active.push_back(l);
    int nc = (int)feet_active.size();
    mass_and_bias(q, v);
    GRF.assign(12, 0.0);
    Vec b(18);
    for (int i = 0; i < 18; ++i) b[i] = tau[i] - nle[i];
    Jc = Mat(3 * nc, 18);
    if (nc == 0) { qdd = chol_solve(Lm, col(b)).a; return; }  // pinocchio::aba == M^-1 (tau - nle)
    wb::Feet F;
    wb::feet_kinematics(P, J, q, v, nullptr, F);
    gamma.assign(3 * nc, 0.0);
    for (int i = 0; i < nc; ++i) {
      int f = feet_active[i];
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 18; ++c) Jc(3 * i + r, c) = F.J[f][r][c];
        /* classical acceleration with qdd = 0 (spatial linear part + w x v) + Baumgarte term (WBM.cpp:392-408) */
        gamma[3 * i + r] = F.acc[f][r] + 2 * BG_alpha * F.v[f][r];
      }
    }
    Mat Minvb = chol_solve(Lm, col(b));
    Mat MinvJt = chol_solve(Lm, transpose(Jc));
    Mat S = mm(Jc, MinvJt);
    for (int i = 0; i < 3 * nc; ++i) S(i, i) += 1e-12;  // inv_damping (WBM.cpp:411)
    Vec rhs = mv(Jc, Minvb.a);
    for (int i = 0; i < 3 * nc; ++i) rhs[i] = -rhs[i] - gamma[i];
    lam = chol_solve(chol_lower(S), col(rhs)).a;
    Vec t = mtv(Jc, lam);
    for (int i = 0; i < 18; ++i) t[i] += b[i];
    qdd = chol_solve(Lm, col(t)).a;
    for (int i = 0; i < nc; ++i) for (int r = 0; r < 3; ++r) GRF[3 * feet_active[i] + r] = lam[3 * i + r];
  }

  /* inverse of [[M, J^T],[J, 0]] by blocks (damping 0) */
  void kkt_inverse(Mat& Ktl, Mat& Ktr, Mat& Kbl, Mat& Kbr) const {
    int nc3 = Jc.r;
    Mat I18(18, 18); I18.identity();
    Mat Minv = chol_solve(Lm, I18);
    if (nc3 == 0) { Ktl = Minv; Ktr = Mat(18, 0); Kbl = Mat(0, 18); Kbr = Mat(0, 0); return; }
    Mat MinvJt = mm(Minv, transpose(Jc));
    Mat S = mm(Jc, MinvJt);
    Mat Is(nc3, nc3); Is.identity();
    Mat Sinv = chol_solve(chol_lower(S), Is);
    Ktr = mm(MinvJt, Sinv);          // 18 x 3nc
    Kbl = transpose(Ktr);            // 3nc x 18
    Kbr = Sinv; for (auto& x : Kbr.a) x = -x;
    Ktl = Minv; madd(Ktl, -1.0, mm(Ktr, transpose(MinvJt)));
  }

  void rnea_derivatives(const double* q, const double* v, const double* a, bool gravity, Mat& dq, Mat& dv) {
    typedef wb::Dual<36> D;
    D qd[18], vd[18], ad[18], tau[18];
    for (int i = 0; i < 18; ++i) { qd[i] = D(q[i]); qd[i].d[i] = 1.0; vd[i] = D(v[i]); vd[i].d[18 + i] = 1.0; ad[i] = D(a[i]); }
    wb::rnea<D>(J, qd, vd, ad, gravity, tau);
    dq = Mat(18, 18); dv = Mat(18, 18);
    for (int r = 0; r < 18; ++r) for (int c = 0; c < 18; ++c) { dq(r, c) = tau[r].d[c]; dv(r, c) = tau[r].d[18 + c]; }
  }

  /* reference CasADi kinematic partials (generated with hip yaw = pi), mapped to the active contacts */
  void casadi_foot(const char* which, const double* q, const double* v, const double* a, Mat out[4]) {
    for (int f = 0; f < 4; ++f) out[f] = Mat(3, 18);
    double* res[4] = {out[0].a.data(), out[1].a.data(), out[2].a.data(), out[3].a.data()};
    if (std::string(which) == "vel") { const double* arg[2] = {q, v}; casadi_call(CASADI_FN(footVelPartialDq), arg, 2, res, 4); }
    else if (std::string(which) == "accq") { const double* arg[3] = {q, v, a}; casadi_call(CASADI_FN(footAccPartialDq), arg, 3, res, 4); }
    else { const double* arg[3] = {q, v, a}; casadi_call(CASADI_FN(footAccPartialDv), arg, 3, res, 4); }
  }
  Mat stack_active(const Mat src[4]) const {
    int nc = (int)feet_active.size();
    Mat o(3 * nc, 18);
    for (int i = 0; i < nc; ++i) for (int r = 0; r < 3; ++r) for (int c = 0; c < 18; ++c) o(3 * i + r, c) = src[feet_active[i]](r, c);
    return o;
  }
  Mat force_partial(const double* q, const double* F12) {  // computeContactForceDerivatives (WBM.cpp:648-675)
    Mat per[4];
    for (int f = 0; f < 4; ++f) per[f] = Mat(18, 18);
    double* res[4] = {per[0].a.data(), per[1].a.data(), per[2].a.data(), per[3].a.data()};
    const double* arg[2] = {q, F12};
    casadi_call(CASADI_FN(footForcePartialDq), arg, 2, res, 4);
    Mat sum(18, 18);
    for (int f : feet_active) madd(sum, 1.0, per[f]);
    return sum;
  }

  void dynamics(Vec& xnext, Vec& y, const Vec& x, const Vec& u, int) override {
    double tau[18] = {0};
    for (int i = 0; i < 12; ++i) tau[6 + i] = u[i];  // SelectionMat (WBM.h:46-47)
    kkt_dynamics(x.data(), x.data() + 18, tau, ph->contact);
    xnext.assign(36, 0.0);
    for (int i = 0; i < 18; ++i) { xnext[i] = x[i] + x[18 + i] * dt; xnext[18 + i] = x[18 + i] + qdd[i] * dt; }  // WBM.cpp:25-26
    y = GRF;
  }

  void dynamics_partial(Mat& A_, Mat& B_, Mat& C_, Mat& D_, const Vec& x, const Vec& u, int) override {
    const double* q = x.data();
    const double* v = x.data() + 18;
    double tau[18] = {0};
    for (int i = 0; i < 12; ++i) tau[6 + i] = u[i];
    kkt_dynamics(q, v, tau, ph->contact);
    int nc = (int)feet_active.size();
    Mat Ktl, Ktr, Kbl, Kbr;
    kkt_inverse(Ktl, Ktr, Kbl, Kbr);
    Mat dtau_dq, dtau_dv;
    rnea_derivatives(q, v, qdd.data(), true, dtau_dq, dtau_dv);
    madd(dtau_dq, -1.0, force_partial(q, GRF.data()));
    Mat dqdd_dq = mm(Ktl, dtau_dq), dqdd_dv = mm(Ktl, dtau_dv);
    for (auto& e : dqdd_dq.a) e = -e;
    for (auto& e : dqdd_dv.a) e = -e;
    Mat dGRF_dq(3 * nc, 18), dGRF_dv(3 * nc, 18);
    if (nc > 0) {
      Mat aq[4], av[4], vq[4];
      casadi_foot("accq", q, v, qdd.data(), aq);
      casadi_foot("accv", q, v, qdd.data(), av);
      casadi_foot("vel", q, v, nullptr, vq);
      Mat da_dq = stack_active(aq), da_dv = stack_active(av), dv_dq = stack_active(vq);
      madd(da_dq, 2 * BG_alpha, dv_dq);  // WBM.cpp:486-488
      madd(da_dv, 2 * BG_alpha, Jc);
      madd(dqdd_dq, -1.0, mm(Ktr, da_dq));
      madd(dqdd_dv, -1.0, mm(Ktr, da_dv));
      dGRF_dq = mm(Kbl, dtau_dq); madd(dGRF_dq, 1.0, mm(Kbr, da_dq));
      dGRF_dv = mm(Kbl, dtau_dv); madd(dGRF_dv, 1.0, mm(Kbr, da_dv));
    }
    /* A = I + dt Ac, B = dt Bc (WBM.cpp:68-69, :122-138) */
    A_.zero(); B_.zero(); C_.zero(); D_.zero();
    for (int i = 0; i < 36; ++i) A_(i, i) = 1.0;
    for (int i = 0; i < 18; ++i) {
      A_(i, 18 + i) += dt;
      for (int c = 0; c < 18; ++c) { A_(18 + i, c) += dqdd_dq(i, c) * dt; A_(18 + i, 18 + c) += dqdd_dv(i, c) * dt; }
      for (int c = 0; c < 12; ++c) B_(18 + i, c) = Ktl(i, 6 + c) * dt;
    }
    for (int i = 0; i < nc; ++i) {
      int f = feet_active[i];
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 18; ++c) { C_(3 * f + r, c) = dGRF_dq(3 * i + r, c); C_(3 * f + r, 18 + c) = dGRF_dv(3 * i + r, c); }
        for (int c = 0; c < 12; ++c) D_(3 * f + r, c) = -Kbl(3 * i + r, 6 + c);  // dGRF_dtau = -Kinv_bl, times SelectionMat
      }
    }
  }

  /* ---- impact (WBM.cpp:178-254, :427-456, :508-543) */
  Vec v_post, impulse_c;  // impulse_c: stacked per active contact
  void kkt_impact(const double* q, const double* v, const int* impact_status) {
    feet_active.clear();
    for (int l = 0; l < 4; ++l) if (impact_status[l] > 0) feet_active.push_back(l);
    int nc = (int)feet_active.size();
    double zero[18] = {0};
    mass_and_bias(q, zero);
    wb::Feet F;
    wb::feet_kinematics(P, J, q, nullptr, nullptr, F);
    Jc = Mat(3 * nc, 18);
    for (int i = 0; i < nc; ++i) for (int r = 0; r < 3; ++r) for (int c = 0; c < 18; ++c) Jc(3 * i + r, c) = F.J[feet_active[i]][r][c];
    Mat MinvJt = chol_solve(Lm, transpose(Jc));
    Mat S = mm(Jc, MinvJt);
    Vec vv(v, v + 18);
    Vec rhs = mv(Jc, vv);
    for (auto& e : rhs) e = -e;
    impulse_c = chol_solve(chol_lower(S), col(rhs)).a;
    Vec dv = mv(MinvJt, impulse_c);
    v_post.assign(18, 0.0);
    for (int i = 0; i < 18; ++i) v_post[i] = v[i] + dv[i];
  }
  Vec impact(const Vec& x) {
    int st[4];
    for (int l = 0; l < 4; ++l) st[l] = (ph->contact[l] == 0 && ph->next_contact[l] == 1) ? 1 : 0;
    kkt_impact(x.data(), x.data() + 18, st);
    Vec xn(36);
    for (int i = 0; i < 18; ++i) { xn[i] = x[i]; xn[18 + i] = v_post[i]; }
    return xn;
  }
  Mat impact_partial(const Vec& x) {
    const double* q = x.data();
    const double* v = x.data() + 18;
    int st[4];
    for (int l = 0; l < 4; ++l) st[l] = (ph->contact[l] == 0 && ph->next_contact[l] == 1) ? 1 : 0;
    kkt_impact(q, v, st);
    int nc = (int)feet_active.size();
    double zero[18] = {0}, dvv[18];
    for (int i = 0; i < 18; ++i) dvv[i] = v_post[i] - v[i];
    Mat dtau_dq, dummy, dgrav, dummy2;
    rnea_derivatives(q, zero, dvv, true, dtau_dq, dummy);
    rnea_derivatives(q, zero, zero, true, dgrav, dummy2);
    madd(dtau_dq, -1.0, dgrav);
    /* impulse scatter uses segment<3>(i), not 3i (WBM.cpp:454) — kept */
    Vec impulse(12, 0.0);
    for (int i = 0; i < nc; ++i) for (int r = 0; r < 3; ++r) impulse[3 * feet_active[i] + r] = impulse_c[i + r];
    madd(dtau_dq, -1.0, force_partial(q, impulse.data()));
    Mat Ktl, Ktr, Kbl, Kbr;
    kkt_inverse(Ktl, Ktr, Kbl, Kbr);
    Mat vq[4];
    casadi_foot("vel", q, v_post.data(), nullptr, vq);
    Mat dv_dq = stack_active(vq);
    Mat dvpost_dq = mm(Ktl, dtau_dq);
    for (auto& e : dvpost_dq.a) e = -e;
    madd(dvpost_dq, -1.0, mm(Ktr, dv_dq));
    Mat dvpost_dv = mm(Ktl, M);
    Mat dP(36, 36);
    for (int i = 0; i < 18; ++i) {
      dP(i, i) = 1.0;
      for (int c = 0; c < 18; ++c) { dP(18 + i, c) = dvpost_dq(i, c); dP(18 + i, 18 + c) = dvpost_dv(i, c); }
    }
    return dP;
  }
  bool any_touchdown() const { for (int l = 0; l < 4; ++l) if (ph->next_contact[l] - ph->contact[l] == 1) return true; return false; }

  Vec resetmap(const Vec& x) override {  // MHPCReset.cpp:4-29
    Vec xn = any_touchdown() ? impact(x) : x;
    if (ph->next_model == CAFE_MODEL_SRB) { Vec s(12); for (int i = 0; i < 6; ++i) { s[i] = xn[i]; s[6 + i] = xn[18 + i]; } return s; }
    return xn;
  }
  Mat resetmap_partial(const Vec& x) override {  // MHPCReset.cpp:31-53
    Mat d(36, 36);
    if (any_touchdown()) d = impact_partial(x); else d.identity();
    if (ph->next_model == CAFE_MODEL_SRB) {
      Mat s(12, 36);
      for (int c = 0; c < 36; ++c) for (int i = 0; i < 6; ++i) { s(i, c) = d(i, c); s(6 + i, c) = d(18 + i, c); }
      return s;
    }
    return d;
  }

  /* ---- costs */
  void foot_terms(const Vec& x, wb::Feet& F) { wb::feet_kinematics(P, J, x.data(), x.data() + 18, nullptr, F); }

  void running_cost(RCost& rc, const Vec& x, const Vec& u, const Vec& y, int k) override {
    const double* r = rec(k);
    (void)y;
    double s = 0, l;
    for (int i = 0; i < 36; ++i) { double dx = x[i] - r[CAFE_REF_XR + i]; s += dx * ph->q[i] * dx; }
    l = 0.5 * s; s = 0;
    for (int i = 0; i < 12; ++i) { double du = u[i] - r[CAFE_REF_UR + i]; s += du * ph->r[i] * du; }
    l += 0.5 * s;  /* S = 0: no output term (SinglePhaseInterface.h:90) */
    l *= dt;
    wb::Feet F;
    foot_terms(x, F);
    double total = l;
    double lreg = 0, lpos = 0, lvel = 0;
    for (int f = 0; f < 4; ++f) {
      bool c = r[CAFE_REF_CONTACT + f] > 0;
      double d[3], q2 = 0;
      for (int a = 0; a < 3; ++a) d[a] = (F.p[f][a] - x[a]) - (r[CAFE_REF_PF + 3 * f + a] - r[CAFE_REF_PCOM + a]);
      const double* w = c ? ph->w_footreg : ph->w_swingpos;
      for (int a = 0; a < 3; ++a) q2 += d[a] * w[a] * d[a];
      double t = .5 * q2; t *= dt;
      if (c) lreg += t; else lpos += t;
      if (!c) {
        double q3 = 0;
        for (int a = 0; a < 3; ++a) { double dv = F.v[f][a] - r[CAFE_REF_VF + 3 * f + a]; q3 += dv * ph->w_swingvel[a] * dv; }
        double t2 = .5 * q3; t2 *= dt;
        lvel += t2;
      }
    }
    total += lreg; total += lpos; total += lvel;
    rc.l = total;
  }

  void running_cost_par(RCost& rc, const Vec& x, const Vec& u, const Vec& y, int k) override {
    const double* r = rec(k);
    (void)y;
    for (int i = 0; i < 36; ++i) { rc.lx[i] += dt * ph->q[i] * (x[i] - r[CAFE_REF_XR + i]); rc.lxx(i, i) += dt * ph->q[i]; }
    for (int i = 0; i < 12; ++i) { rc.lu[i] += dt * ph->r[i] * (u[i] - r[CAFE_REF_UR + i]); rc.luu(i, i) += dt * ph->r[i]; }
    wb::Feet F;
    foot_terms(x, F);
    Mat vq[4];
    casadi_foot("vel", x.data(), x.data() + 18, nullptr, vq);
    for (int f = 0; f < 4; ++f) {
      bool c = r[CAFE_REF_CONTACT + f] > 0;
      double d[3];
      for (int a = 0; a < 3; ++a) d[a] = (F.p[f][a] - x[a]) - (r[CAFE_REF_PF + 3 * f + a] - r[CAFE_REF_PCOM + a]);
      const double* w = c ? ph->w_footreg : ph->w_swingpos;
      /* J with its first three columns zeroed (MHPCCost.cpp:54, :184) */
      for (int i = 3; i < 18; ++i) {
        double g = 0;
        for (int a = 0; a < 3; ++a) g += F.J[f][a][i] * w[a] * d[a];
        rc.lx[i] += g * dt;
        for (int j = 3; j < 18; ++j) { double hh = 0; for (int a = 0; a < 3; ++a) hh += F.J[f][a][i] * w[a] * F.J[f][a][j]; rc.lxx(i, j) += hh * dt; }
      }
      if (!c) {  /* SwingFootVelTracking: J = [dv_dq (CasADi), J_foot] (MHPCCost.cpp:220-246) */
        double dv[3];
        for (int a = 0; a < 3; ++a) dv[a] = F.v[f][a] - r[CAFE_REF_VF + 3 * f + a];
        auto Jx = [&](int a, int i) { return i < 18 ? vq[f](a, i) : F.J[f][a][i - 18]; };
        for (int i = 0; i < 36; ++i) {
          double g = 0;
          for (int a = 0; a < 3; ++a) g += Jx(a, i) * ph->w_swingvel[a] * dv[a];
          rc.lx[i] += g * dt;
          for (int j = 0; j < 36; ++j) { double hh = 0; for (int a = 0; a < 3; ++a) hh += Jx(a, i) * ph->w_swingvel[a] * Jx(a, j); rc.lxx(i, j) += hh * dt; }
        }
      }
    }
  }

  void terminal_cost(TCost& tc, const Vec& x) override {
    const double* r = rec(h);
    double s = 0;
    for (int i = 0; i < 36; ++i) { double dx = x[i] - r[CAFE_REF_XR + i]; s += dx * ph->qf[i] * dx; }
    double phi = s * 0.5;
    wb::Feet F;
    foot_terms(x, F);
    double reg = 0;
    for (int f = 0; f < 4; ++f) {
      if (!(r[CAFE_REF_CONTACT + f] > 0)) continue;  // WBFootPlaceReg::terminal_cost (MHPCCost.cpp:66-88)
      double q2 = 0;
      for (int a = 0; a < 3; ++a) { double d = (F.p[f][a] - x[a]) - (r[CAFE_REF_PF + 3 * f + a] - r[CAFE_REF_PCOM + a]); q2 += d * ph->w_footreg[a] * d; }
      reg += .5 * q2;
    }
    double td = 0;
    for (int i = 0; i < ph->n_td; ++i) { double dv = F.v[ph->td_foot[i]][2]; td += .5 * dv * ph->w_tdvel[2] * dv; }  // TDVelocityPenalty (MHPCCost.cpp:254-268)
    tc.Phi = phi + reg + td;
  }

  void terminal_cost_par(TCost& tc, const Vec& x) override {
    const double* r = rec(h);
    for (int i = 0; i < 36; ++i) { tc.Phix[i] += ph->qf[i] * (x[i] - r[CAFE_REF_XR + i]); tc.Phixx(i, i) += ph->qf[i]; }
    wb::Feet F;
    foot_terms(x, F);
    for (int f = 0; f < 4; ++f) {
      if (!(r[CAFE_REF_CONTACT + f] > 0)) continue;
      double d[3];
      for (int a = 0; a < 3; ++a) d[a] = (F.p[f][a] - x[a]) - (r[CAFE_REF_PF + 3 * f + a] - r[CAFE_REF_PCOM + a]);
      for (int i = 3; i < 18; ++i) {  // factor 2 (MHPCCost.cpp:114-115)
        double g = 0;
        for (int a = 0; a < 3; ++a) g += F.J[f][a][i] * ph->w_footreg[a] * d[a];
        tc.Phix[i] += 2 * g;
        for (int j = 3; j < 18; ++j) { double hh = 0; for (int a = 0; a < 3; ++a) hh += F.J[f][a][i] * ph->w_footreg[a] * F.J[f][a][j]; tc.Phixx(i, j) += 2 * hh; }
      }
    }
    if (ph->n_td > 0) {  // TDVelocityPenalty::terminal_cost_par (MHPCCost.cpp:270-291)
      Mat vq[4];
      casadi_foot("vel", x.data(), x.data() + 18, nullptr, vq);
      for (int t = 0; t < ph->n_td; ++t) {
        int f = ph->td_foot[t];
        double dv = F.v[f][2];
        auto Jz = [&](int i) { return i < 18 ? vq[f](2, i) : F.J[f][2][i - 18]; };
        for (int i = 0; i < 36; ++i) {
          tc.Phix[i] += Jz(i) * ph->w_tdvel[2] * dv;
          for (int j = 0; j < 36; ++j) tc.Phixx(i, j) += Jz(i) * ph->w_tdvel[2] * Jz(j);
        }
      }
    }
  }

  /* ---- constraints */
  void path_constraints(const Vec& x, const Vec& u, const Vec& y, int k) override {
    for (auto& pc : pcon) {
      if (pc.kind == 0) for (int i = 0; i < 12; ++i) { pc.data[k][i].g = -u[i] - (-ph->torque_limit); pc.data[k][12 + i].g = u[i] - (-ph->torque_limit); }
      else if (pc.kind == 1) for (int i = 0; i < 12; ++i) { pc.data[k][i].g = x[6 + i] - ph->joint_lb[i % 3]; pc.data[k][12 + i].g = -x[6 + i] - (-ph->joint_ub[i % 3]); }
      else if (pc.kind == 2) pc.data[k][0].g = x[2] - ph->h_min;
      else if (pc.kind == 4) for (int i = 0; i < 12; ++i) { pc.data[k][i].g = x[24 + i] - ph->jointvel_lb; pc.data[k][12 + i].g = -x[24 + i] - (-ph->jointvel_ub); }
      else {
        int i = 0;
        const double mu = ph->mu;
        for (int l = 0; l < 4; ++l) if (ph->contact[l] > 0) {
          const double fx = y[3 * l], fy = y[3 * l + 1], fz = y[3 * l + 2];
          const double g[5] = {fz, -fx + mu * fz, fx + mu * fz, -fy + mu * fz, fy + mu * fz};
          for (int r = 0; r < 5; ++r) pc.data[k][5 * i + r].g = g[r];
          ++i;
        }
      }
      pc.update_max_violation(k);
    }
  }
  void path_constraints_par(const Vec&, const Vec&, const Vec&, int k) override {
    for (auto& pc : pcon) {
      if (pc.kind == 0) for (int i = 0; i < 12; ++i) { pc.data[k][i].gu.assign(12, 0.0); pc.data[k][i].gu[i] = -1; pc.data[k][12 + i].gu.assign(12, 0.0); pc.data[k][12 + i].gu[i] = 1; }
      else if (pc.kind == 1) for (int i = 0; i < 12; ++i) { pc.data[k][i].gx.assign(36, 0.0); pc.data[k][i].gx[6 + i] = 1; pc.data[k][12 + i].gx.assign(36, 0.0); pc.data[k][12 + i].gx[6 + i] = -1; }
      else if (pc.kind == 2) { pc.data[k][0].gx.assign(36, 0.0); pc.data[k][0].gx[2] = 1; }
      else if (pc.kind == 4) for (int i = 0; i < 12; ++i) { pc.data[k][i].gx.assign(36, 0.0); pc.data[k][i].gx[24 + i] = 1; pc.data[k][12 + i].gx.assign(36, 0.0); pc.data[k][12 + i].gx[24 + i] = -1; }
      else {
        int i = 0;
        const double mu = ph->mu;
        const double Al[5][3] = {{0, 0, 1}, {-1, 0, mu}, {1, 0, mu}, {0, -1, mu}, {0, 1, mu}};
        for (int l = 0; l < 4; ++l) if (ph->contact[l] > 0) {
          for (int r = 0; r < 5; ++r) { pc.data[k][5 * i + r].gy.assign(12, 0.0); for (int c = 0; c < 3; ++c) pc.data[k][5 * i + r].gy[3 * l + c] = Al[r][c]; }
          ++i;
        }
      }
    }
  }
  void terminal_constraints(const Vec& x) override {  // WBTouchDown (MHPCConstraint.cpp:253-277)
    if (tcon.empty()) return;
    wb::Feet F;
    wb::feet_kinematics(P, J, x.data(), nullptr, nullptr, F);
    for (int i = 0; i < tcon[0].size; ++i) tcon[0].data[i].h = F.p[ph->td_foot[i]][2] - ph->ground_height;
    tcon[0].update_max_violation();
  }
  void terminal_constraints_par(const Vec& x) override {  // :279-288
    if (tcon.empty()) return;
    wb::Feet F;
    wb::feet_kinematics(P, J, x.data(), nullptr, nullptr, F);
    for (int i = 0; i < tcon[0].size; ++i) { Vec& hx = tcon[0].data[i].hx; hx.assign(36, 0.0); for (int c = 0; c < 18; ++c) hx[c] = F.J[ph->td_foot[i]][2][c]; }
  }
};

std::unique_ptr<Phase> make_wb_phase(double BG_alpha, double hip_yaw) { return std::unique_ptr<Phase>(new WBPhase(BG_alpha, hip_yaw)); }

/* test hook: continuous-time KKT dynamics (dynamics_continuousTime, WBM.cpp:37-57) */
extern "C" int cafe_oracle_wb_dynamics(double hip_yaw, double BG_alpha, const double* q, const double* v, const double* u,
                                       const int* contact, double* qdd, double* grf) {
  try {
    WBPhase w(BG_alpha, hip_yaw);
    double tau[18] = {0};
    for (int i = 0; i < 12; ++i) tau[6 + i] = u[i];
    w.kkt_dynamics(q, v, tau, contact);
    for (int i = 0; i < 18; ++i) qdd[i] = w.qdd[i];
    for (int i = 0; i < 12; ++i) grf[i] = w.GRF[i];
    return 0;
  } catch (...) { return -1; }
}

}  // namespace oracle
