/*
 * hkd_oracle.cpp — CPU ORACLE (test infrastructure only): the HKD (hybrid kinodynamic) model,
 * costs, constraints and reset map, restated from the reference and evaluating the reference's
 * own CasADi-generated C (oracle/_ref) for every model expression.
 *
 *   HKD::Model::dynamics / dynamics_partial   /root/reference/HKDMPC/HKD-TrajOpt/HKDModel.h:33-61
 *   HKDReset::resetmap / resetmap_partial     /root/reference/HKDMPC/HKD-TrajOpt/HKDReset.h:41-136
 *   GRFConstraint, TouchDownConstraint        /root/reference/HKDMPC/HKD-TrajOpt/HKDConstraints.cpp:7-171
 *   HKDTrackingCost                           /root/reference/HKDMPC/HKD-TrajOpt/HKDCost.h:8-38
 *   HKDFootPlaceReg                           /root/reference/HKDMPC/HKD-TrajOpt/HKDCost.cpp:5-66, HKDCost.h:52-75
 *   problem wiring                            /root/reference/HKDMPC/HKD-TrajOpt/HKDProblem.cpp:224-311
 */
#include "hsddp_oracle.hpp"
#include "casadi_ref.hpp"

namespace oracle {

static Vec foot_position(const Vec& x, int leg) {
  double pos[3] = {x[3], x[4], x[5]}, eul[3] = {x[0], x[1], x[2]}, ql[3] = {x[12 + 3 * leg], x[13 + 3 * leg], x[14 + 3 * leg]};
  double id = (double)leg + 1.0;
  const double* arg[4] = {pos, eul, ql, &id};
  Vec pf(3, 0.0);
  double* res[1] = {pf.data()};
  casadi_call(CASADI_FN(compute_foot_position), arg, 4, res, 1);
  return pf;
}

static Mat foot_jacobian(const Vec& x, int leg) {  // 3 x 18, columns [pos(3), eul(3), qJ(12)]
  double pos[3] = {x[3], x[4], x[5]}, eul[3] = {x[0], x[1], x[2]}, ql[3] = {x[12 + 3 * leg], x[13 + 3 * leg], x[14 + 3 * leg]};
  const double* arg[3] = {pos, eul, ql};
  Mat J(3, 18);
  double* res[1] = {J.a.data()};
  switch (leg) {
    case 0: casadi_call(CASADI_FN(comp_foot_jacob_1), arg, 3, res, 1); break;
    case 1: casadi_call(CASADI_FN(comp_foot_jacob_2), arg, 3, res, 1); break;
    case 2: casadi_call(CASADI_FN(comp_foot_jacob_3), arg, 3, res, 1); break;
    default: casadi_call(CASADI_FN(comp_foot_jacob_4), arg, 3, res, 1); break;
  }
  return J;
}

class HKDPhase : public Phase {
 public:
  Mat Agrf;  // GRFConstraint::A (5 n_c x 24)

  void build_model() override {
    int nc = 0;
    for (int l = 0; l < 4; ++l) nc += ph->contact[l] > 0;
    if (nc > 0) {  // HKDProblem.cpp:259-268
      double mu = ph->mu;
      const double Aleg[5][3] = {{0, 0, 1}, {-1, 0, mu}, {1, 0, mu}, {0, -1, mu}, {0, 1, mu}};
      Agrf = Mat(5 * nc, 24);
      int i = 0;
      for (int l = 0; l < 4; ++l)
        if (ph->contact[l] > 0) {
          for (int r = 0; r < 5; ++r) for (int c = 0; c < 3; ++c) Agrf(5 * i + r, 3 * l + c) = Aleg[r][c];
          ++i;
        }
      PathConstraint pc;
      pc.create(5 * nc, h, n, m, p, ph->reb_grf);
      pcon.push_back(pc);
    }
    if (ph->n_td > 0) {  // HKDProblem.cpp:302-310
      TermConstraint tc;
      tc.create(ph->n_td, n, ph->al_td);
      tcon.push_back(tc);
    }
  }

  void dynamics(Vec& xnext, Vec& y, const Vec& x, const Vec& u, int) override {
    (void)y;
    double c[4] = {(double)ph->contact[0], (double)ph->contact[1], (double)ph->contact[2], (double)ph->contact[3]};
    double dtl = dt;
    const double* arg[4] = {x.data(), u.data(), &dtl, c};
    xnext.assign(24, 0.0);
    double* res[1] = {xnext.data()};
    casadi_call(CASADI_FN(hkinodyn), arg, 4, res, 1);
  }
  void dynamics_partial(Mat& A_, Mat& B_, Mat&, Mat&, const Vec& x, const Vec& u, int) override {
    double c[4] = {(double)ph->contact[0], (double)ph->contact[1], (double)ph->contact[2], (double)ph->contact[3]};
    double dtl = dt;
    const double* arg[4] = {x.data(), u.data(), &dtl, c};
    A_.zero(); B_.zero();
    double* res[2] = {A_.a.data(), B_.a.data()};
    casadi_call(CASADI_FN(hkinodyn_par), arg, 4, res, 2);
  }

  /* d_prel of HKDFootPlaceReg (HKDCost.cpp:10-18) */
  Vec d_prel(const Vec& x, const double* r) const {
    Vec d(12);
    for (int l = 0; l < 4; ++l) for (int a = 0; a < 3; ++a) {
      double prel = x[12 + 3 * l + a] - x[3 + a];
      double prel_r = r[CAFE_REF_PF + 3 * l + a] - r[CAFE_REF_PCOM + a];
      d[3 * l + a] = prel - prel_r;
    }
    return d;
  }
  double qfoot(int i) const { int l = i / 3, a = i % 3; return (a < 2 ? (double)ph->contact[l] : 0.0) * ph->w_footreg[a]; }

  void running_cost(RCost& rc, const Vec& x, const Vec& u, const Vec&, int k) override {
    const double* r = rec(k);
    /* QuadraticTrackingCost::running_cost (SinglePhaseInterface.cpp:21-33, :70-86) */
    double l = 0, s = 0;
    for (int i = 0; i < 24; ++i) { double dx = x[i] - r[CAFE_REF_XR + i]; s += dx * ph->q[i] * dx; }
    l = 0.5 * s;
    s = 0;
    for (int i = 0; i < 24; ++i) { double du = u[i] - r[CAFE_REF_UR + i]; s += du * ph->r[i] * du; }
    l += 0.5 * s;
    l *= dt;
    /* HKDFootPlaceReg::running_cost (HKDCost.cpp:5-20) */
    Vec d = d_prel(x, r);
    double lf = 0;
    for (int i = 0; i < 12; ++i) lf += d[i] * qfoot(i) * d[i];
    lf = .5 * lf;
    lf *= dt;
    rc.l = l + lf;
  }
  void running_cost_par(RCost& rc, const Vec& x, const Vec& u, const Vec&, int k) override {
    const double* r = rec(k);
    for (int i = 0; i < 24; ++i) {  // SinglePhaseInterface.cpp:35-49
      rc.lx[i] += dt * ph->q[i] * (x[i] - r[CAFE_REF_XR + i]);
      rc.lu[i] += dt * ph->r[i] * (u[i] - r[CAFE_REF_UR + i]);
      rc.lxx(i, i) += dt * ph->q[i];
      rc.luu(i, i) += dt * ph->r[i];
    }
    /* HKDCost.cpp:22-36: lx = dt J^T Qfoot d, lxx = dt J^T Qfoot J with J = dprel_dx (HKDCost.h:62-69) */
    Vec d = d_prel(x, r);
    for (int l = 0; l < 4; ++l) for (int a = 0; a < 3; ++a) {
      double c = (double)ph->contact[l];
      double w = qfoot(3 * l + a);
      /* row (3l+a) of J: -c at column 3+a, +c at column 12+3l+a */
      double gd = w * d[3 * l + a];
      rc.lx[3 + a] += dt * (-c) * gd;
      rc.lx[12 + 3 * l + a] += dt * c * gd;
      rc.lxx(3 + a, 3 + a) += dt * c * w * c;
      rc.lxx(12 + 3 * l + a, 12 + 3 * l + a) += dt * c * w * c;
      rc.lxx(3 + a, 12 + 3 * l + a) += dt * (-c) * w * c;
      rc.lxx(12 + 3 * l + a, 3 + a) += dt * c * w * (-c);
    }
  }
  void terminal_cost(TCost& tc, const Vec& x) override {
    const double* r = rec(h);
    double s = 0;
    for (int i = 0; i < 24; ++i) { double dx = x[i] - r[CAFE_REF_XR + i]; s += dx * ph->qf[i] * dx; }
    double phi = s * 0.5;  // SinglePhaseInterface.cpp:52-59
    Vec d = d_prel(x, r);  // HKDCost.cpp:38-50
    double sf = 0;
    for (int i = 0; i < 12; ++i) sf += d[i] * qfoot(i) * d[i];
    tc.Phi = phi + 10 * sf;
  }
  void terminal_cost_par(TCost& tc, const Vec& x) override {
    const double* r = rec(h);
    for (int i = 0; i < 24; ++i) { tc.Phix[i] += ph->qf[i] * (x[i] - r[CAFE_REF_XR + i]); tc.Phixx(i, i) += ph->qf[i]; }
    Vec d = d_prel(x, r);  // HKDCost.cpp:52-66
    for (int l = 0; l < 4; ++l) for (int a = 0; a < 3; ++a) {
      double c = (double)ph->contact[l];
      double w = qfoot(3 * l + a);
      double gd = w * d[3 * l + a];
      tc.Phix[3 + a] += 20 * (-c) * gd;
      tc.Phix[12 + 3 * l + a] += 20 * c * gd;
      tc.Phixx(3 + a, 3 + a) += 20 * c * w * c;
      tc.Phixx(12 + 3 * l + a, 12 + 3 * l + a) += 20 * c * w * c;
      tc.Phixx(3 + a, 12 + 3 * l + a) += 20 * (-c) * w * c;
      tc.Phixx(12 + 3 * l + a, 3 + a) += 20 * c * w * (-c);
    }
  }
  void path_constraints(const Vec&, const Vec& u, const Vec&, int k) override {  // HKDConstraints.cpp:34-52
    if (pcon.empty()) return;
    PathConstraint& pc = pcon[0];
    for (int i = 0; i < pc.size; ++i) { double g = 0; for (int j = 0; j < 24; ++j) g += Agrf(i, j) * u[j]; pc.data[k][i].g = g; }
    pc.update_max_violation(k);
  }
  void path_constraints_par(const Vec&, const Vec&, const Vec&, int k) override {  // :54-66
    if (pcon.empty()) return;
    PathConstraint& pc = pcon[0];
    for (int i = 0; i < pc.size; ++i) for (int j = 0; j < 24; ++j) pc.data[k][i].gu[j] = Agrf(i, j);
  }
  void terminal_constraints(const Vec& x) override {  // HKDConstraints.cpp:79-121
    if (tcon.empty()) return;
    TermConstraint& tc = tcon[0];
    for (int i = 0; i < tc.size; ++i) { Vec pf = foot_position(x, ph->td_foot[i]); tc.data[i].h = pf[2] - ph->ground_height; }
    tc.update_max_violation();
  }
  void terminal_constraints_par(const Vec& x) override {  // HKDConstraints.cpp:123-171
    if (tcon.empty()) return;
    TermConstraint& tc = tcon[0];
    for (int i = 0; i < tc.size; ++i) {
      Mat J = foot_jacobian(x, ph->td_foot[i]);
      Vec& hx = tc.data[i].hx;
      for (int j = 0; j < 3; ++j) { hx[j] = J(2, 3 + j); hx[3 + j] = J(2, j); }
      for (int j = 0; j < 12; ++j) hx[12 + j] = J(2, 6 + j);
    }
  }
  Vec resetmap(const Vec& x) override {  // HKDReset.h:41-76
    Vec xn = x;
    for (int l = 0; l < 4; ++l) {
      int c = ph->contact[l], cn = ph->next_contact[l];
      if (c && !cn) { xn[12 + 3 * l] = 0.0; xn[13 + 3 * l] = -0.8; xn[14 + 3 * l] = 1.7; }
      if (!c && cn) { Vec pf = foot_position(x, l); xn[12 + 3 * l] = 1 * pf[0]; xn[13 + 3 * l] = 1 * pf[1]; xn[14 + 3 * l] = 0 * pf[2]; }
    }
    return xn;
  }
  Mat resetmap_partial(const Vec& x) override {  // HKDReset.h:78-136
    Mat Px(24, 24);
    Px.identity();
    for (int l = 0; l < 4; ++l) {
      int c = ph->contact[l], cn = ph->next_contact[l];
      if (c && !cn) for (int r = 0; r < 3; ++r) for (int j = 0; j < 24; ++j) Px(12 + 3 * l + r, j) = 0;
      if (!c && cn) {
        Mat J = foot_jacobian(x, l);
        const double cmap[3] = {1, 1, 0};
        for (int r = 0; r < 3; ++r) {
          for (int j = 0; j < 3; ++j) { Px(12 + 3 * l + r, j) = cmap[r] * J(r, 3 + j); Px(12 + 3 * l + r, 3 + j) = cmap[r] * J(r, j); }
          for (int j = 0; j < 12; ++j) Px(12 + 3 * l + r, 12 + j) = cmap[r] * J(r, 6 + j);
        }
      }
    }
    return Px;
  }
};

std::unique_ptr<Phase> make_hkd_phase() { return std::unique_ptr<Phase>(new HKDPhase()); }

}  // namespace oracle
