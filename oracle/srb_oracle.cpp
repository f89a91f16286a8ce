/*
 * srb_oracle.cpp — CPU ORACLE (test infrastructure only): single-rigid-body phase of the MHPC problem.
 *   SRBM::Model::dynamics / dynamics_partial    MHPC/MHPC-Trajopt/SRBM.h:43-93   (reference CasADi C from oracle/_ref)
 *   MHPCFootStep::getFootPositions / getContactStatus (reference values at time t)  MHPC/MHPC-Trajopt/MHPCFootStep.h:37-69
 *   SRBTrackingCost                             MHPC/MHPC-Trajopt/MHPCCost.h:207-249, MHPCCostUtil.h:82-110
 *   SRBMMinimumHeight                           MHPC/MHPC-Trajopt/MHPCConstraint.cpp:355-379, MHPCConstraint.h:199
 *   wiring                                      MHPC/MHPC-Trajopt/MHPCProblem.cpp:487-521
 */
#include "hsddp_oracle.hpp"
#include "casadi_ref.hpp"

namespace oracle {

class SRBPhase : public Phase {
 public:
  void build_model() override {
    PathConstraint mh;
    mh.create(1, h, n, m, p, ph->reb_minheight);
    pcon.push_back(mh);
  }
  void dynamics(Vec& xnext, Vec& y, const Vec& x, const Vec& u, int k) override {
    (void)y;
    const double* r = rec(k);
    Vec xdot(12, 0.0);
    const double* arg[4] = {x.data(), u.data(), r + CAFE_REF_PF, r + CAFE_REF_CONTACT};
    double* res[1] = {xdot.data()};
    casadi_call(CASADI_FN(SRBDynamics), arg, 4, res, 1);
    xnext.assign(12, 0.0);
    for (int i = 0; i < 12; ++i) xnext[i] = x[i] + xdot[i] * dt;
  }
  void dynamics_partial(Mat& A_, Mat& B_, Mat&, Mat&, const Vec& x, const Vec& u, int k) override {
    const double* r = rec(k);
    Mat Ac(12, 12), Bc(12, 12);
    const double* arg[4] = {x.data(), u.data(), r + CAFE_REF_PF, r + CAFE_REF_CONTACT};
    double* res[2] = {Ac.a.data(), Bc.a.data()};
    casadi_call(CASADI_FN(SRBDynamicsDerivatives), arg, 4, res, 2);
    A_.identity();
    madd(A_, dt, Ac);
    B_.zero();
    madd(B_, dt, Bc);
  }
  void running_cost(RCost& rc, const Vec& x, const Vec& u, const Vec&, int k) override {
    const double* r = rec(k);
    double s = 0, l;
    for (int i = 0; i < 12; ++i) { double dx = x[i] - r[CAFE_REF_XR + i]; s += dx * ph->q[i] * dx; }
    l = 0.5 * s; s = 0;
    for (int i = 0; i < 12; ++i) { double du = u[i] - r[CAFE_REF_UR + i]; s += du * ph->r[i] * du; }
    l += 0.5 * s;
    l *= dt;
    rc.l = l;
  }
  void running_cost_par(RCost& rc, const Vec& x, const Vec& u, const Vec&, int k) override {
    const double* r = rec(k);
    for (int i = 0; i < 12; ++i) {
      rc.lx[i] += dt * ph->q[i] * (x[i] - r[CAFE_REF_XR + i]); rc.lxx(i, i) += dt * ph->q[i];
      rc.lu[i] += dt * ph->r[i] * (u[i] - r[CAFE_REF_UR + i]); rc.luu(i, i) += dt * ph->r[i];
    }
  }
  void terminal_cost(TCost& tc, const Vec& x) override {
    const double* r = rec(h);
    double s = 0;
    for (int i = 0; i < 12; ++i) { double dx = x[i] - r[CAFE_REF_XR + i]; s += dx * ph->qf[i] * dx; }
    tc.Phi = s * 0.5;
  }
  void terminal_cost_par(TCost& tc, const Vec& x) override {
    const double* r = rec(h);
    for (int i = 0; i < 12; ++i) { tc.Phix[i] += ph->qf[i] * (x[i] - r[CAFE_REF_XR + i]); tc.Phixx(i, i) += ph->qf[i]; }
  }
  void path_constraints(const Vec& x, const Vec&, const Vec&, int k) override { pcon[0].data[k][0].g = x[2] - ph->h_min; pcon[0].update_max_violation(k); }
  void path_constraints_par(const Vec&, const Vec&, const Vec&, int k) override { pcon[0].data[k][0].gx.assign(12, 0.0); pcon[0].data[k][0].gx[2] = 1; }
  void terminal_constraints(const Vec&) override {}
  void terminal_constraints_par(const Vec&) override {}
  Vec resetmap(const Vec& x) override { return x; }
  Mat resetmap_partial(const Vec&) override { Mat I(12, 12); I.identity(); return I; }
};

std::unique_ptr<Phase> make_srb_phase() { return std::unique_ptr<Phase>(new SRBPhase()); }

}  // namespace oracle
