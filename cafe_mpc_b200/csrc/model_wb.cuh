// model_wb.cuh — device-side whole-body (WB) phase of the MHPC problem, thread-per-knot parts: the TERMINAL knot of a phase (terminal
// cost, touchdown constraints, impact map and its Jacobian) and the short single-shooting chains of an MPC tail phase. Rigid-body
// terms come from the generated straight-line functions of gen/wb_gen.h (no Pinocchio, no CasADi), the small dense KKT algebra
// runs per thread on local arrays. The RUNNING knots (the hot path) live in wb_leg_kernels.cu + wb_coop.cuh.
//
// Reference behaviour followed (file:line under /root/reference):
//   WBM::Model::dynamics / dynamics_partial         MHPC/MHPC-Trajopt/WBM.cpp:17-139
//   KKTContactDynamics / ...Derivatives             MHPC/MHPC-Trajopt/WBM.cpp:368-424, :459-505
//   KKTImpact / KKTImpactDerivatives / impact       MHPC/MHPC-Trajopt/WBM.cpp:178-254, :427-456, :508-543
//   MHPCReset::reset_map(_partial)                  MHPC/MHPC-Trajopt/MHPCReset.cpp:4-53
//   WB costs                                        MHPC/MHPC-Trajopt/MHPCCost.cpp:4-291, MHPCCost.h:8-205
//   WB constraints                                  MHPC/MHPC-Trajopt/MHPCConstraint.cpp:9-288
// Pinocchio semantics restated: forwardDynamics (Cholesky(M); J Minv J^T + 1e-12 I; lambda = -(JMinvJt)^-1 (J Minv (tau-nle) + gamma);
// qdd = Minv (tau - nle + J^T lambda)), impulseDynamics (r = 0), KKT matrix inverse (damping 0) — applied here column by column
// instead of forming the inverse:  with R = d(M qdd + nle - J^T lambda - tau)/dz and a = d(J qdd + gamma)/dz,
//   dlambda/dz = S^-1 (J Minv R - a),   dqdd/dz = -Minv (R - J^T dlambda/dz),   S = J Minv J^T.
#pragma once
#include "device_types.cuh"
#include "model_hkd.cuh"

namespace cafe_dev {

// One out-of-line copy of every generated routine (defined in wb_gen_wrappers.cu, compiled separately because ptxas needs
// minutes for these 10^4-op straight-line functions), writing into plain (local-memory) arrays; callers pre-zero the outputs.
__device__ void wbg_terms(const double* q, const double* v, double* nle, double* Mlow, double* J, double* gam, double* pf, double* vf);
__device__ void wbg_feet(const double* q, const double* v, double* pf, double* vf, double* J);
__device__ void wbg_rnea_derivs(const double* q, const double* v, const double* a, double* dq, double* dv, size_t st);
__device__ void wbg_grav_derivs(const double* q, double* dq);
__device__ void wbg_kin_partials(const double* q, const double* v, const double* a, const double* F, double* dvq, double* daq, double* dav, double* djtf, size_t st);
__device__ void wbg_footvel_partial(const double* q, const double* v, double* dvq);

struct WBScratch {
  double L[324];   // M (lower) then its Cholesky factor, column-major, ld 18
  double J[216];   // foot Jacobians, rows 3f+r, ld 12
  double nle[18], gam[12], pf[12], vf[12];
  double Y[216];   // L^-1 Jc^T, 18 x nr, ld 18
  double Ls[144];  // Cholesky of S = Y^T Y (+ damping), nr x nr, ld 12
  double qdd[18], grf[12], lam[12];
  int rows[12];
  int nr;
};

// in-place Cholesky of a lower-stored n x n matrix (ld)
__device__ __forceinline__ void chol_inplace(double* A, int n, int ld) {
  for (int j = 0; j < n; ++j) {
    double d = A[j + ld * j];
    for (int k = 0; k < j; ++k) d -= A[j + ld * k] * A[j + ld * k];
    d = sqrt(d);
    A[j + ld * j] = d;
    const double inv = 1.0 / d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[i + ld * j];
      for (int k = 0; k < j; ++k) s -= A[i + ld * k] * A[j + ld * k];
      A[i + ld * j] = s * inv;
    }
  }
}
__device__ __forceinline__ void fwd_subst(const double* L, int n, int ld, double* x) {  // L y = x
  for (int i = 0; i < n; ++i) { double s = x[i]; for (int k = 0; k < i; ++k) s -= L[i + ld * k] * x[k]; x[i] = s / L[i + ld * i]; }
}
__device__ __forceinline__ void bwd_subst(const double* L, int n, int ld, double* x) {  // L^T y = x
  for (int i = n - 1; i >= 0; --i) { double s = x[i]; for (int k = i + 1; k < n; ++k) s -= L[k + ld * i] * x[k]; x[i] = s / L[i + ld * i]; }
}

struct WBModel {
  static constexpr int N = 36, M = 12, PY = 12;
  // running knots are evaluated by the leg-parallel / cooperative kernels (wb_leg_kernels.cu, wb_coop.cuh); what stays here is the
  // terminal knot (cost, touchdown constraints, impact map and its Jacobian) and the short single-shooting chains of an MPC tail phase
  static constexpr bool COOP = true;

  __device__ static void active_rows(const int* contact, WBScratch& s) {
    s.nr = 0;
    for (int f = 0; f < 4; ++f) if (contact[f] > 0) for (int r = 0; r < 3; ++r) s.rows[s.nr++] = 3 * f + r;
  }

  // M, nle, J, Jdot v, foot positions and velocities at (q, v); Cholesky of M; Y = L^-1 Jc^T; S = Y^T Y
  __device__ static void kkt_setup(const double* x, WBScratch& s, double damping) {
    for (int i = 0; i < 324; ++i) s.L[i] = 0;
    for (int i = 0; i < 216; ++i) s.J[i] = 0;
    wbg_terms(x, x + 18, s.nle, s.L, s.J, s.gam, s.pf, s.vf);
    chol_inplace(s.L, 18, 18);
    const int nr = s.nr;
    for (int c = 0; c < nr; ++c) {
      double* y = s.Y + 18 * c;
      for (int i = 0; i < 18; ++i) y[i] = s.J[s.rows[c] + 12 * i];
      fwd_subst(s.L, 18, 18, y);
    }
    for (int c = 0; c < nr; ++c)
      for (int r = c; r < nr; ++r) {
        double d = 0;
        for (int i = 0; i < 18; ++i) d += s.Y[i + 18 * r] * s.Y[i + 18 * c];
        s.Ls[r + 12 * c] = d + ((r == c) ? damping : 0.0);
      }
    if (nr > 0) chol_inplace(s.Ls, nr, 12);
  }

  // KKTContactDynamics (WBM.cpp:368-424): qdd, GRF for the phase contact set
  __device__ static void forward(const PhaseDev& ph, const double* x, const double* u, WBScratch& s) {
    active_rows(ph.contact, s);
    kkt_setup(x, s, 1e-12);
    double b[18];
    for (int i = 0; i < 18; ++i) b[i] = ((i >= 6) ? u[i - 6] : 0.0) - s.nle[i];
    double mb[18];
    for (int i = 0; i < 18; ++i) mb[i] = b[i];
    fwd_subst(s.L, 18, 18, mb);
    const int nr = s.nr;
    for (int i = 0; i < 12; ++i) s.grf[i] = 0;
    if (nr > 0) {
      // rhs = -Jc Minv b - gamma = -(Y^T L^-1 b) - gamma
      for (int c = 0; c < nr; ++c) {
        double d = 0;
        for (int i = 0; i < 18; ++i) d += s.Y[i + 18 * c] * mb[i];
        const int r = s.rows[c];
        s.lam[c] = -d - (s.gam[r] + 2.0 * ph.BG_alpha * s.vf[r]);
      }
      fwd_subst(s.Ls, nr, 12, s.lam);
      bwd_subst(s.Ls, nr, 12, s.lam);
      for (int c = 0; c < nr; ++c) { s.grf[s.rows[c]] = s.lam[c]; for (int i = 0; i < 18; ++i) mb[i] += s.Y[i + 18 * c] * s.lam[c]; }
    }
    bwd_subst(s.L, 18, 18, mb);  // qdd = L^-T (L^-1 b + Y lambda)
    for (int i = 0; i < 18; ++i) s.qdd[i] = mb[i];
  }

  // running cost from already evaluated foot kinematics (pf, vf) + ReB terms
  __device__ static double running_cost_k(const PhaseDev& ph, const double* rec, const double* x, const double* u, const double* y,
                                          const double* pf, const double* vf, bool reb, double& ming, const RebCtx& rcx) {
    double s = 0;
    for (int i = 0; i < 36; ++i) { const double dx = x[i] - rec[CAFE_REF_XR + i]; s += dx * ph.q[i] * dx; }
    double l = 0.5 * s;
    s = 0;
    for (int i = 0; i < 12; ++i) { const double du = u[i] - rec[CAFE_REF_UR + i]; s += du * ph.r[i] * du; }
    l += 0.5 * s;
    l *= ph.dt;
    double lreg = 0, lpos = 0, lvel = 0;
    for (int f = 0; f < 4; ++f) {
      const bool c = rec[CAFE_REF_CONTACT + f] > 0;
      const double* w = c ? ph.w_footreg : ph.w_swingpos;
      double q2 = 0;
      for (int a = 0; a < 3; ++a) { const double d = (pf[3 * f + a] - x[a]) - (rec[CAFE_REF_PF + 3 * f + a] - rec[CAFE_REF_PCOM + a]); q2 += d * w[a] * d; }
      double t = .5 * q2; t *= ph.dt;
      if (c) lreg += t; else lpos += t;
      if (!c) {
        double q3 = 0;
        for (int a = 0; a < 3; ++a) { const double dv = vf[3 * f + a] - rec[CAFE_REF_VF + 3 * f + a]; q3 += dv * ph.w_swingvel[a] * dv; }
        double t2 = .5 * q3; t2 *= ph.dt;
        lvel += t2;
      }
    }
    l += lreg; l += lpos; l += lvel;
    ming = 0;
    if (true) {
      // path constraints in the reference's order: torque, [joint speed: BarrelRollTO.cpp:190-198], joint, min height, GRF (MHPCProblem.cpp:436-481)
      // element numbering of the per-element barrier parameters: torque 0..23 | joint speed 24..47 | joint 48..71 | min height 72 | GRF 73 + 5 f + i
      double c_t = 0, c_j = 0, c_h = 0, c_g = 0, c_v = 0, m_t = 0, m_j = 0, m_h = 0, m_g = 0, m_v = 0, dl, ep;
      for (int i = 0; i < 12; ++i) { const double g = -u[i] - (-ph.torque_limit); m_t = fmin(m_t, g); rcx.get(ph.reb_torque, i, dl, ep); c_t += ep * reb_value(g, dl); }
      for (int i = 0; i < 12; ++i) { const double g = u[i] - (-ph.torque_limit); m_t = fmin(m_t, g); rcx.get(ph.reb_torque, 12 + i, dl, ep); c_t += ep * reb_value(g, dl); }
      const bool jl = !ph.no_joint_limit, mh = !ph.no_min_height;  // LocoProblem keeps torque + GRF only (LocoProblem.cpp:64-82)
      if (ph.joint_speed_limit) {
        for (int i = 0; i < 12; ++i) { const double g = x[24 + i] - ph.jointvel_lb; m_v = fmin(m_v, g); rcx.get(ph.reb_jointvel, 24 + i, dl, ep); c_v += ep * reb_value(g, dl); }
        for (int i = 0; i < 12; ++i) { const double g = -x[24 + i] - (-ph.jointvel_ub); m_v = fmin(m_v, g); rcx.get(ph.reb_jointvel, 36 + i, dl, ep); c_v += ep * reb_value(g, dl); }
      }
      if (jl) {
        for (int i = 0; i < 12; ++i) { const double g = x[6 + i] - ph.joint_lb[i % 3]; m_j = fmin(m_j, g); rcx.get(ph.reb_joint, 48 + i, dl, ep); c_j += ep * reb_value(g, dl); }
        for (int i = 0; i < 12; ++i) { const double g = -x[6 + i] - (-ph.joint_ub[i % 3]); m_j = fmin(m_j, g); rcx.get(ph.reb_joint, 60 + i, dl, ep); c_j += ep * reb_value(g, dl); }
      }
      if (mh) { const double g = x[2] - ph.h_min; m_h = fmin(m_h, g); rcx.get(ph.reb_minheight, 72, dl, ep); c_h += ep * reb_value(g, dl); }
      bool any = false;
      for (int f = 0; f < 4; ++f)
        if (ph.contact[f] > 0) {
          any = true;
          const double fx = y[3 * f], fy = y[3 * f + 1], fz = y[3 * f + 2], mu = ph.mu;
          const double g[5] = {fz, -fx + mu * fz, fx + mu * fz, -fy + mu * fz, fy + mu * fz};
          for (int i = 0; i < 5; ++i) { m_g = fmin(m_g, g[i]); rcx.get(ph.reb_grf, 73 + 5 * f + i, dl, ep); c_g += ep * reb_value(g[i], dl); }
        }
      ming = fmin(fmin(fmin(m_t, m_j), fmin(m_h, m_g)), m_v);
      if (reb) { l += ph.dt * c_t; if (ph.joint_speed_limit) l += ph.dt * c_v; if (jl) l += ph.dt * c_j; if (mh) l += ph.dt * c_h; if (any) l += ph.dt * c_g; }
    }
    return l;
  }

  __device__ __noinline__ static void roll(const PhaseDev& ph, const double* rec, const double* x, const double* u, double* xn, double* y,
                              bool reb, double& l, double& ming, const RebCtx& rcx) {
    WBScratch s;
    forward(ph, x, u, s);
    for (int i = 0; i < 18; ++i) { xn[i] = x[i] + x[18 + i] * ph.dt; xn[18 + i] = x[18 + i] + s.qdd[i] * ph.dt; }
    for (int i = 0; i < 12; ++i) y[i] = s.grf[i];
    l = running_cost_k(ph, rec, x, u, y, s.pf, s.vf, reb, ming, rcx);
  }

  __device__ static void feet(const double* x, double* pf, double* vf, double* J) {
    for (int i = 0; i < 216; ++i) J[i] = 0;
    wbg_feet(x, x + 18, pf, vf, J);
  }

  __device__ __noinline__ static double terminal_cost(const PhaseDev& ph, const double* rec, const double* x) {
    double pf[12], vf[12], J[216];
    feet(x, pf, vf, J);
    double s = 0;
    for (int i = 0; i < 36; ++i) { const double dx = x[i] - rec[CAFE_REF_XR + i]; s += dx * ph.qf[i] * dx; }
    const double phi = s * 0.5;
    double reg = 0;
    for (int f = 0; f < 4; ++f) {
      if (!(rec[CAFE_REF_CONTACT + f] > 0)) continue;
      double q2 = 0;
      for (int a = 0; a < 3; ++a) { const double d = (pf[3 * f + a] - x[a]) - (rec[CAFE_REF_PF + 3 * f + a] - rec[CAFE_REF_PCOM + a]); q2 += d * ph.w_footreg[a] * d; }
      reg += .5 * q2;
    }
    double td = 0;
    for (int i = 0; i < ph.n_td; ++i) { const double dv = vf[3 * ph.td_foot[i] + 2]; td += .5 * dv * ph.w_tdvel[2] * dv; }
    return phi + reg + td;
  }
  __device__ __noinline__ static void terminal_constraints(const PhaseDev& ph, const double* x, double* hv) {
    double pf[12], vf[12], J[216];
    feet(x, pf, vf, J);
    for (int i = 0; i < ph.n_td; ++i) hv[i] = pf[3 * ph.td_foot[i] + 2] - ph.ground_height;
  }

  // impact (KKTImpact, WBM.cpp:427-456): v+ = v + Minv Jc^T Lambda, Lambda = -(Jc Minv Jc^T)^-1 Jc v
  __device__ static void impact(const PhaseDev& ph, const double* x, WBScratch& s, double* vpost) {
    int st[4];
    for (int f = 0; f < 4; ++f) st[f] = (ph.contact[f] == 0 && ph.next_contact[f] == 1) ? 1 : 0;
    active_rows(st, s);
    double xq[36];
    for (int i = 0; i < 18; ++i) { xq[i] = x[i]; xq[18 + i] = 0.0; }
    kkt_setup(xq, s, 0.0);
    const int nr = s.nr;
    for (int c = 0; c < nr; ++c) { double d = 0; for (int i = 0; i < 18; ++i) d += s.J[s.rows[c] + 12 * i] * x[18 + i]; s.lam[c] = -d; }
    fwd_subst(s.Ls, nr, 12, s.lam);
    bwd_subst(s.Ls, nr, 12, s.lam);
    double t[18];
    for (int i = 0; i < 18; ++i) { double d = 0; for (int c = 0; c < nr; ++c) d += s.Y[i + 18 * c] * s.lam[c]; t[i] = d; }
    bwd_subst(s.L, 18, 18, t);
    for (int i = 0; i < 18; ++i) vpost[i] = x[18 + i] + t[i];
  }
  __device__ static bool any_touchdown(const PhaseDev& ph) {
    for (int f = 0; f < 4; ++f) if (ph.next_contact[f] - ph.contact[f] == 1) return true;
    return false;
  }
  __device__ __noinline__ static void resetmap(const PhaseDev& ph, const double* x, double* xn) {
    double full[36];
    for (int i = 0; i < 36; ++i) full[i] = x[i];
    if (any_touchdown(ph)) { WBScratch s; impact(ph, x, s, full + 18); }
    if (ph.n_next == 12) { for (int i = 0; i < 6; ++i) { xn[i] = full[i]; xn[6 + i] = full[18 + i]; } }
    else for (int i = 0; i < 36; ++i) xn[i] = full[i];
  }

  // (LQ data of a running knot: k_wb_derivs + k_wb_lq)
  __device__ static double lq_knot(const PhaseDev&, int, int, int, const double*, const double*, const double*, const double*, bool) { return 0.0; }

  // ---- terminal cost partials (+AL) and the reset-map Jacobian
  __device__ __noinline__ static void lq_terminal(const PhaseDev& ph, int ldb, int b, const double* rec, const double* x, bool al) {
    double pf[12], vf[12], J[216], dvq[216];
    feet(x, pf, vf, J);
    for (int i = 0; i < 216; ++i) dvq[i] = 0;
    if (ph.n_td > 0) wbg_footvel_partial(x, x + 18, dvq);
    double phix[36];
    for (int i = 0; i < 36; ++i) phix[i] = ph.qf[i] * (x[i] - rec[CAFE_REF_XR + i]);
    double dposw[12];
    for (int f = 0; f < 4; ++f)
      for (int a = 0; a < 3; ++a) dposw[3 * f + a] = ph.w_footreg[a] * ((pf[3 * f + a] - x[a]) - (rec[CAFE_REF_PF + 3 * f + a] - rec[CAFE_REF_PCOM + a]));
    for (int f = 0; f < 4; ++f) {
      if (!(rec[CAFE_REF_CONTACT + f] > 0)) continue;
      for (int i = 3; i < 18; ++i) { double g = 0; for (int a = 0; a < 3; ++a) g += J[3 * f + a + 12 * i] * dposw[3 * f + a]; phix[i] += 2 * g; }
    }
    for (int t = 0; t < ph.n_td; ++t) {
      const int f = ph.td_foot[t];
      const double dv = vf[3 * f + 2];
      for (int i = 0; i < 36; ++i) phix[i] += ((i < 18) ? dvq[3 * f + 2 + 12 * i] : J[3 * f + 2 + 12 * (i - 18)]) * ph.w_tdvel[2] * dv;
    }
    double cg[4] = {0, 0, 0, 0}, ch[4] = {0, 0, 0, 0};
    const int ntd = al ? ph.n_td : 0;
    for (int t = 0; t < ntd; ++t) {
      const double hval = pf[3 * ph.td_foot[t] + 2] - ph.ground_height;
      const double sigma = ph.al_sigma[(size_t)t * ldb + b], lambda = ph.al_lambda[(size_t)t * ldb + b];
      cg[t] = sigma * hval + lambda;
      ch[t] = sigma * (1 + hval) + lambda;
      for (int i = 0; i < 18; ++i) phix[i] += cg[t] * J[3 * ph.td_foot[t] + 2 + 12 * i];
    }
    for (int i = 0; i < 36; ++i) ph.Phix[(size_t)i * ldb + b] = phix[i];
    for (int j = 0; j < 36; ++j)
      for (int i = 0; i < 36; ++i) {
        double v = (i == j) ? ph.qf[i] : 0.0;
        if (i >= 3 && i < 18 && j >= 3 && j < 18)
          for (int f = 0; f < 4; ++f) {
            if (!(rec[CAFE_REF_CONTACT + f] > 0)) continue;
            double hh = 0;
            for (int a = 0; a < 3; ++a) hh += J[3 * f + a + 12 * i] * ph.w_footreg[a] * J[3 * f + a + 12 * j];
            v += 2 * hh;
          }
        for (int t = 0; t < ph.n_td; ++t) {
          const int f = ph.td_foot[t];
          const double ji = (i < 18) ? dvq[3 * f + 2 + 12 * i] : J[3 * f + 2 + 12 * (i - 18)];
          const double jj = (j < 18) ? dvq[3 * f + 2 + 12 * j] : J[3 * f + 2 + 12 * (j - 18)];
          v += ji * ph.w_tdvel[2] * jj;
        }
        if (i < 18 && j < 18) for (int t = 0; t < ntd; ++t) v += ch[t] * J[3 * ph.td_foot[t] + 2 + 12 * i] * J[3 * ph.td_foot[t] + 2 + 12 * j];
        ph.Phixx[(size_t)(i + 36 * j) * ldb + b] = v;
      }
    if (!ph.has_next) return;
    // ---- reset-map Jacobian: impact partial (WBM.cpp:508-543) then the WB->SRB projection (MHPCReset.cpp:31-53)
    const int nn = ph.n_next;
    const bool proj = (nn == 12);
    if (!any_touchdown(ph)) {
      for (int c = 0; c < 36; ++c)
        for (int r = 0; r < nn; ++r) {
          const int src = proj ? (r < 6 ? r : 12 + r) : r;
          ph.Px[(size_t)(r + nn * c) * ldb + b] = (src == c) ? 1.0 : 0.0;
        }
      return;
    }
    WBScratch s;
    double vpost[18];
    impact(ph, x, s, vpost);
    const int nr = s.nr;
    // M itself is needed for dv+/dv = Kinv_tl M; rebuild it from its factor: M = L L^T (column by column below)
    double Rq[324], dvq2[216], dv[18], zero18[18], imp[12];
    for (int i = 0; i < 18; ++i) { dv[i] = vpost[i] - x[18 + i]; zero18[i] = 0; }
    for (int i = 0; i < 324; ++i) Rq[i] = 0;
    {
      // d(M(q) dv + g(q))/dq - dg/dq  (computeRNEADerivatives(q, 0, v+ - v) minus computeGeneralizedGravityDerivatives)
      double tmp[324];
      for (int i = 0; i < 324; ++i) tmp[i] = 0;
      wbg_rnea_derivs(x, zero18, dv, Rq, tmp, 1);
      for (int i = 0; i < 324; ++i) tmp[i] = 0;
      wbg_grav_derivs(x, tmp);
      for (int i = 0; i < 324; ++i) Rq[i] -= tmp[i];
    }
    // impulse scatter with the reference's segment<3>(i) indexing (WBM.cpp:454)
    for (int i = 0; i < 12; ++i) imp[i] = 0;
    {
      int ci = 0;
      for (int f = 0; f < 4; ++f) if (ph.contact[f] == 0 && ph.next_contact[f] == 1) { for (int r = 0; r < 3; ++r) imp[3 * f + r] = s.lam[ci + r]; ++ci; }
    }
    for (int i = 0; i < 216; ++i) dvq2[i] = 0;
    {
      double xv[36];
      for (int i = 0; i < 18; ++i) { xv[i] = x[i]; xv[18 + i] = vpost[i]; }
      double da1[216], da2[216], djtf[324];
      for (int i = 0; i < 216; ++i) { da1[i] = 0; da2[i] = 0; }
      for (int i = 0; i < 324; ++i) djtf[i] = 0;
      wbg_kin_partials(xv, xv + 18, zero18, imp, dvq2, da1, da2, djtf, 1);
      for (int i = 0; i < 324; ++i) Rq[i] -= djtf[i];
    }
    for (int col = 0; col < 36; ++col) {
      double r[18], w[12];
      if (col < 18) for (int i = 0; i < 18; ++i) r[i] = Rq[i + 18 * col];
      else {
        // column of M: M e_c = L (L^T e_c)
        const int c = col - 18;
        double t[18];
        for (int i = 0; i < 18; ++i) t[i] = (i <= c) ? s.L[c + 18 * i] : 0.0;  // (L^T e_c)_i = L[c][i]
        for (int i = 0; i < 18; ++i) { double d = 0; for (int kk = 0; kk <= i; ++kk) d += s.L[i + 18 * kk] * t[kk]; r[i] = -d; }
      }
      fwd_subst(s.L, 18, 18, r);
      for (int c = 0; c < nr; ++c) {
        double d = 0;
        for (int i = 0; i < 18; ++i) d += s.Y[i + 18 * c] * r[i];
        w[c] = d - ((col < 18) ? dvq2[s.rows[c] + 12 * col] : 0.0);
      }
      if (nr > 0) { fwd_subst(s.Ls, nr, 12, w); bwd_subst(s.Ls, nr, 12, w); }
      for (int i = 0; i < 18; ++i) { double d = -r[i]; for (int c = 0; c < nr; ++c) d += s.Y[i + 18 * c] * w[c]; r[i] = d; }
      bwd_subst(s.L, 18, 18, r);  // column of dv+/dq (col < 18) or of Kinv_tl M (col >= 18)
      // rows of dP_dx: [I 0; dv+/dq dv+/dv]
      for (int rr = 0; rr < nn; ++rr) {
        const int src = proj ? (rr < 6 ? rr : 12 + rr) : rr;
        double v;
        if (src < 18) v = (src == col) ? 1.0 : 0.0;
        else v = r[src - 18];
        ph.Px[(size_t)(rr + nn * col) * ldb + b] = v;
      }
    }
  }
};

}  // namespace cafe_dev
