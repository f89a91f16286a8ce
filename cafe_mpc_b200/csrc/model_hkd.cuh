// model_hkd.cuh — device-side HKD (hybrid kinodynamic) model, costs, constraints and reset map.
// One thread evaluates one (problem, knot); all state lives in registers; the model expressions are
// the re-emitted CasADi graphs of gen/hkd_gen.h.
//
// Reference behaviour followed (file:line under /root/reference):
//   HKD::Model::dynamics / dynamics_partial     HKDMPC/HKD-TrajOpt/HKDModel.h:33-61
//   HKDReset::resetmap / resetmap_partial       HKDMPC/HKD-TrajOpt/HKDReset.h:41-136
//   GRFConstraint / TouchDownConstraint         HKDMPC/HKD-TrajOpt/HKDConstraints.cpp:7-171
//   HKDTrackingCost / QuadraticTrackingCost     HKDMPC/HKD-TrajOpt/HKDCost.h:8-38, HSDDPSolver/source/SinglePhaseInterface.cpp:21-133
//   HKDFootPlaceReg                             HKDMPC/HKD-TrajOpt/HKDCost.cpp:5-66
//   ReB / AL folding                            HSDDPSolver/header/ConstraintsBase.h:230-289,400-425,
//                                               HSDDPSolver/source/SinglePhase.cpp:394-450
#pragma once
#include "device_types.cuh"
#include "gen/hkd_gen.h"

namespace cafe_dev {

// relaxed barrier B(g), B'(g), B''(g)   (ConstraintsBase.h:230-289)
__device__ __forceinline__ double reb_value(double g, double delta) {
  if (g > delta) return -log(g);
  double z = (g - 2 * delta) / delta;
  return .5 * (z * z - 1) - log(delta);
}
__device__ __forceinline__ void reb_derivs(double g, double delta, double& bd, double& bdd) {
  if (g > delta) { bd = -1.0 / g; bdd = 1.0 / (g * g); }
  else { bd = (g - 2 * delta) / delta / delta; bdd = 1.0 / (delta * delta); }
}

struct HKDModel {
  static constexpr int N = 24, M = 24, PY = 0;
  static constexpr bool COOP = false;

  __device__ static void foot_position(int leg, const double* x, double* pf) {
    const double pos[3] = {x[3], x[4], x[5]}, eul[3] = {x[0], x[1], x[2]};
    auto st = [&](int i, double v) { pf[i] = v; };
    switch (leg) {
      case 0: { const double ql[3] = {x[12], x[13], x[14]}; cafe_gen_hkd::foot_position_1(pos, eul, ql, (const double*)nullptr, st); break; }
      case 1: { const double ql[3] = {x[15], x[16], x[17]}; cafe_gen_hkd::foot_position_2(pos, eul, ql, (const double*)nullptr, st); break; }
      case 2: { const double ql[3] = {x[18], x[19], x[20]}; cafe_gen_hkd::foot_position_3(pos, eul, ql, (const double*)nullptr, st); break; }
      default: { const double ql[3] = {x[21], x[22], x[23]}; cafe_gen_hkd::foot_position_4(pos, eul, ql, (const double*)nullptr, st); break; }
    }
  }
  // J: 3 x 18 column-major, columns [pos(3), eul(3), qJ(12)]; caller zeroes J
  __device__ static void foot_jacobian(int leg, const double* x, double* J) {
    const double pos[3] = {x[3], x[4], x[5]}, eul[3] = {x[0], x[1], x[2]};
    auto st = [&](int i, double v) { J[i] = v; };
    switch (leg) {
      case 0: { const double ql[3] = {x[12], x[13], x[14]}; cafe_gen_hkd::foot_jacobian_1(pos, eul, ql, st); break; }
      case 1: { const double ql[3] = {x[15], x[16], x[17]}; cafe_gen_hkd::foot_jacobian_2(pos, eul, ql, st); break; }
      case 2: { const double ql[3] = {x[18], x[19], x[20]}; cafe_gen_hkd::foot_jacobian_3(pos, eul, ql, st); break; }
      default: { const double ql[3] = {x[21], x[22], x[23]}; cafe_gen_hkd::foot_jacobian_4(pos, eul, ql, st); break; }
    }
  }

  __device__ static void dynamics(const PhaseDev& ph, const double* rec, const double* x, const double* u, double* xn, double* y) {
    (void)rec; (void)y;
    const double dt = ph.dt;
    const double c[4] = {(double)ph.contact[0], (double)ph.contact[1], (double)ph.contact[2], (double)ph.contact[3]};
    cafe_gen_hkd::hkinodyn(x, u, &dt, c, [&](int i, double v) { xn[i] = v; });
  }

  // running cost l_k = tracking + foot-placement regulariser + dt * ReB(GRF); ming = min(0, g_i)
  __device__ static double running_cost(const PhaseDev& ph, const double* rec, const double* x, const double* u,
                                        const double* y, bool reb, double& ming, const RebCtx& rcx) {
    (void)y;
    double s = 0;
#pragma unroll
    for (int i = 0; i < 24; ++i) { double dx = x[i] - rec[CAFE_REF_XR + i]; s += dx * ph.q[i] * dx; }
    double l = 0.5 * s;
    s = 0;
#pragma unroll
    for (int i = 0; i < 24; ++i) { double du = u[i] - rec[CAFE_REF_UR + i]; s += du * ph.r[i] * du; }
    l += 0.5 * s;
    l *= ph.dt;
    double lf = 0;
#pragma unroll
    for (int leg = 0; leg < 4; ++leg)
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        double d = (x[12 + 3 * leg + a] - x[3 + a]) - (rec[CAFE_REF_PF + 3 * leg + a] - rec[CAFE_REF_PCOM + a]);
        lf += d * ((double)ph.contact[leg] * ph.w_footreg[a]) * d;
      }
    lf = .5 * lf;
    lf *= ph.dt;
    l = l + lf;
    ming = 0;
    double rc = 0;
    bool any = false;
#pragma unroll
    for (int leg = 0; leg < 4; ++leg) {
      if (ph.contact[leg] > 0) {
        any = true;
        const double fx = u[3 * leg], fy = u[3 * leg + 1], fz = u[3 * leg + 2], mu = ph.mu;
        const double g[5] = {fz, -fx + mu * fz, fx + mu * fz, -fy + mu * fz, fy + mu * fz};
#pragma unroll
        for (int i = 0; i < 5; ++i) { double dl, ep; rcx.get(ph.reb_grf, 5 * leg + i, dl, ep); ming = fmin(ming, g[i]); rc += ep * reb_value(g[i], dl); }
      }
    }
    if (reb && any) l += ph.dt * rc;
    return l;
  }

  // dynamics + running cost of one trial knot
  __device__ static void roll(const PhaseDev& ph, const double* rec, const double* x, const double* u, double* xn, double* y,
                              bool reb, double& l, double& ming, const RebCtx& rcx) {
    dynamics(ph, rec, x, u, xn, y);
    l = running_cost(ph, rec, x, u, y, reb, ming, rcx);
  }

  __device__ static double terminal_cost(const PhaseDev& ph, const double* rec, const double* x) {
    double s = 0;
#pragma unroll
    for (int i = 0; i < 24; ++i) { double dx = x[i] - rec[CAFE_REF_XR + i]; s += dx * ph.qf[i] * dx; }
    double phi = s * 0.5;
    double sf = 0;
#pragma unroll
    for (int leg = 0; leg < 4; ++leg)
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        double d = (x[12 + 3 * leg + a] - x[3 + a]) - (rec[CAFE_REF_PF + 3 * leg + a] - rec[CAFE_REF_PCOM + a]);
        sf += d * ((double)ph.contact[leg] * ph.w_footreg[a]) * d;
      }
    return phi + 10 * sf;
  }

  __device__ static void terminal_constraints(const PhaseDev& ph, const double* x, double* hv) {
    for (int i = 0; i < ph.n_td; ++i) { double pf[3]; foot_position(ph.td_foot[i], x, pf); hv[i] = pf[2] - ph.ground_height; }
  }

  __device__ static void resetmap(const PhaseDev& ph, const double* x, double* xn) {
#pragma unroll
    for (int i = 0; i < 24; ++i) xn[i] = x[i];
#pragma unroll
    for (int leg = 0; leg < 4; ++leg) {
      const int c = ph.contact[leg], cn = ph.next_contact[leg];
      if (c && !cn) { xn[12 + 3 * leg] = 0.0; xn[13 + 3 * leg] = -0.8; xn[14 + 3 * leg] = 1.7; }
      if (!c && cn) { double pf[3]; foot_position(leg, x, pf); xn[12 + 3 * leg] = pf[0]; xn[13 + 3 * leg] = pf[1]; xn[14 + 3 * leg] = 0.0 * pf[2]; }
    }
  }

  // LQ data of one running knot, written straight into the batch-major HBM arrays.
  // A, B, lxx, luu were zeroed at allocation; only the (static) non-zero pattern is rewritten.
  __device__ static double lq_knot(const PhaseDev& ph, int k, int ldb, int b, const double* rec, const double* x,
                                   const double* u, const double* y, bool reb) {
    const double dt = ph.dt;
    const RebCtx rcx = reb_ctx(ph, k, ldb, b);
    {
      const double c[4] = {(double)ph.contact[0], (double)ph.contact[1], (double)ph.contact[2], (double)ph.contact[3]};
      double* Ag = ph.A + gix(k, 576, 0, ldb, b);
      double* Bg = ph.Bm + gix(k, 576, 0, ldb, b);
      cafe_gen_hkd::hkinodyn_par(x, u, &dt, c, [&](int i, double v) { Ag[(size_t)i * ldb] = v; },
                                 [&](int i, double v) { Bg[(size_t)i * ldb] = v; });
    }
    double lx[24], lu[24];
    double* lxxg = ph.lxx + gix(k, 576, 0, ldb, b);
    double* luug = ph.luu + gix(k, 576, 0, ldb, b);
#pragma unroll
    for (int i = 0; i < 24; ++i) {
      lx[i] = dt * ph.q[i] * (x[i] - rec[CAFE_REF_XR + i]);
      lu[i] = dt * ph.r[i] * (u[i] - rec[CAFE_REF_UR + i]);
    }
    double dxx[24];
#pragma unroll
    for (int i = 0; i < 24; ++i) dxx[i] = dt * ph.q[i];
#pragma unroll
    for (int leg = 0; leg < 4; ++leg)
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const double c = (double)ph.contact[leg];
        const double w = c * ph.w_footreg[a];
        const double d = (x[12 + 3 * leg + a] - x[3 + a]) - (rec[CAFE_REF_PF + 3 * leg + a] - rec[CAFE_REF_PCOM + a]);
        const double gd = w * d;
        lx[3 + a] += dt * (-c) * gd;
        lx[12 + 3 * leg + a] += dt * c * gd;
        dxx[3 + a] += dt * c * w * c;
        dxx[12 + 3 * leg + a] += dt * c * w * c;
        const double off = dt * (-c) * w * c;
        lxxg[(size_t)((3 + a) + 24 * (12 + 3 * leg + a)) * ldb] = off;
        lxxg[(size_t)((12 + 3 * leg + a) + 24 * (3 + a)) * ldb] = off;
      }
#pragma unroll
    for (int i = 0; i < 24; ++i) lxxg[(size_t)(i + 24 * i) * ldb] = dxx[i];
    // GRF relaxed barrier: gradient and Gauss-Newton Hessian, one 3x3 block per leg
#pragma unroll
    for (int leg = 0; leg < 4; ++leg) {
      double gr[3] = {0, 0, 0}, hs[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
      if (reb && ph.contact[leg] > 0) {
        const double fx = u[3 * leg], fy = u[3 * leg + 1], fz = u[3 * leg + 2], mu = ph.mu;
        const double g[5] = {fz, -fx + mu * fz, fx + mu * fz, -fy + mu * fz, fy + mu * fz};
        const double Al[5][3] = {{0, 0, 1}, {-1, 0, mu}, {1, 0, mu}, {0, -1, mu}, {0, 1, mu}};
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          double bd, bdd, dl, ep;
          rcx.get(ph.reb_grf, 5 * leg + i, dl, ep);
          reb_derivs(g[i], dl, bd, bdd);
          const double e1 = ep * bd, e2 = ep * bdd;
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            gr[r] += e1 * Al[i][r];
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) hs[r][cc] += Al[i][r] * (e2 * Al[i][cc]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        lu[3 * leg + r] += dt * gr[r];
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          double v = dt * hs[r][cc];
          if (r == cc) v += dt * ph.r[3 * leg + r];
          luug[(size_t)((3 * leg + r) + 24 * (3 * leg + cc)) * ldb] = v;
        }
      }
    }
#pragma unroll
    for (int i = 12; i < 24; ++i) luug[(size_t)(i + 24 * i) * ldb] = dt * ph.r[i];
#pragma unroll
    for (int i = 0; i < 24; ++i) {
      ph.lx[gix(k, 24, i, ldb, b)] = lx[i];
      ph.lu[gix(k, 24, i, ldb, b)] = lu[i];
    }
    double ming;
    return running_cost(ph, rec, x, u, y, reb, ming, rcx);
  }

  // terminal cost partials (+ AL terms) and the reset-map Jacobian Px at X[h]
  __device__ static void lq_terminal(const PhaseDev& ph, int ldb, int b, const double* rec, const double* x, bool al) {
    double phix[24], hx[4][24], coefH[4];
    double base_diag[24];
#pragma unroll
    for (int i = 0; i < 24; ++i) { phix[i] = ph.qf[i] * (x[i] - rec[CAFE_REF_XR + i]); base_diag[i] = ph.qf[i]; }
    double off[4][2];
#pragma unroll
    for (int leg = 0; leg < 4; ++leg)
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const double c = (double)ph.contact[leg];
        const double w = c * ph.w_footreg[a];
        const double d = (x[12 + 3 * leg + a] - x[3 + a]) - (rec[CAFE_REF_PF + 3 * leg + a] - rec[CAFE_REF_PCOM + a]);
        const double gd = w * d;
        phix[3 + a] += 20 * (-c) * gd;
        phix[12 + 3 * leg + a] += 20 * c * gd;
        base_diag[3 + a] += 20 * c * w * c;
        base_diag[12 + 3 * leg + a] += 20 * c * w * c;
        off[leg][a] = 20 * (-c) * w * c;
      }
    const int ntd = al ? ph.n_td : 0;
    for (int i = 0; i < ntd; ++i) {
      double J[54];
      for (int j = 0; j < 54; ++j) J[j] = 0;
      foot_jacobian(ph.td_foot[i], x, J);
      for (int j = 0; j < 24; ++j) hx[i][j] = 0;
      for (int j = 0; j < 3; ++j) { hx[i][j] = J[2 + 3 * (3 + j)]; hx[i][3 + j] = J[2 + 3 * j]; }
      for (int j = 0; j < 12; ++j) hx[i][12 + j] = J[2 + 3 * (6 + j)];
      double pf[3];
      foot_position(ph.td_foot[i], x, pf);
      const double hval = pf[2] - ph.ground_height;
      const double sigma = ph.al_sigma[(size_t)i * ldb + b], lambda = ph.al_lambda[(size_t)i * ldb + b];
      const double cg = sigma * hval + lambda;
      coefH[i] = sigma * (1 + hval) + lambda;  // reference quirk kept (ConstraintsBase.h:423)
      for (int j = 0; j < 24; ++j) phix[j] += cg * hx[i][j];
    }
    for (int j = 0; j < 24; ++j) ph.Phix[(size_t)j * ldb + b] = phix[j];
    for (int j = 0; j < 24; ++j)
      for (int i = 0; i < 24; ++i) {
        double v = (i == j) ? base_diag[i] : 0.0;
        for (int leg = 0; leg < 4; ++leg)
          for (int a = 0; a < 2; ++a)
            if ((i == 3 + a && j == 12 + 3 * leg + a) || (j == 3 + a && i == 12 + 3 * leg + a)) v += off[leg][a];
        for (int c = 0; c < ntd; ++c) v += coefH[c] * hx[c][i] * hx[c][j];
        ph.Phixx[(size_t)(i + 24 * j) * ldb + b] = v;
      }
    if (ph.has_next) {
      for (int j = 0; j < 24; ++j)
        for (int i = 0; i < 24; ++i) ph.Px[(size_t)(i + 24 * j) * ldb + b] = (i == j) ? 1.0 : 0.0;
      for (int leg = 0; leg < 4; ++leg) {
        const int c = ph.contact[leg], cn = ph.next_contact[leg];
        if (c && !cn)
          for (int r = 0; r < 3; ++r) ph.Px[(size_t)((12 + 3 * leg + r) + 24 * (12 + 3 * leg + r)) * ldb + b] = 0.0;
        if (!c && cn) {
          double J[54];
          for (int j = 0; j < 54; ++j) J[j] = 0;
          foot_jacobian(leg, x, J);
          const double cmap[3] = {1, 1, 0};
          for (int r = 0; r < 3; ++r) {
            const int row = 12 + 3 * leg + r;
            for (int j = 0; j < 3; ++j) {
              ph.Px[(size_t)(row + 24 * j) * ldb + b] = cmap[r] * J[r + 3 * (3 + j)];
              ph.Px[(size_t)(row + 24 * (3 + j)) * ldb + b] = cmap[r] * J[r + 3 * j];
            }
            for (int j = 0; j < 12; ++j) ph.Px[(size_t)(row + 24 * (12 + j)) * ldb + b] = cmap[r] * J[r + 3 * (6 + j)];
          }
        }
      }
    }
  }
};

}  // namespace cafe_dev
