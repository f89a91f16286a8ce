// model_srb.cuh — device-side single-rigid-body (SRB) phase of the MHPC problem. One thread per (problem, knot).
// Reference behaviour followed (file:line under /root/reference):
//   SRBM::Model::dynamics / dynamics_partial   MHPC/MHPC-Trajopt/SRBM.h:43-93 (x+ = x + dt f, A = I + dt Ac, B = dt Bc)
//   foot positions and contact flags from the reference at the knot's time    MHPC/MHPC-Trajopt/MHPCFootStep.h:37-69
//   SRBTrackingCost (QuadraticTrackingCost)    MHPC/MHPC-Trajopt/MHPCCost.h:207-249, HSDDPSolver/source/SinglePhaseInterface.cpp:21-133
//   SRBMMinimumHeight                          MHPC/MHPC-Trajopt/MHPCConstraint.cpp:355-379
#pragma once
#include "device_types.cuh"
#include "gen/srb_gen.h"
#include "model_hkd.cuh"

namespace cafe_dev {

struct SRBModel {
  static constexpr int N = 12, M = 12, PY = 0;
  static constexpr bool COOP = false;

  __device__ static double running_cost(const PhaseDev& ph, const double* rec, const double* x, const double* u, bool reb, double& ming, const RebCtx& rcx) {
    double s = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) { const double dx = x[i] - rec[CAFE_REF_XR + i]; s += dx * ph.q[i] * dx; }
    double l = 0.5 * s;
    s = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) { const double du = u[i] - rec[CAFE_REF_UR + i]; s += du * ph.r[i] * du; }
    l += 0.5 * s;
    l *= ph.dt;
    const double g = x[2] - ph.h_min;
    ming = fmin(0.0, g);
    if (reb) { double dl, ep; rcx.get(ph.reb_minheight, 0, dl, ep); l += ph.dt * (ep * reb_value(g, dl)); }
    return l;
  }

  // dynamics + running cost of one trial knot
  __device__ static void roll(const PhaseDev& ph, const double* rec, const double* x, const double* u, double* xn, double* y,
                              bool reb, double& l, double& ming, const RebCtx& rcx) {
    (void)y;
    double xd[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) xd[i] = 0;
    cafe_gen_srb::srb_dynamics(x, u, rec + CAFE_REF_PF, rec + CAFE_REF_CONTACT, [&](int i, double v) { xd[i] = v; });
#pragma unroll
    for (int i = 0; i < 12; ++i) xn[i] = x[i] + xd[i] * ph.dt;
    l = running_cost(ph, rec, x, u, reb, ming, rcx);
  }

  __device__ static double terminal_cost(const PhaseDev& ph, const double* rec, const double* x) {
    double s = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) { const double dx = x[i] - rec[CAFE_REF_XR + i]; s += dx * ph.qf[i] * dx; }
    return s * 0.5;
  }
  __device__ static void terminal_constraints(const PhaseDev&, const double*, double*) {}
  __device__ static void resetmap(const PhaseDev&, const double* x, double* xn) {
#pragma unroll
    for (int i = 0; i < 12; ++i) xn[i] = x[i];
  }

  __device__ static double lq_knot(const PhaseDev& ph, int k, int ldb, int b, const double* rec, const double* x, const double* u,
                                   const double* y, bool reb) {
    (void)y;
    const double dt = ph.dt;
    double* Ag = ph.A + gix(k, 144, 0, ldb, b);
    double* Bg = ph.Bm + gix(k, 144, 0, ldb, b);
#pragma unroll
    for (int i = 0; i < 12; ++i) Ag[(size_t)(13 * i) * ldb] = 1.0;
    cafe_gen_srb::srb_dynamics_derivatives(x, u, rec + CAFE_REF_PF, rec + CAFE_REF_CONTACT,
                                           [&](int i, double v) { Ag[(size_t)i * ldb] = ((i % 13) == 0 ? 1.0 : 0.0) + v * dt; },
                                           [&](int i, double v) { Bg[(size_t)i * ldb] = v * dt; });
    double* lxxg = ph.lxx + gix(k, 144, 0, ldb, b);
    double* luug = ph.luu + gix(k, 144, 0, ldb, b);
    const double g = x[2] - ph.h_min;
    double bd = 0, bdd = 0, mh_delta, mh_eps;
    const RebCtx rcx = reb_ctx(ph, k, ldb, b);
    rcx.get(ph.reb_minheight, 0, mh_delta, mh_eps);
    if (reb) reb_derivs(g, mh_delta, bd, bdd);
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      double lx = dt * ph.q[i] * (x[i] - rec[CAFE_REF_XR + i]);
      double lxx = dt * ph.q[i];
      if (i == 2 && reb) { lx += dt * (mh_eps * bd); lxx += dt * (mh_eps * bdd); }
      ph.lx[gix(k, 12, i, ldb, b)] = lx;
      lxxg[(size_t)(13 * i) * ldb] = lxx;
      ph.lu[gix(k, 12, i, ldb, b)] = dt * ph.r[i] * (u[i] - rec[CAFE_REF_UR + i]);
      luug[(size_t)(13 * i) * ldb] = dt * ph.r[i];
    }
    double ming;
    return running_cost(ph, rec, x, u, reb, ming, rcx);
  }

  __device__ static void lq_terminal(const PhaseDev& ph, int ldb, int b, const double* rec, const double* x, bool al) {
    (void)al;
    for (int i = 0; i < 12; ++i) {
      ph.Phix[(size_t)i * ldb + b] = ph.qf[i] * (x[i] - rec[CAFE_REF_XR + i]);
      ph.Phixx[(size_t)(13 * i) * ldb + b] = ph.qf[i];
    }
  }
};

}  // namespace cafe_dev
