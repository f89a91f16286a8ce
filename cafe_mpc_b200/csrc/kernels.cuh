// kernels.cuh — the batched HS-DDP kernels (sm_100a, fp64).
//
//   k_roll    K-ROLL : hybrid rollout of every (problem, knot, step size) + constraint values + costs
//                      <= SinglePhase::hybrid_rollout / compute_cost      HSDDPSolver/source/SinglePhase.cpp:182-262
//                         MultiPhaseDDP::hybrid_rollout                   HSDDPSolver/source/MultiPhaseDDP.cpp:49-92
//   k_select  K-CTRL : merit / Armijo test over the step sizes, accept, exits, AL update, histories
//                      <= MultiPhaseDDP::line_search / solve              MultiPhaseDDP.cpp:95-133, :216-447
//   k_accept          : trial -> current (and -> nominal when accepted)   TrajectoryManagement.cpp:122-127
//   k_lq      K-LQ   : dynamics / cost / constraint partials, ReB + AL folding, per-knot cost
//                      <= SinglePhase::LQ_approximation, compute_cost     SinglePhase.cpp:236-320, :394-450
//   k_bwd     K-BWD + K-LIN : regularised Riccati sweep across phases with impact jumps, then the
//                      multiple-shooting linear rollout and the merit parameter
//                      <= SinglePhase::backward_sweep / linear_rollout    SinglePhase.cpp:145-178, :323-391
//                         MultiPhaseDDP::backward_sweep(_regularized) / linear_rollout / impact_aware_step
//                                                                         MultiPhaseDDP.cpp:12-42, :136-213, :499-503
//
// Mapping: k_roll / k_lq / k_accept run one thread per (problem, knot[, step size]) with the problem
// index fastest, so a warp is 32 problems at the same knot: identical instruction streams, coalesced
// HBM access, no divergence except on per-problem activity flags. k_bwd runs one CTA per group of PB
// problems, 32 workers per problem, sequential over the horizon, all matrices staged in shared memory.
#pragma once
#include "device_types.cuh"
#include "model_hkd.cuh"
#include "model_srb.cuh"
#include "model_wb.cuh"

namespace cafe_dev {

// ------------------------------------------------------------------------------------------- K-ROLL
template <class Model>
__device__ void roll_knot(const SolverDev& S, int pi, int k, int a, int b) {
  constexpr int N = Model::N, M = Model::M, PY = Model::PY;
  const PhaseDev& ph = S.ph[pi];
  const int ldb = S.ldb, h = ph.h;
  const double eps = S.eps[a];
  const double* rec = ph.ref + (size_t)k * CAFE_REF_W;
  const size_t aX = (size_t)a * (h + 1) * N * ldb, aU = (size_t)a * h * M * ldb, aY = (size_t)a * h * PY * ldb;
  const size_t aS = (size_t)a * (h + 1) * ldb;
  double x[N], dlt[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double xb = ph.Xbar[gix(k, N, i, ldb, b)];
    x[i] = xb + eps * ph.dX[gix(k, N, i, ldb, b)];
    dlt[i] = x[i] - xb;
    ph.Xt[aX + gix(k, N, i, ldb, b)] = x[i];
  }
  if (pi == 0 && k == 0) {  // Defect[0] of the first phase: Xsim[0] = x0
    double s = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) { const double d = S.x0[(size_t)i * ldb + b] - x[i]; ph.Dt[aX + gix(0, N, i, ldb, b)] = d; s += d * d; }
    S.c.feas0_t[(size_t)a * ldb + b] = s;
  }
  if (k < h) {
    double u[M];
#pragma unroll
    for (int i = 0; i < M; ++i) u[i] = 0;
    const double* Kg = ph.K + gix(k, M * N, 0, ldb, b);
#pragma unroll
    for (int j = 0; j < N; ++j) {
#pragma unroll
      for (int i = 0; i < M; ++i) u[i] += Kg[(size_t)(i + M * j) * ldb] * dlt[j];
    }
#pragma unroll
    for (int i = 0; i < M; ++i) {
      u[i] = ph.Ubar[gix(k, M, i, ldb, b)] + eps * ph.dU[gix(k, M, i, ldb, b)] + u[i];
      ph.Ut[aU + gix(k, M, i, ldb, b)] = u[i];
    }
    double xn[N], y[PY > 0 ? PY : 1];
    double l, ming;
    Model::roll(ph, rec, x, u, xn, y, S.opt.ReB_active != 0, l, ming);
    double nrm = 0, dsq = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      nrm += xn[i] * xn[i];
      const double xs = ph.Xbar[gix(k + 1, N, i, ldb, b)] + eps * ph.dX[gix(k + 1, N, i, ldb, b)];
      const double d = xn[i] - xs;
      ph.Dt[aX + gix(k + 1, N, i, ldb, b)] = d;
      dsq += d * d;
    }
#pragma unroll
    for (int i = 0; i < PY; ++i) ph.Yt[aY + gix(k, PY, i, ldb, b)] = y[i];
    ph.cost_t[aS + (size_t)k * ldb + b] = l;
    ph.feas_t[aS + (size_t)k * ldb + b] = dsq;
    ph.ming_t[aS + (size_t)k * ldb + b] = ming;
    if (sqrt(nrm) > 1e6) atomicOr(&ph.fail_t[(size_t)a * ldb + b], 1);
  } else {
    double phi = Model::terminal_cost(ph, rec, x);
    double hv[4] = {0, 0, 0, 0}, maxh = 0;
    if (ph.n_td > 0) {
      Model::terminal_constraints(ph, x, hv);
      for (int i = 0; i < ph.n_td; ++i) {
        maxh = fmax(maxh, fabs(hv[i]));
        ph.ht[((size_t)a * 4 + i) * ldb + b] = hv[i];
        if (S.opt.AL_active) {
          const double sg = ph.al_sigma[(size_t)i * ldb + b], lm = ph.al_lambda[(size_t)i * ldb + b];
          double c = 0;
          c += 0.5 * sg * hv[i] * hv[i];
          c += lm * hv[i];
          phi += c;
        }
      }
    }
    ph.maxh_t[(size_t)a * ldb + b] = maxh;
    ph.cost_t[aS + (size_t)h * ldb + b] = phi;
    ph.ming_t[aS + (size_t)h * ldb + b] = 0;
    double dsq = 0;
    if (ph.has_next) {
      const PhaseDev& nx = S.ph[pi + 1];
      double xr[CAFE_MAX_N];
      Model::resetmap(ph, x, xr);
      const size_t aXn = (size_t)a * (nx.h + 1) * nx.n * ldb;
      for (int i = 0; i < nx.n; ++i) {
        const double xs = nx.Xbar[gix(0, nx.n, i, ldb, b)] + eps * nx.dX[gix(0, nx.n, i, ldb, b)];
        const double d = xr[i] - xs;
        nx.Dt[aXn + gix(0, nx.n, i, ldb, b)] = d;
        dsq += d * d;
      }
    }
    ph.feas_t[aS + (size_t)h * ldb + b] = dsq;
  }
}

// step sizes [a0, a1) of the ladder; problems whose line search already succeeded are skipped
__global__ void __launch_bounds__(128) k_roll(const SolverDev* __restrict__ Sp, int a0, int a1) {
  const SolverDev& S = *Sp;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = (int)(t % S.ldb);
  const long long r = t / S.ldb;
  const int gk = (int)(r % S.n_knots), a = a0 + (int)(r / S.n_knots);
  if (a >= a1 || b >= S.B) return;
  if (!S.c.active[b] || !S.c.do_ls[b] || S.c.ls_found[b]) return;
  const int pi = S.knot_phase[gk], k = S.knot_k[gk];
  switch (S.ph[pi].model) {
    case CAFE_MODEL_HKD: roll_knot<HKDModel>(S, pi, k, a, b); break;
    case CAFE_MODEL_WB: roll_knot<WBModel>(S, pi, k, a, b); break;
    case CAFE_MODEL_SRB: roll_knot<SRBModel>(S, pi, k, a, b); break;
    default: break;
  }
}

// --------------------------------------------------------------------------------------------- K-LQ
template <class Model>
__device__ void lq_knot_generic(const SolverDev& S, int pi, int k, int b) {
  constexpr int N = Model::N, M = Model::M, PY = Model::PY;
  const PhaseDev& ph = S.ph[pi];
  const int ldb = S.ldb, h = ph.h;
  const double* rec = ph.ref + (size_t)k * CAFE_REF_W;
  double x[N];
  double dsq = 0;
#pragma unroll
  for (int i = 0; i < N; ++i) { x[i] = ph.X[gix(k, N, i, ldb, b)]; const double d = ph.Defect[gix(k, N, i, ldb, b)]; dsq += d * d; }
  ph.dsq[(size_t)k * ldb + b] = dsq;
  if (k < h) {
    double u[M], y[PY > 0 ? PY : 1];
#pragma unroll
    for (int i = 0; i < M; ++i) u[i] = ph.U[gix(k, M, i, ldb, b)];
#pragma unroll
    for (int i = 0; i < PY; ++i) y[i] = ph.Y[gix(k, PY, i, ldb, b)];
    ph.lk[(size_t)k * ldb + b] = Model::lq_knot(ph, k, ldb, b, rec, x, u, y, S.opt.ReB_active != 0);
  } else {
    double phi = Model::terminal_cost(ph, rec, x);
    if (ph.n_td > 0 && S.opt.AL_active) {
      double hv[4];
      Model::terminal_constraints(ph, x, hv);
      for (int i = 0; i < ph.n_td; ++i) {
        const double sg = ph.al_sigma[(size_t)i * ldb + b], lm = ph.al_lambda[(size_t)i * ldb + b];
        double c = 0;
        c += 0.5 * sg * hv[i] * hv[i];
        c += lm * hv[i];
        phi += c;
      }
    }
    ph.lk[(size_t)h * ldb + b] = phi;
    Model::lq_terminal(ph, ldb, b, rec, x, S.opt.AL_active != 0);
  }
}

__global__ void __launch_bounds__(128) k_lq(const SolverDev* __restrict__ Sp) {
  const SolverDev& S = *Sp;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = (int)(t % S.ldb);
  const int gk = (int)(t / S.ldb);
  if (gk >= S.n_knots || b >= S.B) return;
  if (!S.c.active[b]) return;
  const int pi = S.knot_phase[gk], k = S.knot_k[gk];
  switch (S.ph[pi].model) {
    case CAFE_MODEL_HKD: lq_knot_generic<HKDModel>(S, pi, k, b); break;
    case CAFE_MODEL_WB: lq_knot_generic<WBModel>(S, pi, k, b); break;
    case CAFE_MODEL_SRB: lq_knot_generic<SRBModel>(S, pi, k, b); break;
    default: break;
  }
}

// ------------------------------------------------------------------------------- K-LQ, whole-body dense part
// KKT sensitivities of one WB phase (WBM.cpp:459-505 without forming Kinv): CTA = 4 consecutive problems at one knot, one warp
// per problem, lane = column z of [q(18) v(18) tau(12)]; chol(M), Y = L^-1 Jc^T, chol(S) are read from shared memory at
// warp-uniform addresses (broadcast), the column's right-hand side lives in registers.
//   dlambda/dz = S^-1 (Jc Minv R - a),  dqdd/dz = -Minv (R - Jc^T dlambda/dz);  A = I + dt Ac, B = dt Bc, C/D = dGRF/d(x,u)
// shared memory per problem (doubles): Lt 324 (row-major L, reciprocal diagonal) | Y 216 (18 x 12, ld 18) | Yt 216 (12 x 18, ld 12)
//                                     | Lst 144 (row-major chol(S), reciprocal diagonal) | R 19x36 | a 13x36
#define CAFE_KKT_SM (324 + 216 + 216 + 144 + 684 + 468)
template <int NR>
__global__ void __launch_bounds__(128, 3) k_lq_wb_dense(const SolverDev* __restrict__ Sp, int pi) {
  const SolverDev& S = *Sp;
  const PhaseDev& ph = S.ph[pi];
  extern __shared__ double sm[];
  const int ldb = S.ldb, k = blockIdx.y, b0 = blockIdx.x * 4;
  {
    const int p = threadIdx.x & 3, b = b0 + p;
    double* dst = sm + p * CAFE_KKT_SM;
    const double* src = ph.kkt + gix(k, CAFE_KKT_PACK, 0, ldb, b);
    if (b < S.B && S.c.active[b]) {
      const double bg2 = 2.0 * ph.BG_alpha;
      int rowsA[12];
      { int j = 0; for (int f = 0; f < 4; ++f) if (ph.contact[f] > 0) for (int r = 0; r < 3; ++r) rowsA[j++] = 3 * f + r; for (; j < 12; ++j) rowsA[j] = 0; }
      for (int e = threadIdx.x >> 2; e < 684; e += 32) {  // factors
        const double v = src[(size_t)e * ldb];
        if (e >= CAFE_KKT_LS) { const int idx = e - CAFE_KKT_LS; const int i = idx % 12, j = idx / 12; dst[756 + j + 12 * i] = (i == j) ? 1.0 / v : v; }
        else if (e >= CAFE_KKT_Y) { const int idx = e - CAFE_KKT_Y; const int i = idx % 18, c = idx / 18; dst[324 + idx] = v; dst[540 + c + 12 * i] = v; }
        else { const int i = e % 18, j = e / 18; dst[j + 18 * i] = (i == j) ? 1.0 / v : v; }
      }
      for (int e = threadIdx.x >> 2; e < 648; e += 32) {  // R = [dtau_dq - d(J^T F)/dq | dtau_dv]
        const int i = e % 18, col = e / 18;
        double v = src[(size_t)(CAFE_KKT_RQ + e) * ldb];
        if (col < 18) v -= src[(size_t)(CAFE_KKT_JTF + e) * ldb];
        dst[900 + i + 19 * col] = v;
      }
      for (int e = threadIdx.x >> 2; e < NR * 36; e += 32) {  // a = [da/dq + 2 BG dv/dq | da/dv + 2 BG J] on the active rows
        const int c = e % (NR > 0 ? NR : 1), col = e / (NR > 0 ? NR : 1);
        const int row = rowsA[c];
        double v;
        if (col < 18) v = src[(size_t)(CAFE_KKT_AQ + row + 12 * col) * ldb] + bg2 * src[(size_t)(CAFE_KKT_DVQ + row + 12 * col) * ldb];
        else v = src[(size_t)(CAFE_KKT_AV + row + 12 * (col - 18)) * ldb] + bg2 * src[(size_t)(CAFE_KKT_J + row + 12 * (col - 18)) * ldb];
        dst[1584 + c + 13 * col] = v;
      }
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, b = b0 + warp;
  if (b >= S.B || !S.c.active[b]) return;
  const double* sLt = sm + warp * CAFE_KKT_SM;  // Lt[k + 18 i] = L(i,k), Lt[i + 18 i] = 1 / L(i,i)
  const double* sY = sLt + 324;                 // Y[i + 18 c]
  const double* sYt = sLt + 540;                // Yt[c + 12 i]
  const double* sLst = sLt + 756;               // Lst[k + 12 i] = Ls(i,k), reciprocal diagonal
  const double* sR = sLt + 900;
  const double* sa = sLt + 1584;
  const double dt = ph.dt;
  int foot[4] = {0, 0, 0, 0};
  { int j = 0; for (int f = 0; f < 4; ++f) if (ph.contact[f] > 0) foot[j++] = f; }
  double* Ag = ph.A + gix(k, 1296, 0, ldb, b);
  double* Bg = ph.Bm + gix(k, 432, 0, ldb, b);
  double* Cg = ph.C + gix(k, 432, 0, ldb, b);
  double* Dg = ph.D + gix(k, 144, 0, ldb, b);
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    const int col = pass * 32 + lane;
    if (col >= 48) break;
    double r[18], w[NR > 0 ? NR : 1];
    if (col < 36) {
#pragma unroll
      for (int i = 0; i < 18; ++i) r[i] = sR[i + 19 * col];
    } else {
#pragma unroll
      for (int i = 0; i < 18; ++i) r[i] = (i == 6 + (col - 36)) ? -1.0 : 0.0;
    }
    // r <- L^-1 R  (row i of L is contiguous in Lt)
#pragma unroll
    for (int i = 0; i < 18; ++i) {
      double s = r[i];
#pragma unroll
      for (int kk = 0; kk < i; ++kk) s -= sLt[kk + 18 * i] * r[kk];
      r[i] = s * sLt[i + 18 * i];
    }
    if constexpr (NR > 0) {
#pragma unroll
      for (int c = 0; c < NR; ++c) {
        double d = 0;
#pragma unroll
        for (int i = 0; i < 18; ++i) d += sY[i + 18 * c] * r[i];
        w[c] = d - ((col < 36) ? sa[c + 13 * col] : 0.0);
      }
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        double s = w[i];
#pragma unroll
        for (int kk = 0; kk < i; ++kk) s -= sLst[kk + 12 * i] * w[kk];
        w[i] = s * sLst[i + 12 * i];
      }
      // back substitution with Ls^T: column i of Ls is needed; process as saxpy updates so that rows stay contiguous
#pragma unroll
      for (int i = NR - 1; i >= 0; --i) {
        w[i] *= sLst[i + 12 * i];
#pragma unroll
        for (int kk = 0; kk < i; ++kk) w[kk] -= sLst[kk + 12 * i] * w[i];
      }
    }
#pragma unroll
    for (int i = 0; i < 18; ++i) {
      double d = -r[i];
      if constexpr (NR > 0) {
#pragma unroll
        for (int c = 0; c < NR; ++c) d += sYt[c + 12 * i] * w[c];
      }
      r[i] = d;
    }
    // r <- L^-T r, saxpy form (row i of L contiguous)
#pragma unroll
    for (int i = 17; i >= 0; --i) {
      r[i] *= sLt[i + 18 * i];
#pragma unroll
      for (int kk = 0; kk < i; ++kk) r[kk] -= sLt[kk + 18 * i] * r[i];
    }
    if (col < 36) {
#pragma unroll
      for (int i = 0; i < 18; ++i) Ag[(size_t)((18 + i) + 36 * col) * ldb] = ((col == 18 + i) ? 1.0 : 0.0) + r[i] * dt;
      if constexpr (NR > 0) {
#pragma unroll
        for (int c = 0; c < NR; ++c) Cg[(size_t)((3 * foot[c / 3] + c % 3) + 12 * col) * ldb] = w[c];
      }
    } else {
#pragma unroll
      for (int i = 0; i < 18; ++i) Bg[(size_t)((18 + i) + 36 * (col - 36)) * ldb] = r[i] * dt;
      if constexpr (NR > 0) {
#pragma unroll
        for (int c = 0; c < NR; ++c) Dg[(size_t)((3 * foot[c / 3] + c % 3) + 12 * (col - 36)) * ldb] = w[c];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ K-ACCEPT
__global__ void k_accept(const SolverDev* __restrict__ Sp) {
  const SolverDev& S = *Sp;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = (int)(t % S.ldb);
  const int gk = (int)(t / S.ldb);
  if (gk >= S.n_knots || b >= S.B) return;
  const int a = S.c.sel[b];
  if (a < 0) return;
  const bool acc = S.c.accepted[b] != 0;
  const int pi = S.knot_phase[gk], k = S.knot_k[gk];
  const PhaseDev& ph = S.ph[pi];
  const int ldb = S.ldb, h = ph.h, n = ph.n, m = ph.m, p = ph.p;
  const size_t aX = (size_t)a * (h + 1) * n * ldb, aU = (size_t)a * h * m * ldb, aY = (size_t)a * h * p * ldb;
  for (int i = 0; i < n; ++i) {
    const size_t ix = gix(k, n, i, ldb, b);
    const double v = ph.Xt[aX + ix];
    ph.X[ix] = v;
    ph.Defect[ix] = ph.Dt[aX + ix];
    if (acc) ph.Xbar[ix] = v;
  }
  if (k < h) {
    for (int i = 0; i < m; ++i) {
      const size_t ix = gix(k, m, i, ldb, b);
      const double v = ph.Ut[aU + ix];
      ph.U[ix] = v;
      if (acc) ph.Ubar[ix] = v;
    }
    for (int i = 0; i < p; ++i) { const size_t ix = gix(k, p, i, ldb, b); ph.Y[ix] = ph.Yt[aY + ix]; }
  }
}

// ------------------------------------------------------------------------------------------ K-SELECT
__device__ __forceinline__ void push_hist(const SolverDev& S, int b, double cost, double feas, double mt, double mp) {
  int nh = S.c.n_hist[b];
  if (nh < CAFE_HIST_CAP) {
    double* hp = S.c.hist + ((size_t)nh * 4) * S.ldb + b;
    hp[0] = cost; hp[(size_t)S.ldb] = feas; hp[(size_t)2 * S.ldb] = mt; hp[(size_t)3 * S.ldb] = mp;
  }
  S.c.n_hist[b] = nh + 1;
}

// sum of the per-knot partials of trial a in the reference's order: per phase (sum_k l_k) + Phi, then over phases
__device__ void reduce_trial(const SolverDev& S, int a, int b, double& cost, double& feas, double& max_t, double& max_p, int& fail) {
  const int ldb = S.ldb;
  cost = 0;
  double fs = 0;
  max_t = 0; max_p = 0; fail = 0;
  for (int pi = 0; pi < S.n_phases; ++pi) {
    const PhaseDev& ph = S.ph[pi];
    const size_t aS = (size_t)a * (ph.h + 1) * ldb;
    double pc = 0, pf = 0, pm = 0;
    if (pi == 0) pf += S.c.feas0_t[(size_t)a * ldb + b];
    else pf += S.ph[pi - 1].feas_t[(size_t)a * (S.ph[pi - 1].h + 1) * ldb + (size_t)S.ph[pi - 1].h * ldb + b];
    for (int k = 0; k < ph.h; ++k) {
      pc += ph.cost_t[aS + (size_t)k * ldb + b];
      pf += ph.feas_t[aS + (size_t)k * ldb + b];
      pm = fmin(pm, ph.ming_t[aS + (size_t)k * ldb + b]);
    }
    pc += ph.cost_t[aS + (size_t)ph.h * ldb + b];
    cost += pc;
    fs += pf;
    max_p = fmin(max_p, pm);
    max_t = fmax(max_t, ph.maxh_t[(size_t)a * ldb + b]);
    fail |= ph.fail_t[(size_t)a * ldb + b];
  }
  feas = sqrt(fs);
}

// Armijo test over the step sizes [a0, a1) that k_roll has just evaluated (MultiPhaseDDP::line_search, MultiPhaseDDP.cpp:95-133):
// keeps the FIRST (largest) successful step size, else the last evaluated one; counts the problems that need more trials.
__global__ void k_ls_scan(const SolverDev* __restrict__ Sp, int a0, int a1) {
  const SolverDev& S = *Sp;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= S.B) return;
  const CtrlDev& c = S.c;
  if (!c.active[b] || !c.do_ls[b] || c.ls_found[b]) return;
  const CafeOptions& o = S.opt;
  const double merit_prev = c.merit_prev[b], feas_prev = c.feas[b], rho = c.merit_rho[b], dV1 = c.dV1[b], dV2 = c.dV2[b];
  for (int a = a0; a < a1; ++a) {
    double cost, feas, mt, mp; int fail;
    reduce_trial(S, a, b, cost, feas, mt, mp, fail);
    const double eps = S.eps[a];
    const double merit = cost + rho * feas;
    const double exp_cost_change = eps * dV1 + 0.5 * eps * eps * dV2;
    const double exp_merit_change = exp_cost_change - eps * rho * feas_prev;
    c.sel[b] = a; c.ls_cost[b] = cost; c.ls_feas[b] = feas; c.ls_mt[b] = mt; c.ls_mp[b] = mp; c.ls_fail[b] = fail; c.ls_merit[b] = merit;
    if ((merit <= merit_prev + o.gamma * exp_merit_change) && !fail) { c.ls_found[b] = 1; return; }
  }
  if (a1 < S.NA) atomicAdd(c.n_pending, 1);
}

// mode 0: initial rollout bookkeeping (MultiPhaseDDP.cpp:238-261); mode 1: after a DDP iteration
__global__ void k_select(const SolverDev* __restrict__ Sp, int mode) {
  const SolverDev& S = *Sp;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= S.B) return;
  const CtrlDev& c = S.c;
  const CafeOptions& o = S.opt;
  const int ldb = S.ldb;
  if (!c.active[b]) { c.sel[b] = -1; return; }
  if (mode == 0 || !c.do_ls[b]) c.sel[b] = -1;
  bool inner_done = false;
  if (mode == 0) {
    double cost, feas, mt, mp; int fail;
    reduce_trial(S, 0, b, cost, feas, mt, mp, fail);
    c.sel[b] = 0; c.accepted[b] = 1;
    c.cost[b] = cost; c.feas[b] = feas; c.max_t[b] = mt; c.max_p[b] = mp;
    for (int pi = 0; pi < S.n_phases; ++pi) for (int i = 0; i < S.ph[pi].n_td; ++i) S.ph[pi].hval[(size_t)i * ldb + b] = S.ph[pi].ht[(size_t)i * ldb + b];
    c.n_hist[b] = 0;
    push_hist(S, b, cost, feas, mt, mp);
    if (fail) c.status[b] = CAFE_STATUS_DIVERGED;
    c.iter_ou[b] = 0;
    // enter the outer loop (MultiPhaseDDP.cpp:264-276)
    if (c.iter_ou[b] < o.max_AL_iter) {
      c.iter_ou[b] = 1; c.max_t_prev[b] = mt; c.max_p_prev[b] = mp; c.reg[b] = 0; c.iter_in[b] = 0;
      if (o.max_DDP_iter <= 0) inner_done = true;
    } else { c.active[b] = 0; }
    if (!inner_done) { if (c.active[b]) atomicAdd(c.n_active, 1); return; }
  } else {
    const int it = c.iter[b] - 1;  // trace slot of this iteration
    double* tr = (it >= 0 && it < CAFE_HIST_CAP) ? c.trace + ((size_t)it * 12) * ldb + b : nullptr;
    if (c.do_ls[b]) {
      const double merit_prev = c.merit_prev[b], cost_prev = c.cost_prev[b];
      const bool success = c.ls_found[b] != 0;
      const int sel = c.sel[b];
      const int n_ls = success ? sel + 1 : S.NA;
      const double cost_s = c.ls_cost[b], feas_s = c.ls_feas[b], mt_s = c.ls_mt[b], mp_s = c.ls_mp[b];
      const int fail_s = c.ls_fail[b];
      if (success) c.merit[b] = c.ls_merit[b];
      c.accepted[b] = success ? 1 : 0;
      c.ls_total[b] += n_ls;
      c.feas[b] = feas_s; c.max_t[b] = mt_s; c.max_p[b] = mp_s;
      for (int pi = 0; pi < S.n_phases; ++pi) for (int i = 0; i < S.ph[pi].n_td; ++i) S.ph[pi].hval[(size_t)i * ldb + b] = S.ph[pi].ht[((size_t)sel * 4 + i) * ldb + b];
      if (success) c.cost[b] = cost_s;
      else { c.cost[b] = cost_prev; c.merit[b] = merit_prev; if (fail_s) c.status[b] = CAFE_STATUS_DIVERGED; }
      if (tr) { tr[(size_t)7 * ldb] = n_ls; tr[(size_t)8 * ldb] = success ? 1 : 0; tr[(size_t)9 * ldb] = success ? S.eps[sel] : 0; tr[(size_t)10 * ldb] = c.cost[b]; tr[(size_t)11 * ldb] = c.feas[b]; }
      const double cost_now = c.cost[b];
      if ((fabs((cost_prev - cost_now) / cost_prev) < o.cost_thresh) && (c.feas[b] <= o.dynamics_feas_thresh)) inner_done = true;
      else push_hist(S, b, cost_now, c.feas[b], c.max_t[b], c.max_p[b]);
    } else {
      inner_done = true;  // early exit taken in k_bwd (MultiPhaseDDP.cpp:345-349)
      if (tr) { tr[(size_t)10 * ldb] = c.cost[b]; tr[(size_t)11 * ldb] = c.feas[b]; }
    }
    if (!inner_done && c.iter_in[b] >= o.max_DDP_iter) inner_done = true;
    if (!inner_done) { atomicAdd(c.n_active, 1); return; }
  }
  // ---- end of an outer iteration (MultiPhaseDDP.cpp:394-425)
  while (true) {
    const double mt = c.max_t[b], mp = c.max_p[b], feas = c.feas[b];
    if (mt < o.tconstr_thresh && fabs(mp) < o.pconstr_thresh && feas <= o.dynamics_feas_thresh) { c.active[b] = 0; break; }
    if (fabs(mt - c.max_t_prev[b]) < 0.0001 && fabs(mp - c.max_p_prev[b]) < 0.0001 && feas <= o.dynamics_feas_thresh) { c.active[b] = 0; break; }
    if (o.AL_active) {
      for (int pi = 0; pi < S.n_phases; ++pi) {
        const PhaseDev& ph = S.ph[pi];
        for (int i = 0; i < ph.n_td; ++i) {  // TerminalConstraintBase::update_params (ConstraintsBase.h:375-391)
          const double hv = ph.hval[(size_t)i * ldb + b];
          if (fabs(hv) < o.tconstr_thresh) continue;
          if (fabs(hv) > 0.005) { double sg = ph.al_sigma[(size_t)i * ldb + b] * o.update_penalty; ph.al_sigma[(size_t)i * ldb + b] = fmin(sg, ph.al_td.sigma_max); }
          else ph.al_lambda[(size_t)i * ldb + b] += hv * ph.al_sigma[(size_t)i * ldb + b];
        }
      }
    }
    /* ReB update: with update_relax == update_ReB == 1 and delta >= delta_min (checked at create) it is the identity */
    if (c.iter_ou[b] >= o.max_AL_iter) { c.active[b] = 0; break; }
    c.iter_ou[b] += 1; c.max_t_prev[b] = mt; c.max_p_prev[b] = mp; c.reg[b] = 0; c.iter_in[b] = 0;
    if (o.max_DDP_iter > 0) break;
  }
  if (c.active[b]) atomicAdd(c.n_active, 1);
}

// --------------------------------------------------------------------------------------------- K-BWD
// 32 workers per problem; worker grid 4 x 8 for the matrix products.
#define CAFE_NW 32

// C(MM x NN) = [C0] + op(A) * B over shared-memory operands interleaved over PB problems.
// TA: use A^T (A stored KK x MM). C0 (optional) is read from global memory at c0[(i + MM*j)*ldg].
template <int MM, int NN, int KK, bool TA, int PB>
__device__ __forceinline__ void gemm32(const double* __restrict__ A, int lda, const double* __restrict__ Bm, int ldbm,
                                       double* __restrict__ C, int ldc, const double* __restrict__ c0, size_t ldg, int w, bool run) {
  constexpr int TR = (MM + 3) / 4, TC = (NN + 7) / 8;
  const int r0 = (w & 3) * TR, q0 = (w >> 2) * TC;
  if (!run) return;
  double acc[TR][TC];
#pragma unroll
  for (int r = 0; r < TR; ++r)
#pragma unroll
    for (int q = 0; q < TC; ++q) acc[r][q] = 0;
  for (int l = 0; l < KK; ++l) {
    double av[TR], bv[TC];
#pragma unroll
    for (int r = 0; r < TR; ++r) { const int i = r0 + r; av[r] = (i < MM) ? (TA ? A[(size_t)(l + lda * i) * PB] : A[(size_t)(i + lda * l) * PB]) : 0.0; }
#pragma unroll
    for (int q = 0; q < TC; ++q) { const int j = q0 + q; bv[q] = (j < NN) ? Bm[(size_t)(l + ldbm * j) * PB] : 0.0; }
#pragma unroll
    for (int r = 0; r < TR; ++r)
#pragma unroll
      for (int q = 0; q < TC; ++q) acc[r][q] += av[r] * bv[q];
  }
#pragma unroll
  for (int r = 0; r < TR; ++r)
#pragma unroll
    for (int q = 0; q < TC; ++q) {
      const int i = r0 + r, j = q0 + q;
      if (i < MM && j < NN) {
        double v = acc[r][q];
        if (c0) v += c0[(size_t)(i + MM * j) * ldg];
        C[(size_t)(i + ldc * j) * PB] = v;
      }
    }
}

// shared-memory carve-up (in doubles per problem) for a deck whose largest phase is (NX, MX, PX)
template <int NX, int MX, int PX>
struct BwdLayout {
  static constexpr int oG = 0, oGn = oG + NX, oQx = oGn + NX, oQu = oQx + NX, oDu = oQu + MX, oD = oDu + MX, oLy = oD + NX,
                       oDx = oLy + (PX > 0 ? PX : 0), oDxn = oDx + NX, oDuL = oDxn + NX, oRed = oDuL + MX,
                       oH = oRed + 2 * CAFE_NW, oA = oH + NX * NX, oB = oA + NX * NX, oT = oB + NX * MX,
                       oQxx = oT + NX * (NX + MX), oQux = oQxx + NX * NX, oQuu = oQux + MX * NX, oL = oQuu + MX * MX,
                       oK = oL + MX * MX, oC = oK + MX * NX, oDm = oC + PX * NX, oLyy = oDm + PX * MX,
                       oSC = oLyy + PX * PX, oSD = oSC + PX * NX, total = oSD + PX * MX;
};

template <int N, int M, int PY, int NNEXT, int PB, class L>
__device__ void sweep_phase(const SolverDev& S, int pi, int b, int p, int w, bool run, bool& ok, double reg, double* sm,
                            double& min_piv) {
  const PhaseDev& ph = S.ph[pi];
  const int ldb = S.ldb, h = ph.h;
  const int nthr = CAFE_NW * PB, t = w * PB + p;
  double* sG = sm + L::oG * PB + p;   double* sGn = sm + L::oGn * PB + p; double* sQx = sm + L::oQx * PB + p;
  double* sQu = sm + L::oQu * PB + p; double* sDu = sm + L::oDu * PB + p; double* sD = sm + L::oD * PB + p;
  double* sLy = sm + L::oLy * PB + p;
  double* sH = sm + L::oH * PB + p;   double* sA = sm + L::oA * PB + p;   double* sB = sm + L::oB * PB + p;
  double* sT = sm + L::oT * PB + p;   double* sQxx = sm + L::oQxx * PB + p; double* sQux = sm + L::oQux * PB + p;
  double* sQuu = sm + L::oQuu * PB + p; double* sL = sm + L::oL * PB + p; double* sK = sm + L::oK * PB + p;
  double* sC = sm + L::oC * PB + p;   double* sDm = sm + L::oDm * PB + p; double* sLyy = sm + L::oLyy * PB + p;
  double* sSC = sm + L::oSC * PB + p; double* sSD = sm + L::oSD * PB + p;
  (void)nthr; (void)t; (void)sLy; (void)sC; (void)sDm; (void)sLyy; (void)sSC; (void)sSD;
  const bool act = run && ok;

  // ---- terminal boundary: (G', H') = (Px^T G0+, Px^T H0+ Px) from the next phase (impact_aware_step), else 0
  if (ph.has_next) {
    // sG/sH hold G0+/H0+ of phase pi+1 with dimension NNEXT; Px is NNEXT x N
    if (act) for (int e = w; e < NNEXT * N; e += CAFE_NW) sA[(size_t)e * PB] = ph.Px[(size_t)e * ldb + b];
    __syncthreads();
    gemm32<NNEXT, N, NNEXT, false, PB>(sH, NNEXT, sA, NNEXT, sT, NNEXT, nullptr, 0, w, act);  // T = H0+ * Px
    if (act) for (int j = w; j < N; j += CAFE_NW) { double s = 0; for (int i = 0; i < NNEXT; ++i) s += sA[(size_t)(i + NNEXT * j) * PB] * sG[(size_t)i * PB]; sGn[(size_t)j * PB] = s; }
    __syncthreads();
    gemm32<N, N, NNEXT, true, PB>(sA, NNEXT, sT, NNEXT, sH, N, ph.Phixx + b, ldb, w, act);  // H[h] = Phixx + Px^T T
    if (act) for (int j = w; j < N; j += CAFE_NW) sG[(size_t)j * PB] = ph.Phix[(size_t)j * ldb + b] + sGn[(size_t)j * PB];
  } else {
    if (act) {
      for (int e = w; e < N * N; e += CAFE_NW) sH[(size_t)e * PB] = ph.Phixx[(size_t)e * ldb + b];
      for (int j = w; j < N; j += CAFE_NW) sG[(size_t)j * PB] = ph.Phix[(size_t)j * ldb + b];
    }
  }
  if (act) for (int j = w; j < N; j += CAFE_NW) ph.G[gix(h, N, j, ldb, b)] = sG[(size_t)j * PB];
  __syncthreads();

  for (int k = h - 1; k >= 0; --k) {
    const bool a2 = run && ok;
    // ---- stage A_k, B_k (C_k, D_k, lyy, ly), Defect[k+1]
    if (a2) {
      const double* Ag = ph.A + gix(k, N * N, 0, ldb, b);
      for (int e = w; e < N * N; e += CAFE_NW) sA[(size_t)e * PB] = Ag[(size_t)e * ldb];
      const double* Bg = ph.Bm + gix(k, N * M, 0, ldb, b);
      for (int e = w; e < N * M; e += CAFE_NW) sB[(size_t)e * PB] = Bg[(size_t)e * ldb];
      for (int j = w; j < N; j += CAFE_NW) sD[(size_t)j * PB] = ph.Defect[gix(k + 1, N, j, ldb, b)];
      if constexpr (PY > 0) {
        const double* Cg = ph.C + gix(k, PY * N, 0, ldb, b);
        for (int e = w; e < PY * N; e += CAFE_NW) sC[(size_t)e * PB] = Cg[(size_t)e * ldb];
        const double* Dg = ph.D + gix(k, PY * M, 0, ldb, b);
        for (int e = w; e < PY * M; e += CAFE_NW) sDm[(size_t)e * PB] = Dg[(size_t)e * ldb];
        const double* Lg = ph.lyy + gix(k, PY * PY, 0, ldb, b);
        for (int e = w; e < PY * PY; e += CAFE_NW) sLyy[(size_t)e * PB] = Lg[(size_t)e * ldb];
        for (int j = w; j < PY; j += CAFE_NW) sLy[(size_t)j * PB] = ph.ly[gix(k, PY, j, ldb, b)];
      }
    }
    __syncthreads();
    // ---- Gn = G + H d ; T = H [A B]
    if (a2) for (int i = w; i < N; i += CAFE_NW) { double s = sG[(size_t)i * PB]; for (int j = 0; j < N; ++j) s += sH[(size_t)(i + N * j) * PB] * sD[(size_t)j * PB]; sGn[(size_t)i * PB] = s; }
    gemm32<N, N, N, false, PB>(sH, N, sA, N, sT, N, nullptr, 0, w, a2);
    gemm32<N, M, N, false, PB>(sH, N, sB, N, sT + (size_t)N * N * PB, N, nullptr, 0, w, a2);
    if constexpr (PY > 0) {
      gemm32<PY, N, PY, false, PB>(sLyy, PY, sC, PY, sSC, PY, nullptr, 0, w, a2);
      gemm32<PY, M, PY, false, PB>(sLyy, PY, sDm, PY, sSD, PY, nullptr, 0, w, a2);
    }
    __syncthreads();
    // ---- Q functions
    if (a2) {
      for (int j = w; j < N; j += CAFE_NW) {
        double s = ph.lx[gix(k, N, j, ldb, b)];
        for (int i = 0; i < N; ++i) s += sA[(size_t)(i + N * j) * PB] * sGn[(size_t)i * PB];
        if constexpr (PY > 0) for (int i = 0; i < PY; ++i) s += sC[(size_t)(i + PY * j) * PB] * sLy[(size_t)i * PB];
        sQx[(size_t)j * PB] = s;
      }
      for (int j = w; j < M; j += CAFE_NW) {
        double s = ph.lu[gix(k, M, j, ldb, b)];
        for (int i = 0; i < N; ++i) s += sB[(size_t)(i + N * j) * PB] * sGn[(size_t)i * PB];
        if constexpr (PY > 0) for (int i = 0; i < PY; ++i) s += sDm[(size_t)(i + PY * j) * PB] * sLy[(size_t)i * PB];
        sQu[(size_t)j * PB] = s;
      }
    }
    gemm32<N, N, N, true, PB>(sA, N, sT, N, sQxx, N, ph.lxx + gix(k, N * N, 0, ldb, b), ldb, w, a2);
    gemm32<M, N, N, true, PB>(sB, N, sT, N, sQux, M, nullptr, 0, w, a2);
    gemm32<M, M, N, true, PB>(sB, N, sT + (size_t)N * N * PB, N, sQuu, M, ph.luu + gix(k, M * M, 0, ldb, b), ldb, w, a2);
    __syncthreads();
    if constexpr (PY > 0) {
      // Qxx += C^T (lyy C), Quu += D^T (lyy D), Qux += D^T (lyy C): accumulate through the c0 path is global-only,
      // so do them as explicit per-element sums (PY is small)
      if (a2) {
        for (int e = w; e < N * N; e += CAFE_NW) { const int i = e % N, j = e / N; double s = 0; for (int l = 0; l < PY; ++l) s += sC[(size_t)(l + PY * i) * PB] * sSC[(size_t)(l + PY * j) * PB]; sQxx[(size_t)e * PB] += s; }
        for (int e = w; e < M * M; e += CAFE_NW) { const int i = e % M, j = e / M; double s = 0; for (int l = 0; l < PY; ++l) s += sDm[(size_t)(l + PY * i) * PB] * sSD[(size_t)(l + PY * j) * PB]; sQuu[(size_t)e * PB] += s; }
        for (int e = w; e < M * N; e += CAFE_NW) { const int i = e % M, j = e / M; double s = 0; for (int l = 0; l < PY; ++l) s += sDm[(size_t)(l + PY * i) * PB] * sSC[(size_t)(l + PY * j) * PB]; sQux[(size_t)e * PB] += s; }
      }
      __syncthreads();
    }
    // ---- regularisation, outputs, copy for factorisation
    if (a2) {
      for (int i = w; i < N; i += CAFE_NW) sQxx[(size_t)(i + N * i) * PB] += reg;
      for (int i = w; i < M; i += CAFE_NW) sQuu[(size_t)(i + M * i) * PB] += reg;
    }
    __syncthreads();
    if (a2) {
      double* Quug = ph.Quu + gix(k, M * M, 0, ldb, b);
      for (int e = w; e < M * M; e += CAFE_NW) { const double v = sQuu[(size_t)e * PB]; Quug[(size_t)e * ldb] = v; const int i = e % M, j = e / M; sL[(size_t)e * PB] = (i == j) ? v - 1e-9 : v; }
      double* Quxg = ph.Qux + gix(k, M * N, 0, ldb, b);
      for (int e = w; e < M * N; e += CAFE_NW) { const double v = sQux[(size_t)e * PB]; Quxg[(size_t)e * ldb] = v; sK[(size_t)e * PB] = v; }
      for (int j = w; j < M; j += CAFE_NW) { const double v = sQu[(size_t)j * PB]; ph.Qu[gix(k, M, j, ldb, b)] = v; sDu[(size_t)j * PB] = v; }
    }
    // ---- LDL^T of (Quu - 1e-9 I), right-looking; positive-definiteness test = every pivot > 0
    for (int j = 0; j < M; ++j) {
      __syncthreads();
      if (run && ok) {
        const double d = sL[(size_t)(j + M * j) * PB];
        if (!(d > 0.0)) ok = false;
        else {
          min_piv = fmin(min_piv, d);
          const double inv = 1.0 / d;
          for (int i = j + 1 + w; i < M; i += CAFE_NW) {
            const double vi = sL[(size_t)(i + M * j) * PB] * inv;
            for (int cc = j + 1; cc <= i; ++cc) sL[(size_t)(i + M * cc) * PB] -= vi * sL[(size_t)(cc + M * j) * PB];
          }
          if (j > 0) { const double dp = sL[(size_t)((j - 1) + M * (j - 1)) * PB]; for (int i = j + w; i < M; i += CAFE_NW) sL[(size_t)(i + M * (j - 1)) * PB] /= dp; }
        }
      }
    }
    __syncthreads();
    const bool a3 = run && ok;
    // ---- solve (Quu - 1e-9 I) [K | dU] = -[Qux | Qu], one right-hand side per worker
    if (a3) {
      for (int col = w; col < N + 1; col += CAFE_NW) {
        double* x = (col < N) ? sK + (size_t)(M * col) * PB : sDu;
        for (int i = 0; i < M; ++i) { double s = x[(size_t)i * PB]; for (int l = 0; l < i; ++l) s -= sL[(size_t)(i + M * l) * PB] * x[(size_t)l * PB]; x[(size_t)i * PB] = s; }
        for (int i = 0; i < M; ++i) x[(size_t)i * PB] /= sL[(size_t)(i + M * i) * PB];
        for (int i = M - 1; i >= 0; --i) { double s = x[(size_t)i * PB]; for (int l = i + 1; l < M; ++l) s -= sL[(size_t)(l + M * i) * PB] * x[(size_t)l * PB]; x[(size_t)i * PB] = s; }
        for (int i = 0; i < M; ++i) x[(size_t)i * PB] = -x[(size_t)i * PB];
      }
    }
    __syncthreads();
    // ---- value function: G = Qx + Qux^T dU ; H = sym(Qxx) + Qux^T K
    if (a3) {
      for (int j = w; j < N; j += CAFE_NW) {
        double s = sQx[(size_t)j * PB];
        for (int i = 0; i < M; ++i) s += sQux[(size_t)(i + M * j) * PB] * sDu[(size_t)i * PB];
        sG[(size_t)j * PB] = s;
        ph.G[gix(k, N, j, ldb, b)] = s;
      }
      for (int j = w; j < M; j += CAFE_NW) ph.dU[gix(k, M, j, ldb, b)] = sDu[(size_t)j * PB];
      double* Kg = ph.K + gix(k, M * N, 0, ldb, b);
      for (int e = w; e < M * N; e += CAFE_NW) Kg[(size_t)e * ldb] = sK[(size_t)e * PB];
    }
    gemm32<N, N, M, true, PB>(sQux, M, sK, M, sH, N, nullptr, 0, w, a3);
    __syncthreads();
    if (a3) for (int e = w; e < N * N; e += CAFE_NW) { const int i = e % N, j = e / N; sH[(size_t)e * PB] += (sQxx[(size_t)(i + N * j) * PB] + sQxx[(size_t)(j + N * i) * PB]) / 2; }
    __syncthreads();
  }
  // ---- G[0] += H[0] Defect[0]
  if (run && ok) {
    for (int j = w; j < N; j += CAFE_NW) sD[(size_t)j * PB] = ph.Defect[gix(0, N, j, ldb, b)];
  }
  __syncthreads();
  if (run && ok) for (int i = w; i < N; i += CAFE_NW) { double s = sG[(size_t)i * PB]; for (int j = 0; j < N; ++j) s += sH[(size_t)(i + N * j) * PB] * sD[(size_t)j * PB]; sGn[(size_t)i * PB] = s; }
  __syncthreads();
  if (run && ok) for (int i = w; i < N; i += CAFE_NW) { sG[(size_t)i * PB] = sGn[(size_t)i * PB]; ph.G[gix(0, N, i, ldb, b)] = sGn[(size_t)i * PB]; }
  __syncthreads();
}

// multiple-shooting linear rollout of one phase (SinglePhase::linear_rollout), eps = 1
template <int N, int M, int PB, class L>
__device__ void lin_phase(const SolverDev& S, int pi, int b, int p, int w, bool run, double* sm, double& dV1, double& dV2) {
  const PhaseDev& ph = S.ph[pi];
  const int ldb = S.ldb, h = ph.h;
  double* sDx = sm + L::oDx * PB + p; double* sDxn = sm + L::oDxn * PB + p; double* sDuL = sm + L::oDuL * PB + p;
  double* sRed = sm + L::oRed * PB + p;
  // sDx holds dx_init on entry
  if (run) for (int i = w; i < N; i += CAFE_NW) { const double v = sDx[(size_t)i * PB] + 1.0 * ph.Defect[gix(0, N, i, ldb, b)]; sDx[(size_t)i * PB] = v; ph.dX[gix(0, N, i, ldb, b)] = v; }
  __syncthreads();
  for (int k = 0; k < h; ++k) {
    double part1 = 0, part2 = 0;
    if (run) {
      const double* Kg = ph.K + gix(k, M * N, 0, ldb, b);
      for (int i = w; i < M; i += CAFE_NW) {
        double s = 0;
        for (int j = 0; j < N; ++j) s += Kg[(size_t)(i + M * j) * ldb] * sDx[(size_t)j * PB];
        sDuL[(size_t)i * PB] = 1.0 * ph.dU[gix(k, M, i, ldb, b)] + s;
      }
    }
    __syncthreads();
    if (run) {
      const double* Ag = ph.A + gix(k, N * N, 0, ldb, b);
      const double* Bg = ph.Bm + gix(k, N * M, 0, ldb, b);
      const double* lxxg = ph.lxx + gix(k, N * N, 0, ldb, b);
      const double* luug = ph.luu + gix(k, M * M, 0, ldb, b);
      for (int i = w; i < N; i += CAFE_NW) {
        double s = 0, s2 = 0, q = 0;
        for (int j = 0; j < N; ++j) { const double dxj = sDx[(size_t)j * PB]; s += Ag[(size_t)(i + N * j) * ldb] * dxj; q += lxxg[(size_t)(i + N * j) * ldb] * dxj; }
        for (int j = 0; j < M; ++j) s2 += Bg[(size_t)(i + N * j) * ldb] * sDuL[(size_t)j * PB];
        const double v = s + s2 + 1.0 * ph.Defect[gix(k + 1, N, i, ldb, b)];
        sDxn[(size_t)i * PB] = v;
        ph.dX[gix(k + 1, N, i, ldb, b)] = v;
        const double dxi = sDx[(size_t)i * PB];
        part1 += ph.lx[gix(k, N, i, ldb, b)] * dxi;
        part2 += dxi * q;
      }
      for (int i = w; i < M; i += CAFE_NW) {
        double q = 0;
        for (int j = 0; j < M; ++j) q += luug[(size_t)(i + M * j) * ldb] * sDuL[(size_t)j * PB];
        const double dui = sDuL[(size_t)i * PB];
        part1 += ph.lu[gix(k, M, i, ldb, b)] * dui;
        part2 += dui * q;
      }
    }
    sRed[(size_t)w * PB] = part1;
    sRed[(size_t)(CAFE_NW + w) * PB] = part2;
    __syncthreads();
    if (run) {
      if (w == 0) { double a1 = 0, a2 = 0; for (int i = 0; i < CAFE_NW; ++i) { a1 += sRed[(size_t)i * PB]; a2 += sRed[(size_t)(CAFE_NW + i) * PB]; } dV1 += a1; dV2 += a2; }
      for (int i = w; i < N; i += CAFE_NW) sDx[(size_t)i * PB] = sDxn[(size_t)i * PB];
    }
    __syncthreads();
  }
  // terminal terms
  double part1 = 0, part2 = 0;
  if (run) {
    for (int i = w; i < N; i += CAFE_NW) {
      double q = 0;
      for (int j = 0; j < N; ++j) q += ph.Phixx[(size_t)(i + N * j) * ldb + b] * sDx[(size_t)j * PB];
      const double dxi = sDx[(size_t)i * PB];
      part1 += ph.Phix[(size_t)i * ldb + b] * dxi;
      part2 += dxi * q;
    }
  }
  sRed[(size_t)w * PB] = part1;
  sRed[(size_t)(CAFE_NW + w) * PB] = part2;
  __syncthreads();
  if (run && w == 0) { double a1 = 0, a2 = 0; for (int i = 0; i < CAFE_NW; ++i) { a1 += sRed[(size_t)i * PB]; a2 += sRed[(size_t)(CAFE_NW + i) * PB]; } dV1 += a1; dV2 += a2; }
  // dx_init of the next phase = Px * dX[h]
  if (ph.has_next) {
    const int nn = ph.n_next;
    if (run) for (int i = w; i < nn; i += CAFE_NW) { double s = 0; for (int j = 0; j < N; ++j) s += ph.Px[(size_t)(i + nn * j) * ldb + b] * sDx[(size_t)j * PB]; sDxn[(size_t)i * PB] = s; }
    __syncthreads();
    if (run) for (int i = w; i < nn; i += CAFE_NW) sDx[(size_t)i * PB] = sDxn[(size_t)i * PB];
  }
  __syncthreads();
}

template <int NX, int MX, int PX, int PB>
__global__ void __launch_bounds__(CAFE_NW* PB) k_bwd(const SolverDev* __restrict__ Sp) {
  typedef BwdLayout<NX, MX, PX> L;
  const SolverDev& S = *Sp;
  extern __shared__ double sm[];
  __shared__ double s_reg[PB];
  __shared__ int s_state[PB], s_regiter[PB];  // 0 sweeping, 1 success, 2 gave up, 3 not participating
  const int t = threadIdx.x, p = t % PB, w = t / PB;
  const int b = blockIdx.x * PB + p;
  const CtrlDev& c = S.c;
  const CafeOptions& o = S.opt;
  const int ldb = S.ldb;
  const bool valid = (b < S.B) && c.active[b];
  int it = 0;
  if (w == 0) {
    s_state[p] = valid ? 0 : 3;
    s_regiter[p] = 0;
    s_reg[p] = valid ? c.reg[b] : 0.0;
    if (valid) {
      // compute_cost + measure_dynamics_feasibility on the current (trial) arrays (MultiPhaseDDP.cpp:280-281)
      double cost = 0, fs = 0;
      for (int pi = 0; pi < S.n_phases; ++pi) {
        const PhaseDev& ph = S.ph[pi];
        double pc = 0, pf = 0;
        for (int k = 0; k < ph.h; ++k) pc += ph.lk[(size_t)k * ldb + b];
        pc += ph.lk[(size_t)ph.h * ldb + b];
        for (int k = 0; k <= ph.h; ++k) pf += ph.dsq[(size_t)k * ldb + b];
        cost += pc; fs += pf;
      }
      c.cost[b] = cost; c.feas[b] = sqrt(fs);
      c.iter_in[b] += 1; c.iter[b] += 1;
      it = c.iter[b] - 1;
      if (it < CAFE_HIST_CAP) { double* tr = c.trace + ((size_t)it * 12) * ldb + b; for (int i = 0; i < 12; ++i) tr[(size_t)i * ldb] = 0; tr[0] = cost; tr[(size_t)ldb] = sqrt(fs); }
    }
  }
  __syncthreads();
  double min_piv = 1e300;
  while (true) {
    const bool mine = valid && s_state[p] == 0;
    if (!__syncthreads_or(mine ? 1 : 0)) break;
    bool ok = true;
    const double reg = s_reg[p];
    for (int pi = S.n_phases - 1; pi >= 0; --pi) {
      const int model = S.ph[pi].model, nm = S.ph[pi].has_next ? S.ph[pi + 1].model : -1;
      if constexpr (NX == 24) {
        if (model == CAFE_MODEL_HKD) sweep_phase<24, 24, 0, 24, PB, L>(S, pi, b, p, w, mine, ok, reg, sm, min_piv);
      } else {
        if (model == CAFE_MODEL_SRB) sweep_phase<12, 12, 0, 12, PB, L>(S, pi, b, p, w, mine, ok, reg, sm, min_piv);
        else if (model == CAFE_MODEL_WB && nm == CAFE_MODEL_SRB) sweep_phase<36, 12, 12, 12, PB, L>(S, pi, b, p, w, mine, ok, reg, sm, min_piv);
        else if (model == CAFE_MODEL_WB) sweep_phase<36, 12, 12, 36, PB, L>(S, pi, b, p, w, mine, ok, reg, sm, min_piv);
      }
      (void)nm;
    }
    __syncthreads();
    if (w == 0 && mine) {
      s_regiter[p] += 1;
      if (ok) s_state[p] = 1;
      else {
        double r = fmax(s_reg[p] * o.update_regularization, 1e-03);
        s_reg[p] = r;
        if (r > 1e2) s_state[p] = 2;
      }
    }
    __syncthreads();
  }
  const bool success = valid && s_state[p] == 1;
  // ---- linear rollout (option.MS) and expected cost change
  double dV1 = 0, dV2 = 0;
  {
    double* sDx = sm + L::oDx * PB + p;
    for (int i = w; i < NX; i += CAFE_NW) sDx[(size_t)i * PB] = 0.0;
    __syncthreads();
    for (int pi = 0; pi < S.n_phases; ++pi) {
      const int model = S.ph[pi].model;
      if constexpr (NX == 24) {
        if (model == CAFE_MODEL_HKD) lin_phase<24, 24, PB, L>(S, pi, b, p, w, success, sm, dV1, dV2);
      } else {
        if (model == CAFE_MODEL_WB) lin_phase<36, 12, PB, L>(S, pi, b, p, w, success, sm, dV1, dV2);
        else if (model == CAFE_MODEL_SRB) lin_phase<12, 12, PB, L>(S, pi, b, p, w, success, sm, dV1, dV2);
      }
    }
  }
  if (w == 0 && valid) {
    double r = s_reg[p] / 20;
    if (r < 1e-06) r = 0;
    c.reg[b] = r;
    c.reg_total[b] += s_regiter[p];
    c.min_pivot[b] = fmin(c.min_pivot[b], min_piv);
    it = c.iter[b] - 1;
    double* tr = (it < CAFE_HIST_CAP) ? c.trace + ((size_t)it * 12) * ldb + b : nullptr;
    if (tr) { tr[(size_t)5 * ldb] = r; tr[(size_t)6 * ldb] = s_regiter[p]; }
    if (!success) {
      c.status[b] = CAFE_STATUS_REG_FAIL; c.active[b] = 0; c.do_ls[b] = 0;  // "bad_solve" (MultiPhaseDDP.cpp:317-320)
    } else {
      c.dV1[b] = dV1; c.dV2[b] = dV2;
      const double feas = c.feas[b], cost = c.cost[b];
      const double dV_abs = fabs(dV1 + 0.5 * dV2);
      const double rho = (feas > o.dynamics_feas_thresh) ? dV_abs / ((1 - o.merit_scale) * feas) + o.merit_offset : 0;
      c.merit_rho[b] = rho;
      const double merit = cost + rho * feas;
      c.merit[b] = merit; c.cost_prev[b] = cost; c.merit_prev[b] = merit;
      if (tr) { tr[(size_t)2 * ldb] = dV1; tr[(size_t)3 * ldb] = dV2; tr[(size_t)4 * ldb] = rho; }
      c.do_ls[b] = ((dV_abs < o.cost_thresh) && (feas <= o.dynamics_feas_thresh)) ? 0 : 1;
      c.ls_found[b] = 0;
    }
  }
}

}  // namespace cafe_dev
