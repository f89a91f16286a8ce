// knot_kernels.cuh — the per-(problem, knot) kernels and the per-problem control kernels of the batched HS-DDP path (sm_100a, fp64).
// (The Riccati sweep lives in bwd2.cuh, the whole-body running knots in wb_leg_kernels.cu (straight-line rigid-body routines) and wb_coop.cuh (cooperative KKT kernels).)
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

// resident CTAs per SM the per-(problem,knot) kernels are compiled for (register cap = 65536 / (128 * CAFE_KNOT_MINB))
#ifndef CAFE_KNOT_MINB
#define CAFE_KNOT_MINB 4
#endif
namespace cafe_dev {

// ------------------------------------------------------------------------------------------- K-ROLL
// knots 0 .. kend-1 of phase pi integrated in place from x (the state at knot 0): U = Ubar + eps dU + K (X - Xbar)
template <class Model>
__device__ __noinline__ void chain_phase(const SolverDev& S, int pi, int kend, int a, int b, double* x) {
  constexpr int N = Model::N, M = Model::M, PY = Model::PY;
  const PhaseDev& ph = S.ph[pi];
  const int ldb = S.ldb;
  const double eps = S.eps[a];
  double rec_local[CAFE_REF_W];
  for (int kk = 0; kk < kend; ++kk) {
    const double* rec = knot_record(ph, kk, ldb, b, rec_local);
    double u[M], xn[N], y[PY > 0 ? PY : 1], l, ming;
    for (int i = 0; i < M; ++i) u[i] = 0;
    const double* Kg = ph.K + gix(kk, M * N, 0, ldb, b);
    for (int j = 0; j < N; ++j) {
      const double dj = x[j] - ph.Xbar[gix(kk, N, j, ldb, b)];
      for (int i = 0; i < M; ++i) u[i] += Kg[(size_t)(i + M * j) * ldb] * dj;
    }
    for (int i = 0; i < M; ++i) u[i] = ph.Ubar[gix(kk, M, i, ldb, b)] + eps * ph.dU[gix(kk, M, i, ldb, b)] + u[i];
    Model::roll(ph, rec, x, u, xn, y, S.opt.ReB_active != 0, l, ming, reb_ctx(ph, kk, ldb, b));   // only xn is used
    for (int i = 0; i < N; ++i) x[i] = xn[i];
  }
}

// State of knot k of a phase WITHOUT shooting states (PhaseDev::single_shooting; SinglePhase.cpp:187-221 with an empty SS_set): the
// phase is integrated from the state the previous phase hands over (its trial end state through the reset map; the solver's x0 for
// the first phase), U = Ubar + eps dU + K (X - Xbar). Such a phase is the freshly opened tail of an MPC update, at most dt_mpc / dt
// knots long, so every knot's thread simply repeats the short chain (same routine, same inputs: bit-identical across threads).
template <class Model>
__device__ __noinline__ void single_shooting_state(const SolverDev& S, int pi, int k, int a, int b, double* x) {
  constexpr int N = Model::N, M = Model::M, PY = Model::PY;
  const PhaseDev& ph = S.ph[pi];
  const int ldb = S.ldb;
  const double eps = S.eps[a];
  if (pi == 0) {
    for (int i = 0; i < N; ++i) x[i] = S.x0[(size_t)i * ldb + b];
  } else {
    const PhaseDev& pv = S.ph[pi - 1];
    double xe[N];
    for (int i = 0; i < N; ++i) xe[i] = pv.Xbar[gix(pv.h, N, i, ldb, b)] + eps * pv.dX[gix(pv.h, N, i, ldb, b)];
    Model::resetmap(pv, xe, x);
  }
  chain_phase<Model>(S, pi, k, a, b, x);
}

// Whole-problem single shooting (HSDDP_OPTION::MS = false: MultiPhaseDDP::hybrid_rollout clears every phase's shooting set,
// MultiPhaseDDP.cpp:65-68, so X[0] = x_init and X[k+1] = Xsim[k+1] everywhere, SinglePhase.cpp:187-221): the state of knot k of phase pi is
// the chain from the solver's x0 through every earlier phase and its reset map. Every knot's thread repeats the chain up to its own knot
// (quadratic work: the option is there for completeness - no shipped problem uses it - not for speed).
__device__ __noinline__ void whole_problem_shooting_state(const SolverDev& S, int pi, int k, int a, int b, double* x /*[CAFE_MAX_N]*/) {
  const int ldb = S.ldb;
  for (int i = 0; i < S.ph[0].n; ++i) x[i] = S.x0[(size_t)i * ldb + b];
  for (int q = 0; q <= pi; ++q) {
    const PhaseDev& ph = S.ph[q];
    const int kend = (q == pi) ? k : ph.h;
    double xn[CAFE_MAX_N];
    switch (ph.model) {
      case CAFE_MODEL_HKD: chain_phase<HKDModel>(S, q, kend, a, b, x); if (q < pi) HKDModel::resetmap(ph, x, xn); break;
      case CAFE_MODEL_WB: chain_phase<WBModel>(S, q, kend, a, b, x); if (q < pi) WBModel::resetmap(ph, x, xn); break;
      default: chain_phase<SRBModel>(S, q, kend, a, b, x); if (q < pi) SRBModel::resetmap(ph, x, xn); break;
    }
    if (q < pi) for (int i = 0; i < S.ph[q + 1].n; ++i) x[i] = xn[i];
  }
}

template <class Model>
__device__ void roll_knot(const SolverDev& S, int pi, int k, int a, int b) {
  constexpr int N = Model::N, M = Model::M, PY = Model::PY;
  const PhaseDev& ph = S.ph[pi];
  const int ldb = S.ldb, h = ph.h;
  const double eps = S.eps[a];
  double rec_local[CAFE_REF_W];
  const double* rec = knot_record(ph, k, ldb, b, rec_local);
  const size_t aX = (size_t)a * (h + 1) * N * ldb, aU = (size_t)a * h * M * ldb, aY = (size_t)a * h * PY * ldb;
  const size_t aS = (size_t)a * (h + 1) * ldb;
  double x[N], dlt[N];
  const bool ss = ph.single_shooting != 0;
  if (ss) {
    if (S.opt.MS) single_shooting_state<Model>(S, pi, k, a, b, x);
    else { double xw[CAFE_MAX_N]; whole_problem_shooting_state(S, pi, k, a, b, xw); for (int i = 0; i < N; ++i) x[i] = xw[i]; }
  }
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double xb = ph.Xbar[gix(k, N, i, ldb, b)];
    if (!ss) x[i] = xb + eps * ph.dX[gix(k, N, i, ldb, b)];
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
    // whole-body running knots continue in k_wb_terms (leg-parallel rigid-body terms) and k_wb_fwd (cooperative KKT solve, x+,
    // GRF, cost, defects) from the trial state and control stored above
    if constexpr (Model::COOP) return;
    double xn[N], y[PY > 0 ? PY : 1];
    double l, ming;
    Model::roll(ph, rec, x, u, xn, y, S.opt.ReB_active != 0, l, ming, reb_ctx(ph, k, ldb, b));
    double nrm = 0, dsq = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      nrm += xn[i] * xn[i];
      const double xs = ss ? xn[i] : ph.Xbar[gix(k + 1, N, i, ldb, b)] + eps * ph.dX[gix(k + 1, N, i, ldb, b)];  // X[k+1] = Xsim[k+1] without shooting states
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
        const double xs = nx.single_shooting ? xr[i] : nx.Xbar[gix(0, nx.n, i, ldb, b)] + eps * nx.dX[gix(0, nx.n, i, ldb, b)];
        const double d = xr[i] - xs;
        nx.Dt[aXn + gix(0, nx.n, i, ldb, b)] = d;
        dsq += d * d;
      }
    }
    ph.feas_t[aS + (size_t)h * ldb + b] = dsq;
  }
}

// step sizes [a0, a1) of the ladder; problems whose line search already succeeded are skipped
// thread = (entry j of the index list, knot, step size), j fastest: a warp is 32 listed problems of one knot
__global__ void __launch_bounds__(128, CAFE_KNOT_MINB) k_roll(const SolverDev* __restrict__ Sp, int a0, int a1, const int* __restrict__ list, int n_list) {
  const SolverDev& S = *Sp;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int ldj = (n_list + 31) & ~31;
  const int j = (int)(t % ldj);
  const long long r = t / ldj;
  const int gk = (int)(r % S.n_knots), a = a0 + (int)(r / S.n_knots);
  if (a >= a1 || j >= n_list) return;
  const int b = list[j];
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
  double rec_local[CAFE_REF_W];
  const double* rec = knot_record(ph, k, ldb, b, rec_local);
  double x[N];
  double dsq = 0;
#pragma unroll
  for (int i = 0; i < N; ++i) { x[i] = ph.X[gix(k, N, i, ldb, b)]; const double d = ph.Defect[gix(k, N, i, ldb, b)]; dsq += d * d; }
  ph.dsq[(size_t)k * ldb + b] = dsq;
  if (k < h) {
    if constexpr (Model::COOP) return;   // whole-body running knots: k_wb_derivs + k_wb_lq
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

// the list is padded to a multiple of the CTA size: one knot per CTA (uniform control flow inside a CTA)
__global__ void __launch_bounds__(128, CAFE_KNOT_MINB) k_lq(const SolverDev* __restrict__ Sp, const int* __restrict__ list, int n_list) {
  const SolverDev& S = *Sp;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int ldj = (n_list + 127) & ~127;
  const int j = (int)(t % ldj);
  const int gk = (int)(t / ldj);
  if (gk >= S.n_knots || j >= n_list) return;
  const int b = list[j];
  if (!S.c.active[b]) return;
  const int pi = S.knot_phase[gk], k = S.knot_k[gk];
  switch (S.ph[pi].model) {
    case CAFE_MODEL_HKD: lq_knot_generic<HKDModel>(S, pi, k, b); break;
    case CAFE_MODEL_WB: lq_knot_generic<WBModel>(S, pi, k, b); break;
    case CAFE_MODEL_SRB: lq_knot_generic<SRBModel>(S, pi, k, b); break;
    default: break;
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
  if (gk == 0) S.c.cur_slot[b] = a;   // the rigid-body terms of this trial are the linearisation's (k_wb_lq)
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

// Ordered compaction of a per-problem predicate into an index list (one CTA of 1024 threads; B <= a few 10^4).
// mode 0: problems still iterating -> c.act_list, c.n_active;  mode 1: line searches that need more step sizes -> c.pend_list, c.n_pending
__global__ void __launch_bounds__(1024) k_compact(const SolverDev* __restrict__ Sp, int mode) {
  const SolverDev& S = *Sp;
  const CtrlDev& c = S.c;
  __shared__ int warp_tot[32];
  __shared__ int base;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  int* list = mode == 0 ? c.act_list : c.pend_list;
  if (t == 0) base = 0;
  __syncthreads();
  for (int b0 = 0; b0 < S.B; b0 += 1024) {
    const int b = b0 + t;
    bool p = false;
    if (b < S.B) p = (mode == 0) ? (c.active[b] != 0) : (c.active[b] && c.do_ls[b] && !c.ls_found[b]);
    const unsigned m = __ballot_sync(0xffffffffu, p);
    if (lane == 0) warp_tot[w] = __popc(m);
    __syncthreads();
    int off = base;
    for (int i = 0; i < w; ++i) off += warp_tot[i];
    if (p) list[off + __popc(m & ((1u << lane) - 1u))] = b;
    __syncthreads();
    if (t == 0) { int sum = 0; for (int i = 0; i < 32; ++i) sum += warp_tot[i]; base += sum; }
    __syncthreads();
  }
  if (t == 0) *(mode == 0 ? c.n_active : c.n_pending) = base;
}

// mode 0: initial rollout bookkeeping (MultiPhaseDDP.cpp:238-261); mode 1: after a DDP iteration
__global__ void k_select(const SolverDev* __restrict__ Sp, int mode) {
  const SolverDev& S = *Sp;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= S.B) return;
  const CtrlDev& c = S.c;
  const CafeOptions& o = S.opt;
  const int ldb = S.ldb;
  c.reb_upd[b] = 0;
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
    // PathConstraintBase::update_params (ConstraintsBase.h:194-209, called at :538): decided here, applied per (knot, element) by
    // k_reb_update once k_accept has left the last rollout in X, U, Y - the trajectory the reference's constraint data belong to
    if (o.ReB_active) c.reb_upd[b] += 1;
    if (c.iter_ou[b] >= o.max_AL_iter) { c.active[b] = 0; break; }
    c.iter_ou[b] += 1; c.max_t_prev[b] = mt; c.max_p_prev[b] = mp; c.reg[b] = 0; c.iter_in[b] = 0;
    if (o.max_DDP_iter > 0) break;
  }
  if (c.active[b]) atomicAdd(c.n_active, 1);
}

// ------------------------------------------------------------------------------------------ K-REB-UPDATE
// PathConstraintBase::update_params (ConstraintsBase.h:194-209): every element whose value g of the last rollout is not above
// -pconstr_thresh gets eps *= update_ReB, delta = max(delta * update_relax, delta_min) - here: its update count grows (RebCtx::get
// replays the multiplications). Thread per (problem, running knot); launched only when the parameters can change (reb_dyn).
// g as SinglePhase::hybrid_rollout left it (compute_path_constraints on X[k], U[k], Y[k], SinglePhase.cpp:222): after k_accept
// these are the arrays X, U, Y. Elements that do not exist for a phase (swing feet, switched-off constraint sets) are never read.
__global__ void k_reb_update(const SolverDev* __restrict__ Sp) {
  const SolverDev& S = *Sp;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = (int)(t % S.ldb);
  const int gk = (int)(t / S.ldb);
  if (gk >= S.n_knots || b >= S.B) return;
  const int cnt = S.c.reb_upd[b];
  if (cnt <= 0) return;
  const int pi = S.knot_phase[gk], k = S.knot_k[gk];
  const PhaseDev& ph = S.ph[pi];
  if (k >= ph.h || !ph.reb_dyn) return;
  const int ldb = S.ldb;
  const double thresh = S.opt.pconstr_thresh;
  unsigned char* np = ph.reb_n + ((size_t)k * ph.reb_ne) * (size_t)ldb + b;
  auto hit = [&](int e, double g) {
    if (g > -thresh) return;
    const int v = np[(size_t)e * ldb] + cnt;
    np[(size_t)e * ldb] = (unsigned char)(v > 255 ? 255 : v);
  };
  auto grf5 = [&](int e0, double fx, double fy, double fz) {
    const double mu = ph.mu;
    hit(e0, fz); hit(e0 + 1, -fx + mu * fz); hit(e0 + 2, fx + mu * fz); hit(e0 + 3, -fy + mu * fz); hit(e0 + 4, fy + mu * fz);
  };
  if (ph.model == CAFE_MODEL_HKD) {
    for (int leg = 0; leg < 4; ++leg)
      if (ph.contact[leg] > 0) grf5(5 * leg, ph.U[gix(k, 24, 3 * leg, ldb, b)], ph.U[gix(k, 24, 3 * leg + 1, ldb, b)], ph.U[gix(k, 24, 3 * leg + 2, ldb, b)]);
  } else if (ph.model == CAFE_MODEL_SRB) {
    hit(0, ph.X[gix(k, 12, 2, ldb, b)] - ph.h_min);
  } else {
    for (int i = 0; i < 12; ++i) { const double u = ph.U[gix(k, 12, i, ldb, b)]; hit(i, -u - (-ph.torque_limit)); hit(12 + i, u - (-ph.torque_limit)); }
    if (ph.joint_speed_limit) for (int i = 0; i < 12; ++i) { const double v = ph.X[gix(k, 36, 24 + i, ldb, b)]; hit(24 + i, v - ph.jointvel_lb); hit(36 + i, -v - (-ph.jointvel_ub)); }
    if (!ph.no_joint_limit) for (int i = 0; i < 12; ++i) { const double q = ph.X[gix(k, 36, 6 + i, ldb, b)]; hit(48 + i, q - ph.joint_lb[i % 3]); hit(60 + i, -q - (-ph.joint_ub[i % 3])); }
    if (!ph.no_min_height) hit(72, ph.X[gix(k, 36, 2, ldb, b)] - ph.h_min);
    for (int f = 0; f < 4; ++f)
      if (ph.contact[f] > 0) grf5(73 + 5 * f, ph.Y[gix(k, 12, 3 * f, ldb, b)], ph.Y[gix(k, 12, 3 * f + 1, ldb, b)], ph.Y[gix(k, 12, 3 * f + 2, ldb, b)]);
  }
}

}  // namespace cafe_dev
