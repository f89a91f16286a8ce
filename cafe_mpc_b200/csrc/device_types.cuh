// device_types.cuh — device-side view of a batch of HS-DDP problems that share one phase deck.
//
// HBM layout: every per-problem quantity is stored component-major with the PROBLEM index fastest
// ("SoA over the batch"): element c of knot k of problem b lives at ((k*NC + c)*ldb + b). A warp
// therefore always touches 32 consecutive problems of the same scalar => every global access of the
// per-(problem,knot) kernels is a fully coalesced 256 B transaction, and the per-problem-group
// Riccati kernel reads 8*PB-byte sector-aligned runs (PB problems per CTA).
#pragma once
#include <cuda_runtime.h>
#include "../../include/cafe_deck.h"

#define CAFE_MAX_ALPHAS 12
// problem-major tiles of the whole-body sweep (doubles): [A B] rows 18..35 as 20 x 48 (ld 20), [C D] 12 x 48, K 12 x 36
#define CAFE_WB_AB_TILE 960
#define CAFE_WB_CD_TILE 576
#define CAFE_WB_K_TILE 432
#define CAFE_LXX_MASK_WORDS 21
// whole-body hand-off arrays of the leg-parallel kernels (wb_leg_kernels.cu -> wb_coop.cuh), batch-major [slot][ldb]:
//   terms pack  TM = [trunk: nle 6 | M 6x6 (r + 6c)] [leg f: CAFE_WBL_TM_W compact slots] x 4      (gen/wb_leg_gen.h)
//   deriv pack  DP = [trunk: dtau/dq 3x3 | dtau/dv 3x3 (rows, columns 3..5)] [leg f: CAFE_WBL_DP_W compact slots] x 4
#define CAFE_TM_LEG_W 79
#define CAFE_DP_LEG_W 198
#define CAFE_TM_TRUNK_W 42
#define CAFE_DP_TRUNK_W 18
#define CAFE_TM_W (CAFE_TM_TRUNK_W + 4 * CAFE_TM_LEG_W)
#define CAFE_DP_W (CAFE_DP_TRUNK_W + 4 * CAFE_DP_LEG_W)
#define CAFE_LXX_LIST_CAP 704
#define CAFE_MAX_KNOTS 256
#define CAFE_HIST_CAP 320  // LocoProblem settings: 30 x 10 iterations + the initial entry

struct PhaseDev {
  int model, n, m, p, h, n_next, has_next;
  int contact[4], next_contact[4], n_td, td_foot[4];
  int no_joint_limit, no_min_height, joint_speed_limit, single_shooting;
  CafeRebParam reb_jointvel;  // BarrelRoll::JointSpeedLimit
  double jointvel_lb, jointvel_ub;
   // WB path-constraint set (CafePhase: LocoProblem drops the joint-limit and min-height barriers)
  double dt, mu, ground_height, BG_alpha, h_min, torque_limit, joint_lb[3], joint_ub[3];
  double q[CAFE_MAX_N], r[CAFE_MAX_M], qf[CAFE_MAX_N];
  double w_footreg[3], w_swingpos[3], w_swingvel[3], w_tdvel[3];
  CafeRebParam reb_grf, reb_torque, reb_joint, reb_minheight;
  CafeAlParam al_td;
  const double* ref;     // [h+1][CAFE_REF_W], shared by the batch
  const double* ref_pp;  // optional per-problem records [h+1][CAFE_REF_W][ldb] (same contact schedule), else nullptr
  // trajectories, [(h+1) or h][dim][ldb]
  double *X, *Xbar, *dX, *G, *Defect;
  double *U, *Ubar, *dU, *Qu, *Y;
  // LQ data
  double *A, *Bm, *C, *D;                  // [h][n*n | n*m | p*n | p*m][ldb], column-major per knot
  double *lx, *lu, *ly, *lxx, *luu, *lyy;  // [h][...][ldb]
  double *Phix, *Phixx, *Px;               // [n | n*n | n_next*n][ldb]
  // WB only: rigid-body terms of every line-search trial (k_wb_terms -> k_wb_fwd; the accepted trial's are reused by the next
  // linearisation), joint accelerations of every trial, derivative pieces of the current iterate (k_wb_derivs -> k_wb_lq)
  double *tm;                              // [NA][h][CAFE_TM_W][ldb]
  double *qdd_t;                           // [NA][h][18][ldb]
  double *dp;                              // [h][CAFE_DP_W][ldb]
  // WB only, PROBLEM-major copies laid out exactly like the shared-memory tiles of the backward sweep (bwd2.cuh), so that a CTA
  // fetches them with 16-byte cp.async from contiguous memory (an 8-byte element of a problem-fastest array costs one L1 wavefront
  // each): [b][h][20 x 48] rows 18..35 of [A B] (rows 18, 19 of the tile: zero padding), [b][h][12 x 48] [C D], [b][h][12 x 36] K
  double *ABpm, *CDpm, *Kpm;
  // WB only: structural non-zero pattern of lxx per knot (bit i + 36 j of 21 64-bit words), built on the host from the contact
  // flags of the knot's reference record (the rule of WBModel::lq_knot). The sweep fetches only these entries of lxx; luu is
  // diagonal and lyy has one 3 x 3 block per foot for this model.
  const unsigned long long* lxx_mask;
  // HKD only: structural patterns of A, B, lxx, luu (576-bit masks, 9 words each, the same for every knot and phase)
  const unsigned long long* hkd_mask;
  double *lk, *dsq;                        // [h+1][ldb] per-knot cost (k=h: Phi) and |Defect[k]|^2
  // backward-sweep outputs
  double *K;                               // [h][m*n][ldb]
  double *Quu, *Qux;                       // PROBLEM-major tiles [b][h][ld(m) x m | ld(m) x n] (read by the result packers only)
  // line-search trial slots, one per step size
  double *Xt, *Ut, *Yt, *Dt;               // [NA][(h+1)|h][dim][ldb]
  double *cost_t, *feas_t, *ming_t;        // [NA][h+1][ldb]
  double *maxh_t, *ht;                     // [NA][ldb], [NA][4][ldb]
  int* fail_t;                             // [NA][ldb]
  // augmented-Lagrangian state of the touchdown constraints
  double *al_sigma, *al_lambda, *hval;     // [4][ldb]
  // relaxed-barrier parameters per (running knot, path-constraint element, problem) (PathConstraintBase::params, ConstraintsBase.h:173-209):
  // update_params multiplies eps by update_ReB and delta by update_relax (floored at delta_min) for every element that is violated at the
  // end of an outer iteration, so a parameter is the deck's initial value after n such updates - the count n is what is stored.
  // Elements: HKD 5 per leg (20); SRB 1; WB torque 24 | joint speed 24 | joint 24 | min height 1 | GRF 5 per foot (93).
  unsigned char* reb_n;                    // [h][reb_ne][ldb]
  int reb_ne, reb_dyn;                     // reb_dyn = 0: every parameter keeps its initial value (update factors 1, the shipped settings)
  double reb_br, reb_bw;                   // update_relax, update_ReB
};

// the relaxed-barrier parameters of one (problem, knot): element e -> (delta, eps)
struct RebCtx {
  const unsigned char* n; size_t st; double br, bw;
  __device__ __forceinline__ void get(const CafeRebParam& p0, int e, double& delta, double& eps) const {
    delta = p0.delta; eps = p0.eps;
    if (n) {
      const int c = n[(size_t)e * st];
      for (int i = 0; i < c; ++i) { eps *= bw; delta *= br; delta = fmax(delta, p0.delta_min); }   // update_weight, update_relax (:79-85)
    }
  }
};

struct CtrlDev {
  int *active, *do_ls, *sel, *accepted;
  int *ls_found, *ls_fail;            // staged line search: first successful step size found / divergence flag of the kept trial
  double *ls_cost, *ls_feas, *ls_mt, *ls_mp, *ls_merit;  // summary of the kept (first successful, else last evaluated) trial
  int* n_pending;                     // problems that still need more step sizes
  int *iter_ou, *iter_in, *iter, *ls_total, *reg_total, *n_hist, *status;
  double *reg, *cost, *merit, *feas, *merit_rho, *dV1, *dV2, *cost_prev, *merit_prev;
  double *max_t, *max_p, *max_t_prev, *max_p_prev;
  double *hist;   // [CAFE_HIST_CAP][4][ldb]
  double *trace;  // [CAFE_HIST_CAP][12][ldb]
  double *feas0_t;  // [NA][ldb] |Defect_0[0]|^2 of the first phase per trial
  double *min_pivot;  // [ldb] smallest LDL^T pivot seen (tie monitoring)
  int* n_active;  // single counter
  // index lists (ascending problem index), rebuilt by k_compact: problems still iterating / problems whose line search needs
  // more step sizes. The per-(problem, knot) kernels and the sweep run over these lists, so that warps and CTAs stay full when
  // most of the batch has already converged.
  int *act_list, *pend_list;
  int *cur_slot;  // trial slot whose rollout produced the current iterate X, U (its rigid-body terms are reused by the linearisation)
  int *reb_upd;   // relaxed-barrier updates decided by k_select in this tick (applied per knot by k_reb_update)
};

struct SolverDev {
  int n_phases, B, ldb, NA, n_knots;
  int n_act;  // length of c.act_list (host copy only: set before k_bwd2 is launched with the descriptor by value)
  double eps[CAFE_MAX_ALPHAS];
  CafeOptions opt;
  PhaseDev ph[CAFE_MAX_PHASES];
  CtrlDev c;
  const double* x0;  // [n0][ldb]
  short knot_phase[CAFE_MAX_KNOTS], knot_k[CAFE_MAX_KNOTS];
  short wbk_gk[CAFE_MAX_KNOTS];  // global knot indices of the running whole-body knots (k < h of WB phases)
  int n_wbk;
};

// reference record of knot k as problem b sees it: the deck's shared record, or the problem's own (copied to thread-local storage)
__device__ __forceinline__ const double* knot_record(const PhaseDev& ph, int k, int ldb, int b, double* local) {
  if (!ph.ref_pp) return ph.ref + (size_t)k * CAFE_REF_W;
  for (int c = 0; c < CAFE_REF_W; ++c) local[c] = ph.ref_pp[((size_t)k * CAFE_REF_W + c) * (size_t)ldb + b];
  return local;
}

// element (k, c) of problem b in an array with NC components per knot
__device__ __forceinline__ size_t gix(int k, int NC, int c, int ldb, int b) { return ((size_t)k * NC + c) * (size_t)ldb + b; }
__device__ __forceinline__ RebCtx reb_ctx(const PhaseDev& ph, int k, int ldb, int b) {
  return RebCtx{ph.reb_dyn ? ph.reb_n + ((size_t)k * ph.reb_ne) * (size_t)ldb + b : nullptr, (size_t)ldb, ph.reb_br, ph.reb_bw};
}
__host__ __device__ inline int cafe_reb_elements(int model) { return model == CAFE_MODEL_WB ? 93 : model == CAFE_MODEL_HKD ? 20 : 1; }
