// wb_coop.cuh — warp-cooperative, shared-memory part of the whole-body (WB) knots: one WARP per (problem, knot[, step size]).
//
// The straight-line rigid-body routines run one thread per (problem, knot, leg) in wb_leg_kernels.cu and hand their compact outputs
// over in batch-major arrays (PhaseDev::tm, ::dp). Everything that used to be thread-local dense algebra (8 KB of local memory per
// thread, ncu round 1: 106 M local loads per k_lq launch, 19.5 GB of DRAM traffic) runs here out of shared memory and registers:
//
//   k_wb_fwd  (rollout)        KKTContactDynamics: chol(M), Y = L^-1 Jc^T, chol(Y^T Y + 1e-12 I), lambda, qdd; x+, GRF, running cost,
//                              barrier minima, defects                          <= WBM.cpp:17-57, :368-424; SinglePhase.cpp:200-232
//   k_wb_sens (linearisation)  KKT sensitivities  dlambda/dz = S^-1 (Jc M^-1 R - a), dqdd/dz = -M^-1 (R - Jc^T dlambda/dz) column by
//                              column -> A, B, C, D tiles                      <= WBM.cpp:60-139, :459-505
//   k_wb_cost (linearisation)  cost / barrier partials lx, lu, ly, lxx (structural pattern), luu, lyy, running cost
//                                                                               <= MHPCCost.cpp:4-291; MHPCConstraint.cpp:9-288;
//                                                                                  SinglePhase.cpp:236-320, :405-418
// Arithmetic order: sums that the thread-per-knot code accumulated sequentially (base-block shares of M / nle / dtau, Cholesky
// updates, forward substitutions, cost terms) are accumulated in the same order here; the back substitutions run as column
// updates (descending), which differs from the row form in rounding only.
#pragma once
#include "device_types.cuh"
#include "model_hkd.cuh"
#include "wb_leg_tables.h"

namespace cafe_dev {

#define CAFE_FULL 0xffffffffu
static_assert(CAFE_TM_LEG_W == CAFE_WBL_TM_W && CAFE_DP_LEG_W == CAFE_WBL_DP_W, "device_types.cuh and gen/wb_leg_gen.h disagree");

// compact slot -> (array, local row, local column). The lanes of a warp read DIFFERENT slots: plain global arrays (one coalesced,
// L1-resident load per warp), not __constant__ (a constant-bank access with 32 distinct addresses is replayed 32 times).
static __device__ const unsigned char c_tm_kind[CAFE_WBL_TM_W] = CAFE_WBL_TM_KIND;
static __device__ const unsigned char c_tm_row[CAFE_WBL_TM_W] = CAFE_WBL_TM_ROW;
static __device__ const unsigned char c_tm_col[CAFE_WBL_TM_W] = CAFE_WBL_TM_COL;
static __device__ const unsigned char c_dp_kind[CAFE_WBL_DP_W] = CAFE_WBL_DP_KIND;
static __device__ const unsigned char c_dp_row[CAFE_WBL_DP_W] = CAFE_WBL_DP_ROW;
static __device__ const unsigned char c_dp_col[CAFE_WBL_DP_W] = CAFE_WBL_DP_COL;

// global coordinate of local coordinate l (0..5 base, 6..8 leg) of leg f
__device__ __forceinline__ int wbl_g(int f, int l) { return l < 6 ? l : 3 * f + l; }

// ---- per-warp shared-memory plan (doubles). Three kernels share the first block; occupancy is what bounds these latency-bound kernels,
// so every kernel keeps only what it needs: k_wb_fwd 8.8 KB, k_wb_sens 17.6 KB, k_wb_cost 11 KB per (problem, knot).
struct WbSm {
  static constexpr int ldL = 19, ldJ = 12, ldY = 19, ldS = 13;
  static constexpr int oL = 0;                    // M (lower) -> chol(M), 18 x 18, ld 19
  static constexpr int oDinv = oL + 18 * ldL;     // 1 / L(i,i)
  static constexpr int oJ = oDinv + 18;           // foot Jacobians, rows 3f+r, 12 x 18, ld 12
  static constexpr int oNle = oJ + 216;
  static constexpr int oGam = oNle + 18;
  static constexpr int oPf = oGam + 12;
  static constexpr int oVf = oPf + 12;
  static constexpr int oX = oVf + 12;             // x 36 | u 12 | grf 12
  static constexpr int oU = oX + 36;
  static constexpr int oGrf = oU + 12;
  static constexpr int oStg = oGrf + 12;          // staging of the packs; after the assembly: Y | S | 1/Ls(i,i) | mb
  static constexpr int oY = oStg;                 // Y = L^-1 Jc^T, 18 x nr, ld 19
  static constexpr int oS = oY + 12 * ldY;        // S -> chol(S), nr x nr, ld 13
  static constexpr int oSdinv = oS + 12 * ldS;
  static constexpr int oMb = oSdinv + 12;         // L^-1 (tau - nle)
  static constexpr int szStg = oMb + 18 - oStg;   // 414 = trunk + two legs of the derivative pack
  static_assert(szStg >= CAFE_TM_W && szStg >= CAFE_DP_TRUNK_W + 2 * CAFE_DP_LEG_W, "staging area");
  // rollout: the cost scratch (160) reuses the Jacobian tile, dead once Y is formed
  static constexpr int oScrFwd = oJ;
  static constexpr int totalFwd = (oStg + szStg + 1) & ~1;
  // sensitivities
  static constexpr int oR = totalFwd;                  // R 18 x 36 (ld 18) -> results
  static constexpr int oA = oR + 648;                  // a nr x 36 (ld 12) -> results
  static constexpr int oJt = oA + 432;                 // shared 3 x 3 block of d(J^T F)/dq
  static constexpr int totalSens = (oJt + 10 + 1) & ~1;
  // cost partials
  static constexpr int oDvq = totalFwd;                // dv_foot/dq, compact: foot f at 18 f, element r + 3 (local column - 3)
  static constexpr int oScr = oDvq + 72;               // scratch of the cost routine (160)
  static constexpr int oLx = oScr + 160;               // dposw 12 | dvelw 12 | dg 36
  static constexpr int totalCost = (oLx + 64 + 1) & ~1;
};

// the knot's reference record as a problem sees it: the deck's shared record (stride 1) or the problem's own (stride ldb)
struct RecRef {
  const double* p; size_t st;
  __device__ __forceinline__ double operator[](int i) const { return p[(size_t)i * st]; }
};
__device__ __forceinline__ RecRef wb_rec(const PhaseDev& ph, int k, int ldb, int b) {
  if (ph.ref_pp) return RecRef{ph.ref_pp + (size_t)k * CAFE_REF_W * (size_t)ldb + b, (size_t)ldb};
  return RecRef{ph.ref + (size_t)k * CAFE_REF_W, 1};
}

// stage W elements of a batch-major array (element e at src[e * ldb]) into shared memory, 32 element-threads per problem: every
// load is issued before the first store, so that one memory round trip covers the whole pack
template <int W>
__device__ __forceinline__ void wb_stage(double* __restrict__ dst, const double* __restrict__ src, int ldb, int e0) {
  constexpr int NI = (W + 31) / 32;
  double v[NI];
#pragma unroll
  for (int i = 0; i < NI; ++i) { const int e = e0 + 32 * i; v[i] = (e < W) ? src[(size_t)e * ldb] : 0.0; }
#pragma unroll
  for (int i = 0; i < NI; ++i) { const int e = e0 + 32 * i; if (e < W) dst[e] = v[i]; }
}

// in-place Cholesky of the lower triangle (column-major, ld LD), n <= 32: lane i owns row j + i of the current column
// dinv[j] = 1 / L(j,j)
template <int LD, int NMAX>
__device__ __forceinline__ void chol_warp(double* A, double* dinv, int n, int lane) {
#pragma unroll
  for (int j = 0; j < NMAX; ++j) {
    if (j >= n) break;
    const int i = j + lane;
    double s = 0;
    if (i < n) {
      s = A[i + LD * j];
#pragma unroll
      for (int k = 0; k < j; ++k) s -= A[i + LD * k] * A[j + LD * k];
    }
    const double d = sqrt(__shfl_sync(CAFE_FULL, s, 0));
    const double inv = 1.0 / d;
    if (i < n) A[i + LD * j] = (lane == 0) ? d : s * inv;
    if (lane == 0) dinv[j] = inv;
    __syncwarp();
  }
}

struct WbRows {   // active contact rows of a contact set: row(c) = 3 foot(c / 3) + c % 3 (bit arithmetic: no thread-local arrays)
  int nr; unsigned cm; unsigned long long rp;   // rp: row(c) in nibble c
  __device__ __forceinline__ void set(const int* contact) {
    cm = 0; rp = 0; int j = 0;
    for (int f = 0; f < 4; ++f) if (contact[f] > 0) { cm |= 1u << f; for (int r = 0; r < 3; ++r) rp |= (unsigned long long)(3 * f + r) << (4 * (3 * j + r)); ++j; }
    nr = 3 * j;
  }
  __device__ __forceinline__ int row(int c) const { return (int)((rp >> (4 * c)) & 15ull); }
  // index among the active rows of foot row fr = 3 f + r, or -1
  __device__ __forceinline__ int arow(int fr) const { const int f = fr / 3; return ((cm >> f) & 1u) ? 3 * __popc(cm & ((1u << f) - 1u)) + fr % 3 : -1; }
};

// staged terms pack (stg[0..CAFE_TM_W)) -> M (lower, zero elsewhere), nle, J, gam, pf, vf. Shares of the base block are added in
// the order trunk, leg 0..3 (MassDst / BiasDst of wb_pieces.h onto a zero-initialised destination).
template <bool WITH_M = true>
__device__ __forceinline__ void wb_assemble_terms(double* sm, const double* stg, int lane) {
  double* M = sm + WbSm::oL; double* J = sm + WbSm::oJ; double* nle = sm + WbSm::oNle;
  if (WITH_M) for (int e = lane; e < 18 * WbSm::ldL; e += 32) M[e] = 0.0;
  for (int e = lane; e < 216; e += 32) J[e] = 0.0;
  __syncwarp();
  if (WITH_M) for (int e = lane; e < 36; e += 32) { const int r = e % 6, c = e / 6; if (r >= c) M[r + WbSm::ldL * c] = stg[6 + e]; }
  if (lane < 6) nle[lane] = stg[lane];
  __syncwarp();
  for (int f = 0; f < 4; ++f) {
    const double* lg = stg + CAFE_TM_TRUNK_W + f * CAFE_WBL_TM_W;
    for (int e = lane; e < CAFE_WBL_TM_W; e += 32) {
      const int kind = c_tm_kind[e], r = c_tm_row[e], c = c_tm_col[e];
      const double v = lg[e];
      if (kind == 0) { const int i = wbl_g(f, r); if (r < 6) nle[i] += v; else nle[i] = v; }
      else if (kind == 1) { if (WITH_M) { const int idx = wbl_g(f, r) + WbSm::ldL * wbl_g(f, c); if (r < 6 && c < 6) M[idx] += v; else M[idx] = v; } }
      else if (kind == 2) J[(3 * f + r) + WbSm::ldJ * wbl_g(f, c)] = v;
      else if (kind == 3) sm[WbSm::oGam + 3 * f + r] = v;
      else if (kind == 4) sm[WbSm::oPf + 3 * f + r] = v;
      else sm[WbSm::oVf + 3 * f + r] = v;
    }
    __syncwarp();
  }
}

// chol(M) (in place), 1/diag, Y = L^-1 Jc^T for the rows of `rw` (+ one more right-hand side `extra` -> mb when given),
// S = Y^T Y + damping I, chol(S), 1/diag
__device__ __forceinline__ void wb_factor(double* sm, const WbRows& rw, double damping, bool with_rhs, int lane) {
  double* L = sm + WbSm::oL; const double* J = sm + WbSm::oJ; double* Y = sm + WbSm::oY; double* S = sm + WbSm::oS;
  chol_warp<WbSm::ldL, 18>(L, sm + WbSm::oDinv, 18, lane);
  const int nr = rw.nr;
  const double* dinv = sm + WbSm::oDinv;
  if (lane < nr || (with_rhs && lane == nr)) {
    double y[18];
    if (lane < nr) { const int row = rw.row(lane);
#pragma unroll
      for (int i = 0; i < 18; ++i) y[i] = J[row + WbSm::ldJ * i];
    } else {   // tau - nle, tau = [0; u]
#pragma unroll
      for (int i = 0; i < 18; ++i) y[i] = ((i >= 6) ? sm[WbSm::oU + i - 6] : 0.0) - sm[WbSm::oNle + i];
    }
#pragma unroll
    for (int i = 0; i < 18; ++i) {
      double s = y[i];
#pragma unroll
      for (int k = 0; k < i; ++k) s -= L[i + WbSm::ldL * k] * y[k];
      y[i] = s * dinv[i];
    }
    double* dst = (lane < nr) ? Y + WbSm::ldY * lane : sm + WbSm::oMb;
#pragma unroll
    for (int i = 0; i < 18; ++i) dst[i] = y[i];
  }
  __syncwarp();
  for (int e = lane; e < nr * nr; e += 32) {   // lower triangle of S, one entry per lane and round
    const int r = e % nr, c = e / nr;
    if (r >= c) {
      double d = 0;
#pragma unroll
      for (int i = 0; i < 18; ++i) d += Y[i + WbSm::ldY * r] * Y[i + WbSm::ldY * c];
      S[r + WbSm::ldS * c] = d + ((r == c) ? damping : 0.0);
    }
  }
  __syncwarp();
  chol_warp<WbSm::ldS, 12>(S, sm + WbSm::oSdinv, nr, lane);
}

// running cost of a whole-body knot (QuadraticTrackingCost + foot costs + dt * ReB terms) and the minimum of the path-constraint
// values, from x, u, y (= GRF), pf, vf, rec in shared memory; same terms and summation order as WBModel::running_cost_k
__device__ __forceinline__ double wb_cost_coop(const PhaseDev& ph, const double* sm, const RecRef rec, double* scr, bool reb, int lane, double& ming, const RebCtx& rcx) {
  const double* x = sm + WbSm::oX; const double* u = sm + WbSm::oU; const double* y = sm + WbSm::oGrf;
  const double* pf = sm + WbSm::oPf; const double* vf = sm + WbSm::oVf;
  const double dt = ph.dt;
  for (int i = lane; i < 36; i += 32) { const double dx = x[i] - rec[CAFE_REF_XR + i]; scr[i] = dx * ph.q[i] * dx; }
  if (lane < 12) { const double du = u[lane] - rec[CAFE_REF_UR + lane]; scr[36 + lane] = du * ph.r[lane] * du; }
  if (lane >= 16 && lane < 20) {
    const int f = lane - 16;
    const bool c = rec[CAFE_REF_CONTACT + f] > 0;
    const double* w = c ? ph.w_footreg : ph.w_swingpos;
    double q2 = 0;
    for (int a = 0; a < 3; ++a) { const double d = (pf[3 * f + a] - x[a]) - (rec[CAFE_REF_PF + 3 * f + a] - rec[CAFE_REF_PCOM + a]); q2 += d * w[a] * d; }
    double t = .5 * q2; t *= dt;
    double t2 = 0;
    if (!c) { double q3 = 0; for (int a = 0; a < 3; ++a) { const double dv = vf[3 * f + a] - rec[CAFE_REF_VF + 3 * f + a]; q3 += dv * ph.w_swingvel[a] * dv; } t2 = .5 * q3; t2 *= dt; }
    scr[48 + f] = t; scr[52 + f] = t2;
  }
  // path constraints in the reference's order: torque 24 | joint speed 24 | joint 24 | min height 1 | GRF 5 per foot
  const bool jl = !ph.no_joint_limit, mh = !ph.no_min_height, jv = ph.joint_speed_limit != 0;
  double mn = 0;
  for (int e = lane; e < 93; e += 32) {
    double g = 0, delta = 1, eps = 0; bool on = true;
    const CafeRebParam* p0;
    if (e < 24) { const int i = e % 12; g = (e < 12 ? -u[i] : u[i]) + ph.torque_limit; p0 = &ph.reb_torque; }
    else if (e < 48) { const int i = (e - 24) % 12; on = jv; g = (e < 36) ? x[24 + i] - ph.jointvel_lb : -x[24 + i] + ph.jointvel_ub; p0 = &ph.reb_jointvel; }
    else if (e < 72) { const int i = (e - 48) % 12; on = jl; g = (e < 60) ? x[6 + i] - ph.joint_lb[i % 3] : -x[6 + i] + ph.joint_ub[i % 3]; p0 = &ph.reb_joint; }
    else if (e == 72) { on = mh; g = x[2] - ph.h_min; p0 = &ph.reb_minheight; }
    else {
      const int f = (e - 73) / 5, i = (e - 73) % 5;
      on = ph.contact[f] > 0;
      const double fx = y[3 * f], fy = y[3 * f + 1], fz = y[3 * f + 2], mu = ph.mu;
      g = (i == 0) ? fz : (i == 1) ? -fx + mu * fz : (i == 2) ? fx + mu * fz : (i == 3) ? -fy + mu * fz : fy + mu * fz;
      p0 = &ph.reb_grf;
    }
    double val = 0;
    if (on) { rcx.get(*p0, e, delta, eps); mn = fmin(mn, g); val = eps * reb_value(g, delta); }
    scr[56 + e] = val;
  }
  for (int o = 16; o > 0; o >>= 1) mn = fmin(mn, __shfl_xor_sync(CAFE_FULL, mn, o));
  ming = mn;
  __syncwarp();
  // six independent partial sums, each sequential in the reference's order
  double part = 0;
  if (lane == 0) { for (int i = 0; i < 36; ++i) part += scr[i]; }
  else if (lane == 1) { for (int i = 0; i < 12; ++i) part += scr[36 + i]; }
  else if (lane == 2) { for (int i = 0; i < 24; ++i) part += scr[56 + i]; }
  else if (lane == 3) { for (int i = 0; i < 24; ++i) part += scr[80 + i]; }
  else if (lane == 4) { for (int i = 0; i < 24; ++i) part += scr[104 + i]; }
  else if (lane == 5) { for (int i = 0; i < 20; ++i) part += scr[129 + i]; }
  const double sx = __shfl_sync(CAFE_FULL, part, 0), su = __shfl_sync(CAFE_FULL, part, 1), ct = __shfl_sync(CAFE_FULL, part, 2),
               cv = __shfl_sync(CAFE_FULL, part, 3), cj = __shfl_sync(CAFE_FULL, part, 4), cg = __shfl_sync(CAFE_FULL, part, 5);
  double l = 0.5 * sx;
  l += 0.5 * su;
  l *= dt;
  double lreg = 0, lpos = 0, lvel = 0;
  bool any = false;
  for (int f = 0; f < 4; ++f) {
    if (rec[CAFE_REF_CONTACT + f] > 0) lreg += scr[48 + f]; else { lpos += scr[48 + f]; lvel += scr[52 + f]; }
    any = any || ph.contact[f] > 0;
  }
  l += lreg; l += lpos; l += lvel;
  if (reb) { l += dt * ct; if (jv) l += dt * cv; if (jl) l += dt * cj; if (mh) l += dt * scr[128]; if (any) l += dt * cg; }
  __syncwarp();
  return l;
}

// cooperative staging: the four warps of a CTA serve four consecutive list entries; thread t loads elements t/4, t/4 + 32, ... of
// problem t % 4, so that the four 8-byte words of every 32-byte sector of the batch-major arrays are requested together
#define CAFE_WB_PKS 4

// ------------------------------------------------------------------------------------------------ K-ROLL, whole-body running knots
// grid (ceil(n_list / 4), n_wbk, a1 - a0), 128 threads
__global__ void __launch_bounds__(128, 6) k_wb_fwd(const SolverDev* __restrict__ Sp, int a0, const int* __restrict__ list, int n_list) {
  const SolverDev& S = *Sp;
  extern __shared__ __align__(16) double smem[];
  const int t = threadIdx.x, w = t >> 5, lane = t & 31;
  const int gk = S.wbk_gk[blockIdx.y], pi = S.knot_phase[gk], k = S.knot_k[gk], a = a0 + blockIdx.z;
  const PhaseDev& ph = S.ph[pi];
  const int ldb = S.ldb, h = ph.h;
  const size_t aX = (size_t)a * (h + 1) * 36 * ldb, aU = (size_t)a * h * 12 * ldb;
  {
    const int p = t & 3, jp = blockIdx.x * CAFE_WB_PKS + p;
    const int b = jp < n_list ? list[jp] : -1;
    if (b >= 0 && S.c.active[b] && S.c.do_ls[b] && !S.c.ls_found[b]) {
      double* sm = smem + p * WbSm::totalFwd;
      double* stg = sm + WbSm::oStg;
      const double* tm = ph.tm + ((size_t)(a * h + k) * CAFE_TM_W) * ldb + b;
      wb_stage<CAFE_TM_W>(stg, tm, ldb, t >> 2);
      for (int e = t >> 2; e < 36; e += 32) sm[WbSm::oX + e] = ph.Xt[aX + gix(k, 36, e, ldb, b)];
      for (int e = t >> 2; e < 12; e += 32) sm[WbSm::oU + e] = ph.Ut[aU + gix(k, 12, e, ldb, b)];
    }
  }
  __syncthreads();
  const int j = blockIdx.x * CAFE_WB_PKS + w;
  if (j >= n_list) return;
  const int b = list[j];
  if (!S.c.active[b] || !S.c.do_ls[b] || S.c.ls_found[b]) return;
  double* sm = smem + w * WbSm::totalFwd;
  wb_assemble_terms(sm, sm + WbSm::oStg, lane);
  WbRows rw; rw.set(ph.contact);
  const int nr = rw.nr;
  wb_factor(sm, rw, 1e-12, true, lane);
  const double* L = sm + WbSm::oL; const double* Y = sm + WbSm::oY; const double* Ss = sm + WbSm::oS;
  // lambda = -(Jc Minv (tau - nle)) - (gamma + 2 BG v_foot), then (Ls Ls^T) lambda
  double lam = 0;
  if (lane < nr) {
    double d = 0;
#pragma unroll
    for (int i = 0; i < 18; ++i) d += Y[i + WbSm::ldY * lane] * sm[WbSm::oMb + i];
    const int r = rw.row(lane);
    lam = -d - (sm[WbSm::oGam + r] + 2.0 * ph.BG_alpha * sm[WbSm::oVf + r]);
  }
  const double* sdinv = sm + WbSm::oSdinv; const double* dinv = sm + WbSm::oDinv;
  for (int i = 0; i < nr; ++i) {
    const double xi = __shfl_sync(CAFE_FULL, lam, i) * sdinv[i];
    if (lane == i) lam = xi; else if (lane > i && lane < nr) lam -= Ss[lane + WbSm::ldS * i] * xi;
  }
  for (int i = nr - 1; i >= 0; --i) {
    const double xi = __shfl_sync(CAFE_FULL, lam, i) * sdinv[i];
    if (lane == i) lam = xi; else if (lane < i) lam -= Ss[i + WbSm::ldS * lane] * xi;
  }
  // qdd = L^-T (L^-1 b + Y lambda)
  double mb = lane < 18 ? sm[WbSm::oMb + lane] : 0.0;
  for (int c = 0; c < nr; ++c) { const double lc = __shfl_sync(CAFE_FULL, lam, c); if (lane < 18) mb += Y[lane + WbSm::ldY * c] * lc; }
  for (int i = 17; i >= 0; --i) {
    const double xi = __shfl_sync(CAFE_FULL, mb, i) * dinv[i];
    if (lane == i) mb = xi; else if (lane < i) mb -= L[i + WbSm::ldL * lane] * xi;
  }
  // outputs: GRF (y), qdd, x+ and the defect against the next shooting state
  const double dt = ph.dt, eps = S.eps[a];
  const size_t aY = (size_t)a * h * 12 * ldb, aS = (size_t)a * (h + 1) * ldb;
  if (lane < 12) sm[WbSm::oGrf + lane] = 0.0;
  __syncwarp();
  if (lane < nr) sm[WbSm::oGrf + rw.row(lane)] = lam;
  __syncwarp();
  if (lane < 12) ph.Yt[aY + gix(k, 12, lane, ldb, b)] = sm[WbSm::oGrf + lane];
  double* scr = sm + WbSm::oScrFwd;   // the Jacobian tile is dead (Y is formed)
  const double* x = sm + WbSm::oX;
  if (lane < 18) {
    ph.qdd_t[((size_t)(a * h + k) * 18 + lane) * ldb + b] = mb;
    scr[lane] = x[lane] + x[18 + lane] * dt;
    scr[18 + lane] = x[18 + lane] + mb * dt;
  }
  __syncwarp();
  const bool ss = ph.single_shooting != 0;
  for (int i = lane; i < 36; i += 32) {
    const double xn = scr[i];
    const double xs = ss ? xn : ph.Xbar[gix(k + 1, 36, i, ldb, b)] + eps * ph.dX[gix(k + 1, 36, i, ldb, b)];
    const double d = xn - xs;
    ph.Dt[aX + gix(k + 1, 36, i, ldb, b)] = d;
    scr[36 + i] = d;
  }
  __syncwarp();
  if (lane == 0) { double dsq = 0; for (int i = 0; i < 36; ++i) dsq += scr[36 + i] * scr[36 + i]; ph.feas_t[aS + (size_t)k * ldb + b] = dsq; }
  if (lane == 1) { double nrm = 0; for (int i = 0; i < 36; ++i) nrm += scr[i] * scr[i]; if (sqrt(nrm) > 1e6) atomicOr(&ph.fail_t[(size_t)a * ldb + b], 1); }
  __syncwarp();
  double ming;
  const double l = wb_cost_coop(ph, sm, wb_rec(ph, k, ldb, b), scr, S.opt.ReB_active != 0, lane, ming, reb_ctx(ph, k, ldb, b));
  if (lane == 0) { ph.cost_t[aS + (size_t)k * ldb + b] = l; ph.ming_t[aS + (size_t)k * ldb + b] = ming; }
}

// column z of [q v tau] (lane = column, two passes over the 48 columns): r <- L^-1 R_z; w = Ls^-T Ls^-1 (Y^T r - a_z); r <- L^-T (Y w - r).
// NR = number of contact rows (compile time: the loops unroll without predicates). State columns go back into the lane's (dead) columns
// of R / a, control columns straight to the problem-major tiles.
template <int NR>
__device__ __forceinline__ void wb_colsolve(double* sm, double* __restrict__ ABt, double* __restrict__ CDt, const WbRows& rw, double dt, int lane) {
  // volatile: the ~700 operand loads of a pass stay in program order next to their use (fully unrolled, ptxas otherwise hoists
  // them all to the top of the pass and spills 4 KB per thread)
  const volatile double* L = sm + WbSm::oL; const volatile double* dinv = sm + WbSm::oDinv; const volatile double* Y = sm + WbSm::oY;
  const volatile double* Ls = sm + WbSm::oS; const volatile double* sdinv = sm + WbSm::oSdinv;
  double* R = sm + WbSm::oR; double* Aa = sm + WbSm::oA;
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    const int col = pass * 32 + lane;
    if (col >= 48) break;
    double r[18], wv[NR > 0 ? NR : 1];
    if (col < 36) {
#pragma unroll
      for (int i = 0; i < 18; ++i) r[i] = R[i + 18 * col];
    } else {
#pragma unroll
      for (int i = 0; i < 18; ++i) r[i] = (i == 6 + (col - 36)) ? -1.0 : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 18; ++i) {
      double s = r[i];
#pragma unroll
      for (int kk = 0; kk < i; ++kk) s -= L[i + WbSm::ldL * kk] * r[kk];
      r[i] = s * dinv[i];
    }
    if constexpr (NR > 0) {
#pragma unroll
      for (int c = 0; c < NR; ++c) {
        double d = 0;
#pragma unroll
        for (int i = 0; i < 18; ++i) d += Y[i + WbSm::ldY * c] * r[i];
        wv[c] = d - ((col < 36) ? Aa[c + 12 * col] : 0.0);
      }
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        double s = wv[i];
#pragma unroll
        for (int kk = 0; kk < i; ++kk) s -= Ls[i + WbSm::ldS * kk] * wv[kk];
        wv[i] = s * sdinv[i];
        }
#pragma unroll
      for (int i = NR - 1; i >= 0; --i) {
        wv[i] *= sdinv[i];
#pragma unroll
        for (int kk = 0; kk < i; ++kk) wv[kk] -= Ls[i + WbSm::ldS * kk] * wv[i];
      }
    }
#pragma unroll
    for (int i = 0; i < 18; ++i) {
      double d = -r[i];
      if constexpr (NR > 0) {
#pragma unroll
        for (int c = 0; c < NR; ++c) d += Y[i + WbSm::ldY * c] * wv[c];
      }
      r[i] = d;
    }
#pragma unroll
    for (int i = 17; i >= 0; --i) {
      r[i] *= dinv[i];
#pragma unroll
      for (int kk = 0; kk < i; ++kk) r[kk] -= L[i + WbSm::ldL * kk] * r[i];
    }
    if (col < 36) {
#pragma unroll
      for (int i = 0; i < 18; ++i) R[i + 18 * col] = ((col == 18 + i) ? 1.0 : 0.0) + r[i] * dt;
      if constexpr (NR > 0) {
#pragma unroll
        for (int c = 0; c < NR; ++c) Aa[c + 12 * col] = wv[c];
      }
    } else {
#pragma unroll
      for (int i = 0; i < 18; ++i) ABt[i + 20 * col] = r[i] * dt;
      if constexpr (NR > 0) {
#pragma unroll
        for (int c = 0; c < NR; ++c) CDt[rw.row(c) + 12 * col] = wv[c];
      }
    }
  }
}

// --------------------------------------------------------------------------------------------------- K-LQ, whole-body running knots (1)
// KKT sensitivities -> problem-major [A B], [C D] tiles. grid (ceil(n_list / 4), n_wbk), 128 threads.
__global__ void __launch_bounds__(128, 3) k_wb_sens(const SolverDev* __restrict__ Sp, const int* __restrict__ list, int n_list) {
  const SolverDev& S = *Sp;
  extern __shared__ __align__(16) double smem[];
  const int t = threadIdx.x, w = t >> 5, lane = t & 31;
  const int gk = S.wbk_gk[blockIdx.y], pi = S.knot_phase[gk], k = S.knot_k[gk];
  const PhaseDev& ph = S.ph[pi];
  const int ldb = S.ldb, h = ph.h;
  const int p4 = t & 3, jp = blockIdx.x * CAFE_WB_PKS + p4;
  const int bp = jp < n_list ? list[jp] : -1;
  const int j = blockIdx.x * CAFE_WB_PKS + w;
  const bool live = j < n_list;
  const int b = live ? list[j] : 0;
  double* sm = smem + w * WbSm::totalSens;
  WbRows rw; rw.set(ph.contact);
  const int nr = rw.nr;
  double* R = sm + WbSm::oR; double* Aa = sm + WbSm::oA; double* jt = sm + WbSm::oJt;
  const double bg2 = 2.0 * ph.BG_alpha;
  const double* dpp = ph.dp + ((size_t)k * CAFE_DP_W) * ldb;
  // ---- R = [dtau/dq - d(J^T F)/dq | dtau/dv] (18 x 36), a = [da/dq + 2 BG dv/dq | da/dv (+ 2 BG J below)] on the active rows.
  //      The derivative pack is staged in two halves (trunk + legs 0, 1; legs 2, 3): shared memory is what bounds the occupancy.
  for (int half = 0; half < 2; ++half) {
    if (bp >= 0) {
      double* stg = smem + p4 * WbSm::totalSens + WbSm::oStg;
      if (half == 0) wb_stage<CAFE_DP_TRUNK_W + 2 * CAFE_DP_LEG_W>(stg, dpp + bp, ldb, t >> 2);
      else wb_stage<2 * CAFE_DP_LEG_W>(stg + CAFE_DP_TRUNK_W, dpp + (size_t)(CAFE_DP_TRUNK_W + 2 * CAFE_DP_LEG_W) * ldb + bp, ldb, t >> 2);
    }
    __syncthreads();
    if (live) {
      const double* stg = sm + WbSm::oStg;
      if (half == 0) {
        for (int e = lane; e < 648; e += 32) R[e] = 0.0;
        for (int e = lane; e < 432; e += 32) Aa[e] = 0.0;
        if (lane < 9) jt[lane] = 0.0;
        __syncwarp();
        if (lane < 18) { const int dv = lane >= 9, idx = lane % 9; R[(3 + idx % 3) + 18 * (3 + idx / 3 + (dv ? 18 : 0))] = stg[lane]; }
        __syncwarp();
      }
      for (int fl = 0; fl < 2; ++fl) {   // dtau shares in the order trunk, leg 0..3 (RneaDst of wb_pieces.h); foot rows of a
        const int f = 2 * half + fl;
        const double* lg = stg + CAFE_DP_TRUNK_W + fl * CAFE_WBL_DP_W;
        for (int e = lane; e < CAFE_WBL_DP_JTF; e += 32) {
          const int kind = c_dp_kind[e], r = c_dp_row[e], c = c_dp_col[e];
          const double v = lg[e];
          if (kind <= 1) {
            const int idx = wbl_g(f, r) + 18 * (wbl_g(f, c) + (kind == 1 ? 18 : 0));
            const bool shared = r < 6 && c >= 3 && c <= 5, first = (f == 0 && r < 3);
            if (shared && !first) R[idx] += v; else R[idx] = v;
          } else if (kind >= 3) {
            const int ar = rw.arow(3 * f + r);
            if (ar >= 0) Aa[ar + 12 * (wbl_g(f, c) + (kind == 4 ? 18 : 0))] = v;
          }
        }
        __syncwarp();
        for (int e = CAFE_WBL_DP_DVQ + lane; e < CAFE_WBL_DP_DAQ; e += 32) {   // + 2 BG dv_foot/dq
          const int r = c_dp_row[e], c = c_dp_col[e];
          const int ar = rw.arow(3 * f + r);
          if (ar >= 0) Aa[ar + 12 * wbl_g(f, c)] += bg2 * lg[e];
        }
        // d(J^T F)/dq: private entries leave R at once, the shared block (rows, columns 3..5) is summed foot 0..3 first (JtfDst)
        for (int e = CAFE_WBL_DP_JTF + lane; e < CAFE_WBL_DP_W; e += 32) {
          const int r = c_dp_row[e], c = c_dp_col[e];
          const double v = lg[e];
          const bool shared = r >= 3 && r <= 5 && c >= 3 && c <= 5;
          if (shared) { if (f == 0) jt[(r - 3) + 3 * (c - 3)] = v; else jt[(r - 3) + 3 * (c - 3)] += v; }
          else R[wbl_g(f, r) + 18 * wbl_g(f, c)] -= v;
        }
        __syncwarp();
      }
    }
    __syncthreads();
  }
  if (live && lane < 9) R[(3 + lane % 3) + 18 * (3 + lane / 3)] -= jt[lane];
  // ---- rigid-body terms of the trial that produced the current iterate
  if (bp >= 0) {
    double* stg = smem + p4 * WbSm::totalSens + WbSm::oStg;
    const int a = S.c.cur_slot[bp];
    wb_stage<CAFE_TM_W>(stg, ph.tm + ((size_t)(a * h + k) * CAFE_TM_W) * ldb + bp, ldb, t >> 2);
  }
  __syncthreads();
  if (!live) return;
  wb_assemble_terms(sm, sm + WbSm::oStg, lane);
  {
    const double* J = sm + WbSm::oJ;
    for (int e = lane; e < nr * 18; e += 32) { const int c = e % nr, col = e / nr; Aa[c + 12 * (18 + col)] += bg2 * J[rw.row(c) + WbSm::ldJ * col]; }
  }
  wb_factor(sm, rw, 0.0, false, lane);   // damping 0 for the sensitivities (computeKKTContactDynamicMatrixInverse, WBM.cpp:467)
  __syncwarp();
  const double dt = ph.dt;
  double* ABt = ph.ABpm + ((size_t)b * h + k) * CAFE_WB_AB_TILE;
  double* CDt = ph.CDpm + ((size_t)b * h + k) * CAFE_WB_CD_TILE;
  switch (nr) {   // CTA-uniform (a phase has one contact set)
    case 0: wb_colsolve<0>(sm, ABt, CDt, rw, dt, lane); break;
    case 3: wb_colsolve<3>(sm, ABt, CDt, rw, dt, lane); break;
    case 6: wb_colsolve<6>(sm, ABt, CDt, rw, dt, lane); break;
    case 9: wb_colsolve<9>(sm, ABt, CDt, rw, dt, lane); break;
    default: wb_colsolve<12>(sm, ABt, CDt, rw, dt, lane); break;
  }
  __syncwarp();
  for (int e = lane; e < 18 * 36; e += 32) { const int i = e % 18, c = e / 18; ABt[i + 20 * c] = R[e]; }
  for (int e = lane; e < nr * 36; e += 32) { const int c = e % nr, cc = e / nr; CDt[rw.row(c) + 12 * cc] = Aa[c + 12 * cc]; }
}

// --------------------------------------------------------------------------------------------------- K-LQ, whole-body running knots (2)
// cost and barrier partials lx, lu, ly, lxx (structural pattern), luu, lyy and the running cost. grid (ceil(n_list / 4), n_wbk), 128 threads.
__global__ void __launch_bounds__(128, 5) k_wb_cost(const SolverDev* __restrict__ Sp, const int* __restrict__ list, int n_list) {
  const SolverDev& S = *Sp;
  extern __shared__ __align__(16) double smem[];
  const int t = threadIdx.x, w = t >> 5, lane = t & 31;
  const int gk = S.wbk_gk[blockIdx.y], pi = S.knot_phase[gk], k = S.knot_k[gk];
  const PhaseDev& ph = S.ph[pi];
  const int ldb = S.ldb, h = ph.h;
  {
    const int p4 = t & 3, jp = blockIdx.x * CAFE_WB_PKS + p4;
    const int bp = jp < n_list ? list[jp] : -1;
    if (bp >= 0) {
      double* smp = smem + p4 * WbSm::totalCost;
      const int a = S.c.cur_slot[bp];
      wb_stage<CAFE_TM_W>(smp + WbSm::oStg, ph.tm + ((size_t)(a * h + k) * CAFE_TM_W) * ldb + bp, ldb, t >> 2);
      const double* dpp = ph.dp + ((size_t)k * CAFE_DP_W) * ldb + bp;
      for (int e = t >> 2; e < 72; e += 32) smp[WbSm::oDvq + e] = dpp[(size_t)(CAFE_DP_TRUNK_W + (e / 18) * CAFE_WBL_DP_W + CAFE_WBL_DP_DVQ + e % 18) * ldb];
      for (int e = t >> 2; e < 36; e += 32) smp[WbSm::oX + e] = ph.X[gix(k, 36, e, ldb, bp)];
      for (int e = t >> 2; e < 12; e += 32) { smp[WbSm::oU + e] = ph.U[gix(k, 12, e, ldb, bp)]; smp[WbSm::oGrf + e] = ph.Y[gix(k, 12, e, ldb, bp)]; }
    }
  }
  __syncthreads();
  const int j = blockIdx.x * CAFE_WB_PKS + w;
  if (j >= n_list) return;
  const int b = list[j];
  double* sm = smem + w * WbSm::totalCost;
  wb_assemble_terms<false>(sm, sm + WbSm::oStg, lane);
  const double dt = ph.dt;
  const RecRef rec = wb_rec(ph, k, ldb, b);
  const double* dvqc = sm + WbSm::oDvq;
  // dv_foot(3f+a)/dq_i from the compact store (non-zero for the base-rotation columns 3..5 and the leg's own three columns)
  auto dvq_at = [&](int f, int a, int i) -> double {
    const int lc = (i >= 3 && i < 6) ? i - 3 : (i >= 6 + 3 * f && i < 9 + 3 * f) ? 3 + i - (6 + 3 * f) : -1;
    return lc >= 0 ? dvqc[18 * f + a + 3 * lc] : 0.0;
  };
  // ---- cost and barrier partials
  const double* x = sm + WbSm::oX; const double* u = sm + WbSm::oU; const double* grf = sm + WbSm::oGrf;
  const double* J = sm + WbSm::oJ; const double* pf = sm + WbSm::oPf; const double* vf = sm + WbSm::oVf;
  const bool reb = S.opt.ReB_active != 0;
  const RebCtx rcx = reb_ctx(ph, k, ldb, b);
  double* dposw = sm + WbSm::oLx; double* dvelw = dposw + 12; double* dg = dposw + 24;
  // lu, luu (diagonal): tracking + torque-limit barrier
  if (lane < 12) {
    const int i = lane;
    double lu = dt * ph.r[i] * (u[i] - rec[CAFE_REF_UR + i]);
    double luu = dt * ph.r[i];
    if (reb) {
      double bd1, bdd1, bd2, bdd2, dl1, ep1, dl2, ep2;
      rcx.get(ph.reb_torque, i, dl1, ep1); rcx.get(ph.reb_torque, 12 + i, dl2, ep2);
      reb_derivs(-u[i] + ph.torque_limit, dl1, bd1, bdd1);
      reb_derivs(u[i] + ph.torque_limit, dl2, bd2, bdd2);
      lu += dt * (ep1 * bd1 * (-1.0) + ep2 * bd2);
      luu += dt * (ep1 * bdd1 + ep2 * bdd2);
    }
    ph.lu[gix(k, 12, i, ldb, b)] = lu;
    ph.luu[gix(k, 144, 13 * i, ldb, b)] = luu;
    // weighted foot residuals W d
    const int f = i / 3, a = i % 3;
    const bool c = rec[CAFE_REF_CONTACT + f] > 0;
    const double* wq = c ? ph.w_footreg : ph.w_swingpos;
    const double d = (pf[i] - x[a]) - (rec[CAFE_REF_PF + i] - rec[CAFE_REF_PCOM + a]);
    dposw[i] = wq[a] * d;
    dvelw[i] = c ? 0.0 : ph.w_swingvel[a] * (vf[i] - rec[CAFE_REF_VF + i]);
  } else if (lane >= 16 && lane < 20) {
    // ly, lyy: GRF barrier on the output (3 x 3 block per stance foot)
    const int f = lane - 16;
    double gr[3] = {0, 0, 0}, hs[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    if (reb && ph.contact[f] > 0) {
      const double fx = grf[3 * f], fy = grf[3 * f + 1], fz = grf[3 * f + 2], mu = ph.mu;
      const double g[5] = {fz, -fx + mu * fz, fx + mu * fz, -fy + mu * fz, fy + mu * fz};
      const double Al[5][3] = {{0, 0, 1}, {-1, 0, mu}, {1, 0, mu}, {0, -1, mu}, {0, 1, mu}};
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        double bd, bdd, dl, ep;
        rcx.get(ph.reb_grf, 73 + 5 * f + i, dl, ep);
        reb_derivs(g[i], dl, bd, bdd);
        const double e1 = ep * bd, e2 = ep * bdd;
#pragma unroll
        for (int r = 0; r < 3; ++r) { gr[r] += e1 * Al[i][r];
#pragma unroll
          for (int c = 0; c < 3; ++c) hs[r][c] += Al[i][r] * (e2 * Al[i][c]); }
      }
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      ph.ly[gix(k, 12, 3 * f + r, ldb, b)] = dt * gr[r];
#pragma unroll
      for (int c = 0; c < 3; ++c) ph.lyy[gix(k, 144, (3 * f + r) + 12 * (3 * f + c), ldb, b)] = dt * hs[r][c];
    }
  }
  __syncwarp();
  const bool jl = reb && !ph.no_joint_limit, mh = reb && !ph.no_min_height, jv = reb && ph.joint_speed_limit;
  // lx and the diagonal of lxx (tracking + joint-limit / min-height / joint-speed barriers)
  for (int i = lane; i < 36; i += 32) {
    double lx = dt * ph.q[i] * (x[i] - rec[CAFE_REF_XR + i]);
    for (int f = 0; f < 4; ++f) {
      const bool c = rec[CAFE_REF_CONTACT + f] > 0;
      if (i >= 3 && i < 18) { double g = 0; for (int a = 0; a < 3; ++a) g += J[3 * f + a + WbSm::ldJ * i] * dposw[3 * f + a]; lx += g * dt; }
      if (!c) {
        double g = 0;
        for (int a = 0; a < 3; ++a) g += ((i < 18) ? dvq_at(f, a, i) : J[3 * f + a + WbSm::ldJ * (i - 18)]) * dvelw[3 * f + a];
        lx += g * dt;
      }
    }
    double v = dt * ph.q[i];
    if (jl && i >= 6 && i < 18) {
      double b1, d1, b2, d2, dl1, ep1, dl2, ep2;
      rcx.get(ph.reb_joint, 48 + (i - 6), dl1, ep1); rcx.get(ph.reb_joint, 60 + (i - 6), dl2, ep2);
      reb_derivs(x[i] - ph.joint_lb[(i - 6) % 3], dl1, b1, d1);
      reb_derivs(-x[i] + ph.joint_ub[(i - 6) % 3], dl2, b2, d2);
      lx += dt * (ep1 * b1 - ep2 * b2);
      v += dt * (ep1 * d1 + ep2 * d2);
    }
    if (mh && i == 2) {
      double b1, d1, dl1, ep1;
      rcx.get(ph.reb_minheight, 72, dl1, ep1);
      reb_derivs(x[2] - ph.h_min, dl1, b1, d1);
      lx += dt * (ep1 * b1);
      v += dt * (ep1 * d1);
    }
    if (jv && i >= 24) {
      double b1, b2, d1, d2, dl1, ep1, dl2, ep2;
      rcx.get(ph.reb_jointvel, 24 + (i - 24), dl1, ep1); rcx.get(ph.reb_jointvel, 36 + (i - 24), dl2, ep2);
      reb_derivs(x[i] - ph.jointvel_lb, dl1, b1, d1);
      reb_derivs(-x[i] + ph.jointvel_ub, dl2, b2, d2);
      lx += dt * (ep1 * b1 - ep2 * b2);
      v += dt * (ep1 * d1 + ep2 * d2);
    }
    ph.lx[gix(k, 36, i, ldb, b)] = lx;
    dg[i] = v;
  }
  __syncwarp();
  // lxx, structural pattern only (see WBModel::lq_knot): a foot Jacobian has the base-rotation columns 3..5 and its own three joint
  // columns, the swing-foot velocity Jacobian [dv/dq | J] additionally the six base columns and the leg columns of the velocity half.
  double* lxxg = ph.lxx + gix(k, 1296, 0, ldb, b);
  auto jxv = [&](int f, int a, int i) -> double { return (i < 18) ? dvq_at(f, a, i) : J[3 * f + a + WbSm::ldJ * (i - 18)]; };
  // (a) the 9 x 9 block of the base columns {3,4,5,18..23}, shared by the feet: accumulated foot by foot
  for (int e = lane; e < 81; e += 32) {
    const int p = e % 9, q = e / 9;
    const int i = p < 3 ? 3 + p : 15 + p, jq = q < 3 ? 3 + q : 15 + q;
    double val = (p == q) ? dg[i] : 0.0;
    for (int f = 0; f < 4; ++f) {
      const bool c = rec[CAFE_REF_CONTACT + f] > 0;
      const double* wq = c ? ph.w_footreg : ph.w_swingpos;
      if (p < 3 && q < 3) { double hh = 0; for (int a = 0; a < 3; ++a) hh += J[3 * f + a + WbSm::ldJ * i] * wq[a] * J[3 * f + a + WbSm::ldJ * jq]; val += hh * dt; }
      if (!c) { double hh = 0; for (int a = 0; a < 3; ++a) hh += jxv(f, a, i) * ph.w_swingvel[a] * jxv(f, a, jq); val += hh * dt; }
    }
    lxxg[(size_t)(i + 36 * jq) * ldb] = val;
  }
  // (b) entries private to one foot
  for (int f = 0; f < 4; ++f) {
    const bool c = rec[CAFE_REF_CONTACT + f] > 0;
    const double* wq = c ? ph.w_footreg : ph.w_swingpos;
    const int nc = c ? 6 : 15;
    for (int e = lane; e < nc * nc; e += 32) {
      const int ii = e % nc, jj = e / nc;
      // local column list: 0..2 base rotation, 3..5 leg q, 6..11 base velocity, 12..14 leg v
      const bool bi = ii < 3 || (ii >= 6 && ii < 12), bj = jj < 3 || (jj >= 6 && jj < 12);
      if (bi && bj) continue;
      const int i = ii < 3 ? 3 + ii : ii < 6 ? 6 + 3 * f + (ii - 3) : ii < 12 ? 18 + (ii - 6) : 24 + 3 * f + (ii - 12);
      const int jq = jj < 3 ? 3 + jj : jj < 6 ? 6 + 3 * f + (jj - 3) : jj < 12 ? 18 + (jj - 6) : 24 + 3 * f + (jj - 12);
      double val = (i == jq) ? dg[i] : 0.0;
      if (ii < 6 && jj < 6) { double hh = 0; for (int a = 0; a < 3; ++a) hh += J[3 * f + a + WbSm::ldJ * i] * wq[a] * J[3 * f + a + WbSm::ldJ * jq]; val += hh * dt; }
      if (!c) { double hh = 0; for (int a = 0; a < 3; ++a) hh += jxv(f, a, i) * ph.w_swingvel[a] * jxv(f, a, jq); val += hh * dt; }
      lxxg[(size_t)(i + 36 * jq) * ldb] = val;
    }
    if (c && lane < 3) { const int i = 24 + 3 * f + lane; lxxg[(size_t)(37 * i) * ldb] = dg[i]; }   // leg-velocity diagonal of a stance foot
  }
  if (lane < 3) lxxg[(size_t)(37 * lane) * ldb] = dg[lane];
  // ---- running cost at the current iterate (compute_cost, SinglePhase.cpp:236-262)
  double ming;
  const double l = wb_cost_coop(ph, sm, rec, sm + WbSm::oScr, reb, lane, ming, rcx);
  if (lane == 0) ph.lk[(size_t)k * ldb + b] = l;
}

}  // namespace cafe_dev
