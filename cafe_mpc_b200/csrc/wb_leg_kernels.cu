// wb_leg_kernels.cu — the straight-line rigid-body routines of the whole-body knots, one thread per (problem, knot, PIECE).
//
// A piece is the share of one inertia group / one foot (RNEA is linear in the link inertias, a foot only involves the base and its
// own leg; see wb_pieces.h): the four legs run the SAME leg-generic routine (gen/wb_leg_gen.h) with the leg's nine mirrored
// constants read from constant memory. grid = (list blocks, running WB knots, pieces [x step sizes]); the problem index is the
// fastest thread index, so a warp is 32 problems in the same routine (no divergence) and every access to the batch-major arrays is
// one coalesced 256-byte transaction. Outputs are the routines' compact non-zero lists, stored once into PhaseDev::tm / ::dp; the
// cooperative kernels of wb_coop.cuh take it from there (no thread-local arrays here: the register file is the only scratch).
//
//   k_wb_terms   (rollout)         pieces 0..3: leg share of M, nle + foot J, Jdot v, position, velocity; 4: trunk share
//                                  <= crba / nonLinearEffects / frame Jacobians etc., MHPC/MHPC-Trajopt/WBM.cpp:375-411
//   k_wb_derivs  (linearisation)   per leg nine forward-mode routines: columns of the leg's share of dtau/dq, dtau/dv
//                                  (computeRNEADerivatives, WBM.cpp:474) for the directions q3 | q4 | q5 | q6..8 | v3..5 | v6..8, and of the
//                                  foot's kinematic partials (the reference's CasADi functions, WBM.cpp:479-487) for q3..5 | q6..8 | v3..8;
//                                  + the trunk share: 37 pieces
// The routines that carry CAFE_GEN_SYNC markers (a CTA barrier every 1024 operations: the four warps of a CTA then share
// instruction-cache lines) require that every thread of the CTA runs the same routine and none exits early: the piece is a
// grid dimension, and the lanes past the end of the list repeat its last entry (same values, stores predicated off).
#define CAFE_GEN_SYNC __syncthreads();
#define CAFE_HD __device__ __forceinline__
#include "gen/wb_leg_gen.h"
#include "gen/wb_gen.h"
#include "device_types.cuh"
#include "launchers.h"

#ifndef CAFE_LEG_MINB
#define CAFE_LEG_MINB 2
#endif

namespace cafe_dev {

static __constant__ double c_wbl_P[4][CAFE_WBL_NP] = CAFE_WBL_LEG_CONSTANTS;

// output functor: compact slot idx of a routine -> batch-major array (element stride ldb), stores predicated
struct SlotDst {
  double* p; size_t st; bool on;
  __device__ __forceinline__ void operator()(int idx, double x) const { if (on) p[idx * st] = x; }
};
// trunk shares: the routines of gen/wb_gen.h pass dense indices of 18-vectors / 18 x 18 matrices (literals: the maps fold away)
struct TrunkVec6 {   // nle[i], i < 6
  double* p; size_t st; bool on;
  __device__ __forceinline__ void operator()(int idx, double x) const { if (on) p[idx * st] = x; }
};
struct TrunkMat6 {   // M(r, c), r, c < 6 -> r + 6 c
  double* p; size_t st; bool on;
  __device__ __forceinline__ void operator()(int idx, double x) const { if (on) p[((idx % 18) + 6 * (idx / 18)) * st] = x; }
};
struct TrunkMat3 {   // dtau(r, c), r, c in 3..5 -> (r - 3) + 3 (c - 3)
  double* p; size_t st; bool on;
  __device__ __forceinline__ void operator()(int idx, double x) const { if (on) p[((idx % 18 - 3) + 3 * (idx / 18 - 3)) * st] = x; }
};

__global__ void __launch_bounds__(128, CAFE_LEG_MINB) k_wb_terms(const SolverDev* __restrict__ Sp, int a0, const int* __restrict__ list, int n_list) {
  const SolverDev& S = *Sp;
  const int j = blockIdx.x * 128 + threadIdx.x;
  if (j >= n_list) return;
  const int b = list[j];
  if (!S.c.active[b] || !S.c.do_ls[b] || S.c.ls_found[b]) return;
  const int gk = S.wbk_gk[blockIdx.y], pi = S.knot_phase[gk], k = S.knot_k[gk];
  const int piece = blockIdx.z % 5, a = a0 + blockIdx.z / 5;
  const PhaseDev& ph = S.ph[pi];
  const int ldb = S.ldb, h = ph.h;
  const size_t st = (size_t)ldb;
  const double* xt = ph.Xt + (size_t)a * (h + 1) * 36 * ldb + gix(k, 36, 0, ldb, b);   // element i at xt[i * ldb]
  double* tm = ph.tm + ((size_t)(a * h + k) * CAFE_TM_W) * ldb + b;
  if (piece == 4) {
    double q[18], v[18];
#pragma unroll
    for (int i = 0; i < 18; ++i) { q[i] = 0.0; v[i] = 0.0; }
#pragma unroll
    for (int i = 0; i < 6; ++i) { q[i] = xt[(size_t)i * st]; v[i] = xt[(size_t)(18 + i) * st]; }
    cafe_gen_wb::wb_terms_trunk(q, v, TrunkVec6{tm, st, true}, TrunkMat6{tm + 6 * st, st, true});
    return;
  }
  const int f = piece;
  double ql[9], vl[9];
#pragma unroll
  for (int i = 0; i < 6; ++i) { ql[i] = xt[(size_t)i * st]; vl[i] = xt[(size_t)(18 + i) * st]; }
#pragma unroll
  for (int i = 0; i < 3; ++i) { ql[6 + i] = xt[(size_t)(6 + 3 * f + i) * st]; vl[6 + i] = xt[(size_t)(24 + 3 * f + i) * st]; }
  double P[CAFE_WBL_NP];
#pragma unroll
  for (int i = 0; i < CAFE_WBL_NP; ++i) P[i] = c_wbl_P[f][i];
  double* o = tm + (size_t)(CAFE_TM_TRUNK_W + f * CAFE_WBL_TM_W) * st;
  cafe_gen_wbl::wbl_terms_leg(ql, vl, P, SlotDst{o + CAFE_WBL_TM_NLE * st, st, true}, SlotDst{o + CAFE_WBL_TM_M * st, st, true},
                              SlotDst{o + CAFE_WBL_TM_J * st, st, true}, SlotDst{o + CAFE_WBL_TM_GAM * st, st, true},
                              SlotDst{o + CAFE_WBL_TM_PF * st, st, true}, SlotDst{o + CAFE_WBL_TM_VF * st, st, true});
}

// piece = blockIdx.z: 9 routines x 4 legs (routine-major: the four legs of a routine are neighbours in the grid), then the trunk
#define CAFE_WB_DERIV_PIECES 37
// the list is padded to the CTA size by repeating its last entry (stores off): every thread of a CTA runs the same routine to its end
__global__ void __launch_bounds__(128, CAFE_LEG_MINB) k_wb_derivs(const SolverDev* __restrict__ Sp, const int* __restrict__ list, int n_list) {
  const SolverDev& S = *Sp;
  const int j0 = blockIdx.x * 128 + threadIdx.x;
  const bool on = j0 < n_list;
  const int b = list[on ? j0 : n_list - 1];
  const int gk = S.wbk_gk[blockIdx.y], pi = S.knot_phase[gk], k = S.knot_k[gk];
  const int piece = blockIdx.z;
  const PhaseDev& ph = S.ph[pi];
  const int ldb = S.ldb, h = ph.h;
  const size_t st = (size_t)ldb;
  const double* xc = ph.X + gix(k, 36, 0, ldb, b);
  const double* qdd = ph.qdd_t + ((size_t)(S.c.cur_slot[b] * h + k) * 18) * ldb + b;
  double* dp = ph.dp + ((size_t)k * CAFE_DP_W) * ldb + b;
  if (piece == CAFE_WB_DERIV_PIECES - 1) {
    double q[18], v[18], a[18];
#pragma unroll
    for (int i = 0; i < 18; ++i) { q[i] = 0.0; v[i] = 0.0; a[i] = 0.0; }
#pragma unroll
    for (int i = 0; i < 6; ++i) { q[i] = xc[(size_t)i * st]; v[i] = xc[(size_t)(18 + i) * st]; a[i] = qdd[(size_t)i * st]; }
    cafe_gen_wb::wb_rnea_derivs_trunk(q, v, a, TrunkMat3{dp, st, on}, TrunkMat3{dp + 9 * st, st, on});
    return;
  }
  const int f = piece & 3, routine = piece >> 2;
  double ql[9], vl[9], al[9];
#pragma unroll
  for (int i = 0; i < 6; ++i) { ql[i] = xc[(size_t)i * st]; vl[i] = xc[(size_t)(18 + i) * st]; al[i] = qdd[(size_t)i * st]; }
#pragma unroll
  for (int i = 0; i < 3; ++i) { ql[6 + i] = xc[(size_t)(6 + 3 * f + i) * st]; vl[6 + i] = xc[(size_t)(24 + 3 * f + i) * st]; al[6 + i] = qdd[(size_t)(6 + 3 * f + i) * st]; }
  double P[CAFE_WBL_NP];
#pragma unroll
  for (int i = 0; i < CAFE_WBL_NP; ++i) P[i] = c_wbl_P[f][i];
  const SlotDst o{dp + (size_t)(CAFE_DP_TRUNK_W + f * CAFE_WBL_DP_W) * st, st, on};
  double F[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) F[i] = ph.Y[gix(k, 12, 3 * f + i, ldb, b)];
  switch (routine) {   // CTA-uniform
    case 0: cafe_gen_wbl::wbl_rnea_q4(ql, vl, al, P, o); break;
    case 1: cafe_gen_wbl::wbl_rnea_q678(ql, vl, al, P, o); break;
    case 2: cafe_gen_wbl::wbl_rnea_q5(ql, vl, al, P, o); break;
    case 3: cafe_gen_wbl::wbl_rnea_v345(ql, vl, al, P, o); break;
    case 4: cafe_gen_wbl::wbl_rnea_q3(ql, vl, al, P, o); break;
    case 5: cafe_gen_wbl::wbl_kin_q345(ql, vl, al, F, P, o); break;
    case 6: cafe_gen_wbl::wbl_rnea_v678(ql, vl, al, P, o); break;
    case 7: cafe_gen_wbl::wbl_kin_q678(ql, vl, al, F, P, o); break;
    default: cafe_gen_wbl::wbl_kin_v345678(ql, vl, al, F, P, o); break;
  }
}

void launch_wb_terms(const SolverDev* dS, int n_wbk, cudaStream_t st, int a0, int a1, const int* list, int n_list) {
  if (n_list <= 0 || n_wbk <= 0 || a1 <= a0) return;
  const dim3 grid((n_list + 127) / 128, n_wbk, 5 * (a1 - a0));
  k_wb_terms<<<grid, 128, 0, st>>>(dS, a0, list, n_list);
}
void launch_wb_derivs(const SolverDev* dS, int n_wbk, cudaStream_t st, const int* list, int n_list) {
  if (n_list <= 0 || n_wbk <= 0) return;
  const dim3 grid((n_list + 127) / 128, n_wbk, CAFE_WB_DERIV_PIECES);
  k_wb_derivs<<<grid, 128, 0, st>>>(dS, list, n_list);
}

}  // namespace cafe_dev
