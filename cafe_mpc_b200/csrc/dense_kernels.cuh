// dense_kernels.cuh — cooperative shared-memory part of the whole-body linearisation (see model_wb.cuh::lq_knot for the producer).
#pragma once
#include "device_types.cuh"

namespace cafe_dev {

// ------------------------------------------------------------------------------- K-LQ, whole-body dense part
// KKT sensitivities of one WB phase (WBM.cpp:459-505 without forming Kinv): CTA = 4 consecutive entries of the active list at one knot,
// one warp per problem, lane = column z of [q(18) v(18) tau(12)]; chol(M), Y = L^-1 Jc^T, chol(S) are read from shared memory at
// warp-uniform addresses (broadcast), the column's right-hand side lives in registers.
//   dlambda/dz = S^-1 (Jc Minv R - a),  dqdd/dz = -Minv (R - Jc^T dlambda/dz);  A = I + dt Ac, B = dt Bc, C/D = dGRF/d(x,u)
// shared memory per problem (doubles): Lt 324 (row-major L, reciprocal diagonal) | Y 216 (18 x 12, ld 18) | Yt 216 (12 x 18, ld 12)
//                                     | Lst 144 (row-major chol(S), reciprocal diagonal) | R 19x36 | a 13x36
#define CAFE_KKT_SM (324 + 216 + 216 + 144 + 684 + 468)
// one warp per problem, two passes over the 48 columns (one thread per column, 192 threads, measured slower: 30.9 vs 27.7 ms per batch)
#define CAFE_DENSE_NT 128
template <int NR>
__global__ void __launch_bounds__(CAFE_DENSE_NT, 3) k_lq_wb_dense(const SolverDev* __restrict__ Sp, int pi, const int* __restrict__ list, int n_list) {
  const SolverDev& S = *Sp;
  const PhaseDev& ph = S.ph[pi];
  extern __shared__ double sm[];
  const int ldb = S.ldb, k = blockIdx.y, b0 = blockIdx.x * 4;
  {
    const int p = threadIdx.x & 3, jl = b0 + p, b = jl < n_list ? list[jl] : 0;   // entries b0..b0+3 of the active list
    double* dst = sm + p * CAFE_KKT_SM;
    const double* src = ph.kkt + gix(k, CAFE_KKT_PACK, 0, ldb, b);
    if (jl < n_list && S.c.active[b]) {
      const double bg2 = 2.0 * ph.BG_alpha;
      int rowsA[12];
      { int j = 0; for (int f = 0; f < 4; ++f) if (ph.contact[f] > 0) for (int r = 0; r < 3; ++r) rowsA[j++] = 3 * f + r; for (; j < 12; ++j) rowsA[j] = 0; }
      // every loop below first issues a batch of independent global loads and only then consumes them: one element per iteration
      // made the CTA wait a full memory latency ~60 times in a row (ncu: long-scoreboard 13.6 cycles per issue)
      constexpr int ST = CAFE_DENSE_NT / 4, UB = 8;
      for (int e0 = threadIdx.x >> 2; e0 < 684; e0 += ST * UB) {  // factors
        double v[UB];
#pragma unroll
        for (int u = 0; u < UB; ++u) { const int e = e0 + ST * u; v[u] = e < 684 ? src[(size_t)e * ldb] : 0.0; }
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          const int e = e0 + ST * u;
          if (e >= 684) break;
          if (e >= CAFE_KKT_LS) { const int idx = e - CAFE_KKT_LS; const int i = idx % 12, j = idx / 12; dst[756 + j + 12 * i] = (i == j) ? 1.0 / v[u] : v[u]; }
          else if (e >= CAFE_KKT_Y) { const int idx = e - CAFE_KKT_Y; const int i = idx % 18, c = idx / 18; dst[324 + idx] = v[u]; dst[540 + c + 12 * i] = v[u]; }
          else { const int i = e % 18, j = e / 18; dst[j + 18 * i] = (i == j) ? 1.0 / v[u] : v[u]; }
        }
      }
      for (int e0 = threadIdx.x >> 2; e0 < 648; e0 += ST * UB) {  // R = [dtau_dq - d(J^T F)/dq | dtau_dv]
        double v[UB], w[UB];
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          const int e = e0 + ST * u;
          v[u] = e < 648 ? src[(size_t)(CAFE_KKT_RQ + e) * ldb] : 0.0;
          w[u] = e < 324 ? src[(size_t)(CAFE_KKT_JTF + e) * ldb] : 0.0;   // columns 0..17 only
        }
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          const int e = e0 + ST * u;
          if (e >= 648) break;
          const int i = e % 18, col = e / 18;
          dst[900 + i + 19 * col] = (col < 18) ? v[u] - w[u] : v[u];
        }
      }
      for (int e0 = threadIdx.x >> 2; e0 < NR * 36; e0 += ST * UB) {  // a = [da/dq + 2 BG dv/dq | da/dv + 2 BG J] on the active rows
        double v[UB], w[UB];
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          const int e = e0 + ST * u;
          const int c = e % (NR > 0 ? NR : 1), col = e / (NR > 0 ? NR : 1);
          const int row = rowsA[c];
          const bool in = e < NR * 36;
          const int o1 = (col < 18) ? CAFE_KKT_AQ + row + 12 * col : CAFE_KKT_AV + row + 12 * (col - 18);
          const int o2 = (col < 18) ? CAFE_KKT_DVQ + row + 12 * col : CAFE_KKT_J + row + 12 * (col - 18);
          v[u] = in ? src[(size_t)o1 * ldb] : 0.0;
          w[u] = in ? src[(size_t)o2 * ldb] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          const int e = e0 + ST * u;
          if (e >= NR * 36) break;
          const int c = e % (NR > 0 ? NR : 1), col = e / (NR > 0 ? NR : 1);
          dst[1584 + c + 13 * col] = v[u] + bg2 * w[u];
        }
      }
    }
  }
  __syncthreads();
  const int prob = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (b0 + prob >= n_list) return;
  const int b = list[b0 + prob];
  if (!S.c.active[b]) return;
  const double* sLt = sm + prob * CAFE_KKT_SM;  // Lt[k + 18 i] = L(i,k), Lt[i + 18 i] = 1 / L(i,i)
  const double* sY = sLt + 324;                 // Y[i + 18 c]
  const double* sYt = sLt + 540;                // Yt[c + 12 i]
  const double* sLst = sLt + 756;               // Lst[k + 12 i] = Ls(i,k), reciprocal diagonal
  double* sR = sm + prob * CAFE_KKT_SM + 900;    // column `col` is read by its own lane only and then reused for that lane's results
  double* sa = sm + prob * CAFE_KKT_SM + 1584;
  const double dt = ph.dt;
  int foot[4] = {0, 0, 0, 0};
  { int j = 0; for (int f = 0; f < 4; ++f) if (ph.contact[f] > 0) foot[j++] = f; }
  // outputs go to the problem-major tiles the backward sweep stages with 16-byte copies: column `col` of [A B] rows 18..35 / [C D]
  double* ABt = ph.ABpm + ((size_t)b * ph.h + k) * CAFE_WB_AB_TILE;
  double* CDt = ph.CDpm + ((size_t)b * ph.h + k) * CAFE_WB_CD_TILE;
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
      // the 36 state columns go back into this lane's (dead) R / a columns and leave the SM as contiguous runs below: an 8-byte
      // store per lane at a 160-byte stride cost a DRAM sector each (ncu: 3.6 GB written per launch for 0.55 GB of results)
#pragma unroll
      for (int i = 0; i < 18; ++i) sR[i + 19 * col] = ((col == 18 + i) ? 1.0 : 0.0) + r[i] * dt;
      if constexpr (NR > 0) {
#pragma unroll
        for (int c = 0; c < NR; ++c) sa[c + 13 * col] = w[c];
      }
    } else {   // the 12 control columns (no shared-memory slot left): direct stores
#pragma unroll
      for (int i = 0; i < 18; ++i) ABt[i + 20 * col] = r[i] * dt;
      if constexpr (NR > 0) {
#pragma unroll
        for (int c = 0; c < NR; ++c) CDt[(3 * foot[c / 3] + c % 3) + 12 * col] = w[c];
      }
    }
  }
  __syncwarp();
  for (int e = lane; e < 18 * 36; e += 32) { const int i = e % 18, c = e / 18; ABt[i + 20 * c] = sR[i + 19 * c]; }
  if constexpr (NR > 0) {
    for (int e = lane; e < NR * 36; e += 32) { const int c = e % NR, cc = e / NR; CDt[(3 * foot[c / 3] + c % 3) + 12 * cc] = sa[c + 13 * cc]; }
  }
}

}  // namespace cafe_dev
