// bwd2.cuh — K-BWD v2 (+ K-LIN + merit parameter): one CTA of NT threads per PROBLEM.
//
//   SinglePhase::backward_sweep / linear_rollout                      HSDDPSolver/source/SinglePhase.cpp:145-178, :323-391
//   MultiPhaseDDP::backward_sweep(_regularized) / linear_rollout / impact_aware_step
//                                                                     HSDDPSolver/source/MultiPhaseDDP.cpp:12-42, :136-213, :499-503
//
// Why v2: v1 interleaved PB problems per CTA and needed ~70 KB of shared memory per problem, i.e. ONE latency-bound
// 96..128-thread CTA per SM and dozens of CTA-wide barriers per knot (ncu: 5 % fp64 pipe, 4.7 % warps active). Here
//   * a CTA owns one problem (barriers couple only its NT/32 warps), 4-6 CTAs are resident per SM,
//   * matrices live in padded (odd leading dimension) column-major shared-memory tiles, all GEMMs are register tiled over
//     an 8 x NT/8 thread grid (25-36 independent accumulators per thread),
//   * the whole-body phases use the block structure A = [[I, dt I],[A21, A22]], B = [0; B2]: only P = H[:,18:36] enters the
//     products (T = P [A2 B2] + epilogue, Qxx = A2^T T2 + epilogue, ...): ~40 % fewer flops, half the operand footprint,
//   * LDL^T of (Quu - 1e-9 I) runs in one warp with warp-level barriers, the solves for [K | dU] keep each right-hand-side
//     column in registers (thread = column) and use reciprocal pivots.
#pragma once
#include "device_types.cuh"

namespace cafe_dev {

// C(i,j) = epi(i,j, sum_{l<KK} opA(i,l) B(l,j) + sum_{l<KK2} A2[l + lda2*i] B2[l + ldb2*j])   for i < MM, j < NN
//   opA(i,l) = TA ? A[l + lda*i] : A[i + lda*l];  B(l,j) = B[l + ldb*j]; thread grid 8 x NT/8
template <int MM, int NN, int KK, bool TA, int KK2, int NT, class Epi>
__device__ __forceinline__ void gemm_nt(const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb,
                                        const double* __restrict__ A2, int lda2, const double* __restrict__ B2, int ldb2, int t, bool run, Epi epi) {
  constexpr int WR = 8, WC = NT / 8;
  constexpr int TR = (MM + WR - 1) / WR, TC = (NN + WC - 1) / WC;
  const int r0 = (t % WR) * TR, q0 = (t / WR) * TC;
  if (!run) return;
  double acc[TR][TC];
#pragma unroll
  for (int r = 0; r < TR; ++r)
#pragma unroll
    for (int q = 0; q < TC; ++q) acc[r][q] = 0;
#pragma unroll 2
  for (int l = 0; l < KK; ++l) {
    double av[TR], bv[TC];
#pragma unroll
    for (int r = 0; r < TR; ++r) { const int i = r0 + r; av[r] = (i < MM) ? (TA ? A[l + lda * i] : A[i + lda * l]) : 0.0; }
#pragma unroll
    for (int q = 0; q < TC; ++q) { const int j = q0 + q; bv[q] = (j < NN) ? B[l + ldb * j] : 0.0; }
#pragma unroll
    for (int r = 0; r < TR; ++r)
#pragma unroll
      for (int q = 0; q < TC; ++q) acc[r][q] += av[r] * bv[q];
  }
  if constexpr (KK2 > 0) {
#pragma unroll 2
    for (int l = 0; l < KK2; ++l) {
      double av[TR], bv[TC];
#pragma unroll
      for (int r = 0; r < TR; ++r) { const int i = r0 + r; av[r] = (i < MM) ? A2[l + lda2 * i] : 0.0; }
#pragma unroll
      for (int q = 0; q < TC; ++q) { const int j = q0 + q; bv[q] = (j < NN) ? B2[l + ldb2 * j] : 0.0; }
#pragma unroll
      for (int r = 0; r < TR; ++r)
#pragma unroll
        for (int q = 0; q < TC; ++q) acc[r][q] += av[r] * bv[q];
    }
  }
#pragma unroll
  for (int r = 0; r < TR; ++r)
#pragma unroll
    for (int q = 0; q < TC; ++q) {
      const int i = r0 + r, j = q0 + q;
      if (i < MM && j < NN) epi(i, j, acc[r][q]);
    }
}

// shared-memory plan (doubles) of one problem whose largest phase is (NX, MX, PX); WBS: structured whole-body storage
template <int NX, int MX, int PX, bool WBS>
struct Bwd2Layout {
  static constexpr int ldH = NX | 1, KA = WBS ? 18 : NX, ldA = KA | 1, ldM = MX | 1, ldP = (PX > 0 ? PX : 1) | 1;
  // vectors
  static constexpr int vG = 0, vGn = NX, vQx = 2 * NX, vD = 3 * NX, vDx = 4 * NX, vDxn = 5 * NX, vQu = 6 * NX, vDu = 6 * NX + MX,
                       vDuL = 6 * NX + 2 * MX, vLy = 6 * NX + 3 * MX, vRed = vLy + PX + 1, nVec = vRed + 2 * 128 + 1;
  static constexpr int oH = nVec;                       // H / Qxx / H_new      NX x NX (ldH)
  static constexpr int oAB = oH + ldH * NX;             // [A B] rows KA (ldA) x (NX+MX); reused for [K | dU] (MX x (NX+1), ldM)
  static constexpr int szAB = (ldA * (NX + MX) > ldM * (NX + 1)) ? ldA * (NX + MX) : ldM * (NX + 1);
  static constexpr int oT = oAB + szAB;                 // T = H [A B]  NX x (NX+MX) (ldH); reused for L (MX x MX, ldM)
  static constexpr int oQux = oT + ldH * (NX + MX);     // Qux MX x NX (ldM)
  static constexpr int oQuu = oQux + ldM * NX;          // Quu MX x MX (ldM)
  static constexpr int oCD = oQuu + ldM * MX;           // [C D]  PX x (NX+MX) (ldP)
  static constexpr int oSCD = oCD + (PX > 0 ? ldP * (NX + MX) : 0);   // lyy [C D]
  static constexpr int oLyy = oSCD + (PX > 0 ? ldP * (NX + MX) : 0);  // lyy PX x PX (ldP)
  static constexpr int total = oLyy + (PX > 0 ? ldP * PX : 0) + 2;
};

// One phase of the sweep. N, M, PY: phase dimensions; NNEXT: state dimension of the next phase (for the jump); WB: use the
// whole-body block structure. On entry, when the phase has a successor, sG/sH hold G0+/H0+ of the successor (ld ldH).
template <int N, int M, int PY, int NNEXT, bool WB, int NT, class L>
__device__ void sweep_phase2(const SolverDev& S, int pi, int b, int t, bool run, bool& ok, double reg, double* sm, double& min_piv) {
  const PhaseDev& ph = S.ph[pi];
  const int ldb = S.ldb, h = ph.h;
  constexpr int ldH = L::ldH, ldA = L::ldA, ldM = L::ldM, ldP = L::ldP;
  constexpr int KA = WB ? 18 : N;   // stored rows of [A B]
  constexpr int R0 = WB ? 18 : 0;   // first stored row
  double* sG = sm + L::vG; double* sGn = sm + L::vGn; double* sQx = sm + L::vQx; double* sD = sm + L::vD;
  double* sQu = sm + L::vQu; double* sDu = sm + L::vDu; double* sLy = sm + L::vLy;
  double* sH = sm + L::oH; double* sAB = sm + L::oAB; double* sT = sm + L::oT; double* sQux = sm + L::oQux; double* sQuu = sm + L::oQuu;
  double* sCD = sm + L::oCD; double* sSCD = sm + L::oSCD; double* sLyy = sm + L::oLyy;
  double* sK = sAB;   // [K | dU] after A, B are dead
  double* sL = sT;    // LDL^T factor after T is dead
  (void)sLy; (void)sCD; (void)sSCD; (void)sLyy;
  const double dt = ph.dt;

  // ---- boundary: (G', H') = (Px^T G0+, Px^T H0+ Px) (impact_aware_step), G[h] = Phix + G', H[h] = Phixx + H'
  if (ph.has_next) {
    // Px (NNEXT x N, ld = NNEXT|1) and the product H0+ Px live in the contiguous [A B | T] area
    double* sPx = sAB;
    constexpr int ldX = NNEXT | 1;
    double* sTj = sPx + ldX * N;
    static_assert(ldX * N + ldH * N <= L::oQux - L::oAB, "Px and H0+ Px do not fit");
    if (run && ok) for (int e = t; e < NNEXT * N; e += NT) sPx[(e % NNEXT) + ldX * (e / NNEXT)] = ph.Px[(size_t)e * ldb + b];
    __syncthreads();
    gemm_nt<NNEXT, N, NNEXT, false, 0, NT>(sH, ldH, sPx, ldX, nullptr, 0, nullptr, 0, t, run && ok,
                                           [&](int i, int j, double v) { sTj[i + ldH * j] = v; });  // T = H0+ Px
    if (run && ok) for (int j = t; j < N; j += NT) { double s = 0; for (int i = 0; i < NNEXT; ++i) s += sPx[i + ldX * j] * sG[i]; sGn[j] = s; }
    __syncthreads();
    gemm_nt<N, N, NNEXT, true, 0, NT>(sPx, ldX, sTj, ldH, nullptr, 0, nullptr, 0, t, run && ok,
                                      [&](int i, int j, double v) { sH[i + ldH * j] = v + ph.Phixx[(size_t)(i + N * j) * ldb + b]; });
    if (run && ok) for (int j = t; j < N; j += NT) sG[j] = ph.Phix[(size_t)j * ldb + b] + sGn[j];
  } else {
    if (run && ok) {
      for (int e = t; e < N * N; e += NT) sH[(e % N) + ldH * (e / N)] = ph.Phixx[(size_t)e * ldb + b];
      for (int j = t; j < N; j += NT) sG[j] = ph.Phix[(size_t)j * ldb + b];
    }
  }
  __syncthreads();
  if (run && ok) for (int j = t; j < N; j += NT) ph.G[gix(h, N, j, ldb, b)] = sG[j];

  for (int k = h - 1; k >= 0; --k) {
    const bool a2 = run && ok;
    // ---- stage [A B] (stored rows), [C D], lyy, ly, Defect[k+1]
    if (a2) {
      const double* Ag = ph.A + gix(k, N * N, 0, ldb, b);
      for (int e = t; e < KA * N; e += NT) { const int i = e % KA, j = e / KA; sAB[i + ldA * j] = Ag[(size_t)((R0 + i) + N * j) * ldb]; }
      const double* Bg = ph.Bm + gix(k, N * M, 0, ldb, b);
      for (int e = t; e < KA * M; e += NT) { const int i = e % KA, j = e / KA; sAB[i + ldA * (N + j)] = Bg[(size_t)((R0 + i) + N * j) * ldb]; }
      for (int j = t; j < N; j += NT) sD[j] = ph.Defect[gix(k + 1, N, j, ldb, b)];
      if constexpr (PY > 0) {
        const double* Cg = ph.C + gix(k, PY * N, 0, ldb, b);
        for (int e = t; e < PY * N; e += NT) sCD[(e % PY) + ldP * (e / PY)] = Cg[(size_t)e * ldb];
        const double* Dg = ph.D + gix(k, PY * M, 0, ldb, b);
        for (int e = t; e < PY * M; e += NT) sCD[(e % PY) + ldP * (N + e / PY)] = Dg[(size_t)e * ldb];
        const double* Lg = ph.lyy + gix(k, PY * PY, 0, ldb, b);
        for (int e = t; e < PY * PY; e += NT) sLyy[(e % PY) + ldP * (e / PY)] = Lg[(size_t)e * ldb];
        for (int j = t; j < PY; j += NT) sLy[j] = ph.ly[gix(k, PY, j, ldb, b)];
      }
    }
    __syncthreads();
    // ---- Gn = G + H d ; T = H [A B] ; S[C D] = lyy [C D]
    if (a2) for (int i = t; i < N; i += NT) { double s = sG[i]; for (int j = 0; j < N; ++j) s += sH[i + ldH * j] * sD[j]; sGn[i] = s; }
    if constexpr (WB) {
      // T = P [A2 B2] + [H(:,0:18), dt H(:,0:18), 0],  P = H(:,18:36)
      gemm_nt<N, N + M, 18, false, 0, NT>(sH + ldH * 18, ldH, sAB, ldA, nullptr, 0, nullptr, 0, t, a2, [&](int i, int j, double v) {
        if (j < 18) v += sH[i + ldH * j];
        else if (j < 36) v += dt * sH[i + ldH * (j - 18)];
        sT[i + ldH * j] = v;
      });
    } else {
      gemm_nt<N, N + M, N, false, 0, NT>(sH, ldH, sAB, ldA, nullptr, 0, nullptr, 0, t, a2, [&](int i, int j, double v) { sT[i + ldH * j] = v; });
    }
    if constexpr (PY > 0)
      gemm_nt<PY, N + M, PY, false, 0, NT>(sLyy, ldP, sCD, ldP, nullptr, 0, nullptr, 0, t, a2, [&](int i, int j, double v) { sSCD[i + ldP * j] = v; });
    __syncthreads();
    // ---- Q functions (H is dead from here on: Qxx is written over it)
    if (a2) {
      for (int j = t; j < N + M; j += NT) {
        double s = (j < N) ? ph.lx[gix(k, N, j, ldb, b)] : ph.lu[gix(k, M, j - N, ldb, b)];
        for (int i = 0; i < KA; ++i) s += sAB[i + ldA * j] * sGn[R0 + i];
        if constexpr (WB) { if (j < 18) s += sGn[j]; else if (j < 36) s += dt * sGn[j - 18]; }
        if constexpr (PY > 0) for (int i = 0; i < PY; ++i) s += sCD[i + ldP * j] * sLy[i];
        if (j < N) sQx[j] = s; else sQu[j - N] = s;
      }
    }
    {
      const double* lxxg = ph.lxx + gix(k, N * N, 0, ldb, b);
      const double* luug = ph.luu + gix(k, M * M, 0, ldb, b);
      // Qxx = lxx + A^T T_A (+ C^T S_C)
      gemm_nt<N, N, KA, true, PY, NT>(sAB, ldA, sT + R0, ldH, sCD, ldP, sSCD, ldP, t, a2, [&](int i, int j, double v) {
        if constexpr (WB) { if (i < 18) v += sT[i + ldH * j]; else v += dt * sT[(i - 18) + ldH * j]; }
        v += lxxg[(size_t)(i + N * j) * ldb];
        if (i == j) v += reg;
        sH[i + ldH * j] = v;
      });
      // Qux = B^T T_A (+ D^T S_C)
      gemm_nt<M, N, KA, true, PY, NT>(sAB + ldA * N, ldA, sT + R0, ldH, sCD + ldP * N, ldP, sSCD, ldP, t, a2,
                                      [&](int i, int j, double v) { sQux[i + ldM * j] = v; });
      // Quu = luu + B^T T_B (+ D^T S_D) + reg I
      gemm_nt<M, M, KA, true, PY, NT>(sAB + ldA * N, ldA, sT + R0 + ldH * N, ldH, sCD + ldP * N, ldP, sSCD + ldP * N, ldP, t, a2, [&](int i, int j, double v) {
        v += luug[(size_t)(i + M * j) * ldb];
        if (i == j) v += reg;
        sQuu[i + ldM * j] = v;
      });
    }
    __syncthreads();
    // ---- outputs Quu, Qux, Qu; copies for the factorisation / solves (A, B, T are dead)
    if (a2) {
      double* Quug = ph.Quu + gix(k, M * M, 0, ldb, b);
      for (int e = t; e < M * M; e += NT) { const int i = e % M, j = e / M; const double v = sQuu[i + ldM * j]; Quug[(size_t)e * ldb] = v; sL[i + ldM * j] = (i == j) ? v - 1e-9 : v; }
      double* Quxg = ph.Qux + gix(k, M * N, 0, ldb, b);
      for (int e = t; e < M * N; e += NT) { const int i = e % M, j = e / M; const double v = sQux[i + ldM * j]; Quxg[(size_t)e * ldb] = v; sK[i + ldM * j] = v; }
      for (int j = t; j < M; j += NT) { const double v = sQu[j]; ph.Qu[gix(k, M, j, ldb, b)] = v; sK[j + ldM * N] = v; }
    }
    __syncthreads();
    // ---- LDL^T of (Quu - 1e-9 I) by warp 0 (lane = row), pivots replaced by their reciprocals; PD test = every pivot > 0
    if (t < 32 && run && ok) {
      bool okw = true;
      for (int j = 0; j < M; ++j) {
        const double d = sL[j + ldM * j];
        if (!(d > 0.0)) { okw = false; break; }
        min_piv = fmin(min_piv, d);
        const double inv = 1.0 / d;
        const int i = j + 1 + t;
        double vi = 0;
        if (i < M) {
          vi = sL[i + ldM * j] * inv;
          for (int cc = j + 1; cc <= i; ++cc) sL[i + ldM * cc] -= vi * sL[cc + ldM * j];  // column j is read unscaled by every lane
        }
        __syncwarp();
        if (i < M) sL[i + ldM * j] = vi;
        if (t == 0) sL[j + ldM * j] = inv;
        __syncwarp();
      }
      if (t == 0) sm[L::vRed] = okw ? 1.0 : 0.0;
    }
    __syncthreads();
    if (run && ok && sm[L::vRed] == 0.0) ok = false;
    const bool a3 = run && ok;
    // ---- [K | dU] = -(Quu - 1e-9 I)^-1 [Qux | Qu], thread = column
    if (a3 && t < N + 1) {
      double x[M];
      double* col = sK + ldM * t;
#pragma unroll
      for (int i = 0; i < M; ++i) x[i] = col[i];
#pragma unroll
      for (int i = 0; i < M; ++i) {
#pragma unroll
        for (int l = 0; l < i; ++l) x[i] -= sL[i + ldM * l] * x[l];
      }
#pragma unroll
      for (int i = 0; i < M; ++i) x[i] *= sL[i + ldM * i];
#pragma unroll
      for (int i = M - 1; i >= 0; --i) {
#pragma unroll
        for (int l = i + 1; l < M; ++l) x[i] -= sL[l + ldM * i] * x[l];
      }
#pragma unroll
      for (int i = 0; i < M; ++i) col[i] = -x[i];
    }
    __syncthreads();
    // ---- value function: G = Qx + Qux^T dU ; H = sym(Qxx) + Qux^T K
    if (a3) {
      const double* dUs = sK + ldM * N;
      for (int j = t; j < N; j += NT) {
        double s = sQx[j];
        for (int i = 0; i < M; ++i) s += sQux[i + ldM * j] * dUs[i];
        sG[j] = s;
        ph.G[gix(k, N, j, ldb, b)] = s;
      }
      for (int j = t; j < M; j += NT) ph.dU[gix(k, M, j, ldb, b)] = dUs[j];
      double* Kg = ph.K + gix(k, M * N, 0, ldb, b);
      for (int e = t; e < M * N; e += NT) Kg[(size_t)e * ldb] = sK[(e % M) + ldM * (e / M)];
      // symmetrise Qxx in place: disjoint (i<j) pairs
      for (int e = t; e < N * N; e += NT) { const int i = e % N, j = e / N; if (i < j) { const double s = (sH[i + ldH * j] + sH[j + ldH * i]) / 2; sH[i + ldH * j] = s; sH[j + ldH * i] = s; } }
    }
    __syncthreads();
    gemm_nt<N, N, M, true, 0, NT>(sQux, ldM, sK, ldM, nullptr, 0, nullptr, 0, t, a3, [&](int i, int j, double v) { sH[i + ldH * j] += v; });
    __syncthreads();
  }
  // ---- G[0] += H[0] Defect[0]
  if (run && ok) for (int j = t; j < N; j += NT) sD[j] = ph.Defect[gix(0, N, j, ldb, b)];
  __syncthreads();
  if (run && ok) for (int i = t; i < N; i += NT) { double s = sG[i]; for (int j = 0; j < N; ++j) s += sH[i + ldH * j] * sD[j]; sGn[i] = s; }
  __syncthreads();
  if (run && ok) for (int i = t; i < N; i += NT) { sG[i] = sGn[i]; ph.G[gix(0, N, i, ldb, b)] = sGn[i]; }
  __syncthreads();
}

// multiple-shooting linear rollout of one phase (SinglePhase::linear_rollout), eps = 1; NT threads split the rows
template <int N, int M, int NT, class L>
__device__ void lin_phase2(const SolverDev& S, int pi, int b, int t, bool run, double* sm, double& dV1, double& dV2) {
  const PhaseDev& ph = S.ph[pi];
  const int ldb = S.ldb, h = ph.h;
  double* sDx = sm + L::vDx; double* sDxn = sm + L::vDxn; double* sDuL = sm + L::vDuL; double* sRed = sm + L::vRed;
  if (run) for (int i = t; i < N; i += NT) { const double v = sDx[i] + 1.0 * ph.Defect[gix(0, N, i, ldb, b)]; sDx[i] = v; ph.dX[gix(0, N, i, ldb, b)] = v; }
  __syncthreads();
  for (int k = 0; k < h; ++k) {
    double part1 = 0, part2 = 0;
    if (run) {
      const double* Kg = ph.K + gix(k, M * N, 0, ldb, b);
      for (int i = t; i < M; i += NT) {
        double s = 0;
        for (int j = 0; j < N; ++j) s += Kg[(size_t)(i + M * j) * ldb] * sDx[j];
        sDuL[i] = 1.0 * ph.dU[gix(k, M, i, ldb, b)] + s;
      }
    }
    __syncthreads();
    if (run) {
      const double* Ag = ph.A + gix(k, N * N, 0, ldb, b);
      const double* Bg = ph.Bm + gix(k, N * M, 0, ldb, b);
      const double* lxxg = ph.lxx + gix(k, N * N, 0, ldb, b);
      const double* luug = ph.luu + gix(k, M * M, 0, ldb, b);
      for (int i = t; i < N; i += NT) {
        double s = 0, s2 = 0, q = 0;
        for (int j = 0; j < N; ++j) { const double dxj = sDx[j]; s += Ag[(size_t)(i + N * j) * ldb] * dxj; q += lxxg[(size_t)(i + N * j) * ldb] * dxj; }
        for (int j = 0; j < M; ++j) s2 += Bg[(size_t)(i + N * j) * ldb] * sDuL[j];
        const double v = s + s2 + 1.0 * ph.Defect[gix(k + 1, N, i, ldb, b)];
        sDxn[i] = v;
        ph.dX[gix(k + 1, N, i, ldb, b)] = v;
        const double dxi = sDx[i];
        part1 += ph.lx[gix(k, N, i, ldb, b)] * dxi;
        part2 += dxi * q;
      }
      for (int i = t; i < M; i += NT) {
        double q = 0;
        for (int j = 0; j < M; ++j) q += luug[(size_t)(i + M * j) * ldb] * sDuL[j];
        const double dui = sDuL[i];
        part1 += ph.lu[gix(k, M, i, ldb, b)] * dui;
        part2 += dui * q;
      }
    }
    sRed[t] = part1;
    sRed[128 + t] = part2;
    __syncthreads();
    if (run) {
      if (t == 0) { double a1 = 0, a2 = 0; for (int i = 0; i < NT; ++i) { a1 += sRed[i]; a2 += sRed[128 + i]; } dV1 += a1; dV2 += a2; }
      for (int i = t; i < N; i += NT) sDx[i] = sDxn[i];
    }
    __syncthreads();
  }
  double part1 = 0, part2 = 0;
  if (run) {
    for (int i = t; i < N; i += NT) {
      double q = 0;
      for (int j = 0; j < N; ++j) q += ph.Phixx[(size_t)(i + N * j) * ldb + b] * sDx[j];
      const double dxi = sDx[i];
      part1 += ph.Phix[(size_t)i * ldb + b] * dxi;
      part2 += dxi * q;
    }
  }
  sRed[t] = part1;
  sRed[128 + t] = part2;
  __syncthreads();
  if (run && t == 0) { double a1 = 0, a2 = 0; for (int i = 0; i < NT; ++i) { a1 += sRed[i]; a2 += sRed[128 + i]; } dV1 += a1; dV2 += a2; }
  if (ph.has_next) {
    const int nn = ph.n_next;
    if (run) for (int i = t; i < nn; i += NT) { double s = 0; for (int j = 0; j < N; ++j) s += ph.Px[(size_t)(i + nn * j) * ldb + b] * sDx[j]; sDxn[i] = s; }
    __syncthreads();
    if (run) for (int i = t; i < nn; i += NT) sDx[i] = sDxn[i];
  }
  __syncthreads();
}

// DECK: 0 = HKD phases only (24,24,0); 1 = MHPC (WB 36,12,12 + SRB 12,12,0)
template <int DECK, int NT>
__global__ void __launch_bounds__(NT, 4) k_bwd2(const SolverDev* __restrict__ Sp) {
  typedef Bwd2Layout<(DECK == 0 ? 24 : 36), (DECK == 0 ? 24 : 12), (DECK == 0 ? 0 : 12), (DECK == 1)> L;
  constexpr int NX = (DECK == 0 ? 24 : 36);
  const SolverDev& S = *Sp;
  extern __shared__ double sm[];
  __shared__ double s_reg;
  __shared__ int s_state, s_regiter;  // 0 sweeping, 1 success, 2 gave up
  const int t = threadIdx.x, b = blockIdx.x;
  const CtrlDev& c = S.c;
  const CafeOptions& o = S.opt;
  const int ldb = S.ldb;
  if (b >= S.B || !c.active[b]) return;  // whole CTA
  int it = 0;
  if (t == 0) {
    s_state = 0; s_regiter = 0; s_reg = c.reg[b];
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
  __syncthreads();
  double min_piv = 1e300;
  while (s_state == 0) {
    bool ok = true;
    const double reg = s_reg;
    for (int pi = S.n_phases - 1; pi >= 0; --pi) {
      const int model = S.ph[pi].model, nm = S.ph[pi].has_next ? S.ph[pi + 1].model : -1;
      if constexpr (DECK == 0) {
        if (model == CAFE_MODEL_HKD) sweep_phase2<24, 24, 0, 24, false, NT, L>(S, pi, b, t, true, ok, reg, sm, min_piv);
      } else {
        if (model == CAFE_MODEL_SRB) sweep_phase2<12, 12, 0, 12, false, NT, L>(S, pi, b, t, true, ok, reg, sm, min_piv);
        else if (model == CAFE_MODEL_WB && nm == CAFE_MODEL_SRB) sweep_phase2<36, 12, 12, 12, true, NT, L>(S, pi, b, t, true, ok, reg, sm, min_piv);
        else if (model == CAFE_MODEL_WB) sweep_phase2<36, 12, 12, 36, true, NT, L>(S, pi, b, t, true, ok, reg, sm, min_piv);
      }
      (void)nm;
    }
    __syncthreads();
    if (t == 0) {
      s_regiter += 1;
      if (ok) s_state = 1;
      else {
        const double r = fmax(s_reg * o.update_regularization, 1e-03);
        s_reg = r;
        if (r > 1e2) s_state = 2;
      }
    }
    __syncthreads();
  }
  const bool success = s_state == 1;
  double dV1 = 0, dV2 = 0;
  {
    double* sDx = sm + L::vDx;
    for (int i = t; i < NX; i += NT) sDx[i] = 0.0;
    __syncthreads();
    for (int pi = 0; pi < S.n_phases; ++pi) {
      const int model = S.ph[pi].model;
      if constexpr (DECK == 0) {
        if (model == CAFE_MODEL_HKD) lin_phase2<24, 24, NT, L>(S, pi, b, t, success, sm, dV1, dV2);
      } else {
        if (model == CAFE_MODEL_WB) lin_phase2<36, 12, NT, L>(S, pi, b, t, success, sm, dV1, dV2);
        else if (model == CAFE_MODEL_SRB) lin_phase2<12, 12, NT, L>(S, pi, b, t, success, sm, dV1, dV2);
      }
    }
  }
  if (t == 0) {
    double r = s_reg / 20;
    if (r < 1e-06) r = 0;
    c.reg[b] = r;
    c.reg_total[b] += s_regiter;
    c.min_pivot[b] = fmin(c.min_pivot[b], min_piv);
    it = c.iter[b] - 1;
    double* tr = (it < CAFE_HIST_CAP) ? c.trace + ((size_t)it * 12) * ldb + b : nullptr;
    if (tr) { tr[(size_t)5 * ldb] = r; tr[(size_t)6 * ldb] = s_regiter; }
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
