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
//   * LDL^T of (Quu - 1e-9 I) and the solves for [K | dU] are one register-resident elimination (thread = column of
//     [Quu | Qux | Qu], one 64-thread named barrier per pivot), the other warps write the outputs and prefetch meanwhile,
//   * the solver descriptor is a __grid_constant__ kernel parameter: phase pointers and dimensions are constant-bank operands
//     (the L1 left beside ~210 KB of shared memory is too small to keep a descriptor in global memory resident),
//   * every per-knot operand is fetched with 8-byte cp.async (LDGSTS) straight into its shared-memory tile: the sweep issues the
//     copies of knot k-1 as soon as the Q functions of knot k are formed (A, B, C, D tiles are dead then) and waits for them
//     only at the top of the next iteration, the linear rollout double-buffers whole knots. No registers are spent on
//     staging and the ~1 us global latency is hidden behind the factorisation / solves.
#pragma once
#include <cooperative_groups.h>
#include "device_types.cuh"

namespace cafe_dev {

__device__ __forceinline__ void cp_async8(double* dst_smem, const double* src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(double* dst_smem, const double* src) {   // both 16-byte aligned; .cg: straight from L2
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }
// pacing barrier of the CTA cluster (no memory ordering needed: it only keeps the four problems that share every 32-byte
// sector of the problem-fastest arrays at the same knot, so that one DRAM fetch serves all four through L2)
// the CTAs of a cluster meet every CAFE_PACE-th knot (power of two): the L2 keeps a few knots of four problems resident, so
// a drift of a few knots still lets one DRAM fetch serve all four, and the barrier cost drops by that factor
#ifndef CAFE_PACE
#define CAFE_PACE 4
#endif
#ifndef CAFE_BWD_MINB
#define CAFE_BWD_MINB 4   // resident sweep CTAs per SM the kernel is compiled for (register cap 65536 / (NT * CAFE_BWD_MINB)); shared memory allows 4
#endif
__device__ __forceinline__ void cluster_pace() { asm volatile("barrier.cluster.arrive.relaxed.aligned;\nbarrier.cluster.wait.aligned;\n" ::: "memory"); }
// bit e of a structural-pattern mask (read-only, the same for every problem: served by L1)
__device__ __forceinline__ bool mask_bit(const unsigned long long* __restrict__ m, int e) { return (__ldg(m + (e >> 6)) >> (e & 63)) & 1ULL; }
__device__ __forceinline__ void prefetch_l1(const double* p) { asm volatile("prefetch.global.L1 [%0];\n" ::"l"(p)); }

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
#pragma unroll 6
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
#pragma unroll 6
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

// ---- fp64 tensor-core path (DMMA, mma.sync.m8n8k4.f64) -----------------------------------------------------------------
// The register-tiled gemm_nt above is bound by shared-memory operand traffic: every k step issues 8 LDS.64 (two wavefronts
// each, mostly broadcast) for 15 DFMAs, ncu shows the LSU data pipe at 45 % and the fp64 pipe at 10 %. One DMMA multiplies an
// 8x4 by a 4x8 fragment (256 FMAs) from ONE operand double per lane and matrix: 4 wavefronts per 256 FMAs instead of 16 per
// 480, an eighth of the issue slots, and every lane fetches a distinct element (no bandwidth lost to broadcasts).
// Fragment layout (PTX ISA, m8n8k4 .f64): g = lane/4, tg = lane%4:  a = A(g, tg),  b = B(tg, g),  c0,c1 = C(g, 2 tg + {0,1}).
// With every leading dimension = 4 (mod 8) doubles the 16 lanes of a half warp (4 values of g x 4 of tg) hit 32 distinct
// banks for both operand orientations (element + ld * k and k + ld * element).
#ifndef CAFE_BWD_MMA
#define CAFE_BWD_MMA 1
#endif
__host__ __device__ constexpr int ld_mma(int n) { return CAFE_BWD_MMA ? ((n % 8) <= 4 ? n - n % 8 + 4 : n - n % 8 + 12) : (n | 1); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  // volatile: the instruction is warp-collective; it must never be duplicated into lane-predicated copies
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// tile lists: which 8x8 output tiles a product needs, in column-major order
struct NoPre { __device__ __forceinline__ double operator()(int, int) const { return 0.0; } };
template <int MM, int NN>
struct TilesFull {
  static constexpr int MT = (MM + 7) / 8, COUNT = MT * ((NN + 7) / 8);
  static __device__ __forceinline__ void map(int q, int& m0, int& n0) { m0 = (q % MT) * 8; n0 = (q / MT) * 8; }
};
// the (N+M) x (N+M) product [A B]^T T without its upper right block (rows < N, columns >= N): Qxx, Qux, Quu but not Qxu
template <int N, int M>
struct TilesQ {
  static constexpr int MT = (N + M + 7) / 8, NFULL = (N + 7) / 8, MS = N / 8, NREST = (N + M + 7) / 8 - NFULL;
  static constexpr int COUNT = NFULL * MT + NREST * (MT - MS);
  static __device__ __forceinline__ void map(int q, int& m0, int& n0) {
    if (q < NFULL * MT) { m0 = (q % MT) * 8; n0 = (q / MT) * 8; }
    else { const int r = q - NFULL * MT; m0 = (MS + r % (MT - MS)) * 8; n0 = (NFULL + r / (MT - MS)) * 8; }
  }
};

// C(i,j) = epi(i,j, sum_{l<KK} opA(i,l) B(l,j) + sum_{l<KK2} A2[l + lda2*i] B2[l + ldb2*j])   for i < MM, j < NN, tiles of TL
//   opA(i,l) = TA ? A[l + lda*i] : A[i + lda*l];  B(l,j) = B[l + ldb*j].  The NT/32 warps take the tiles round robin, four at a
//   time (independent accumulators). Rows / columns past MM / NN are read from whatever follows in shared memory and only
//   reach outputs that are dropped. pre(i,j) is added to the product (a global-memory addend, fetched before the k loop so that
//   its latency hides behind the DMMAs). k runs to the next multiple of 4 WITHOUT a mask (a per-lane select around a collective
//   instruction invites lane-predicated code): the caller guarantees that the k padding of one operand is zero and that of the
//   other finite (Bwd2Layout: zero pad rows of [A B], zero pad columns behind H; shared memory is cleared at kernel start).
template <int MM, int NN, int KK, bool TA, int KK2, int NT, class TL, class Pre, class Epi>
__device__ __forceinline__ void gemm_mma(const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb,
                                         const double* __restrict__ A2, int lda2, const double* __restrict__ B2, int ldb2, int t, bool run, Pre pre, Epi epi) {
  constexpr int NW = NT / 32, G = (NT > 128) ? 2 : 4, KS = (KK + 3) / 4, KS2 = (KK2 + 3) / 4;
  if (!run) return;
  __syncwarp();
  const int w = t >> 5, lane = t & 31, g = lane >> 2, tg = lane & 3;
  for (int q0 = w; q0 < TL::COUNT; q0 += NW * G) {
    double c[G][2];
    int m0[G], n0[G];
    const double* ap[G]; const double* bp[G]; const double* ap2[G]; const double* bp2[G];
#pragma unroll
    for (int u = 0; u < G; ++u) {
      const int q = q0 + u * NW;
      TL::map(q < TL::COUNT ? q : q0, m0[u], n0[u]);   // a pass that runs out of tiles repeats its first one (result dropped)
      ap[u] = TA ? A + tg + lda * (m0[u] + g) : A + (m0[u] + g) + lda * tg;
      bp[u] = B + tg + ldb * (n0[u] + g);
      if constexpr (KK2 > 0) { ap2[u] = A2 + tg + lda2 * (m0[u] + g); bp2[u] = B2 + tg + ldb2 * (n0[u] + g); }
      c[u][0] = 0; c[u][1] = 0;
    }
    // addends of the epilogue that come from global memory: requested now, consumed after the k loop
    double pv[G][2];
#pragma unroll
    for (int u = 0; u < G; ++u) {
      const int i = m0[u] + g, j = n0[u] + 2 * tg;
      pv[u][0] = (i < MM && j < NN) ? pre(i, j) : 0.0;
      pv[u][1] = (i < MM && j + 1 < NN) ? pre(i, j + 1) : 0.0;
    }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int u = 0; u < G; ++u) dmma884(c[u][0], c[u][1], TA ? ap[u][ks * 4] : ap[u][ks * 4 * lda], bp[u][ks * 4]);
    }
    if constexpr (KK2 > 0) {
#pragma unroll
      for (int ks = 0; ks < KS2; ++ks) {
#pragma unroll
        for (int u = 0; u < G; ++u) dmma884(c[u][0], c[u][1], ap2[u][ks * 4], bp2[u][ks * 4]);
      }
    }
#pragma unroll
    for (int u = 0; u < G; ++u) {
      if (q0 + u * NW < TL::COUNT) {
        const int i = m0[u] + g, j = n0[u] + 2 * tg;
        if (i < MM && j < NN) epi(i, j, c[u][0] + pv[u][0]);
        if (i < MM && j + 1 < NN) epi(i, j + 1, c[u][1] + pv[u][1]);
      }
    }
  }
}

// ---- the same product with PAIRS of tiles that share one operand fragment: 3 LDS.64 per two DMMAs instead of 4 (the k loops' shared-memory
// wavefronts are a quarter of the kernel's LSU traffic). SH = 1: two row tiles of one tile column share b; SH = 2: two column tiles of one
// tile row share a. Every tile still accumulates its k steps in the same order from zero, so the results are bit-identical to gemm_mma.
#ifndef CAFE_BWD_PAIR
#define CAFE_BWD_PAIR 1
#endif
template <int MM, int NN, int SH>
struct PairsFull {
  static constexpr int MT = (MM + 7) / 8, NTC = (NN + 7) / 8, MP = SH == 1 ? (MT + 1) / 2 : MT, NP = SH == 2 ? (NTC + 1) / 2 : NTC, COUNT = MP * NP;
  static __device__ __forceinline__ void map(int p, int& m0, int& n0, bool& two) {
    const int pm = p % MP, pn = p / MP;
    if constexpr (SH == 1) { m0 = pm * 16; n0 = pn * 8; two = 2 * pm + 1 < MT; }
    else { m0 = pm * 8; n0 = pn * 16; two = 2 * pn + 1 < NTC; }
  }
};
// TilesQ in pairs of row tiles (needs an even number of row tiles in both column groups)
template <int N, int M>
struct PairsQ {
  static constexpr int MT = (N + M + 7) / 8, NFULL = (N + 7) / 8, MS = N / 8, NREST = MT - NFULL;
  static constexpr bool OK = MT % 2 == 0 && MS % 2 == 0;
  static constexpr int PF = MT / 2, PR = (MT - MS) / 2, COUNT = NFULL * PF + NREST * PR;
  static __device__ __forceinline__ void map(int p, int& m0, int& n0, bool& two) {
    two = true;
    if (p < NFULL * PF) { m0 = (p % PF) * 16; n0 = (p / PF) * 8; }
    else { const int r = p - NFULL * PF; m0 = MS * 8 + (r % PR) * 16; n0 = (NFULL + r / PR) * 8; }
  }
};
// the pair lists cover exactly the tiles of the tile lists they replace (whole-body phase: N = 36, M = 12)
static_assert(PairsQ<36, 12>::OK && 2 * PairsQ<36, 12>::COUNT == TilesQ<36, 12>::COUNT, "Q product: 16 pairs = 32 tiles");
static_assert(PairsFull<36, 48, 2>::COUNT == 15 && 2 * PairsFull<36, 48, 2>::COUNT == TilesFull<36, 48>::COUNT, "T = P [A2 B2]: 15 pairs = 30 tiles");
static_assert(PairsFull<36, 36, 2>::COUNT == 15 && TilesFull<36, 36>::COUNT == 25, "H += Qux^T K: 15 pairs (5 of them half) = 25 tiles");
template <int MM, int NN, int KK, bool TA, int KK2, int NT, int SH, class PL, class Pre, class Epi>
__device__ __forceinline__ void gemm_mma_pair(const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb,
                                              const double* __restrict__ A2, int lda2, const double* __restrict__ B2, int ldb2, int t, bool run, Pre pre, Epi epi) {
  constexpr int NW = NT / 32, G2 = (NT > 128) ? 1 : 2, KS = (KK + 3) / 4, KS2 = (KK2 + 3) / 4;
  static_assert(SH == 1 || SH == 2, "which operand is shared");
  if (!run) return;
  __syncwarp();
  const int w = t >> 5, lane = t & 31, g = lane >> 2, tg = lane & 3;
  for (int p0 = w; p0 < PL::COUNT; p0 += NW * G2) {
    double c[G2][2][2];
    int m0[G2][2], n0[G2][2];
    bool two[G2];
    const double* ap[G2][2]; const double* bp[G2][2]; const double* ap2[G2][2]; const double* bp2[G2][2];
#pragma unroll
    for (int u = 0; u < G2; ++u) {
      const int p = p0 + u * NW;
      int mm, nn; bool tw;
      PL::map(p < PL::COUNT ? p : p0, mm, nn, tw);   // a pass that runs out of pairs repeats its first one (result dropped)
      two[u] = tw;
      m0[u][0] = mm; n0[u][0] = nn;
      m0[u][1] = (SH == 1 && tw) ? mm + 8 : mm;        // half a pair: the second tile repeats the first (result dropped)
      n0[u][1] = (SH == 2 && tw) ? nn + 8 : nn;
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        ap[u][v] = TA ? A + tg + lda * (m0[u][v] + g) : A + (m0[u][v] + g) + lda * tg;
        bp[u][v] = B + tg + ldb * (n0[u][v] + g);
        if constexpr (KK2 > 0) { ap2[u][v] = A2 + tg + lda2 * (m0[u][v] + g); bp2[u][v] = B2 + tg + ldb2 * (n0[u][v] + g); }
        c[u][v][0] = 0; c[u][v][1] = 0;
      }
    }
    double pv[G2][2][2];
#pragma unroll
    for (int u = 0; u < G2; ++u)
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        const int i = m0[u][v] + g, j = n0[u][v] + 2 * tg;
        const bool on = v == 0 || two[u];
        pv[u][v][0] = (on && i < MM && j < NN) ? pre(i, j) : 0.0;
        pv[u][v][1] = (on && i < MM && j + 1 < NN) ? pre(i, j + 1) : 0.0;
      }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int u = 0; u < G2; ++u) {
        if constexpr (SH == 1) {
          const double b = bp[u][0][ks * 4];
          const double a0 = TA ? ap[u][0][ks * 4] : ap[u][0][ks * 4 * lda], a1 = TA ? ap[u][1][ks * 4] : ap[u][1][ks * 4 * lda];
          dmma884(c[u][0][0], c[u][0][1], a0, b);
          dmma884(c[u][1][0], c[u][1][1], a1, b);
        } else {
          const double a = TA ? ap[u][0][ks * 4] : ap[u][0][ks * 4 * lda];
          const double b0 = bp[u][0][ks * 4], b1 = bp[u][1][ks * 4];
          dmma884(c[u][0][0], c[u][0][1], a, b0);
          dmma884(c[u][1][0], c[u][1][1], a, b1);
        }
      }
    }
    if constexpr (KK2 > 0) {
#pragma unroll
      for (int ks = 0; ks < KS2; ++ks) {
#pragma unroll
        for (int u = 0; u < G2; ++u) {
          if constexpr (SH == 1) {
            const double b = bp2[u][0][ks * 4];
            const double a0 = ap2[u][0][ks * 4], a1 = ap2[u][1][ks * 4];
            dmma884(c[u][0][0], c[u][0][1], a0, b);
            dmma884(c[u][1][0], c[u][1][1], a1, b);
          } else {
            const double a = ap2[u][0][ks * 4];
            const double b0 = bp2[u][0][ks * 4], b1 = bp2[u][1][ks * 4];
            dmma884(c[u][0][0], c[u][0][1], a, b0);
            dmma884(c[u][1][0], c[u][1][1], a, b1);
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < G2; ++u) {
      if (p0 + u * NW < PL::COUNT) {
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          if (v == 0 || two[u]) {
            const int i = m0[u][v] + g, j = n0[u][v] + 2 * tg;
            if (i < MM && j < NN) epi(i, j, c[u][v][0] + pv[u][v][0]);
            if (i < MM && j + 1 < NN) epi(i, j + 1, c[u][v][1] + pv[u][v][1]);
          }
        }
      }
    }
  }
}

// shared-memory plan (doubles) of one problem whose largest phase is (NX, MX, PX); WBS: structured whole-body storage
template <int NX, int MX, int PX, bool WBS>
struct Bwd2Layout {
  static constexpr int ldH = ld_mma(NX), KA = WBS ? 18 : NX, ldA = ld_mma(KA), ldM = ld_mma(MX), ldP = ld_mma(PX > 0 ? PX : 1);
  // vectors
  static constexpr int vG = 0, vGn = NX, vQx = 2 * NX, vD = 3 * NX, vDx = 4 * NX, vDxn = 5 * NX, vQu = 6 * NX, vDu = 6 * NX + MX,
                       vDuL = 6 * NX + 2 * MX, vLy = 6 * NX + 3 * MX, vRed = vLy + PX + 1, nVec = (vRed + 32 + 1) & ~1;   // vRed: flag + up to 24 pivots (the two NT-wide
                       // rows of the final dV reduction live in the linear rollout's stage buffers, dead by then: oRed)
  static constexpr int oH = nVec;                       // H / Qxx / H_new      NX x NX (ldH)
  static constexpr int oAB = oH + ldH * (NX + 2);       // [A B] rows KA (ldA) x (NX+MX); two zero pad columns behind H (k padding of P = H(:,18:36))
  static constexpr int szAB = ldA * (NX + MX);
  static constexpr int oT = oAB + szAB;                 // T = H [A B]  NX x (NX+MX) (ldH); reused for L (MX x MX, ldM)
  static constexpr int oQux = oT + ldH * (NX + MX);     // Qux MX x NX (ldM)
  static constexpr int oQuu = oQux + ldM * NX;          // Quu MX x MX (ldM)
  static constexpr int oCD = oQuu + ldM * MX;           // [C D]  PX x (NX+MX) (ldP)
  static constexpr int oSCD = oCD + (PX > 0 ? ldP * (NX + MX) : 0);   // lyy [C D]; reused for [K | dU] when PX > 0
  static constexpr int oLyy = oSCD + (PX > 0 ? ldP * (NX + MX) : 0);  // lyy PX x PX (ldP)
  static constexpr int szK = ldM * (NX + 1);            // [K | dU]  MX x (NX+1) (ldM): its own tile when there is no lyy [C D] tile
  static constexpr int oK = PX > 0 ? oSCD : oLyy;
  static constexpr int endSweep = (PX > 0 ? oLyy + ldP * PX : oK + szK) + 2;
  static_assert(PX == 0 || szK <= ldP * (NX + MX), "[K | dU] does not fit the lyy [C D] tile");
  // linear rollout: two stages of {K, A(stored rows), B, lxx, luu, lx, lu, d, dU}
  static constexpr int lK = 0, lA = lK + ldM * NX, lB = lA + ldA * NX, lLxx = lB + ldA * MX, lLuu = lLxx + ldH * NX, lLx = lLuu + ldM * MX,
                       lLu = lLx + NX, lD = lLu + MX, lDU = lD + NX, szLin = (lDU + MX + 2) & ~1;   // even: the second stage stays 16-byte aligned
  static constexpr int oLin = nVec;
  static constexpr int endLin = oLin + 2 * szLin;
  static constexpr int oRed = oLin;   // 2 x 256 doubles
  static_assert(2 * szLin >= 2 * 256, "reduction rows");
  // slack behind the sweep's tiles only: fragment loads of partial edge tiles run past the last tile. (HKD: 57.3 KB, four problems per SM -
  // with the reduction rows in nVec and the slack behind the larger linear-rollout plan it was 61.7 KB, three per SM.)
  static constexpr int total = (endSweep + 64 > endLin ? endSweep + 64 : endLin);
};

// One phase of the sweep. N, M, PY: phase dimensions; NNEXT: state dimension of the next phase (for the jump); WB: use the
// whole-body block structure. On entry, when the phase has a successor, sG/sH hold G0+/H0+ of the successor (ld ldH).
template <int N, int M, int PY, int NNEXT, bool WB, int NT, class L>
__device__ void sweep_phase2(const SolverDev& S, int pi, int b, int t, bool run, bool& ok, double reg, double* sm, double& min_piv, double& dv1, double& dv2) {
  const PhaseDev& ph = S.ph[pi];
  const int ldb = S.ldb, h = ph.h;
  constexpr int ldH = L::ldH, ldA = L::ldA, ldM = L::ldM, ldP = L::ldP;
  constexpr int KA = WB ? 18 : N;   // stored rows of [A B]
  constexpr int R0 = WB ? 18 : 0;   // first stored row
  double* sG = sm + L::vG; double* sGn = sm + L::vGn; double* sQx = sm + L::vQx; double* sD = sm + L::vD;
  double* sQu = sm + L::vQu; double* sDu = sm + L::vDu; double* sLy = sm + L::vLy;
  double* sH = sm + L::oH; double* sAB = sm + L::oAB; double* sT = sm + L::oT; double* sQux = sm + L::oQux; double* sQuu = sm + L::oQuu;
  double* sCD = sm + L::oCD; double* sSCD = sm + L::oSCD; double* sLyy = sm + L::oLyy;
  double* sK = sm + L::oK;   // [K | dU] (over lyy [C D], dead once the Q functions are formed)
  double* sL = sT;    // LDL^T factor after T is dead
  (void)sLy; (void)sCD; (void)sSCD; (void)sLyy;
  const double dt = ph.dt;
  // expected cost change from the sweep itself, used when there is no linear rollout (MS = false): dV_k = -Qu^T dU, dV_1 -= dV_k, dV_2 += dV_k
  // per phase, phases summed last to first (SinglePhase.cpp:383-387, MultiPhaseDDP.cpp:174-213); thread 0 keeps the sums
  double pdv1 = 0, pdv2 = 0;

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

  // asynchronous staging of [A B] (stored rows), [C D], lyy, ly, Defect[k+1] of knot k into their tiles
  const unsigned long long* hm = WB ? nullptr : ph.hkd_mask;   // HKD: structural patterns of A, B, lxx, luu (else nullptr)
  auto stage = [&](int k, int t0, int nt) {
    if constexpr (WB) {
      // the producer (k_wb_sens) wrote these two tiles problem-major in exactly this layout: contiguous 16-byte copies
      static_assert(ldA * (N + M) == CAFE_WB_AB_TILE && ldP * (N + M) == CAFE_WB_CD_TILE && L::oAB % 2 == 0 && L::oCD % 2 == 0, "tile layout");
      const double* ABt = ph.ABpm + ((size_t)b * h + k) * CAFE_WB_AB_TILE;
      for (int e = t0; e < CAFE_WB_AB_TILE / 2; e += nt) cp_async16(sAB + 2 * e, ABt + 2 * e);
      const double* CDt = ph.CDpm + ((size_t)b * h + k) * CAFE_WB_CD_TILE;
      for (int e = t0; e < CAFE_WB_CD_TILE / 2; e += nt) cp_async16(sCD + 2 * e, CDt + 2 * e);
    } else {
      const double* Ag = ph.A + gix(k, N * N, 0, ldb, b);
      const double* Bg = ph.Bm + gix(k, N * M, 0, ldb, b);
      if (hm != nullptr) {   // structural non-zeros only (86 + 60 of 1152 for the HKD model)
        for (int e = t0; e < N * N; e += nt) if (mask_bit(hm, e)) cp_async8(sAB + (e % N) + ldA * (e / N), Ag + (size_t)e * ldb);
        for (int e = t0; e < N * M; e += nt) if (mask_bit(hm + 9, e)) cp_async8(sAB + (e % N) + ldA * (N + e / N), Bg + (size_t)e * ldb);
      } else {
        for (int e = t0; e < KA * N; e += nt) { const int i = e % KA, j = e / KA; cp_async8(sAB + i + ldA * j, Ag + (size_t)((R0 + i) + N * j) * ldb); }
        for (int e = t0; e < KA * M; e += nt) { const int i = e % KA, j = e / KA; cp_async8(sAB + i + ldA * (N + j), Bg + (size_t)((R0 + i) + N * j) * ldb); }
      }
    }
    for (int j = t0; j < N; j += nt) cp_async8(sD + j, ph.Defect + gix(k + 1, N, j, ldb, b));
    if constexpr (PY > 0) {
      const double* Lg = ph.lyy + gix(k, PY * PY, 0, ldb, b);
      // lyy: one 3 x 3 block per foot (GRF barrier on the output); the rest of the tile stays zero (cleared at kernel start)
      for (int e = t0; e < PY * 3; e += nt) { const int j = e / 3, i = 3 * (j / 3) + e % 3; cp_async8(sLyy + i + ldP * j, Lg + (size_t)(i + PY * j) * ldb); }
      for (int j = t0; j < PY; j += nt) cp_async8(sLy + j, ph.ly + gix(k, PY, j, ldb, b));
    }
    cp_async_commit();
  };
  if constexpr (CAFE_BWD_MMA && ldA > KA) {
    // k padding of [A B] (rows KA..ldA-1, never staged): zero again, the jump above used the area for Px
    constexpr int PAD = ldA - KA;
    if (run && ok) for (int e = t; e < PAD * (N + M); e += NT) sAB[KA + (e % PAD) + ldA * (e / PAD)] = 0.0;
  }
  if (hm != nullptr) {
    // only the structural non-zeros of [A B] are staged: the rest of the tile (used for Px above) must be zero
    __syncthreads();
    if (run && ok) for (int e = t; e < ldA * (N + M); e += NT) sAB[e] = 0.0;
    __syncthreads();
  }
  if (run && ok && h > 0) stage(h - 1, t, NT);
  for (int k = h - 1; k >= 0; --k) {
    const bool a2 = run && ok;
    cp_async_wait_all();
    __syncthreads();
    if ((k & (CAFE_PACE - 1)) == 0) cluster_pace();
    // ---- Gn = G + H d ; T = H [A B] ; S[C D] = lyy [C D]
    if (a2) for (int i = NT - 1 - t; i < N; i += NT) { double s = sG[i]; for (int j = 0; j < N; ++j) s += sH[i + ldH * j] * sD[j]; sGn[i] = s; }   // last warps: fewest tiles
#if CAFE_BWD_MMA
    if constexpr (WB) {
      // T = P [A2 B2] + [H(:,0:18), dt H(:,0:18), 0],  P = H(:,18:36)
      auto epiT = [&](int i, int j, double v) {
        if (j < 18) v += sH[i + ldH * j];
        else if (j < 36) v += dt * sH[i + ldH * (j - 18)];
        sT[i + ldH * j] = v;
      };
#if CAFE_BWD_PAIR
      gemm_mma_pair<N, N + M, 18, false, 0, NT, 2, PairsFull<N, N + M, 2>>(sH + ldH * 18, ldH, sAB, ldA, nullptr, 0, nullptr, 0, t, a2, NoPre(), epiT);
#else
      gemm_mma<N, N + M, 18, false, 0, NT, TilesFull<N, N + M>>(sH + ldH * 18, ldH, sAB, ldA, nullptr, 0, nullptr, 0, t, a2, NoPre(), epiT);
#endif
    } else {
      gemm_mma<N, N + M, N, false, 0, NT, TilesFull<N, N + M>>(sH, ldH, sAB, ldA, nullptr, 0, nullptr, 0, t, a2, NoPre(), [&](int i, int j, double v) { sT[i + ldH * j] = v; });
    }
    if constexpr (PY > 0)
      gemm_mma<PY, N + M, PY, false, 0, NT, TilesFull<PY, N + M>>(sLyy, ldP, sCD, ldP, nullptr, 0, nullptr, 0, t, a2, NoPre(), [&](int i, int j, double v) { sSCD[i + ldP * j] = v; });
#else
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
#endif
    __syncthreads();
    // ---- Q functions (H is dead from here on: Qxx is written over it)
    if (a2) {
#if !CAFE_BWD_MMA
      // the epilogues below add lxx / luu from global memory: start pulling them into L1 now
      const double* lxxp = ph.lxx + gix(k, N * N, 0, ldb, b);
      for (int e = t; e < N * N; e += NT) prefetch_l1(lxxp + (size_t)e * ldb);
      const double* luup = ph.luu + gix(k, M * M, 0, ldb, b);
      for (int e = t; e < M * M; e += NT) prefetch_l1(luup + (size_t)e * ldb);
#endif
      for (int j = t; j < N + M; j += NT) {
        double s = (j < N) ? ph.lx[gix(k, N, j, ldb, b)] : ph.lu[gix(k, M, j - N, ldb, b)];
        for (int i = 0; i < KA; ++i) s += sAB[i + ldA * j] * sGn[R0 + i];
        if constexpr (WB) { if (j < 18) s += sGn[j]; else if (j < 36) s += dt * sGn[j - 18]; }
        if constexpr (PY > 0) for (int i = 0; i < PY; ++i) s += sCD[i + ldP * j] * sLy[i];
        if (j < N) sQx[j] = s; else sQu[j - N] = s;
      }
    }
#if CAFE_BWD_MMA
    {
      // [Qxx Qxu; Qux Quu] = [A B]^T T_rows (+ [C D]^T lyy [C D]) is ONE (N+M)^2 product; its upper right block is not needed
      const double* lxxg = ph.lxx + gix(k, N * N, 0, ldb, b);
      const double* luug = ph.luu + gix(k, M * M, 0, ldb, b);
      const unsigned long long* lmask = WB ? ph.lxx_mask + (size_t)k * CAFE_LXX_MASK_WORDS : nullptr;
      (void)lmask;
      auto preQ = [&](int i, int j) -> double {   // lxx / luu: 8-byte loads of a problem-fastest array, issued before the k loop
          if constexpr (WB) {           // only the structural non-zeros: lxx by its per-knot pattern, luu is diagonal
            if (j < N) return (i < N && mask_bit(lmask, i + N * j)) ? lxxg[(size_t)(i + N * j) * ldb] : 0.0;
            return (i >= N && i == j) ? luug[(size_t)((i - N) + M * (j - N)) * ldb] : 0.0;
          } else {
            if (j < N) return (i < N && (hm == nullptr || mask_bit(hm + 18, i + N * j))) ? lxxg[(size_t)(i + N * j) * ldb] : 0.0;
            return (i >= N && (hm == nullptr || mask_bit(hm + 27, (i - N) + M * (j - N)))) ? luug[(size_t)((i - N) + M * (j - N)) * ldb] : 0.0;
          }
        };
      auto epiQ = [&](int i, int j, double v) {
        if (j < N) {
          if (i < N) {   // Qxx = lxx + A^T T_A (+ C^T S_C) + reg I   (H is dead: written over it)
            if constexpr (WB) { if (i < 18) v += sT[i + ldH * j]; else v += dt * sT[(i - 18) + ldH * j]; }
            if (i == j) v += reg;
            sH[i + ldH * j] = v;
          } else {       // Qux = B^T T_A (+ D^T S_C)
            sQux[(i - N) + ldM * j] = v;
          }
        } else if (i >= N) {   // Quu = luu + B^T T_B (+ D^T S_D) + reg I
          const int iu = i - N, ju = j - N;
          if (iu == ju) v += reg;
          sQuu[iu + ldM * ju] = v;
        }
      };
      if constexpr (CAFE_BWD_PAIR && WB && PairsQ<N, M>::OK)
        gemm_mma_pair<N + M, N + M, KA, true, PY, NT, 1, PairsQ<N, M>>(sAB, ldA, sT + R0, ldH, sCD, ldP, sSCD, ldP, t, a2, preQ, epiQ);
      else
        gemm_mma<N + M, N + M, KA, true, PY, NT, TilesQ<N, M>>(sAB, ldA, sT + R0, ldH, sCD, ldP, sSCD, ldP, t, a2, preQ, epiQ);
    }
#else
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
#endif
    __syncthreads();
    // ---- factorisation and solves, fused. Thread c < M+N+1 of the first NE threads keeps column c of
    //      [Quu - 1e-9 I | Qux | Qu] in registers; at step j the owner of column j publishes the pivot and the multipliers
    //      l_ij, every later column is updated in registers (Schur update of the matrix columns = forward substitution of the
    //      right-hand sides). PD test (Eigen's isPositive on Quu - 1e-9 I, SinglePhase.cpp:157-165) = every pivot > 0.
    //      The remaining warps meanwhile write Quu, Qux, Qu and start fetching knot k-1 (A, B, C, D, lyy, ly, d tiles are dead).
    constexpr int NC = M + N + 1, NE = (NC + 31) / 32 * 32;
    constexpr bool SPLIT = NT > NE;
    if (a2) {
      if (!SPLIT || t >= NE) {
        const int t0 = SPLIT ? t - NE : t;
        constexpr int nt = SPLIT ? NT - NE : NT;
        if (k > 0) stage(k - 1, t0, nt);
        // Quu, Qux leave as whole tiles (problem-major, 16-byte stores): only the result packers read them
        static_assert(L::oQuu % 2 == 0 && L::oQux % 2 == 0 && (ldM * M) % 2 == 0 && (ldM * N) % 2 == 0, "Quu / Qux tile layout");
        double2* Quut = reinterpret_cast<double2*>(ph.Quu + ((size_t)b * h + k) * (ldM * M));
        for (int e = t0; e < ldM * M / 2; e += nt) Quut[e] = reinterpret_cast<const double2*>(sQuu)[e];
        double2* Quxt = reinterpret_cast<double2*>(ph.Qux + ((size_t)b * h + k) * (ldM * N));
        for (int e = t0; e < ldM * N / 2; e += nt) Quxt[e] = reinterpret_cast<const double2*>(sQux)[e];
        for (int j = t0; j < M; j += nt) ph.Qu[gix(k, M, j, ldb, b)] = sQu[j];
      }
      if (t < NE) {
        double x[M];
        const double* colp = (t < M) ? sQuu + ldM * t : (t < M + N) ? sQux + ldM * (t - M) : sQu;
#pragma unroll
        for (int i = 0; i < M; ++i) x[i] = (t < NC) ? colp[i] : 0.0;
#pragma unroll
        for (int i = 0; i < M; ++i) if (i == t) x[i] -= 1e-9;
        bool pd = true;
#pragma unroll
        for (int j = 0; j < M; ++j) {
          if (t == j) {
            const double d = x[j];
            sm[L::vRed + 1 + j] = d;
            const double inv = 1.0 / d;
            sL[j + ldM * j] = inv;
#pragma unroll
            for (int i = j + 1; i < M; ++i) sL[i + ldM * j] = x[i] * inv;
          }
          asm volatile("bar.sync 1, %0;" ::"n"(NE) : "memory");
          const double d = sm[L::vRed + 1 + j];
          if (!(d > 0.0)) { pd = false; break; }
          min_piv = fmin(min_piv, d);
          if (t > j) {
#pragma unroll
            for (int i = j + 1; i < M; ++i) x[i] -= sL[i + ldM * j] * x[j];
          }
        }
        if (pd && t >= M && t < NC) {
          // [K | dU] = -(L D L^T)^-1 [Qux | Qu]: x holds L^-1 b
#pragma unroll
          for (int i = 0; i < M; ++i) x[i] *= sL[i + ldM * i];
#pragma unroll
          for (int i = M - 1; i >= 0; --i) {
#pragma unroll
            for (int l = i + 1; l < M; ++l) x[i] -= sL[l + ldM * i] * x[l];
          }
          double* col = sK + ldM * (t - M);
#pragma unroll
          for (int i = 0; i < M; ++i) col[i] = -x[i];
        }
        if (t == 0) sm[L::vRed] = pd ? 1.0 : 0.0;
      }
    }
    __syncthreads();
    if (run && ok && sm[L::vRed] == 0.0) ok = false;
    const bool a3 = run && ok;
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
      if (t == 0 && !S.opt.MS) {
        double d = 0;
        for (int i = 0; i < M; ++i) d += sQu[i] * dUs[i];
        const double dV_k = -d;
        pdv1 -= dV_k; pdv2 += dV_k;
      }
      double* Kg = ph.K + gix(k, M * N, 0, ldb, b);
      for (int e = t; e < M * N; e += NT) Kg[(size_t)e * ldb] = sK[(e % M) + ldM * (e / M)];
      {   // second, problem-major copy (the ld x N tile as it is) for the linear rollout of this kernel: 16-byte stores / copies
        static_assert(L::oK % 2 == 0 && (ldM * N) % 2 == 0, "K tile layout");
        double2* Kt = reinterpret_cast<double2*>(ph.Kpm + ((size_t)b * h + k) * (ldM * N));
        for (int e = t; e < ldM * N / 2; e += NT) Kt[e] = reinterpret_cast<const double2*>(sK)[e];
      }
      // symmetrise Qxx in place: disjoint (i<j) pairs
      {
        int i = t % N, j = t / N;   // element (i, j), advanced by NT per step without divisions
        for (int e = t; e < N * N; e += NT) {
          if (i < j) { const double s = (sH[i + ldH * j] + sH[j + ldH * i]) / 2; sH[i + ldH * j] = s; sH[j + ldH * i] = s; }
          i += NT % N; j += NT / N;
          if (i >= N) { i -= N; ++j; }
        }
      }
    }
    __syncthreads();
#if CAFE_BWD_MMA
    if constexpr (CAFE_BWD_PAIR && WB)
      gemm_mma_pair<N, N, M, true, 0, NT, 2, PairsFull<N, N, 2>>(sQux, ldM, sK, ldM, nullptr, 0, nullptr, 0, t, a3, NoPre(), [&](int i, int j, double v) { sH[i + ldH * j] += v; });
    else
      gemm_mma<N, N, M, true, 0, NT, TilesFull<N, N>>(sQux, ldM, sK, ldM, nullptr, 0, nullptr, 0, t, a3, NoPre(), [&](int i, int j, double v) { sH[i + ldH * j] += v; });
#else
    gemm_nt<N, N, M, true, 0, NT>(sQux, ldM, sK, ldM, nullptr, 0, nullptr, 0, t, a3, [&](int i, int j, double v) { sH[i + ldH * j] += v; });
#endif
    __syncthreads();
  }
  // ---- G[0] += H[0] Defect[0]
  if (run && ok) for (int j = t; j < N; j += NT) sD[j] = ph.Defect[gix(0, N, j, ldb, b)];
  __syncthreads();
  if (run && ok) for (int i = t; i < N; i += NT) { double s = sG[i]; for (int j = 0; j < N; ++j) s += sH[i + ldH * j] * sD[j]; sGn[i] = s; }
  __syncthreads();
  if (run && ok) for (int i = t; i < N; i += NT) { sG[i] = sGn[i]; ph.G[gix(0, N, i, ldb, b)] = sGn[i]; }
  __syncthreads();
  dv1 += pdv1; dv2 += pdv2;
}

// multiple-shooting linear rollout of one phase (SinglePhase::linear_rollout), eps = 1. Every knot's operands are staged with
// cp.async into one of two shared-memory buffers while the previous knot is processed; dx ping-pongs between sDx and sDxn.
// part1 / part2 accumulate this thread's share of dV_1 / dV_2 over all knots (reduced once by the caller).
template <int N, int M, bool WB, int NT, class L>
__device__ void lin_phase2(const SolverDev& S, int pi, int b, int t, bool run, double* sm, double& part1, double& part2) {
  const PhaseDev& ph = S.ph[pi];
  const int ldb = S.ldb, h = ph.h;
  constexpr int ldH = L::ldH, ldA = L::ldA, ldM = L::ldM;
  constexpr int KA = WB ? 18 : N, R0 = WB ? 18 : 0;
  const double dt = ph.dt;
  double* dxb[2] = {sm + L::vDx, sm + L::vDxn};
  double* sDuL = sm + L::vDuL;
  const unsigned long long* hm = WB ? nullptr : ph.hkd_mask;
  if (hm != nullptr) {
    // the HKD tiles are staged at their structural non-zeros only and consumed densely: everything else must be zero
    __syncthreads();
    if (run) for (int e = t; e < 2 * L::szLin; e += NT) sm[L::oLin + e] = 0.0;
    __syncthreads();
  }
  auto stage = [&](int k) {
    double* B0 = sm + L::oLin + (k & 1) * L::szLin;
    static_assert(L::oLin % 2 == 0 && L::szLin % 2 == 0 && L::lK % 2 == 0 && L::lA % 2 == 0, "linear-rollout tile layout");
    {
      const double* Kt = ph.Kpm + ((size_t)b * h + k) * (ldM * N);   // the sweep's problem-major copy of the ld x N tile
      for (int e = t; e < ldM * N / 2; e += NT) cp_async16(B0 + L::lK + 2 * e, Kt + 2 * e);
    }
    if constexpr (WB) {
      static_assert(L::lB == L::lA + ldA * N, "[A2 | B2] adjacent");
      const double* ABt = ph.ABpm + ((size_t)b * h + k) * CAFE_WB_AB_TILE;   // [A2 | B2] are adjacent in the stage buffer as well
      for (int e = t; e < CAFE_WB_AB_TILE / 2; e += NT) cp_async16(B0 + L::lA + 2 * e, ABt + 2 * e);
    } else {
      const double* Ag = ph.A + gix(k, N * N, 0, ldb, b);
      const double* Bg = ph.Bm + gix(k, N * M, 0, ldb, b);
      if (hm != nullptr) {
        for (int e = t; e < N * N; e += NT) if (mask_bit(hm, e)) cp_async8(B0 + L::lA + (e % N) + ldA * (e / N), Ag + (size_t)e * ldb);
        for (int e = t; e < N * M; e += NT) if (mask_bit(hm + 9, e)) cp_async8(B0 + L::lB + (e % N) + ldA * (e / N), Bg + (size_t)e * ldb);
      } else {
        for (int e = t; e < KA * N; e += NT) { const int i = e % KA, j = e / KA; cp_async8(B0 + L::lA + i + ldA * j, Ag + (size_t)((R0 + i) + N * j) * ldb); }
        for (int e = t; e < KA * M; e += NT) { const int i = e % KA, j = e / KA; cp_async8(B0 + L::lB + i + ldA * j, Bg + (size_t)((R0 + i) + N * j) * ldb); }
      }
    }
    const double* lxxg = ph.lxx + gix(k, N * N, 0, ldb, b);
    const double* luug = ph.luu + gix(k, M * M, 0, ldb, b);
    if constexpr (WB) {
      // structural non-zeros only (the consumers below skip the other tile entries, which may hold another knot's values)
      const unsigned long long* lmask = ph.lxx_mask + (size_t)k * CAFE_LXX_MASK_WORDS;
      int i = t % N, j = t / N;
      for (int e = t; e < N * N; e += NT) {
        if (mask_bit(lmask, e)) cp_async8(B0 + L::lLxx + i + ldH * j, lxxg + (size_t)e * ldb);
        i += NT % N; j += NT / N;
        if (i >= N) { i -= N; ++j; }
      }
      for (int e = t; e < M; e += NT) cp_async8(B0 + L::lLuu + e + ldM * e, luug + (size_t)(e + M * e) * ldb);
    } else if (hm != nullptr) {
      for (int e = t; e < N * N; e += NT) if (mask_bit(hm + 18, e)) cp_async8(B0 + L::lLxx + (e % N) + ldH * (e / N), lxxg + (size_t)e * ldb);
      for (int e = t; e < M * M; e += NT) if (mask_bit(hm + 27, e)) cp_async8(B0 + L::lLuu + (e % M) + ldM * (e / M), luug + (size_t)e * ldb);
    } else {
      for (int e = t; e < N * N; e += NT) cp_async8(B0 + L::lLxx + (e % N) + ldH * (e / N), lxxg + (size_t)e * ldb);
      for (int e = t; e < M * M; e += NT) cp_async8(B0 + L::lLuu + (e % M) + ldM * (e / M), luug + (size_t)e * ldb);
    }
    for (int j = t; j < N; j += NT) {
      cp_async8(B0 + L::lLx + j, ph.lx + gix(k, N, j, ldb, b));
      cp_async8(B0 + L::lD + j, ph.Defect + gix(k + 1, N, j, ldb, b));
    }
    for (int j = t; j < M; j += NT) {
      cp_async8(B0 + L::lLu + j, ph.lu + gix(k, M, j, ldb, b));
      cp_async8(B0 + L::lDU + j, ph.dU + gix(k, M, j, ldb, b));
    }
    cp_async_commit();
  };
  if (run && h > 0) stage(0);
  // on entry sDx holds the state perturbation handed over by the previous phase (or zero)
  if (run) for (int i = t; i < N; i += NT) { const double v = dxb[0][i] + 1.0 * ph.Defect[gix(0, N, i, ldb, b)]; dxb[0][i] = v; ph.dX[gix(0, N, i, ldb, b)] = v; }
  for (int k = 0; k < h; ++k) {
    const double* B0 = sm + L::oLin + (k & 1) * L::szLin;
    const double* dx = dxb[k & 1];
    double* dxn = dxb[(k + 1) & 1];
    cp_async_wait_all();
    __syncthreads();
    if ((k & (CAFE_PACE - 1)) == 0) cluster_pace();
    if (run) {
      if (k + 1 < h) stage(k + 1);
      // du = dU + K dx (threads 0..M-1);  dV terms of the state (threads of the upper warps)
      if (t < M) {
        double s = 0;
        for (int j = 0; j < N; ++j) s += B0[L::lK + t + ldM * j] * dx[j];
        sDuL[t] = 1.0 * B0[L::lDU + t] + s;
      }
      constexpr int T1 = (NT >= 64 + N) ? 64 : (NT >= 32 + N ? 32 : 0);  // first thread of the lxx rows
      for (int i = t - T1; i >= 0 && i < N; i += NT) {
        double q = 0;
        if constexpr (WB) {
          const unsigned long long* lmask = ph.lxx_mask + (size_t)k * CAFE_LXX_MASK_WORDS;
          for (int j = 0; j < N; ++j) if (mask_bit(lmask, i + N * j)) q += B0[L::lLxx + i + ldH * j] * dx[j];
        } else {
          for (int j = 0; j < N; ++j) q += B0[L::lLxx + i + ldH * j] * dx[j];
        }
        const double dxi = dx[i];
        part1 += B0[L::lLx + i] * dxi;
        part2 += dxi * q;
      }
    }
    __syncthreads();
    if (run) {
      // dx+ = A dx + B du + d
      for (int i = t; i < N; i += NT) {
        double s = 0, s2 = 0;
        if (WB && i < 18) {
          s = dx[i] + dt * dx[18 + i];
        } else {
          const int r = i - R0;
          for (int j = 0; j < N; ++j) s += B0[L::lA + r + ldA * j] * dx[j];
          for (int j = 0; j < M; ++j) s2 += B0[L::lB + r + ldA * j] * sDuL[j];
        }
        const double v = s + s2 + 1.0 * B0[L::lD + i];
        dxn[i] = v;
        ph.dX[gix(k + 1, N, i, ldb, b)] = v;
      }
      constexpr int T2 = (NT >= 64 + M) ? 64 : (NT >= 32 + M ? 32 : 0);
      for (int i = t - T2; i >= 0 && i < M; i += NT) {
        double q = 0;
        if constexpr (WB) q = B0[L::lLuu + i + ldM * i] * sDuL[i];   // luu is diagonal
        else for (int j = 0; j < M; ++j) q += B0[L::lLuu + i + ldM * j] * sDuL[j];
        const double dui = sDuL[i];
        part1 += B0[L::lLu + i] * dui;
        part2 += dui * q;
      }
    }
    // the next iteration's top barrier orders dxn / sDuL / buffer reuse
  }
  __syncthreads();
  const double* dx = dxb[h & 1];
  if (run) {
    for (int i = t; i < N; i += NT) {
      double q = 0;
      for (int j = 0; j < N; ++j) q += ph.Phixx[(size_t)(i + N * j) * ldb + b] * dx[j];
      const double dxi = dx[i];
      part1 += ph.Phix[(size_t)i * ldb + b] * dxi;
      part2 += dxi * q;
    }
  }
  // hand the perturbation over to the next phase in sDx
  double carry = 0;
  const int nn = ph.has_next ? ph.n_next : N;
  if (run && t < nn) {
    if (ph.has_next) { double s = 0; for (int j = 0; j < N; ++j) s += ph.Px[(size_t)(t + nn * j) * ldb + b] * dx[j]; carry = s; }
    else carry = dx[t];
  }
  __syncthreads();
  if (run && t < nn) dxb[0][t] = carry;
  __syncthreads();
}

// DECK: 0 = HKD phases only (24,24,0); 1 = MHPC (WB 36,12,12 + SRB 12,12,0)
template <int DECK, int NT>
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(NT, CAFE_BWD_MINB) k_bwd2(const __grid_constant__ SolverDev S) {
  typedef Bwd2Layout<(DECK == 0 ? 24 : 36), (DECK == 0 ? 24 : 12), (DECK == 0 ? 0 : 12), (DECK == 1)> L;
  constexpr int NX = (DECK == 0 ? 24 : 36);
  extern __shared__ __align__(16) double sm[];
  __shared__ double s_reg;
  __shared__ int s_state, s_regiter;  // 0 sweeping, 1 success, 2 gave up
  // CTA i sweeps the i-th problem that is still iterating (c.act_list, ascending); the tail CTAs of the last cluster idle
  const CtrlDev& c = S.c;
  const bool listed = (int)blockIdx.x < S.n_act;
  const int t = threadIdx.x, b = listed ? c.act_list[blockIdx.x] : 0;
  const CafeOptions& o = S.opt;
  const int ldb = S.ldb;
  // The four CTAs of a cluster own four consecutive problems and advance knot by knot together (cluster_pace); a CTA whose
  // problem is inactive, or whose sweep is already done while a neighbour repeats it with more regularisation, keeps pace only.
  namespace cg = cooperative_groups;
  cg::cluster_group cl = cg::this_cluster();
  const bool mine = listed && c.active[b];
  int it = 0;
  if (t == 0) {
    s_state = mine ? 0 : 3; s_regiter = 0; s_reg = mine ? c.reg[b] : 0.0;
    if (mine) {
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
  // every word of the tiles is finite from here on (the fragment loads of partial edge tiles read neighbouring tiles)
  for (int e = t; e < L::total; e += NT) sm[e] = 0.0;
  double min_piv = 1e300;
  double sdv1 = 0, sdv2 = 0;   // thread 0: expected cost change summed by the sweep (MS = false)
  for (int round = 0;; ++round) {
    cl.sync();
    bool any = false;
    for (unsigned r = 0; r < cl.num_blocks(); ++r) any |= (*cl.map_shared_rank(&s_state, r) == 0);
    cl.sync();
    if (!any) {
      if (round == 0) return;  // the whole cluster is inactive
      break;
    }
    const bool sweeping = s_state == 0;
    bool ok = true;
    const double reg = s_reg;
    if (sweeping) { sdv1 = 0; sdv2 = 0; }   // a sweep that is repeated with more regularisation starts its sums again
    for (int pi = S.n_phases - 1; pi >= 0; --pi) {
      const int model = S.ph[pi].model, nm = S.ph[pi].has_next ? S.ph[pi + 1].model : -1;
      if constexpr (DECK == 0) {
        if (model == CAFE_MODEL_HKD) sweep_phase2<24, 24, 0, 24, false, NT, L>(S, pi, b, t, sweeping, ok, reg, sm, min_piv, sdv1, sdv2);
      } else {
        if (model == CAFE_MODEL_SRB) sweep_phase2<12, 12, 0, 12, false, NT, L>(S, pi, b, t, sweeping, ok, reg, sm, min_piv, sdv1, sdv2);
        else if (model == CAFE_MODEL_WB && nm == CAFE_MODEL_SRB) sweep_phase2<36, 12, 12, 12, true, NT, L>(S, pi, b, t, sweeping, ok, reg, sm, min_piv, sdv1, sdv2);
        else if (model == CAFE_MODEL_WB) sweep_phase2<36, 12, 12, 36, true, NT, L>(S, pi, b, t, sweeping, ok, reg, sm, min_piv, sdv1, sdv2);
      }
      (void)nm;
    }
    __syncthreads();
    if (t == 0 && sweeping) {
      s_regiter += 1;
      if (ok) s_state = 1;
      else {
        const double r = fmax(s_reg * o.update_regularization, 1e-03);
        s_reg = r;
        if (r > 1e2) s_state = 2;
      }
    }
  }
  const bool success = s_state == 1;
  double dV1 = 0, dV2 = 0;
  {
    double* sDx = sm + L::vDx;
    double* sRed = sm + L::oRed;
    for (int i = t; i < NX; i += NT) sDx[i] = 0.0;
    __syncthreads();
    double part1 = 0, part2 = 0;
    // MS = false: no linear rollout (MultiPhaseDDP.cpp:330-333) - dX stays zero, the expected cost change is the sweep's
    for (int pi = 0; pi < S.n_phases && S.opt.MS; ++pi) {
      const int model = S.ph[pi].model;
      if constexpr (DECK == 0) {
        if (model == CAFE_MODEL_HKD) lin_phase2<24, 24, false, NT, L>(S, pi, b, t, success, sm, part1, part2);
      } else {
        if (model == CAFE_MODEL_WB) lin_phase2<36, 12, true, NT, L>(S, pi, b, t, success, sm, part1, part2);
        else if (model == CAFE_MODEL_SRB) lin_phase2<12, 12, false, NT, L>(S, pi, b, t, success, sm, part1, part2);
      }
    }
    // dV_1, dV_2: fixed-order tree over the NT per-thread shares
    sRed[t] = part1;
    sRed[NT + t] = part2;
    __syncthreads();
    for (int w = NT / 2; w > 0; w >>= 1) {
      if (t < w) { sRed[t] += sRed[t + w]; sRed[NT + t] += sRed[NT + t + w]; }
      __syncthreads();
    }
    dV1 = sRed[0]; dV2 = sRed[NT];
    if (!S.opt.MS) { dV1 = sdv1; dV2 = sdv2; }
  }
  if (t == 0 && mine) {
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
