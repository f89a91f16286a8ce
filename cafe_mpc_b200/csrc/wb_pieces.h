// wb_pieces.h — how the per-inertia-group / per-foot pieces of the generated whole-body derivatives (gen/wb_gen.h) add up.
//
// RNEA is linear in the link inertias, so dtau/d(q,v) = [trunk] + sum_f [leg f]; the contact-force term is a sum over feet,
// d(J^T F)/dq = sum_f d(J_f^T F_f)/dq. Piece f only involves the base and leg f: its non-zeros are rows/columns {base, leg f}, and
// the pieces overlap only in a small base block, which the destination functors below accumulate (first writer stores, the
// others add), everything else is a plain store of a static non-zero pattern. The index an output functor receives is a literal
// in the generated code, so the case distinction folds away at compile time.
#pragma once
#include <cstddef>
#ifndef CAFE_HD
#ifdef __CUDACC__
#define CAFE_HD __host__ __device__ __forceinline__
#else
#define CAFE_HD inline
#endif
#endif

namespace cafe_gen_wb {

// share of an 18x18 RNEA derivative matrix (column-major, element stride st). PIECE 0 = trunk (rows 3..5 x columns 3..5 only),
// PIECE 1..4 = legs; shared block = rows 0..5 x columns 3..5 (trunk stores rows 3..5 first, leg 0 stores rows 0..2 first)
template <int PIECE>
struct RneaDst {
  double* p; size_t st;
  CAFE_HD void operator()(int idx, double x) const {
    const int r = idx % 18, c = idx / 18;
    const bool shared = r < 6 && c >= 3 && c <= 5;
    const bool first = PIECE == 0 || (PIECE == 1 && r < 3);
    if (shared && !first) p[idx * st] += x; else p[idx * st] = x;
  }
};

// share of d(J^T F)/dq of foot FOOT; shared block = rows 3..5 x columns 3..5 (foot 0 stores first)
template <int FOOT>
struct JtfDst {
  double* p; size_t st;
  CAFE_HD void operator()(int idx, double x) const {
    const int r = idx % 18, c = idx / 18;
    const bool shared = r >= 3 && r <= 5 && c >= 3 && c <= 5;
    if (shared && FOOT != 0) p[idx * st] += x; else p[idx * st] = x;
  }
};

// shares of the mass matrix (lower triangle, ld 18) and of the bias vector: rows/columns 0..5 are common to all pieces and are
// accumulated onto a zero-initialised destination, the rest is private to one leg
struct MassDst {
  double* p;
  CAFE_HD void operator()(int idx, double x) const { if (idx % 18 < 6 && idx / 18 < 6) p[idx] += x; else p[idx] = x; }
};
struct BiasDst {
  double* p;
  CAFE_HD void operator()(int idx, double x) const { if (idx < 6) p[idx] += x; else p[idx] = x; }
};

struct PlainDst {
  double* p; size_t st;
  CAFE_HD void operator()(int idx, double x) const { p[idx * st] = x; }
};

}  // namespace cafe_gen_wb
