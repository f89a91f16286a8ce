// wb_gen_wrappers.cu — out-of-line instantiations of the generated whole-body routines (gen/wb_gen.h), one copy each,
// with plain array outputs. Separate translation unit (relocatable device code) so that the slow ptxas pass over these
// straight-line functions only re-runs when the generated header changes.
// These per-leg instantiations serve the terminal knots (impact map and its Jacobian) and the single-shooting chains only; the running
// knots use the leg-generic routines of gen/wb_leg_gen.h (wb_leg_kernels.cu). No lock-step markers here.
#define CAFE_GEN_SYNC
#define CAFE_HD __device__ __forceinline__   // device-only instantiations here (the host tests include the header on their own)
#include "gen/wb_gen.h"
#include "wb_pieces.h"

namespace cafe_dev {

// M (lower, ld 18; the caller zero-initialises it), nle, foot Jacobians, Jdot v, foot positions / velocities: trunk + one piece per leg
__device__ __noinline__ void wbg_terms_trunk(const double* q, const double* v, double* nle, double* Mlow) {
  cafe_gen_wb::wb_terms_trunk(q, v, cafe_gen_wb::BiasDst{nle}, cafe_gen_wb::MassDst{Mlow});
}
#define CAFE_TERMS_LEG(F)                                                                                                                        \
  __device__ __noinline__ void wbg_terms_leg##F(const double* q, const double* v, double* nle, double* Mlow, double* J, double* gam, double* pf, \
                                                double* vf) {                                                                                    \
    cafe_gen_wb::wb_terms_leg##F(q, v, cafe_gen_wb::BiasDst{nle}, cafe_gen_wb::MassDst{Mlow}, cafe_gen_wb::PlainDst{J, 1},                       \
                                 cafe_gen_wb::PlainDst{gam, 1}, cafe_gen_wb::PlainDst{pf, 1}, cafe_gen_wb::PlainDst{vf, 1});                     \
  }
CAFE_TERMS_LEG(0) CAFE_TERMS_LEG(1) CAFE_TERMS_LEG(2) CAFE_TERMS_LEG(3)
__device__ void wbg_terms(const double* q, const double* v, double* nle, double* Mlow, double* J, double* gam, double* pf, double* vf) {
  for (int i = 0; i < 6; ++i) nle[i] = 0.0;
  wbg_terms_trunk(q, v, nle, Mlow);
  wbg_terms_leg0(q, v, nle, Mlow, J, gam, pf, vf); wbg_terms_leg1(q, v, nle, Mlow, J, gam, pf, vf);
  wbg_terms_leg2(q, v, nle, Mlow, J, gam, pf, vf); wbg_terms_leg3(q, v, nle, Mlow, J, gam, pf, vf);
}
__device__ __noinline__ void wbg_feet(const double* q, const double* v, double* pf, double* vf, double* J) {
  cafe_gen_wb::wb_feet(q, v, [&](int i, double x) { pf[i] = x; }, [&](int i, double x) { vf[i] = x; }, [&](int i, double x) { J[i] = x; });
}
// st: element stride of the outputs (1 for thread-local arrays, ldb for the batch-major arrays in HBM). One out-of-line function
// per piece: ptxas allocates registers per function, and a piece's live set (base + one leg) fits where the monolithic routine spilled.
__device__ __noinline__ void wbg_rnea_trunk(const double* q, const double* v, const double* a, double* dq, double* dv, size_t st) {
  cafe_gen_wb::wb_rnea_derivs_trunk(q, v, a, cafe_gen_wb::RneaDst<0>{dq, st}, cafe_gen_wb::RneaDst<0>{dv, st});
}
#define CAFE_RNEA_LEG(F)                                                                                                              \
  __device__ __noinline__ void wbg_rnea_leg##F(const double* q, const double* v, const double* a, double* dq, double* dv, size_t st) { \
    cafe_gen_wb::wb_rnea_derivs_leg##F(q, v, a, cafe_gen_wb::RneaDst<F + 1>{dq, st}, cafe_gen_wb::RneaDst<F + 1>{dv, st});             \
  }
CAFE_RNEA_LEG(0) CAFE_RNEA_LEG(1) CAFE_RNEA_LEG(2) CAFE_RNEA_LEG(3)
__device__ void wbg_rnea_derivs(const double* q, const double* v, const double* a, double* dq, double* dv, size_t st) {
  wbg_rnea_trunk(q, v, a, dq, dv, st);
  wbg_rnea_leg0(q, v, a, dq, dv, st); wbg_rnea_leg1(q, v, a, dq, dv, st); wbg_rnea_leg2(q, v, a, dq, dv, st); wbg_rnea_leg3(q, v, a, dq, dv, st);
}
__device__ __noinline__ void wbg_grav_derivs(const double* q, double* dq) {
  cafe_gen_wb::wb_grav_derivs(q, [&](int i, double x) { dq[i] = x; });
}
#define CAFE_KIN_FOOT(F)                                                                                                                   \
  __device__ __noinline__ void wbg_kin_foot##F(const double* q, const double* v, const double* a, const double* Fc, double* dvq, double* daq, \
                                               double* dav, double* djtf, size_t st) {                                                      \
    cafe_gen_wb::wb_kin_partials_foot##F(q, v, a, Fc, cafe_gen_wb::PlainDst{dvq, st}, cafe_gen_wb::PlainDst{daq, st},                        \
                                         cafe_gen_wb::PlainDst{dav, st}, cafe_gen_wb::JtfDst<F>{djtf, st});                                 \
  }
CAFE_KIN_FOOT(0) CAFE_KIN_FOOT(1) CAFE_KIN_FOOT(2) CAFE_KIN_FOOT(3)
__device__ void wbg_kin_partials(const double* q, const double* v, const double* a, const double* F, double* dvq, double* daq, double* dav, double* djtf, size_t st) {
  wbg_kin_foot0(q, v, a, F, dvq, daq, dav, djtf, st); wbg_kin_foot1(q, v, a, F, dvq, daq, dav, djtf, st);
  wbg_kin_foot2(q, v, a, F, dvq, daq, dav, djtf, st); wbg_kin_foot3(q, v, a, F, dvq, daq, dav, djtf, st);
}
__device__ __noinline__ void wbg_footvel_partial(const double* q, const double* v, double* dvq) {
  cafe_gen_wb::wb_footvel_partial(q, v, [&](int i, double x) { dvq[i] = x; });
}

}  // namespace cafe_dev
