// wb_gen_wrappers.cu — out-of-line instantiations of the generated whole-body routines (gen/wb_gen.h), one copy each,
// with plain array outputs. Separate translation unit (relocatable device code) so that the slow ptxas pass over these
// straight-line functions only re-runs when the generated header changes.
#include "gen/wb_gen.h"

namespace cafe_dev {

__device__ __noinline__ void wbg_terms(const double* q, const double* v, double* nle, double* Mlow, double* J, double* gam, double* pf, double* vf) {
  cafe_gen_wb::wb_terms(q, v, [&](int i, double x) { nle[i] = x; }, [&](int i, double x) { Mlow[i] = x; }, [&](int i, double x) { J[i] = x; },
                        [&](int i, double x) { gam[i] = x; }, [&](int i, double x) { pf[i] = x; }, [&](int i, double x) { vf[i] = x; });
}
__device__ __noinline__ void wbg_feet(const double* q, const double* v, double* pf, double* vf, double* J) {
  cafe_gen_wb::wb_feet(q, v, [&](int i, double x) { pf[i] = x; }, [&](int i, double x) { vf[i] = x; }, [&](int i, double x) { J[i] = x; });
}
// st: element stride of the outputs (1 for thread-local arrays, ldb for the batch-major arrays in HBM)
__device__ __noinline__ void wbg_rnea_derivs(const double* q, const double* v, const double* a, double* dq, double* dv, size_t st) {
  cafe_gen_wb::wb_rnea_derivs(q, v, a, [&](int i, double x) { dq[i * st] = x; }, [&](int i, double x) { dv[i * st] = x; });
}
__device__ __noinline__ void wbg_grav_derivs(const double* q, double* dq) {
  cafe_gen_wb::wb_grav_derivs(q, [&](int i, double x) { dq[i] = x; });
}
__device__ __noinline__ void wbg_kin_partials(const double* q, const double* v, const double* a, const double* F, double* dvq, double* daq, double* dav, double* djtf, size_t st) {
  cafe_gen_wb::wb_kin_partials(q, v, a, F, [&](int i, double x) { dvq[i * st] = x; }, [&](int i, double x) { daq[i * st] = x; }, [&](int i, double x) { dav[i * st] = x; },
                               [&](int i, double x) { djtf[i * st] = x; });
}
__device__ __noinline__ void wbg_footvel_partial(const double* q, const double* v, double* dvq) {
  cafe_gen_wb::wb_footvel_partial(q, v, [&](int i, double x) { dvq[i] = x; });
}

}  // namespace cafe_dev
