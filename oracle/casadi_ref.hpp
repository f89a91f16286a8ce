/*
 * casadi_ref.hpp — ORACLE side only. Declares the reference's CasADi-generated C functions
 * (compiled UNCHANGED from /root/reference into oracle/_ref/libcafe_ref_casadi.so by
 * oracle/Makefile) and a caller that restates common/casadi_interface.cpp:5-81
 * (evaluate, then scatter the CCS non-zeros into dense column-major outputs that the
 * caller has zeroed).
 */
#pragma once
#include <vector>
#include <cstring>

typedef long long int casadi_int;
typedef int (*casadi_fn_t)(const double**, double**, casadi_int*, double*, int);
typedef const casadi_int* (*casadi_sp_t)(casadi_int);
typedef int (*casadi_work_t)(casadi_int*, casadi_int*, casadi_int*, casadi_int*);

#define CASADI_DECL(name)                                                         \
  int name(const double** arg, double** res, casadi_int* iw, double* w, int mem); \
  const casadi_int* name##_sparsity_out(casadi_int i);                            \
  const casadi_int* name##_sparsity_in(casadi_int i);                             \
  int name##_work(casadi_int* sz_arg, casadi_int* sz_res, casadi_int* sz_iw, casadi_int* sz_w);

extern "C" {
CASADI_DECL(hkinodyn)
CASADI_DECL(hkinodyn_par)
CASADI_DECL(compute_foot_position)
CASADI_DECL(comp_foot_jacob_1)
CASADI_DECL(comp_foot_jacob_2)
CASADI_DECL(comp_foot_jacob_3)
CASADI_DECL(comp_foot_jacob_4)
CASADI_DECL(SRBDynamics)
CASADI_DECL(SRBDynamicsDerivatives)
CASADI_DECL(footVelPartialDq)
CASADI_DECL(footAccPartialDq)
CASADI_DECL(footAccPartialDv)
CASADI_DECL(footForcePartialDq)
}

namespace oracle {

struct CasadiFn {
  casadi_fn_t f;
  casadi_sp_t sp_out, sp_in;
  casadi_work_t work;
};
#define CASADI_FN(name) oracle::CasadiFn{name, name##_sparsity_out, name##_sparsity_in, name##_work}

inline int ccs_nnz(const casadi_int* sp) { return (int)sp[2 + sp[1]]; }

/* RES[o] must be zero-initialised dense nrow*ncol buffers. */
inline void casadi_call(const CasadiFn& fn, const double* const* ARG, int n_arg, double* const* RES, int n_res) {
  casadi_int sz_arg = 0, sz_res = 0, sz_iw = 0, sz_w = 0;
  fn.work(&sz_arg, &sz_res, &sz_iw, &sz_w);
  std::vector<const double*> arg((size_t)std::max<casadi_int>(sz_arg, n_arg), nullptr);
  std::vector<double*> res((size_t)std::max<casadi_int>(sz_res, n_res), nullptr);
  std::vector<std::vector<double>> nz(n_res);
  for (int i = 0; i < n_arg; ++i) arg[i] = ARG[i];
  for (int o = 0; o < n_res; ++o) { nz[o].assign((size_t)ccs_nnz(fn.sp_out(o)) + 1, 0.0); res[o] = nz[o].data(); }
  std::vector<casadi_int> iw((size_t)sz_iw + 1);
  std::vector<double> w((size_t)sz_w + 1);
  fn.f(arg.data(), res.data(), iw.data(), w.data(), 0);
  for (int o = 0; o < n_res; ++o) {
    const casadi_int* sp = fn.sp_out(o);
    casadi_int nrow = sp[0], ncol = sp[1];
    const casadi_int* colind = sp + 2;
    const casadi_int* row = colind + ncol + 1;
    for (casadi_int c = 0; c < ncol; ++c)
      for (casadi_int k = colind[c]; k < colind[c + 1]; ++k) RES[o][row[k] + nrow * c] = nz[o][k];
  }
}

}  // namespace oracle
