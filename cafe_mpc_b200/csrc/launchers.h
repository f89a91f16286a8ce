// launchers.h — host-callable launch wrappers of the kernels that live in knot_kernels.cu
#pragma once
#include <cuda_runtime.h>
#include "device_types.cuh"

namespace cafe_dev {
// list / n_list: index list of the problems to process (CtrlDev::act_list or pend_list); n_knots: knots per problem
void launch_roll(const SolverDev* dS, int n_knots, cudaStream_t st, int a0, int a1, const int* list, int n_list);
void launch_lq(const SolverDev* dS, int n_knots, cudaStream_t st, const int* list, int n_list);
void launch_compact(const SolverDev* dS, cudaStream_t st, int mode);
void launch_accept(const SolverDev* dS, long long nthreads, cudaStream_t st);
void launch_ls_scan(const SolverDev* dS, int B, cudaStream_t st, int a0, int a1);
void launch_select(const SolverDev* dS, int B, cudaStream_t st, int mode);
void launch_reb_update(const SolverDev* dS, long long nthreads, cudaStream_t st);
// whole-body running knots (n_wbk of them per problem): leg-parallel rigid-body routines (wb_leg_kernels.cu) and the cooperative
// shared-memory kernels that consume them (wb_coop.cu)
void launch_wb_terms(const SolverDev* dS, int n_wbk, cudaStream_t st, int a0, int a1, const int* list, int n_list);
void launch_wb_fwd(const SolverDev* dS, int n_wbk, cudaStream_t st, int a0, int a1, const int* list, int n_list);
void launch_wb_derivs(const SolverDev* dS, int n_wbk, cudaStream_t st, const int* list, int n_list);
void launch_wb_sens(const SolverDev* dS, int n_wbk, cudaStream_t st, const int* list, int n_list);
void launch_wb_cost(const SolverDev* dS, int n_wbk, cudaStream_t st, const int* list, int n_list);
size_t knot_kernels_local_bytes();
int wb_coop_configure();   // opt-in shared-memory sizes of the cooperative kernels (0 = ok)
}  // namespace cafe_dev
