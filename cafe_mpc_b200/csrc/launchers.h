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
}  // namespace cafe_dev
