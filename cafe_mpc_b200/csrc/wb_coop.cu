// wb_coop.cu — translation unit of the warp-cooperative whole-body kernels (wb_coop.cuh) and their launchers.
#include "wb_coop.cuh"
#include "launchers.h"

namespace cafe_dev {

int wb_coop_configure() {
  cudaError_t e = cudaFuncSetAttribute(k_wb_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(CAFE_WB_PKS * WbSm::totalFwd * sizeof(double)));
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_wb_sens, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(CAFE_WB_PKS * WbSm::totalSens * sizeof(double)));
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k_wb_cost, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(CAFE_WB_PKS * WbSm::totalCost * sizeof(double)));
  return (int)e;
}
void launch_wb_fwd(const SolverDev* dS, int n_wbk, cudaStream_t st, int a0, int a1, const int* list, int n_list) {
  if (n_list <= 0 || n_wbk <= 0 || a1 <= a0) return;
  const dim3 grid((n_list + CAFE_WB_PKS - 1) / CAFE_WB_PKS, n_wbk, a1 - a0);
  k_wb_fwd<<<grid, 32 * CAFE_WB_PKS, CAFE_WB_PKS * WbSm::totalFwd * sizeof(double), st>>>(dS, a0, list, n_list);
}
void launch_wb_sens(const SolverDev* dS, int n_wbk, cudaStream_t st, const int* list, int n_list) {
  if (n_list <= 0 || n_wbk <= 0) return;
  const dim3 grid((n_list + CAFE_WB_PKS - 1) / CAFE_WB_PKS, n_wbk);
  k_wb_sens<<<grid, 32 * CAFE_WB_PKS, CAFE_WB_PKS * WbSm::totalSens * sizeof(double), st>>>(dS, list, n_list);
}
void launch_wb_cost(const SolverDev* dS, int n_wbk, cudaStream_t st, const int* list, int n_list) {
  if (n_list <= 0 || n_wbk <= 0) return;
  const dim3 grid((n_list + CAFE_WB_PKS - 1) / CAFE_WB_PKS, n_wbk);
  k_wb_cost<<<grid, 32 * CAFE_WB_PKS, CAFE_WB_PKS * WbSm::totalCost * sizeof(double), st>>>(dS, list, n_list);
}

}  // namespace cafe_dev
