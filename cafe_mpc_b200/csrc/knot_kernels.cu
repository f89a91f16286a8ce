// knot_kernels.cu — translation unit of the per-(problem, knot) kernels (k_roll, k_lq, k_accept) and the control kernels
// (k_ls_scan, k_select). Compiled with -maxrregcount=128 (see build.py): these kernels are dominated by thread-local
// memory traffic of the generated model routines and need 4 resident CTAs of 128 threads per SM for latency hiding.
#include "knot_kernels.cuh"
#include "launchers.h"

namespace cafe_dev {

void launch_roll(const SolverDev* dS, int n_knots, cudaStream_t st, int a0, int a1, const int* list, int n_list) {
  if (n_list <= 0) return;
  const long long nthreads = (long long)((n_list + 31) & ~31) * n_knots * (a1 - a0);
  k_roll<<<(unsigned)((nthreads + 127) / 128), 128, 0, st>>>(dS, a0, a1, list, n_list);
}
void launch_lq(const SolverDev* dS, int n_knots, cudaStream_t st, const int* list, int n_list) {
  if (n_list <= 0) return;
  const long long nthreads = (long long)((n_list + 127) & ~127) * n_knots;
  k_lq<<<(unsigned)((nthreads + 127) / 128), 128, 0, st>>>(dS, list, n_list);
}
// largest per-thread local-memory frame of the thread-per-knot kernels (terminal whole-body knots keep arrays there)
size_t knot_kernels_local_bytes() {
  size_t m = 0;
  cudaFuncAttributes a;
  if (cudaFuncGetAttributes(&a, k_lq) == cudaSuccess && a.localSizeBytes > m) m = a.localSizeBytes;
  if (cudaFuncGetAttributes(&a, k_roll) == cudaSuccess && a.localSizeBytes > m) m = a.localSizeBytes;
  return m;
}
void launch_compact(const SolverDev* dS, cudaStream_t st, int mode) { k_compact<<<1, 1024, 0, st>>>(dS, mode); }
void launch_accept(const SolverDev* dS, long long nthreads, cudaStream_t st) { k_accept<<<(unsigned)((nthreads + 127) / 128), 128, 0, st>>>(dS); }
void launch_ls_scan(const SolverDev* dS, int B, cudaStream_t st, int a0, int a1) { k_ls_scan<<<(B + 127) / 128, 128, 0, st>>>(dS, a0, a1); }
void launch_select(const SolverDev* dS, int B, cudaStream_t st, int mode) { k_select<<<(B + 127) / 128, 128, 0, st>>>(dS, mode); }
void launch_reb_update(const SolverDev* dS, long long nthreads, cudaStream_t st) { k_reb_update<<<(unsigned)((nthreads + 127) / 128), 128, 0, st>>>(dS); }

}  // namespace cafe_dev
