// Micro-benchmarks that decide the design of the backward-sweep GEMMs (k_bwd2):
//   1. DFMA vs DMMA (mma.sync m8n8k4 / m16n8k8 / m16n8k16 .f64) throughput, register operands only
//   2. shared-memory cost of the operand patterns: LDS.64 / LDS.128 with 32, 8 and 4 distinct addresses per warp
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_ubench fp64_ubench.cu ; run: ./fp64_ubench
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("cuda error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void k_dfma(double* out, int iters) {
  double a[8];
  const double x = 1.0 + 1e-9 * threadIdx.x, y = 1e-9;
  for (int i = 0; i < 8; ++i) a[i] = i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], x, y);
  double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma1684(double (&c)[4], const double (&a)[2], double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int NACC>
__global__ void k_dmma884(double* out, int iters) {
  double c[NACC][2];
  const double a = 1e-3 * threadIdx.x, b = 1e-3;
  for (int i = 0; i < NACC; ++i) { c[i][0] = i; c[i][1] = -i; }
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma884(c[i][0], c[i][1], a, b);
  double s = 0; for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_dmma1684(double* out, int iters) {
  double c[NACC][4]; double a[2] = {1e-3 * threadIdx.x, 2e-3}; const double b = 1e-3;
  for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) c[i][j] = i + j;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma1684(c[i], a, b);
  double s = 0; for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_dmma1688(double* out, int iters) {
  double c[NACC][4]; double a[4] = {1e-3 * threadIdx.x, 2e-3, 3e-3, 4e-3}; double b[2] = {1e-3, 2e-3};
  for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) c[i][j] = i + j;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma1688(c[i], a, b);
  double s = 0; for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_dmma16816(double* out, int iters) {
  double c[NACC][4]; double a[8]; double b[4];
  for (int j = 0; j < 8; ++j) a[j] = 1e-3 * (threadIdx.x + j);
  for (int j = 0; j < 4; ++j) b[j] = 1e-3 * j;
  for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) c[i][j] = i + j;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma16816(c[i], a, b);
  double s = 0; for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// shared-memory patterns. MODE 0: 32 distinct doubles (lane), 1: 8 distinct (lane%8)*5, 2: 4 distinct (lane/8)*3*37, 3: 1 address
//   W = 8 -> LDS.64, W = 16 -> LDS.128 (pairs of doubles; distinct pairs per the same modes)
template <int MODE, int W>
__global__ void k_lds(double* out, int iters) {
  extern __shared__ double sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 1e-6;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  int idx = MODE == 0 ? lane : MODE == 1 ? (lane % 8) * 5 : MODE == 2 ? (lane / 8) * 111 : 0;
  if (W == 16) idx *= 2;
  idx += (threadIdx.x / 32) * 8;   // warps start on different rows (still 64-byte aligned)
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int o = (idx + u * 64 + (it & 7) * 512) & 4095;   // unknown to the compiler beyond alignment
      if (W == 8) {
        double v;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((unsigned)__cvta_generic_to_shared(sm + o)));
        if (u & 1) s1 += v; else s0 += v;
      } else {
        double2 v;
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"((unsigned)__cvta_generic_to_shared(sm + (o & ~1))));
        if (u & 1) { s1 += v.x; s3 += v.y; } else { s0 += v.x; s2 += v.y; }
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + s2 + s3;
}

template <class F>
float time_it(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int nsm = p.multiProcessorCount;
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz_nominal\": %d}\n", p.name, nsm, clk_khz);
  double* out; CK(cudaMalloc(&out, sizeof(double) * nsm * 8 * 1024));
  const int iters = 20000;
  for (int wps = 4; wps <= 16; wps *= 2) {   // warps per SM
    const int threads = 32 * wps;
    {
      float ms = time_it([&] { k_dfma<<<nsm, threads>>>(out, iters); });
      printf("{\"test\": \"dfma\", \"warps_per_sm\": %d, \"tflops\": %.2f}\n", wps, 2.0 * 8 * iters * threads * nsm / ms / 1e9);
    }
#define RUN_MMA(NAME, KERN, FMAS_PER_WARP_INSTR, NACC)                                                                      \
    { float ms = time_it([&] { KERN<NACC><<<nsm, threads>>>(out, iters); });                                                 \
      printf("{\"test\": \"%s\", \"acc\": %d, \"warps_per_sm\": %d, \"tflops\": %.2f}\n", NAME, NACC, wps,                 \
             2.0 * FMAS_PER_WARP_INSTR * NACC * (double)iters * wps * nsm / ms / 1e9); }
    RUN_MMA("dmma.m8n8k4", k_dmma884, 256, 4)
    RUN_MMA("dmma.m8n8k4", k_dmma884, 256, 8)
    RUN_MMA("dmma.m16n8k4", k_dmma1684, 512, 4)
    RUN_MMA("dmma.m16n8k8", k_dmma1688, 1024, 4)
    RUN_MMA("dmma.m16n8k16", k_dmma16816, 2048, 4)
  }
  CK(cudaGetLastError());
  const int li = 4000;
#define RUN_LDS(MODE, W)                                                                                                     \
  { CK(cudaFuncSetAttribute(k_lds<MODE, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 8));                          \
    float ms = time_it([&] { k_lds<MODE, W><<<nsm, 512, 4096 * 8>>>(out, li); });                                             \
    const double n_lds_per_sm = 16.0 * 8 * li;                                                                               \
    printf("{\"test\": \"lds\", \"bytes\": %d, \"mode\": %d, \"ns_per_warp_lds_per_sm\": %.3f}\n", W, MODE, ms * 1e6 / n_lds_per_sm); }
  RUN_LDS(0, 8) RUN_LDS(1, 8) RUN_LDS(2, 8) RUN_LDS(3, 8)
  RUN_LDS(0, 16) RUN_LDS(1, 16) RUN_LDS(2, 16) RUN_LDS(3, 16)
  CK(cudaDeviceSynchronize());
  return 0;
}
