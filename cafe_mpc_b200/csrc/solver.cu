// solver.cu — GPU part of the C ABI (include/cafe_gpu.h): device arena, tick loop, result packing.
//
// One "tick" advances EVERY still-active problem of the batch by one DDP iteration
// (MultiPhaseDDP::solve inner-loop body, HSDDPSolver/source/MultiPhaseDDP.cpp:277-386):
//     k_lq -> k_bwd (sweep + linear rollout + merit parameter) -> k_roll (all step sizes at once)
//          -> k_select (Armijo over step sizes, exits, AL update) -> k_accept
// Per-problem control state lives on the device; the host only polls one counter per tick.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/cafe_gpu.h"
#include "bwd2.cuh"
#include "launchers.h"
#include "host/problem_builders.h"

namespace cafe { void set_last_error(const std::string& s); }

#define CUDA_OK(call)                                                                        \
  do {                                                                                       \
    cudaError_t e_ = (call);                                                                 \
    if (e_ != cudaSuccess) {                                                                 \
      cafe::set_last_error(std::string(#call) + ": " + cudaGetErrorString(e_));              \
      return CAFE_ERR_CUDA;                                                                  \
    }                                                                                        \
  } while (0)

// one array of the packed records. pm_ld > 0: the source is a problem-major tile array [b][pm_h][pm_ld x (nc / pm_rows)] (the sweep's
// Quu / Qux outputs), element e of a knot = row e % pm_rows, column e / pm_rows; else batch-major [knot][nc][ldb]
struct PackSeg { const double* src; int knots, nc; long dst; int pm_ld, pm_rows, pm_h; };

struct CafeHandle {
  int device = 0, max_batch = 0, ldb = 0, B = 0, NA = 0;
  CafeDeck deck;
  std::vector<double> ref_host;
  SolverDev S;            // host copy (device pointers inside)
  SolverDev* dS = nullptr;
  char* arena = nullptr;  // everything per-problem
  size_t arena_bytes = 0, zero_bytes = 0;  // [0, zero_bytes) is re-zeroed at every solve
  double* d_ref = nullptr;
  double* d_ref_pp = nullptr; int ref_pp_B = 0;
  unsigned long long* d_lxx_mask = nullptr;
  unsigned long long* d_hkd_mask = nullptr;   // structural A / B / lxx / luu patterns of the HKD model   // structural lxx patterns of the whole-body knots
  double* d_guess = nullptr; size_t guess_bytes = 0; int guess_B = 0;  // packed initial guesses [B][solution_size] (warm start)  // per-problem reference records [n_records][CAFE_REF_W][ldb]
  double* d_x0raw = nullptr;
  // augmented-Lagrangian parameters the next solve starts from instead of the deck's (the MPC loop's carry-over: the reference never resets
  // them, ConstraintsBase.h:367-374): [(phase * 4 + element) * 2 + {sigma, lambda}][ldb]; al_B = 0: deck values
  double* d_al = nullptr; int al_B = 0;
  // relaxed-barrier update counts carried with the knots over an MPC update (PathConstraintBase::pop_front / push_back, ConstraintsBase.h:296-306;
  // reset_params is empty): laid out for the CURRENT deck, phase i at reb_carry_off[i], [h][reb_ne][ldb] like PhaseDev::reb_n; reb_B = 0: none
  unsigned char* d_reb_carry = nullptr; size_t reb_carry_bytes = 0, reb_carry_off[CAFE_MAX_PHASES] = {0}; int reb_B = 0;
  int* d_fail = nullptr; size_t fail_bytes = 0;
  int* h_nactive = nullptr;  // pinned, mapped: the list lengths are stored into it by the device (k_publish_int) - no copy engine on the tick path,
  int* d_nactive_map = nullptr;   // so a bulk D2H of the previous solve's records (asynchronous collection) cannot delay a tick; its device address
  double* d_pack = nullptr; size_t pack_bytes = 0;
  // asynchronous collection (cafe_gpu_get_commands_async / cafe_gpu_gather_commands_async): two slots, so that the records of solve i travel to
  // the host while solve i + 1 runs. Per slot: a pack buffer, "packed" (solver stream) and "landed" (copy stream) events
  cudaStream_t stream_copy = nullptr;
  double* d_pack_async[2] = {nullptr, nullptr}; size_t pack_async_bytes[2] = {0, 0};
  cudaEvent_t ev_packed[2] = {nullptr, nullptr}, ev_landed[2] = {nullptr, nullptr};
  bool async_used[2] = {false, false};
  PackSeg* d_segs = nullptr; int max_segs = 0;
  // small host tables for the pack / unpack / shift kernels: one growable device buffer + one pinned staging buffer per handle, copies
  // stream-ordered on `stream` (no allocation, no host synchronisation in the MPC loop)
  char* d_tab = nullptr; char* h_tab = nullptr; size_t tab_cap = 0; cudaEvent_t ev_tab = nullptr; bool tab_busy = false;
  size_t ref_cap = 0, lxx_mask_cap = 0, ref_pp_cap = 0;   // capacities (elements) of d_ref / d_lxx_mask: cafe_gpu_update_deck re-uses them
  cudaStream_t stream = nullptr;
  // second stream of a tick: the active list is cut in two and the LQ -> dense -> sweep -> first rollout chains of the halves run
  // on two streams, so that one half's kernels fill the wave tails of the other's (per-problem results do not depend on it)
  cudaStream_t stream2[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
  // linearisation of one part on two streams (k_lq | k_wb_derivs -> k_wb_sens, k_wb_cost): fork / derivatives-ready / join events per part
  cudaEvent_t ev_lqf[2] = {nullptr, nullptr}, ev_lqd[2] = {nullptr, nullptr}, ev_lqj[2] = {nullptr, nullptr};
  cudaEvent_t ev_lqc = nullptr;   // unsplit ticks: k_wb_cost on a third stream (it waits for the derivatives only, not for k_lq)
  bool lq_overlap = true;   // CAFE_LQ_OVERLAP=0 keeps the linearisation on one stream
  int split_n = 2;       // number of parts (CAFE_SPLIT_N, 2..4)
  int split_min = 1024;  // smallest active list that is cut (CAFE_SPLIT_MIN; 0 = never)
  bool profiling = false;
  double ms[CAFE_NKERNELS] = {0};
  long launches[CAFE_NKERNELS] = {0};
  double units[CAFE_NKERNELS] = {0};   // work items launched per slot: (problem, knot[, step size]) for the knot kernels, problems for the sweep
  int ticks = 0;
  int bwd_variant = 0;  // 0: HKD deck (24,24,0)   1: MHPC deck (36,24->12,12)
  size_t bwd_smem = 0;
  int bwd_nt = 128;     // threads per problem of the MHPC sweep (dev switch CAFE_BWD_NT)
  int bwd_small = 296;  // lists up to this length (one wave of two 256-thread CTAs per SM) are swept with 256 threads per problem: with the
                        // machine under-filled the latency of one problem's sweep is what counts (CAFE_BWD_SMALL; 0 = never). Same bits either way.
  int bwd_pb = 4;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evs = nullptr, eve = nullptr;
  double total_ms = 0;
};

namespace {

using namespace cafe_dev;

__global__ void k_init(const SolverDev* __restrict__ Sp, const double* __restrict__ x0raw, int n0) {
  const SolverDev& S = *Sp;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = (int)(t % S.ldb);
  const int gk = (int)(t / S.ldb);
  if (gk >= S.n_knots || b >= S.B) return;
  const int pi = S.knot_phase[gk], k = S.knot_k[gk];
  const PhaseDev& ph = S.ph[pi];
  double rec_local[CAFE_REF_W];
  const double* rec = knot_record(ph, k, S.ldb, b, rec_local);
  for (int i = 0; i < ph.n; ++i) { const double v = rec[CAFE_REF_XR + i]; ph.Xbar[gix(k, ph.n, i, S.ldb, b)] = v; ph.X[gix(k, ph.n, i, S.ldb, b)] = v; }
  if (k == 0) for (int i = 0; i < ph.n_td; ++i) { ph.al_sigma[(size_t)i * S.ldb + b] = ph.al_td.sigma; ph.al_lambda[(size_t)i * S.ldb + b] = ph.al_td.lambda; }
  if (ph.model == CAFE_MODEL_WB && k < ph.h) {   // rows 0..17 of A = [I, dt I] never change (rows 18..35 live in the problem-major tiles)
    double* Ag = ph.A + gix(k, 1296, 0, S.ldb, b);
    for (int i = 0; i < 18; ++i) { Ag[(size_t)(i + 36 * i) * S.ldb] = 1.0; Ag[(size_t)(i + 36 * (18 + i)) * S.ldb] = ph.dt; }
  }
  if (gk == 0) {
    double* x0 = const_cast<double*>(S.x0);
    for (int i = 0; i < n0; ++i) x0[(size_t)i * S.ldb + b] = x0raw[(size_t)b * n0 + i];
    S.c.active[b] = 1; S.c.do_ls[b] = 1; S.c.sel[b] = -1; S.c.min_pivot[b] = 1e300;
  }
}

// carried augmented-Lagrangian parameters replace the deck's initial values (after k_init)
__global__ void k_apply_al(const SolverDev* __restrict__ Sp, const double* __restrict__ al) {
  const SolverDev& S = *Sp;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= S.B) return;
  for (int pi = 0; pi < S.n_phases; ++pi) {
    const PhaseDev& ph = S.ph[pi];
    for (int i = 0; i < ph.n_td; ++i) {
      ph.al_sigma[(size_t)i * S.ldb + b] = al[(size_t)((pi * 4 + i) * 2) * S.ldb + b];
      ph.al_lambda[(size_t)i * S.ldb + b] = al[(size_t)((pi * 4 + i) * 2 + 1) * S.ldb + b];
    }
  }
}
// the MPC update's carry-over: new phase pi continues old phase src[pi] (or -1: its constraint starts from the deck's values init[pi])
struct AlCarry { int n_phases, src[CAFE_MAX_PHASES], n_td[CAFE_MAX_PHASES]; double sigma0[CAFE_MAX_PHASES], lambda0[CAFE_MAX_PHASES]; };
__global__ void k_carry_al(const SolverDev* __restrict__ Sp, const __grid_constant__ AlCarry c, int B, int ldb_dst, double* __restrict__ al) {
  const SolverDev& S = *Sp;   // the OLD layout
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  for (int pi = 0; pi < c.n_phases; ++pi)
    for (int i = 0; i < 4; ++i) {
      double sg = 0, lm = 0;
      if (i < c.n_td[pi]) {
        if (c.src[pi] >= 0) { const PhaseDev& ph = S.ph[c.src[pi]]; sg = ph.al_sigma[(size_t)i * S.ldb + b]; lm = ph.al_lambda[(size_t)i * S.ldb + b]; }
        else { sg = c.sigma0[pi]; lm = c.lambda0[pi]; }
      }
      al[(size_t)((pi * 4 + i) * 2) * ldb_dst + b] = sg;
      al[(size_t)((pi * 4 + i) * 2 + 1) * ldb_dst + b] = lm;
    }
}

// one word from device memory to mapped page-locked host memory (a posted store over PCIe): how the host learns a list length
__global__ void k_publish_int(const int* __restrict__ src, int* __restrict__ dst_mapped) { *dst_mapped = *src; __threadfence_system(); }

// relaxed-barrier update counts of one knot of the new deck: copied from a knot of the previous plan, or zero (a constraint that did not exist)
struct RebCarryEntry { const unsigned char* src; size_t dst; int ne; };
__global__ void k_carry_reb(const RebCarryEntry* __restrict__ ent, int ldb_src, int ldb_dst, int B, unsigned char* __restrict__ out) {
  const RebCarryEntry en = ent[blockIdx.y];
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < en.ne * ldb_dst; t += gridDim.x * blockDim.x) {
    const int e = t / ldb_dst, b = t % ldb_dst;
    out[en.dst + (size_t)e * ldb_dst + b] = (en.src && b < B) ? en.src[(size_t)e * ldb_src + b] : (unsigned char)0;
  }
}

__global__ void k_init_devx0(const SolverDev* __restrict__ Sp, const double* __restrict__ x0dev, int ldx, int n0) {
  const SolverDev& S = *Sp;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= S.B) return;
  double* x0 = const_cast<double*>(S.x0);
  for (int i = 0; i < n0; ++i) x0[(size_t)i * S.ldb + b] = x0dev[(size_t)i * ldx + b];
}

__global__ void k_pack(const PackSeg* __restrict__ segs, int nseg, int ldb, int b0, int nb, long sol_size, double* __restrict__ out) {
  const int s = blockIdx.y;
  if (s >= nseg) return;
  const PackSeg sg = segs[s];
  const long total = (long)sg.knots * sg.nc;
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total * nb; t += (long)gridDim.x * blockDim.x) {
    const int bb = (int)(t % nb);
    const long e = t / nb;
    double v;
    if (sg.pm_ld > 0) {
      const int kk = (int)(e / sg.nc), c = (int)(e % sg.nc);
      v = sg.src[((size_t)(b0 + bb) * sg.pm_h + kk) * ((size_t)sg.pm_ld * (sg.nc / sg.pm_rows)) + (c % sg.pm_rows) + sg.pm_ld * (c / sg.pm_rows)];
    } else {
      v = sg.src[(size_t)e * ldb + b0 + bb];
    }
    out[(size_t)bb * sol_size + sg.dst + e] = v;
  }
}

// inverse of k_pack for the warm start: packed records -> batch-major arrays (optionally two destinations: Xbar and X, Ubar and U)
struct UnpackSeg { double* dst; double* dst2; int knots, nc; long src; };
__global__ void k_unpack(const UnpackSeg* __restrict__ segs, int nseg, int ldb, int nb, long sol_size, const double* __restrict__ in) {
  const int s = blockIdx.y;
  if (s >= nseg) return;
  const UnpackSeg sg = segs[s];
  const long total = (long)sg.knots * sg.nc;
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total * nb; t += (long)gridDim.x * blockDim.x) {
    const int bb = (int)(t % nb);
    const long e = t / nb;
    const double v = in[(size_t)bb * sol_size + sg.src + e];
    sg.dst[(size_t)e * ldb + bb] = v;
    if (sg.dst2) sg.dst2[(size_t)e * ldb + bb] = v;
  }
}

// Receding-horizon shift on the device: one entry = one knot of one array (Xbar, Ubar or K) of the NEW deck's packed guess record,
// filled from a knot of the previous solver's arrays (src, batch-major with its own ldb) or, for states of a phase the old plan did
// not have yet, from the new deck's reference record (ref, optionally per problem). Entries without a source stay zero.
struct ShiftEntry { const double* src; const double* ref; const double* ref_pp; int src_knot, nc, ref_knot; long dst; };
__global__ void k_shift_guess(const ShiftEntry* __restrict__ ent, int n_ent, int ldb_src, int ldb_dst, int nb, long sol_size, double* __restrict__ out) {
  const int s = blockIdx.y;
  if (s >= n_ent) return;
  const ShiftEntry e = ent[s];
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < (long)e.nc * nb; t += (long)gridDim.x * blockDim.x) {
    const int bb = (int)(t % nb), c = (int)(t / nb);
    double v;
    if (e.src) v = e.src[((size_t)e.src_knot * e.nc + c) * ldb_src + bb];
    else if (e.ref_pp) v = e.ref_pp[((size_t)e.ref_knot * CAFE_REF_W + CAFE_REF_XR + c) * ldb_dst + bb];
    else v = e.ref[(size_t)e.ref_knot * CAFE_REF_W + CAFE_REF_XR + c];
    out[(size_t)bb * sol_size + e.dst + c] = v;
  }
}

// float32 wire record (MHPC_Command_lcmt field order): one segment = nk knots x w components taken from components
// [c0, c0+w) of an array with nc_src components per knot, starting at knot k0 of its phase
struct PackSegF { const double* src; int nc_src, c0, w, k0, nk; long dst; int pm_ld, pm_rows, pm_h; int tr_cols, tr_ld; };   // pm_*: as in PackSeg
// tr_cols > 0: the w components are a row-major (w / tr_cols) x tr_cols block read from a column-major matrix with tr_ld rows (wire [m][n] = K(m, n))
__global__ void k_pack_lcm(const PackSegF* __restrict__ segs, int nseg, int ldb, int nb, long rec_size, float* __restrict__ out) {
  const int s = blockIdx.y;
  if (s >= nseg) return;
  const PackSegF sg = segs[s];
  const long total = (long)sg.nk * sg.w;
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total * nb; t += (long)gridDim.x * blockDim.x) {
    const int bb = (int)(t % nb);
    const long e = t / nb;
    const int kk = (int)(e / sg.w), c = (int)(e % sg.w);
    double v;
    if (sg.pm_ld > 0) {
      const int cc = sg.c0 + c;
      v = sg.src[((size_t)bb * sg.pm_h + sg.k0 + kk) * ((size_t)sg.pm_ld * (sg.nc_src / sg.pm_rows)) + (cc % sg.pm_rows) + sg.pm_ld * (cc / sg.pm_rows)];
    } else {
      const int cc = sg.tr_cols > 0 ? sg.c0 + (c / sg.tr_cols) + sg.tr_ld * (c % sg.tr_cols) : sg.c0 + c;
      v = sg.src[((size_t)(sg.k0 + kk) * sg.nc_src + cc) * ldb + bb];
    }
    out[(size_t)bb * rec_size + sg.dst + e] = (float)v;
  }
}

__global__ void k_fp64_peak(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 0.999999, c = 1e-7;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// the same measurement on the fp64 tensor pipe: 8 independent m8n8k4 accumulator pairs per warp (256 FMAs per instruction)
__global__ void k_fp64_mma_peak(double* out, int iters) {
  double c[8][2];
  const double a = 1e-3 * (threadIdx.x & 31), b = 1e-3;
  for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; }
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) cafe_dev::dmma884(c[i][0], c[i][1], a, b);
  double s = 0;
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

struct Carver {
  size_t off = 0;
  char* base = nullptr;
  template <class T>
  T* take(size_t count) {
    off = (off + 255) & ~(size_t)255;
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += count * sizeof(T);
    return p;
  }
};

// lays out every per-problem array; called twice (size pass with base == nullptr, then for real)
void carve(CafeHandle* H, Carver& cv, size_t& zero_bytes) {
  SolverDev& S = H->S;
  const size_t ldb = H->ldb;
  const int NA = H->NA;
  // ---- region re-zeroed at every solve
  CtrlDev& c = S.c;
  c.active = cv.take<int>(ldb); c.do_ls = cv.take<int>(ldb); c.sel = cv.take<int>(ldb); c.accepted = cv.take<int>(ldb);
  c.ls_found = cv.take<int>(ldb); c.ls_fail = cv.take<int>(ldb); c.n_pending = cv.take<int>(64);
  c.ls_cost = cv.take<double>(ldb); c.ls_feas = cv.take<double>(ldb); c.ls_mt = cv.take<double>(ldb); c.ls_mp = cv.take<double>(ldb); c.ls_merit = cv.take<double>(ldb);
  c.iter_ou = cv.take<int>(ldb); c.iter_in = cv.take<int>(ldb); c.iter = cv.take<int>(ldb); c.ls_total = cv.take<int>(ldb);
  c.reg_total = cv.take<int>(ldb); c.n_hist = cv.take<int>(ldb); c.status = cv.take<int>(ldb);
  c.n_active = cv.take<int>(64);
  c.act_list = cv.take<int>(ldb); c.pend_list = cv.take<int>(ldb); c.cur_slot = cv.take<int>(ldb); c.reb_upd = cv.take<int>(ldb);
  c.reg = cv.take<double>(ldb); c.cost = cv.take<double>(ldb); c.merit = cv.take<double>(ldb); c.feas = cv.take<double>(ldb);
  c.merit_rho = cv.take<double>(ldb); c.dV1 = cv.take<double>(ldb); c.dV2 = cv.take<double>(ldb);
  c.cost_prev = cv.take<double>(ldb); c.merit_prev = cv.take<double>(ldb);
  c.max_t = cv.take<double>(ldb); c.max_p = cv.take<double>(ldb); c.max_t_prev = cv.take<double>(ldb); c.max_p_prev = cv.take<double>(ldb);
  c.min_pivot = cv.take<double>(ldb);
  c.feas0_t = cv.take<double>((size_t)NA * ldb);
  c.hist = cv.take<double>((size_t)CAFE_HIST_CAP * 4 * ldb);
  c.trace = cv.take<double>((size_t)CAFE_HIST_CAP * 12 * ldb);
  S.x0 = cv.take<double>((size_t)CAFE_MAX_N * ldb);
  // fail flags of all phases, contiguous so that one memset clears them
  {
    int* f = cv.take<int>((size_t)S.n_phases * NA * ldb);
    H->d_fail = f; H->fail_bytes = (size_t)S.n_phases * NA * ldb * sizeof(int);
    for (int i = 0; i < S.n_phases; ++i) S.ph[i].fail_t = f ? f + (size_t)i * NA * ldb : nullptr;
  }
  for (int i = 0; i < S.n_phases; ++i) {
    PhaseDev& ph = S.ph[i];
    const size_t n = ph.n, m = ph.m, p = ph.p, h = ph.h;
    ph.X = cv.take<double>((h + 1) * n * ldb); ph.Xbar = cv.take<double>((h + 1) * n * ldb); ph.dX = cv.take<double>((h + 1) * n * ldb);
    ph.G = cv.take<double>((h + 1) * n * ldb); ph.Defect = cv.take<double>((h + 1) * n * ldb);
    ph.U = cv.take<double>(h * m * ldb); ph.Ubar = cv.take<double>(h * m * ldb); ph.dU = cv.take<double>(h * m * ldb);
    ph.Qu = cv.take<double>(h * m * ldb); ph.Y = cv.take<double>(h * p * ldb + 1);
    ph.K = cv.take<double>(h * m * n * ldb);
    ph.lk = cv.take<double>((h + 1) * ldb); ph.dsq = cv.take<double>((h + 1) * ldb);
    ph.al_sigma = cv.take<double>(4 * ldb); ph.al_lambda = cv.take<double>(4 * ldb); ph.hval = cv.take<double>(4 * ldb);
    ph.maxh_t = cv.take<double>((size_t)NA * ldb); ph.ht = cv.take<double>((size_t)NA * 4 * ldb);
    ph.reb_ne = cafe_reb_elements(ph.model);
    ph.reb_n = cv.take<unsigned char>(h * (size_t)ph.reb_ne * ldb + 1);
  }
  zero_bytes = cv.off;
  // ---- region whose zero pattern is static (zeroed once at create)
  for (int i = 0; i < S.n_phases; ++i) {
    PhaseDev& ph = S.ph[i];
    const size_t n = ph.n, m = ph.m, p = ph.p, h = ph.h, nn = ph.n_next > 0 ? ph.n_next : 1;
    ph.A = cv.take<double>(h * n * n * ldb); ph.Bm = cv.take<double>(h * n * m * ldb);
    ph.C = cv.take<double>(h * p * n * ldb + 1); ph.D = cv.take<double>(h * p * m * ldb + 1);
    ph.lx = cv.take<double>(h * n * ldb); ph.lu = cv.take<double>(h * m * ldb); ph.ly = cv.take<double>(h * p * ldb + 1);
    ph.lxx = cv.take<double>(h * n * n * ldb); ph.luu = cv.take<double>(h * m * m * ldb); ph.lyy = cv.take<double>(h * p * p * ldb + 1);
    ph.Phix = cv.take<double>(n * ldb); ph.Phixx = cv.take<double>(n * n * ldb); ph.Px = cv.take<double>(nn * n * ldb);
    const bool wb = ph.model == CAFE_MODEL_WB;
    ph.tm = cv.take<double>(wb ? (size_t)NA * h * CAFE_TM_W * ldb : 1); ph.qdd_t = cv.take<double>(wb ? (size_t)NA * h * 18 * ldb : 1);
    ph.dp = cv.take<double>(wb ? h * (size_t)CAFE_DP_W * ldb : 1);
    ph.ABpm = cv.take<double>(wb ? h * (size_t)CAFE_WB_AB_TILE * ldb : 2); ph.CDpm = cv.take<double>(wb ? h * (size_t)CAFE_WB_CD_TILE * ldb : 2);
    ph.Kpm = cv.take<double>(h * (size_t)cafe_dev::ld_mma((int)m) * n * ldb + 2);
    ph.Quu = cv.take<double>(h * (size_t)cafe_dev::ld_mma((int)m) * m * ldb + 2); ph.Qux = cv.take<double>(h * (size_t)cafe_dev::ld_mma((int)m) * n * ldb + 2);   // problem-major tiles [b][h][ld(m) x m | n]
    ph.Xt = cv.take<double>((size_t)NA * (h + 1) * n * ldb); ph.Ut = cv.take<double>((size_t)NA * h * m * ldb);
    ph.Yt = cv.take<double>((size_t)NA * h * p * ldb + 1); ph.Dt = cv.take<double>((size_t)NA * (h + 1) * n * ldb);
    ph.cost_t = cv.take<double>((size_t)NA * (h + 1) * ldb); ph.feas_t = cv.take<double>((size_t)NA * (h + 1) * ldb);
    ph.ming_t = cv.take<double>((size_t)NA * (h + 1) * ldb);
  }
}

int compute_alphas(const CafeOptions& o, double* eps) {  // MultiPhaseDDP.cpp:109-130: eps = 1; while (eps > 1e-3) {...; eps *= alpha;}
  int n = 0;
  double e = 1;
  while (e > 1e-3 && n < CAFE_MAX_ALPHAS) { eps[n++] = e; e *= o.alpha; }
  return n;
}

template <class F>
int timed(CafeHandle* H, int slot, F&& launch) {
  if (H->profiling) cudaEventRecord(H->ev0, H->stream);
  launch();
  H->launches[slot]++;
  if (H->profiling) {
    cudaEventRecord(H->ev1, H->stream);
    cudaEventSynchronize(H->ev1);
    float ms = 0;
    cudaEventElapsedTime(&ms, H->ev0, H->ev1);
    H->ms[slot] += ms;
  }
  return 0;
}

int launch_bwd(CafeHandle* H, int first = 0, int n_list = -1, cudaStream_t st = nullptr) {
  if (n_list < 0) { n_list = H->S.n_act; st = H->stream; }
  const unsigned grid = (n_list + 3) / 4 * 4;   // clusters of four listed problems
  if (grid == 0) return 0;
  SolverDev S = H->S;   // the kernel's descriptor: entries [first, first + n_list) of the active list
  S.c.act_list += first; S.n_act = n_list;
  if (H->bwd_variant == 0) k_bwd2<0, 128><<<grid, 128, H->bwd_smem, st>>>(S);
  else if (H->bwd_nt == 256 || n_list <= H->bwd_small) k_bwd2<1, 256><<<grid, 256, H->bwd_smem, st>>>(S);
  else k_bwd2<1, 128><<<grid, 128, H->bwd_smem, st>>>(S);
  return 0;
}

}  // namespace

// inside cafe_gpu_create after the handle exists: a failure releases everything created so far
#define CUDA_OK_H(call)                                                                      \
  do {                                                                                       \
    cudaError_t e_ = (call);                                                                 \
    if (e_ != cudaSuccess) {                                                                 \
      cafe::set_last_error(std::string(#call) + ": " + cudaGetErrorString(e_));              \
      cafe_gpu_destroy(H);                                                                   \
      return CAFE_ERR_CUDA;                                                                  \
    }                                                                                        \
  } while (0)

// device copy of a host table, ordered on H->stream behind everything already queued there
template <class T>
static int upload_table(CafeHandle* H, const std::vector<T>& v, T** out) {
  const size_t bytes = v.size() * sizeof(T);
  if (bytes > H->tab_cap) {
    if (H->tab_busy) { CUDA_OK(cudaEventSynchronize(H->ev_tab)); H->tab_busy = false; }
    cudaFree(H->d_tab); if (H->h_tab) cudaFreeHost(H->h_tab);
    H->d_tab = nullptr; H->h_tab = nullptr; H->tab_cap = 0;
    const size_t cap = std::max<size_t>(2 * bytes, 256 * 1024);
    CUDA_OK(cudaMalloc(&H->d_tab, cap));
    CUDA_OK(cudaMallocHost(&H->h_tab, cap));
    H->tab_cap = cap;
  }
  if (!H->ev_tab) CUDA_OK(cudaEventCreateWithFlags(&H->ev_tab, cudaEventDisableTiming));
  if (H->tab_busy) CUDA_OK(cudaEventSynchronize(H->ev_tab));   // the staging buffer may still be read by the previous copy
  std::memcpy(H->h_tab, v.data(), bytes);
  CUDA_OK(cudaMemcpyAsync(H->d_tab, H->h_tab, bytes, cudaMemcpyHostToDevice, H->stream));
  CUDA_OK(cudaEventRecord(H->ev_tab, H->stream));
  H->tab_busy = true;
  *out = reinterpret_cast<T*>(H->d_tab);
  return 0;
}

static int validate_deck(const CafeDeck* deck, bool& all_hkd, int& n_knots) {
  if (!deck || deck->n_phases <= 0 || deck->n_phases > CAFE_MAX_PHASES) { cafe::set_last_error("bad argument"); return CAFE_ERR_ARG; }
  all_hkd = true; n_knots = 0;
  for (int i = 0; i < deck->n_phases; ++i) { all_hkd = all_hkd && deck->phase[i].model == CAFE_MODEL_HKD; n_knots += deck->phase[i].horizon + 1; }
  bool any_hkd = false;
  for (int i = 0; i < deck->n_phases; ++i) any_hkd = any_hkd || deck->phase[i].model == CAFE_MODEL_HKD;
  if (!all_hkd && any_hkd) { cafe::set_last_error("a deck cannot mix HKD phases with WB/SRB phases"); return CAFE_ERR_UNSUPPORTED; }
  for (int i = 0; i + 1 < deck->n_phases; ++i)
    if (deck->phase[i].model == CAFE_MODEL_SRB) { cafe::set_last_error("an SRB phase must be the last phase"); return CAFE_ERR_UNSUPPORTED; }
  if (n_knots > CAFE_MAX_KNOTS) { cafe::set_last_error("horizon too long"); return CAFE_ERR_UNSUPPORTED; }
  for (int i = 0; i < deck->n_phases; ++i) {
    const CafePhase& p = deck->phase[i];
    if (p.single_shooting && i > 0 && (deck->phase[i - 1].model != p.model || deck->phase[i - 1].single_shooting)) {
      cafe::set_last_error("a single-shooting phase must follow a shooting phase of the same model"); return CAFE_ERR_UNSUPPORTED;
    }
  }
  return 0;
}

// everything of a handle that depends on the deck: descriptor, arena layout, reference records, structural masks. Called by
// cafe_gpu_create and, with the allocations of the previous deck re-used where they are large enough, by cafe_gpu_update_deck.
static int configure(CafeHandle* H, const CafeDeck* deck, bool all_hkd, int n_knots) {
  H->deck = *deck;
  H->ref_host.assign(deck->ref, deck->ref + (size_t)deck->n_records * CAFE_REF_W);
  H->deck.ref = H->ref_host.data();
  H->NA = CAFE_MAX_ALPHAS;  // trial slots are sized for the longest step-size ladder
  // the ladder depends on option.alpha; size for alpha <= 0.5 (10 trials) -- see solve
  H->NA = 10;
  SolverDev& S = H->S;
  std::memset(&S, 0, sizeof(S));
  S.n_phases = deck->n_phases; S.ldb = H->ldb; S.n_knots = n_knots;
  int gk = 0;
  for (int i = 0; i < deck->n_phases; ++i) {
    const CafePhase& p = deck->phase[i];
    PhaseDev& d = S.ph[i];
    d.model = p.model; d.n = cafe_model_n(p.model); d.m = cafe_model_m(p.model); d.p = cafe_model_p(p.model); d.h = p.horizon;
    d.has_next = (i < deck->n_phases - 1) ? 1 : 0;
    d.n_next = d.has_next ? cafe_model_n(deck->phase[i + 1].model) : 0;
    for (int l = 0; l < 4; ++l) { d.contact[l] = p.contact[l]; d.next_contact[l] = p.next_contact[l]; d.td_foot[l] = p.td_foot[l]; }
    d.n_td = p.n_td; d.dt = p.dt; d.mu = p.mu; d.ground_height = p.ground_height; d.BG_alpha = deck->BG_alpha;
    d.h_min = p.h_min; d.torque_limit = p.torque_limit; d.no_joint_limit = p.no_joint_limit; d.no_min_height = p.no_min_height;
    d.single_shooting = p.single_shooting;
    d.joint_speed_limit = p.joint_speed_limit; d.reb_jointvel = p.reb_jointvel; d.jointvel_lb = p.jointvel_lb; d.jointvel_ub = p.jointvel_ub;
    for (int l = 0; l < 3; ++l) { d.joint_lb[l] = p.joint_lb[l]; d.joint_ub[l] = p.joint_ub[l]; }
    std::memcpy(d.q, p.q, sizeof(d.q)); std::memcpy(d.r, p.r, sizeof(d.r)); std::memcpy(d.qf, p.qf, sizeof(d.qf));
    std::memcpy(d.w_footreg, p.w_footreg, sizeof(d.w_footreg)); std::memcpy(d.w_swingpos, p.w_swingpos, sizeof(d.w_swingpos));
    std::memcpy(d.w_swingvel, p.w_swingvel, sizeof(d.w_swingvel)); std::memcpy(d.w_tdvel, p.w_tdvel, sizeof(d.w_tdvel));
    d.reb_grf = p.reb_grf; d.reb_torque = p.reb_torque; d.reb_joint = p.reb_joint; d.reb_minheight = p.reb_minheight; d.al_td = p.al_td;
    for (int k = 0; k <= p.horizon; ++k) {
      if (p.model == CAFE_MODEL_WB && k < p.horizon) S.wbk_gk[S.n_wbk++] = (short)gk;
      S.knot_phase[gk] = (short)i; S.knot_k[gk] = (short)k; ++gk;
    }
  }
  Carver sz;
  size_t zb = 0;
  carve(H, sz, zb);
  const size_t need_arena = sz.off + 256;
  if (need_arena > H->arena_bytes) {
    // a horizon window moves a few knots between phases from one MPC step to the next: room for that, so that updates re-use the arena
    cudaFree(H->arena); H->arena = nullptr; H->arena_bytes = 0;
    const size_t want = need_arena + need_arena / 16;
    cudaError_t e = cudaMalloc(&H->arena, want);
    if (e != cudaSuccess) { cafe::set_last_error(std::string("cudaMalloc arena: ") + cudaGetErrorString(e)); H->arena = nullptr; return CAFE_ERR_CUDA; }
    H->arena_bytes = want;
  }
  Carver cv; cv.base = H->arena;
  carve(H, cv, H->zero_bytes);
  CUDA_OK(cudaMemsetAsync(H->arena, 0, need_arena, H->stream));   // behind whatever still reads the previous layout on this stream
  if (H->ref_host.size() > H->ref_cap) {
    cudaFree(H->d_ref); H->d_ref = nullptr; H->ref_cap = 0;
    CUDA_OK(cudaMalloc(&H->d_ref, (H->ref_host.size() + 8 * CAFE_REF_W) * sizeof(double)));
    H->ref_cap = H->ref_host.size() + 8 * CAFE_REF_W;
  }
  CUDA_OK(cudaMemcpyAsync(H->d_ref, H->ref_host.data(), H->ref_host.size() * sizeof(double), cudaMemcpyHostToDevice, H->stream));
  for (int i = 0; i < deck->n_phases; ++i) { S.ph[i].ref = H->d_ref + (size_t)deck->phase[i].knot_offset * CAFE_REF_W; S.ph[i].ref_pp = nullptr; }
  {
    // structural pattern of the whole-body lxx per knot (cafe::wb_lxx_pattern, host/mhpc_problem.cpp)
    std::vector<unsigned long long> masks;
    std::vector<size_t> first(deck->n_phases, 0);
    for (int i = 0; i < deck->n_phases; ++i) {
      first[i] = masks.size();
      if (deck->phase[i].model != CAFE_MODEL_WB) continue;
      for (int k = 0; k < deck->phase[i].horizon; ++k) {
        unsigned long long w[CAFE_LXX_MASK_WORDS];
        cafe::wb_lxx_pattern(H->ref_host.data() + ((size_t)deck->phase[i].knot_offset + k) * CAFE_REF_W, w);
        masks.insert(masks.end(), w, w + CAFE_LXX_MASK_WORDS);
      }
    }
    if (!masks.empty()) {
      if (masks.size() > H->lxx_mask_cap) {
        cudaFree(H->d_lxx_mask); H->d_lxx_mask = nullptr; H->lxx_mask_cap = 0;
        CUDA_OK(cudaMalloc(&H->d_lxx_mask, (masks.size() + 8 * CAFE_LXX_MASK_WORDS) * sizeof(unsigned long long)));
        H->lxx_mask_cap = masks.size() + 8 * CAFE_LXX_MASK_WORDS;
      }
      CUDA_OK(cudaMemcpyAsync(H->d_lxx_mask, masks.data(), masks.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, H->stream));
    }
    for (int i = 0; i < deck->n_phases; ++i) S.ph[i].lxx_mask = (deck->phase[i].model == CAFE_MODEL_WB && H->d_lxx_mask) ? H->d_lxx_mask + first[i] : nullptr;
  }
  if (all_hkd && !H->d_hkd_mask) {
    unsigned long long hm[36];
    cafe::hkd_lq_patterns(hm);
    CUDA_OK(cudaMalloc(&H->d_hkd_mask, sizeof(hm)));
    CUDA_OK(cudaMemcpy(H->d_hkd_mask, hm, sizeof(hm), cudaMemcpyHostToDevice));
  }
  for (int i = 0; i < deck->n_phases; ++i) S.ph[i].hkd_mask = (deck->phase[i].model == CAFE_MODEL_HKD) ? H->d_hkd_mask : nullptr;
  CUDA_OK(cudaStreamSynchronize(H->stream));   // `masks` and the caller's deck must outlive the copies
  H->guess_B = 0; H->ref_pp_B = 0; H->B = 0;
  return 0;
}

extern "C" int cafe_gpu_create(const CafeDeck* deck, int device, int max_batch, CafeHandle** out) {
  if (!deck || !out || max_batch <= 0) { cafe::set_last_error("bad argument"); return CAFE_ERR_ARG; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cafe::set_last_error("no CUDA device: this library has no CPU fallback"); return CAFE_ERR_CUDA; }
  if (device < 0 || device >= ndev) { cafe::set_last_error("bad device index"); return CAFE_ERR_ARG; }
  bool all_hkd = true; int n_knots = 0;
  if (int rc = validate_deck(deck, all_hkd, n_knots)) return rc;
  CUDA_OK(cudaSetDevice(device));
  CafeHandle* H = new CafeHandle();
  H->device = device; H->max_batch = max_batch; H->ldb = (max_batch + 31) / 32 * 32;
  CUDA_OK_H(cudaStreamCreate(&H->stream));
  if (int rc = configure(H, deck, all_hkd, n_knots)) { cafe_gpu_destroy(H); return rc; }
  CUDA_OK_H(cudaMalloc(&H->d_x0raw, (size_t)H->ldb * CAFE_MAX_N * sizeof(double)));
  CUDA_OK_H(cudaMalloc(&H->d_al, (size_t)H->ldb * CAFE_MAX_PHASES * 8 * sizeof(double)));
  CUDA_OK_H(cudaMalloc(&H->dS, sizeof(SolverDev)));
  CUDA_OK_H(cudaHostAlloc(&H->h_nactive, 64, cudaHostAllocMapped));
  CUDA_OK_H(cudaHostGetDevicePointer(&H->d_nactive_map, H->h_nactive, 0));
  for (int i = 0; i < 3; ++i) { CUDA_OK_H(cudaStreamCreate(&H->stream2[i])); CUDA_OK_H(cudaEventCreateWithFlags(&H->ev_join[i], cudaEventDisableTiming)); }
  CUDA_OK_H(cudaEventCreateWithFlags(&H->ev_fork, cudaEventDisableTiming));
  for (int i = 0; i < 2; ++i) { CUDA_OK_H(cudaEventCreateWithFlags(&H->ev_lqf[i], cudaEventDisableTiming)); CUDA_OK_H(cudaEventCreateWithFlags(&H->ev_lqd[i], cudaEventDisableTiming)); CUDA_OK_H(cudaEventCreateWithFlags(&H->ev_lqj[i], cudaEventDisableTiming)); }
  CUDA_OK_H(cudaEventCreateWithFlags(&H->ev_lqc, cudaEventDisableTiming));
  if (const char* e = getenv("CAFE_LQ_OVERLAP")) H->lq_overlap = atoi(e) != 0;
  if (const char* e = getenv("CAFE_SPLIT_MIN")) H->split_min = atoi(e);
  if (const char* e = getenv("CAFE_SPLIT_N")) { H->split_n = atoi(e); if (H->split_n < 2) H->split_n = 2; if (H->split_n > 4) H->split_n = 4; }
  CUDA_OK_H(cudaEventCreate(&H->ev0));
  CUDA_OK_H(cudaEventCreate(&H->ev1));
  CUDA_OK_H(cudaEventCreate(&H->evs));
  CUDA_OK_H(cudaEventCreate(&H->eve));
  H->max_segs = 9 * CAFE_MAX_PHASES;
  CUDA_OK_H(cudaMalloc(&H->d_segs, H->max_segs * sizeof(PackSeg)));
  if (all_hkd) {
    H->bwd_variant = 0; H->bwd_pb = 1;
    H->bwd_smem = (size_t)cafe_dev::Bwd2Layout<24, 24, 0, false>::total * sizeof(double);
    CUDA_OK_H(cudaFuncSetAttribute(cafe_dev::k_bwd2<0, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)H->bwd_smem));
  } else {
    H->bwd_variant = 1; H->bwd_pb = 1;
    H->bwd_smem = (size_t)cafe_dev::Bwd2Layout<36, 12, 12, true>::total * sizeof(double);
    if (const char* e = std::getenv("CAFE_BWD_SMEM_PAD")) H->bwd_smem += (size_t)std::atol(e);   // dev switch: fewer sweep CTAs per SM (room for a concurrent kernel)
    CUDA_OK_H(cudaFuncSetAttribute(cafe_dev::k_bwd2<1, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)H->bwd_smem));
    CUDA_OK_H(cudaFuncSetAttribute(cafe_dev::k_bwd2<1, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)H->bwd_smem));
    if (const char* e = std::getenv("CAFE_BWD_NT")) H->bwd_nt = std::atoi(e) == 256 ? 256 : 128;
    if (const char* e = std::getenv("CAFE_BWD_SMALL")) H->bwd_small = std::atoi(e);
  }
  if (!all_hkd && cafe_dev::wb_coop_configure() != 0) { cafe::set_last_error("cudaFuncSetAttribute (cooperative whole-body kernels) failed"); cafe_gpu_destroy(H); return CAFE_ERR_CUDA; }
  // Thread-local arrays survive only in the terminal-knot code of the whole-body model (impact map and its Jacobian) and in the
  // single-shooting chains; their out-of-line callees need call-stack room. The limit is a per-context setting shared with every
  // other CUDA user of the process, so it is only ever RAISED, and only to what these kernels need (their own frame as reported by
  // cudaFuncGetAttributes plus the deepest callee chain, bounded by CAFE_STACK_BYTES, default 24 KiB).
  if (!all_hkd) {
    size_t need = cafe_dev::knot_kernels_local_bytes() + 8192, cur = 0;
    size_t cap = 24 * 1024;
    if (const char* e = getenv("CAFE_STACK_BYTES")) cap = (size_t)atol(e);
    if (need < cap) need = cap;
    CUDA_OK_H(cudaDeviceGetLimit(&cur, cudaLimitStackSize));
    if (cur < need) CUDA_OK_H(cudaDeviceSetLimit(cudaLimitStackSize, need));
  }
  *out = H;
  return 0;
}

extern "C" int cafe_gpu_destroy(CafeHandle* H) {
  if (!H) return 0;
  cafe_gpu_comm_destroy(H);
  cudaSetDevice(H->device);
  if (H->tab_busy && H->ev_tab) cudaEventSynchronize(H->ev_tab);
  cudaFree(H->d_tab); if (H->h_tab) cudaFreeHost(H->h_tab); if (H->ev_tab) cudaEventDestroy(H->ev_tab);
  cudaFree(H->arena); cudaFree(H->d_ref); cudaFree(H->d_ref_pp); cudaFree(H->d_lxx_mask); cudaFree(H->d_hkd_mask); cudaFree(H->d_guess); cudaFree(H->d_x0raw); cudaFree(H->d_al); cudaFree(H->d_reb_carry); cudaFree(H->d_pack_async[0]); cudaFree(H->d_pack_async[1]);
  if (H->stream_copy) cudaStreamDestroy(H->stream_copy);
  for (int i = 0; i < 2; ++i) { if (H->ev_packed[i]) cudaEventDestroy(H->ev_packed[i]); if (H->ev_landed[i]) cudaEventDestroy(H->ev_landed[i]); }
  cudaFree(H->dS); cudaFree(H->d_pack); cudaFree(H->d_segs);
  if (H->h_nactive) cudaFreeHost(H->h_nactive);
  if (H->stream) cudaStreamDestroy(H->stream);
  for (int i = 0; i < 3; ++i) { if (H->stream2[i]) cudaStreamDestroy(H->stream2[i]); if (H->ev_join[i]) cudaEventDestroy(H->ev_join[i]); }
  if (H->ev_fork) cudaEventDestroy(H->ev_fork);
  for (int i = 0; i < 2; ++i) { if (H->ev_lqf[i]) cudaEventDestroy(H->ev_lqf[i]); if (H->ev_lqd[i]) cudaEventDestroy(H->ev_lqd[i]); if (H->ev_lqj[i]) cudaEventDestroy(H->ev_lqj[i]); }
  if (H->ev_lqc) cudaEventDestroy(H->ev_lqc);
  if (H->ev0) cudaEventDestroy(H->ev0);
  if (H->ev1) cudaEventDestroy(H->ev1);
  if (H->evs) cudaEventDestroy(H->evs);
  if (H->eve) cudaEventDestroy(H->eve);
  delete H;
  return 0;
}

static int solve_common(CafeHandle* H, const double* x0_host, const double* x0_dev, int ldx, int B, const CafeOptions* opt) {
  if (!H || !opt || B <= 0 || B > H->max_batch) { cafe::set_last_error("bad argument"); return CAFE_ERR_ARG; }
  if (opt->max_AL_iter > 250) { cafe::set_last_error("more than 250 outer iterations: the relaxed-barrier update counts are 8 bits wide"); return CAFE_ERR_UNSUPPORTED; }
  if (opt->max_AL_iter * opt->max_DDP_iter + 1 > CAFE_HIST_CAP) { cafe::set_last_error("iteration caps exceed the history capacity"); return CAFE_ERR_UNSUPPORTED; }
  if (H->guess_B > 0 && B > H->guess_B) { cafe::set_last_error("batch larger than the initial-guess set"); return CAFE_ERR_ARG; }
  if (H->al_B > 0 && B > H->al_B) { cafe::set_last_error("batch larger than the carried augmented-Lagrangian parameter set"); return CAFE_ERR_ARG; }
  if (H->S.ph[0].ref_pp && B > H->ref_pp_B) { cafe::set_last_error("batch larger than the per-problem reference set"); return CAFE_ERR_ARG; }
  CUDA_OK(cudaSetDevice(H->device));
  SolverDev& S = H->S;
  S.B = B; S.opt = *opt;
  {
    // the relaxed-barrier parameters stay at their initial values unless an update can change one (ConstraintsBase.h:79-85, :194-209):
    // a factor different from 1, or an initial delta below its floor delta_min
    bool dyn = opt->ReB_active && (opt->update_relax != 1.0 || opt->update_ReB != 1.0);
    for (int i = 0; i < S.n_phases && opt->ReB_active; ++i) {
      const CafePhase& p = H->deck.phase[i];
      dyn = dyn || p.reb_grf.delta < p.reb_grf.delta_min || p.reb_torque.delta < p.reb_torque.delta_min || p.reb_joint.delta < p.reb_joint.delta_min ||
            p.reb_minheight.delta < p.reb_minheight.delta_min || (p.joint_speed_limit && p.reb_jointvel.delta < p.reb_jointvel.delta_min);
    }
    for (int i = 0; i < S.n_phases; ++i) { S.ph[i].reb_dyn = dyn ? 1 : 0; S.ph[i].reb_br = opt->update_relax; S.ph[i].reb_bw = opt->update_ReB; }
  }
  // MS = false: MultiPhaseDDP::hybrid_rollout clears every phase's shooting set (MultiPhaseDDP.cpp:65-68) - whole-problem single shooting
  for (int i = 0; i < S.n_phases; ++i) S.ph[i].single_shooting = (H->deck.phase[i].single_shooting || !opt->MS) ? 1 : 0;
  const bool reb_dyn = S.ph[0].reb_dyn != 0;
  S.NA = compute_alphas(*opt, S.eps);
  if (S.NA > H->NA) { cafe::set_last_error("step-size ladder longer than the allocated trial slots"); return CAFE_ERR_UNSUPPORTED; }
  // trial-slot strides are fixed by the allocation (H->NA slots), the ladder may be shorter
  H->B = B;
  for (int i = 0; i < CAFE_NKERNELS; ++i) { H->ms[i] = 0; H->launches[i] = 0; H->units[i] = 0; }
  H->ticks = 0;
  cudaStream_t st = H->stream;
  CUDA_OK(cudaEventRecord(H->evs, st));
  CUDA_OK(cudaMemcpyAsync(H->dS, &S, sizeof(SolverDev), cudaMemcpyHostToDevice, st));
  CUDA_OK(cudaMemsetAsync(H->arena, 0, H->zero_bytes, st));
  const int n0 = S.ph[0].n;
  if (x0_host) CUDA_OK(cudaMemcpyAsync(H->d_x0raw, x0_host, (size_t)B * n0 * sizeof(double), cudaMemcpyHostToDevice, st));
  const long long nthr_knots = (long long)S.ldb * S.n_knots;
  const int tpb = 128;
  const unsigned g_knots = (unsigned)((nthr_knots + tpb - 1) / tpb);
  timed(H, CAFE_K_MISC, [&] { k_init<<<g_knots, tpb, 0, st>>>(H->dS, H->d_x0raw, n0); });
  if (x0_dev) timed(H, CAFE_K_MISC, [&] { k_init_devx0<<<(B + 127) / 128, 128, 0, st>>>(H->dS, x0_dev, ldx, n0); });
  if (H->al_B > 0) timed(H, CAFE_K_MISC, [&] { k_apply_al<<<(B + 127) / 128, 128, 0, st>>>(H->dS, H->d_al); });
  if (H->reb_B > 0 && reb_dyn) {
    if (B > H->reb_B) { cafe::set_last_error("batch larger than the carried relaxed-barrier parameter set"); return CAFE_ERR_ARG; }
    for (int i = 0; i < S.n_phases; ++i)
      if (S.ph[i].h > 0) CUDA_OK(cudaMemcpyAsync(S.ph[i].reb_n, H->d_reb_carry + H->reb_carry_off[i], (size_t)S.ph[i].h * S.ph[i].reb_ne * S.ldb, cudaMemcpyDeviceToDevice, st));
  }
  if (H->guess_B > 0) {
    // warm start: Xbar (= X), Ubar (= U) and K of the caller's guess replace the cold-start values; the first rollout (eps = 0)
    // then applies U = Ubar + K (X - Xbar) around it, which is how the reference re-solves after MHPCProblem::update
    std::vector<UnpackSeg> segs;
    long off = 0;
    for (int i = 0; i < S.n_phases; ++i) {
      PhaseDev& ph = S.ph[i];
      const int n = ph.n, m = ph.m, p = ph.p, h = ph.h;
      segs.push_back(UnpackSeg{ph.Xbar, ph.X, h + 1, n, off}); off += (long)(h + 1) * n;
      segs.push_back(UnpackSeg{ph.Ubar, ph.U, h, m, off}); off += (long)h * m;
      off += (long)h * p + (long)h * m;  // Y, dU
      segs.push_back(UnpackSeg{ph.K, nullptr, h, m * n, off}); off += (long)h * m * n;
      off += (long)h * m + (long)h * m * m + (long)h * m * n + (long)(h + 1) * n;  // Qu, Quu, Qux, G
    }
    UnpackSeg* d_us = nullptr;
    if (int rc = upload_table(H, segs, &d_us)) return rc;
    dim3 grid(592, (unsigned)segs.size());
    timed(H, CAFE_K_MISC, [&] { k_unpack<<<grid, 256, 0, st>>>(d_us, (int)segs.size(), H->ldb, B, off, H->d_guess); });
  }
  // one rollout group = the thread-per-knot kernel (trial states and controls; SRB / HKD / terminal knots completely) followed, on
  // whole-body decks, by the leg-parallel rigid-body terms and the cooperative KKT solve of the running whole-body knots
  const int n_wbk = S.n_wbk;
  auto roll_group = [&](cudaStream_t sq, int a0, int a1, const int* list, int n_list, bool tm) {
    auto run = [&](int slot, auto&& f) { if (tm) timed(H, slot, f); else { f(); H->launches[slot]++; } };
    run(CAFE_K_ROLL, [&] { cafe_dev::launch_roll(H->dS, S.n_knots, sq, a0, a1, list, n_list); });
    H->units[CAFE_K_ROLL] += (double)n_list * S.n_knots * (a1 - a0);
    H->units[CAFE_K_WB_TERMS] += (double)n_list * n_wbk * (a1 - a0); H->units[CAFE_K_WB_FWD] += (double)n_list * n_wbk * (a1 - a0);
    if (n_wbk > 0) {
      run(CAFE_K_WB_TERMS, [&] { cafe_dev::launch_wb_terms(H->dS, n_wbk, sq, a0, a1, list, n_list); });
      run(CAFE_K_WB_FWD, [&] { cafe_dev::launch_wb_fwd(H->dS, n_wbk, sq, a0, a1, list, n_list); });
    }
  };
  // sb != sq: the thread-per-knot kernel (SRB / HKD / terminal knots: one long dependent chain per thread, latency bound) and the cost
  // partials of the whole-body knots run on the companion stream sb beside derivatives -> sensitivities on sq (slot e of the event arrays);
  // k_wb_cost and k_wb_sens both read the derivative pack, nothing else is shared between the two chains
  // sc != sb: a third stream for k_wb_cost, so that it starts with the derivatives instead of queueing behind the latency-bound k_lq on sb
  auto lq_group = [&](cudaStream_t sq, cudaStream_t sb, int e, const int* list, int n_list, bool tm, cudaStream_t sc = nullptr) {
    auto run = [&](int slot, auto&& f) { if (tm) timed(H, slot, f); else { f(); H->launches[slot]++; } };
    H->units[CAFE_K_LQ] += (double)n_list * S.n_knots; H->units[CAFE_K_BWD] += (double)n_list;
    H->units[CAFE_K_WB_DERIVS] += (double)n_list * n_wbk; H->units[CAFE_K_WB_SENS] += (double)n_list * n_wbk; H->units[CAFE_K_WB_COST] += (double)n_list * n_wbk;
    const bool two = sb != sq && n_wbk > 0;
    if (two) { cudaEventRecord(H->ev_lqf[e], sq); cudaStreamWaitEvent(sb, H->ev_lqf[e], 0); }
    run(CAFE_K_LQ, [&] { cafe_dev::launch_lq(H->dS, S.n_knots, two ? sb : sq, list, n_list); });
    if (n_wbk > 0) {
      run(CAFE_K_WB_DERIVS, [&] { cafe_dev::launch_wb_derivs(H->dS, n_wbk, sq, list, n_list); });
      const bool three = two && sc != nullptr && sc != sb && sc != sq;
      cudaStream_t s_cost = three ? sc : (two ? sb : sq);
      if (two) { cudaEventRecord(H->ev_lqd[e], sq); cudaStreamWaitEvent(s_cost, H->ev_lqd[e], 0); }
      run(CAFE_K_WB_SENS, [&] { cafe_dev::launch_wb_sens(H->dS, n_wbk, sq, list, n_list); });
      run(CAFE_K_WB_COST, [&] { cafe_dev::launch_wb_cost(H->dS, n_wbk, s_cost, list, n_list); });
      if (three) { cudaEventRecord(H->ev_lqc, sc); cudaStreamWaitEvent(sq, H->ev_lqc, 0); }
      if (two) { cudaEventRecord(H->ev_lqj[e], sb); cudaStreamWaitEvent(sq, H->ev_lqj[e], 0); }
    }
  };
  // ---- initial rollout (eps = 0) and bookkeeping
  {
    SolverDev S0 = S;  // same pointers, ladder {0}
    S0.NA = 1; S0.eps[0] = 0.0;
    CUDA_OK(cudaMemcpyAsync(H->dS, &S0, sizeof(SolverDev), cudaMemcpyHostToDevice, st));
    timed(H, CAFE_K_SELECT, [&] { cafe_dev::launch_compact(H->dS, st, 0); });   // every problem is active: the identity list
    roll_group(st, 0, 1, S.c.act_list, B, true);
    CUDA_OK(cudaMemsetAsync(S.c.n_active, 0, sizeof(int), st));
    timed(H, CAFE_K_SELECT, [&] { cafe_dev::launch_select(H->dS, B, st, 0); });
    timed(H, CAFE_K_SELECT, [&] { cafe_dev::launch_compact(H->dS, st, 0); });
    timed(H, CAFE_K_ACCEPT, [&] { cafe_dev::launch_accept(H->dS, nthr_knots, st); });
    if (reb_dyn) timed(H, CAFE_K_ACCEPT, [&] { cafe_dev::launch_reb_update(H->dS, nthr_knots, st); });
    CUDA_OK(cudaStreamSynchronize(st));  // S0 must stay alive until the copy has been consumed
    CUDA_OK(cudaMemcpyAsync(H->dS, &S, sizeof(SolverDev), cudaMemcpyHostToDevice, st));
  }
  const int max_ticks = opt->max_AL_iter * opt->max_DDP_iter + 2;
  for (int tick = 0; tick < max_ticks; ++tick) {
    k_publish_int<<<1, 1, 0, st>>>(S.c.n_active, H->d_nactive_map);
    CUDA_OK(cudaStreamSynchronize(st));
    if (*(volatile int*)H->h_nactive == 0) break;
    H->ticks++;
    const int n_act = *H->h_nactive;   // = length of c.act_list (k_compact)
    H->S.n_act = n_act;
    CUDA_OK(cudaMemsetAsync(H->d_fail, 0, H->fail_bytes, st));
    const bool split = !H->profiling && H->split_min > 0 && n_act >= H->split_min;
    const int a1_first = 1 < S.NA ? 1 : S.NA;
    if (split) {
      // the active list in split_n parts on as many streams: LQ -> sweep -> first line-search group per part
      const int np = H->split_n;
      const int part = ((n_act + np - 1) / np + 127) & ~127;
      const int* lst = S.c.act_list;
      CUDA_OK(cudaEventRecord(H->ev_fork, st));
      for (int q = 0; q < np; ++q) {
        const int first = q * part, cnt = std::min(part, n_act - first);
        if (cnt <= 0) break;
        cudaStream_t sq = q == 0 ? st : H->stream2[q - 1];
        if (q > 0) CUDA_OK(cudaStreamWaitEvent(sq, H->ev_fork, 0));
        // companion streams: part 0 -> stream2[1], part 1 -> stream2[2] (two parts); more parts keep one stream each
        cudaStream_t sb = (H->lq_overlap && np == 2) ? H->stream2[1 + q] : sq;
        lq_group(sq, sb, q & 1, lst + first, cnt, false);
        launch_bwd(H, first, cnt, sq);
        H->launches[CAFE_K_BWD]++;
        roll_group(sq, 0, a1_first, lst + first, cnt, false);
        if (q > 0) { CUDA_OK(cudaEventRecord(H->ev_join[q - 1], sq)); CUDA_OK(cudaStreamWaitEvent(st, H->ev_join[q - 1], 0)); }
      }
    } else {
      lq_group(st, (H->lq_overlap && !H->profiling) ? H->stream2[0] : st, 0, S.c.act_list, n_act, true, (H->lq_overlap && !H->profiling) ? H->stream2[1] : nullptr);
      timed(H, CAFE_K_BWD, [&] { launch_bwd(H); });
    }
    // staged line search: step sizes are evaluated in growing groups; most problems accept one of the first
    for (int a0 = 0, width = 1; a0 < S.NA; a0 += width, width *= 2) {
      const int a1 = (a0 + width < S.NA) ? a0 + width : S.NA;
      // the first group runs over the active list (problems that skip the line search return at once), later groups over
      // the list of line searches that still need step sizes
      if (a0 == 0) { if (!split) roll_group(st, a0, a1, S.c.act_list, n_act, true); }
      else roll_group(st, a0, a1, S.c.pend_list, H->h_nactive[1], true);
      CUDA_OK(cudaMemsetAsync(S.c.n_pending, 0, sizeof(int), st));
      timed(H, CAFE_K_SELECT, [&] { cafe_dev::launch_ls_scan(H->dS, B, st, a0, a1); });
      if (a1 >= S.NA) break;
      timed(H, CAFE_K_SELECT, [&] { cafe_dev::launch_compact(H->dS, st, 1); });
      k_publish_int<<<1, 1, 0, st>>>(S.c.n_pending, H->d_nactive_map + 1);
      CUDA_OK(cudaStreamSynchronize(st));
      if (H->h_nactive[1] == 0) break;
    }
    CUDA_OK(cudaMemsetAsync(S.c.n_active, 0, sizeof(int), st));
    timed(H, CAFE_K_SELECT, [&] { cafe_dev::launch_select(H->dS, B, st, 1); });
    timed(H, CAFE_K_SELECT, [&] { cafe_dev::launch_compact(H->dS, st, 0); });
    timed(H, CAFE_K_ACCEPT, [&] { cafe_dev::launch_accept(H->dS, nthr_knots, st); });
    if (reb_dyn) timed(H, CAFE_K_ACCEPT, [&] { cafe_dev::launch_reb_update(H->dS, nthr_knots, st); });
  }
  CUDA_OK(cudaEventRecord(H->eve, st));
  CUDA_OK(cudaStreamSynchronize(st));
  CUDA_OK(cudaGetLastError());
  { float ms = 0; cudaEventElapsedTime(&ms, H->evs, H->eve); H->total_ms = ms; }
  return 0;
}

extern "C" int cafe_gpu_solve_batch(CafeHandle* H, const double* x0, int B, const CafeOptions* opt) {
  if (!x0) { cafe::set_last_error("null x0"); return CAFE_ERR_ARG; }
  return solve_common(H, x0, nullptr, 0, B, opt);
}
extern "C" int cafe_gpu_solve_batch_device(CafeHandle* H, const double* x0_dev, int ldx, int B, const CafeOptions* opt) {
  if (!x0_dev || ldx < B) { cafe::set_last_error("bad x0_dev"); return CAFE_ERR_ARG; }
  return solve_common(H, nullptr, x0_dev, ldx, B, opt);
}

extern "C" int cafe_gpu_set_profiling(CafeHandle* H, int on) { if (!H) return CAFE_ERR_ARG; H->profiling = on != 0; return 0; }
extern "C" int cafe_gpu_get_timing(CafeHandle* H, double ms[CAFE_NKERNELS], long launches[CAFE_NKERNELS], int* ticks) {
  if (!H) return CAFE_ERR_ARG;
  for (int i = 0; i < CAFE_NKERNELS; ++i) { if (ms) ms[i] = H->ms[i]; if (launches) launches[i] = H->launches[i]; }
  if (ticks) *ticks = H->ticks;
  return 0;
}
extern "C" int cafe_gpu_get_units(CafeHandle* H, double units[CAFE_NKERNELS]) {
  if (!H || !units) return CAFE_ERR_ARG;
  for (int i = 0; i < CAFE_NKERNELS; ++i) units[i] = H->units[i];
  return 0;
}
extern "C" int cafe_gpu_get_solve_ms(CafeHandle* H, double* ms) { if (!H || !ms) return CAFE_ERR_ARG; *ms = H->total_ms; return 0; }

template <class T>
static int fetch(CafeHandle* H, const T* dev, size_t count, std::vector<T>& host) {
  host.resize(count);
  CUDA_OK(cudaMemcpy(host.data(), dev, count * sizeof(T), cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int cafe_gpu_get_info(CafeHandle* H, CafeInfo* info) {
  if (!H || !info) { cafe::set_last_error("bad argument"); return CAFE_ERR_ARG; }
  CUDA_OK(cudaSetDevice(H->device));
  const int B = H->B;
  const CtrlDev& c = H->S.c;
  std::vector<int> st, it, ls, rg, ou, nh;
  std::vector<double> cost, feas, mt, mp;
  int rc;
  if ((rc = fetch(H, c.status, B, st)) || (rc = fetch(H, c.iter, B, it)) || (rc = fetch(H, c.ls_total, B, ls)) || (rc = fetch(H, c.reg_total, B, rg)) ||
      (rc = fetch(H, c.iter_ou, B, ou)) || (rc = fetch(H, c.n_hist, B, nh)) || (rc = fetch(H, c.cost, B, cost)) || (rc = fetch(H, c.feas, B, feas)) ||
      (rc = fetch(H, c.max_t, B, mt)) || (rc = fetch(H, c.max_p, B, mp)))
    return rc;
  for (int b = 0; b < B; ++b) {
    info[b].status = st[b]; info[b].iter = it[b]; info[b].ls_iter_total = ls[b]; info[b].reg_iter_total = rg[b];
    info[b].outer_iter = ou[b]; info[b].n_hist = nh[b]; info[b].cost = cost[b]; info[b].feas = feas[b];
    info[b].max_tconstr = mt[b]; info[b].max_pconstr = mp[b];
  }
  return 0;
}

static int get_table(CafeHandle* H, const double* dev, int width, double* out, int cap) {
  const int B = H->B, ldb = H->ldb;
  const int rows = cap < CAFE_HIST_CAP ? cap : CAFE_HIST_CAP;
  std::vector<double> tmp;
  int rc = fetch(H, dev, (size_t)rows * width * ldb, tmp);
  if (rc) return rc;
  for (int b = 0; b < B; ++b)
    for (int r = 0; r < rows; ++r)
      for (int j = 0; j < width; ++j) out[((size_t)b * cap + r) * width + j] = tmp[((size_t)r * width + j) * ldb + b];
  return 0;
}
extern "C" int cafe_gpu_get_history(CafeHandle* H, double* hist, int cap) {
  if (!H || !hist || cap <= 0) { cafe::set_last_error("bad argument"); return CAFE_ERR_ARG; }
  CUDA_OK(cudaSetDevice(H->device));
  return get_table(H, H->S.c.hist, 4, hist, cap);
}
extern "C" int cafe_gpu_get_trace(CafeHandle* H, double* trace, int cap) {
  if (!H || !trace || cap <= 0) { cafe::set_last_error("bad argument"); return CAFE_ERR_ARG; }
  CUDA_OK(cudaSetDevice(H->device));
  return get_table(H, H->S.c.trace, 12, trace, cap);
}

static int run_pack(CafeHandle* H, const std::vector<PackSeg>& segs, long rec_size, int b0, int nb, double* out, double* dev_out = nullptr) {
  const size_t need = (size_t)nb * rec_size * sizeof(double);
  if (dev_out) {  // pack straight into a caller-owned device buffer (e.g. the NCCL gather source)
    if ((int)segs.size() > H->max_segs) { cafe::set_last_error("too many pack segments"); return CAFE_ERR_ARG; }
    CUDA_OK(cudaMemcpyAsync(H->d_segs, segs.data(), segs.size() * sizeof(PackSeg), cudaMemcpyHostToDevice, H->stream));
    dim3 grid(592, (unsigned)segs.size());
    k_pack<<<grid, 256, 0, H->stream>>>(H->d_segs, (int)segs.size(), H->ldb, b0, nb, rec_size, dev_out);
    CUDA_OK(cudaStreamSynchronize(H->stream));
    CUDA_OK(cudaGetLastError());
    return 0;
  }
  if (need > H->pack_bytes) {
    cudaFree(H->d_pack); H->d_pack = nullptr; H->pack_bytes = 0;
    CUDA_OK(cudaMalloc(&H->d_pack, need));
    H->pack_bytes = need;
  }
  if ((int)segs.size() > H->max_segs) { cafe::set_last_error("too many pack segments"); return CAFE_ERR_ARG; }
  CUDA_OK(cudaMemcpyAsync(H->d_segs, segs.data(), segs.size() * sizeof(PackSeg), cudaMemcpyHostToDevice, H->stream));
  dim3 grid(592, (unsigned)segs.size());
  k_pack<<<grid, 256, 0, H->stream>>>(H->d_segs, (int)segs.size(), H->ldb, b0, nb, rec_size, H->d_pack);
  CUDA_OK(cudaMemcpyAsync(out, H->d_pack, need, cudaMemcpyDeviceToHost, H->stream));
  CUDA_OK(cudaStreamSynchronize(H->stream));
  CUDA_OK(cudaGetLastError());
  return 0;
}

// pack on the solver's stream into `dev_dst` without waiting for it (the caller orders whatever reads dev_dst behind ev_packed[slot])
static int run_pack_nowait(CafeHandle* H, const std::vector<PackSeg>& segs, long rec_size, int nb, double* dev_dst, int slot) {
  if ((int)segs.size() > H->max_segs) { cafe::set_last_error("too many pack segments"); return CAFE_ERR_ARG; }
  if (!H->stream_copy) {
    CUDA_OK(cudaStreamCreateWithFlags(&H->stream_copy, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) { CUDA_OK(cudaEventCreateWithFlags(&H->ev_packed[i], cudaEventDisableTiming)); CUDA_OK(cudaEventCreateWithFlags(&H->ev_landed[i], cudaEventDisableTiming)); }
  }
  // the slot's previous transfer has finished reading its buffers before they are written again
  if (H->async_used[slot]) CUDA_OK(cudaStreamWaitEvent(H->stream, H->ev_landed[slot], 0));
  // the segment table is shared with the blocking packers: staged from pageable memory, i.e. copied out of `segs` before this call returns
  CUDA_OK(cudaMemcpyAsync(H->d_segs, segs.data(), segs.size() * sizeof(PackSeg), cudaMemcpyHostToDevice, H->stream));
  dim3 grid(592, (unsigned)segs.size());
  k_pack<<<grid, 256, 0, H->stream>>>(H->d_segs, (int)segs.size(), H->ldb, 0, nb, rec_size, dev_dst);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaEventRecord(H->ev_packed[slot], H->stream));
  CUDA_OK(cudaStreamWaitEvent(H->stream_copy, H->ev_packed[slot], 0));
  H->async_used[slot] = true;
  return 0;
}

extern "C" int cafe_gpu_get_solution(CafeHandle* H, int b0, int nb, double* sol) {
  if (!H || !sol || b0 < 0 || nb <= 0 || b0 + nb > H->B) { cafe::set_last_error("bad argument"); return CAFE_ERR_ARG; }
  CUDA_OK(cudaSetDevice(H->device));
  std::vector<PackSeg> segs;
  long off = 0;
  for (int i = 0; i < H->S.n_phases; ++i) {
    const PhaseDev& ph = H->S.ph[i];
    const int n = ph.n, m = ph.m, p = ph.p, h = ph.h;
    auto add = [&](const double* src, int knots, int nc) { if (knots * nc > 0) segs.push_back(PackSeg{src, knots, nc, off, 0, 0, 0}); off += (long)knots * nc; };
    auto add_pm = [&](const double* src, int knots, int nc) { if (knots * nc > 0) segs.push_back(PackSeg{src, knots, nc, off, cafe_dev::ld_mma(m), m, h}); off += (long)knots * nc; };
    add(ph.Xbar, h + 1, n); add(ph.Ubar, h, m); add(ph.Y, h, p); add(ph.dU, h, m); add(ph.K, h, m * n);
    add(ph.Qu, h, m); add_pm(ph.Quu, h, m * m); add_pm(ph.Qux, h, m * n); add(ph.G, h + 1, n);
  }
  return run_pack(H, segs, off, b0, nb, sol);
}

static int commands_impl(CafeHandle* H, int n_gain_knots, double* cmd, double* dev_out);
extern "C" int cafe_gpu_get_commands(CafeHandle* H, int n_gain_knots, double* cmd) {
  if (!H || !cmd || n_gain_knots < 0) { cafe::set_last_error("bad argument"); return CAFE_ERR_ARG; }
  return commands_impl(H, n_gain_knots, cmd, nullptr);
}
extern "C" int cafe_gpu_get_commands_device(CafeHandle* H, int n_gain_knots, double* cmd_dev) {
  if (!H || !cmd_dev || n_gain_knots < 0) { cafe::set_last_error("bad argument"); return CAFE_ERR_ARG; }
  return commands_impl(H, n_gain_knots, nullptr, cmd_dev);
}
static long command_segments(CafeHandle* H, int n_gain_knots, std::vector<PackSeg>& segs);
static int commands_impl(CafeHandle* H, int n_gain_knots, double* cmd, double* dev_out) {
  CUDA_OK(cudaSetDevice(H->device));
  std::vector<PackSeg> segs;
  const long off = command_segments(H, n_gain_knots, segs);
  return run_pack(H, segs, off, 0, H->B, cmd, dev_out);
}
// Asynchronous collection of the command records: packed on the solver's stream (the next solve may be started at once: it is ordered behind
// the pack), copied to `cmd` (page-locked host memory, [B][cafe_command_size]) on a copy stream. slot = 0 / 1: two collections may be in flight.
extern "C" int cafe_gpu_get_commands_async(CafeHandle* H, int n_gain_knots, double* cmd, int slot) {
  if (!H || !cmd || n_gain_knots < 0 || slot < 0 || slot > 1 || H->B <= 0) { cafe::set_last_error("bad argument"); return CAFE_ERR_ARG; }
  CUDA_OK(cudaSetDevice(H->device));
  std::vector<PackSeg> segs;
  const long rec = command_segments(H, n_gain_knots, segs);
  const size_t need = (size_t)H->B * rec * sizeof(double);
  if (need > H->pack_async_bytes[slot]) {
    if (H->async_used[slot]) CUDA_OK(cudaEventSynchronize(H->ev_landed[slot]));
    cudaFree(H->d_pack_async[slot]); H->d_pack_async[slot] = nullptr; H->pack_async_bytes[slot] = 0;
    CUDA_OK(cudaMalloc(&H->d_pack_async[slot], need));
    H->pack_async_bytes[slot] = need;
  }
  if (int rc = run_pack_nowait(H, segs, rec, H->B, H->d_pack_async[slot], slot)) return rc;
  CUDA_OK(cudaMemcpyAsync(cmd, H->d_pack_async[slot], need, cudaMemcpyDeviceToHost, H->stream_copy));
  CUDA_OK(cudaEventRecord(H->ev_landed[slot], H->stream_copy));
  return 0;
}
// blocks until the records of the slot's last asynchronous collection have landed
extern "C" int cafe_gpu_commands_wait(CafeHandle* H, int slot) {
  if (!H || slot < 0 || slot > 1) { cafe::set_last_error("bad argument"); return CAFE_ERR_ARG; }
  if (!H->async_used[slot]) return 0;
  CUDA_OK(cudaSetDevice(H->device));
  CUDA_OK(cudaEventSynchronize(H->ev_landed[slot]));
  return 0;
}
static long command_segments(CafeHandle* H, int n_gain_knots, std::vector<PackSeg>& segs) {
  long off = 0;
  int left = n_gain_knots;
  for (int i = 0; i < H->S.n_phases; ++i) {
    const PhaseDev& ph = H->S.ph[i];
    const int n = ph.n, m = ph.m, p = ph.p, h = ph.h;
    auto add = [&](const double* src, int knots, int nc) { if (knots * nc > 0) segs.push_back(PackSeg{src, knots, nc, off, 0, 0, 0}); off += (long)knots * nc; };
    auto add_pm = [&](const double* src, int knots, int nc) { if (knots * nc > 0) segs.push_back(PackSeg{src, knots, nc, off, cafe_dev::ld_mma(m), m, h}); off += (long)knots * nc; };
    add(ph.Xbar, h + 1, n); add(ph.Ubar, h, m); add(ph.Y, h, p);
    const int g = left < h ? left : h;
    add(ph.K, g, m * n); add(ph.Qu, g, m); add_pm(ph.Quu, g, m * m); add_pm(ph.Qux, g, m * n);
    left -= g;
  }
  return off;
}

// ---- warm start: initial Xbar / Ubar / K per problem, in the packed solution layout (the other arrays of the record are ignored)
extern "C" int cafe_gpu_set_initial_guess(CafeHandle* H, const double* guess, int B) {
  if (!H) { cafe::set_last_error("bad argument"); return CAFE_ERR_ARG; }
  CUDA_OK(cudaSetDevice(H->device));
  if (!guess) { H->guess_B = 0; return 0; }  // back to the cold start (Xbar = reference, zero controls and gains)
  if (B <= 0 || B > H->max_batch) { cafe::set_last_error("bad batch size for the initial-guess set"); return CAFE_ERR_ARG; }
  const size_t need = (size_t)B * (size_t)cafe_solution_size(&H->deck) * sizeof(double);
  if (need > H->guess_bytes) {
    cudaFree(H->d_guess); H->d_guess = nullptr; H->guess_bytes = 0;
    CUDA_OK(cudaMalloc(&H->d_guess, need));
    H->guess_bytes = need;
  }
  CUDA_OK(cudaMemcpy(H->d_guess, guess, need, cudaMemcpyHostToDevice));
  H->guess_B = B;
  return 0;
}

// ---- receding-horizon warm start without leaving the device: what MHPCProblem::update / update_WB_plan /
//      update_SRB_plan do to the trajectories (MHPCProblem.cpp:252-397, TrajectoryManagement.cpp:130-228), as a map from the knots of
//      the new deck (start offset dst_k0) to the knots of the previous plan (start offset src_k0) held by `src`:
//        * a whole-body phase continues the old phase with the same stance that overlaps it in absolute time (popped knots vanish),
//        * knots past the old plan's end repeat its last state (push_back_state(X.back())), zero control and gain,
//        * a phase the old plan did not have starts from the new deck's reference states (zero control and gain),
//        * the trailing reduced-order phase is kept as it is while its horizon is unchanged (update_SRB_plan: nsteps = 0).
// the phases of the deck the shifted guess is for: dimensions, stance, reference records on the device
struct ShiftDst { int n_phases; struct P { int model, n, m, p, h, contact[4]; const double* ref; const double* ref_pp; } ph[CAFE_MAX_PHASES]; };
static ShiftDst shift_dst_of(const SolverDev& S) {
  ShiftDst d; d.n_phases = S.n_phases;
  for (int i = 0; i < S.n_phases; ++i) {
    const PhaseDev& p = S.ph[i];
    d.ph[i] = ShiftDst::P{p.model, p.n, p.m, p.p, p.h, {p.contact[0], p.contact[1], p.contact[2], p.contact[3]}, p.ref, p.ref_pp};
  }
  return d;
}
static int build_shift(const SolverDev& src, int src_k0, const ShiftDst& dst, int dst_k0, std::vector<ShiftEntry>& ent, long& sol_size) {
  struct Range { int idx, s, e; int contact[4]; };
  std::vector<Range> old_r, new_r;
  {
    int s = src_k0;
    const int lead = src.ph[0].model;  // whole-body for MHPC decks, hybrid kinodynamic for HKD decks (HKDProblem::update, HKDProblem.cpp:117-222)
    if (lead != CAFE_MODEL_SRB)
      for (int i = 0; i < src.n_phases && src.ph[i].model == lead; ++i) {
        Range r{i, s, s + src.ph[i].h, {0, 0, 0, 0}};
        for (int f = 0; f < 4; ++f) r.contact[f] = src.ph[i].contact[f];
        old_r.push_back(r);
        s += src.ph[i].h;
      }
    s = dst_k0;
    const int lead2 = dst.ph[0].model;
    if (lead2 != CAFE_MODEL_SRB)
      for (int i = 0; i < dst.n_phases && dst.ph[i].model == lead2; ++i) {
        Range r{i, s, s + dst.ph[i].h, {0, 0, 0, 0}};
        for (int f = 0; f < 4; ++f) r.contact[f] = dst.ph[i].contact[f];
        new_r.push_back(r);
        s += dst.ph[i].h;
      }
  }
  if (old_r.empty() || new_r.empty() || src.ph[0].model != dst.ph[0].model) { cafe::set_last_error("both decks must start with full-order phases of the same model"); return CAFE_ERR_UNSUPPORTED; }
  const int lead_model = dst.ph[0].model;
  const PhaseDev& last = src.ph[old_r.back().idx];
  const int old_end = old_r.back().e;
  long off = 0;
  for (int i = 0; i < dst.n_phases; ++i) {
    const ShiftDst::P& ph = dst.ph[i];
    const int n = ph.n, m = ph.m, p = ph.p, h = ph.h;
    const long oX = off, oU = oX + (long)(h + 1) * n, oK = oU + (long)h * m + (long)h * p + (long)h * m;
    off = oK + (long)h * m * n + (long)h * m + (long)h * m * m + (long)h * m * n + (long)(h + 1) * n;
    const double* ref = ph.ref; const double* ref_pp = ph.ref_pp;
    auto x_from = [&](const PhaseDev* sp, int sk, int k) { ent.push_back(ShiftEntry{sp ? sp->Xbar : nullptr, ref, ref_pp, sk, n, k, oX + (long)k * n}); };
    auto uk_from = [&](const PhaseDev& sp, int sk, int k) {
      // HKDProblem.cpp:220: the first control of the front trajectory is zeroed by every HKD update (the record is pre-zeroed)
      if (!(lead_model == CAFE_MODEL_HKD && i == 0 && k == 0)) ent.push_back(ShiftEntry{sp.Ubar, nullptr, nullptr, sk, m, 0, oU + (long)k * m});
      ent.push_back(ShiftEntry{sp.K, nullptr, nullptr, sk, m * n, 0, oK + (long)k * m * n});
    };
    if (ph.model != lead_model) {
      const int j = (int)old_r.size() + (i - (int)new_r.size());
      const bool keep = j >= 0 && j < src.n_phases && src.ph[j].model == ph.model && src.ph[j].h == h;
      for (int k = 0; k <= h; ++k) { x_from(keep ? &src.ph[j] : nullptr, k, k); if (keep && k < h) uk_from(src.ph[j], k, k); }
      continue;
    }
    const Range& nr = new_r[i];
    const Range* sr = nullptr;
    for (const Range& r : old_r) {
      bool same = true;
      for (int f = 0; f < 4; ++f) same = same && r.contact[f] == nr.contact[f];
      if (same && r.s <= nr.e && r.e >= nr.s) { sr = &r; break; }
    }
    const bool continues_last = sr && sr->idx == old_r.back().idx;
    for (int k = 0; k <= h; ++k) {
      const int a = nr.s + k;
      if (sr && sr->s <= a && a <= sr->e) x_from(&src.ph[sr->idx], a - sr->s, k);
      else if (continues_last && a > old_end) x_from(&last, last.h, k);
      else x_from(nullptr, 0, k);
      if (k < h && sr && sr->s <= a && a < sr->e) uk_from(src.ph[sr->idx], a - sr->s, k);
    }
  }
  sol_size = off;
  return 0;
}
// which old phase's touchdown constraint a phase of the new deck inherits (the rules of cafe_mpc_b200/mpc.py::shift_al): the reference keeps a
// phase's TouchDownConstraint object, and with it sigma / lambda, for as long as the phase lives - TerminalConstraintBase::reset_params, called
// by every update (HKDProblem.cpp:208, MHPCProblem.cpp:363), is an empty function (ConstraintsBase.h:367-374). A phase that continues an
// old phase (same stance, overlapping in absolute time) with the same touchdown feet inherits; any other constraint starts from the deck's values.
static void build_al_carry(const CafeDeck& od, int old_k0, const CafeDeck& nd, int new_k0, AlCarry& c) {
  c.n_phases = nd.n_phases;
  const int lead_o = od.phase[0].model, lead_n = nd.phase[0].model;
  int ns = new_k0;
  for (int i = 0; i < nd.n_phases; ++i) {
    const CafePhase& np_ = nd.phase[i];
    c.src[i] = -1; c.n_td[i] = np_.n_td; c.sigma0[i] = np_.al_td.sigma; c.lambda0[i] = np_.al_td.lambda;
    if (np_.model != lead_n || lead_n != lead_o || lead_n == CAFE_MODEL_SRB) continue;
    const int ne = ns + np_.horizon;
    int os = old_k0;
    for (int j = 0; j < od.n_phases && od.phase[j].model == lead_o; ++j) {
      const CafePhase& op = od.phase[j];
      const int oe = os + op.horizon;
      bool same = true;
      for (int f = 0; f < 4; ++f) same = same && op.contact[f] == np_.contact[f];
      if (same && os <= ne && oe >= ns) {
        bool feet = op.n_td > 0 && op.n_td == np_.n_td;
        for (int f = 0; feet && f < op.n_td; ++f) feet = op.td_foot[f] == np_.td_foot[f];
        if (feet) c.src[i] = j;
        break;
      }
      os = oe;
    }
    ns = ne;
  }
}
static int run_al_carry(CafeHandle* owner, const CafeHandle* src, int src_k0, const CafeDeck& nd, int dst_k0, int B) {
  AlCarry c;
  build_al_carry(src->deck, src_k0, nd, dst_k0, c);
  k_carry_al<<<(B + 127) / 128, 128, 0, owner->stream>>>(src->dS, c, B, owner->ldb, owner->d_al);
  CUDA_OK(cudaGetLastError());
  owner->al_B = B;
  return 0;
}

// the relaxed-barrier update counts travel with the knots (cafe_mpc_b200/mpc.py::shift_reb states the rules and where the reference has them from):
// overlapping knots of a continued phase keep theirs, a knot appended at the tail copies the LAST knot's, the trailing reduced-order phase keeps
// its data while its horizon is unchanged, everything else starts at zero updates. Only when the previous solve could change them (reb_dyn).
static int run_reb_carry(CafeHandle* owner, const CafeHandle* src, int src_k0, const CafeDeck& nd, int dst_k0, int B) {
  owner->reb_B = 0;
  if (!src->S.ph[0].reb_dyn) return 0;
  const CafeDeck& od = src->deck;
  const size_t ldb = (size_t)owner->ldb;
  size_t total = 0;
  for (int i = 0; i < nd.n_phases; ++i) { owner->reb_carry_off[i] = total; total += (size_t)nd.phase[i].horizon * cafe_reb_elements(nd.phase[i].model) * ldb; }
  if (total > owner->reb_carry_bytes) {
    CUDA_OK(cudaStreamSynchronize(owner->stream));
    cudaFree(owner->d_reb_carry); owner->d_reb_carry = nullptr; owner->reb_carry_bytes = 0;
    CUDA_OK(cudaMalloc(&owner->d_reb_carry, total + total / 4 + 256));
    owner->reb_carry_bytes = total + total / 4 + 256;
  }
  std::vector<RebCarryEntry> ent;
  const int lead_o = od.phase[0].model, lead_n = nd.phase[0].model;
  int n_lead_o = 0, n_lead_n = 0;
  while (n_lead_o < od.n_phases && od.phase[n_lead_o].model == lead_o) ++n_lead_o;
  while (n_lead_n < nd.n_phases && nd.phase[n_lead_n].model == lead_n) ++n_lead_n;
  int ns = dst_k0;
  for (int i = 0; i < nd.n_phases; ++i) {
    const CafePhase& np_ = nd.phase[i];
    const int ne = cafe_reb_elements(np_.model), h = np_.horizon;
    auto dst_of = [&](int k) { return owner->reb_carry_off[i] + (size_t)k * ne * ldb; };
    if (i >= n_lead_n || lead_n != lead_o || lead_n == CAFE_MODEL_SRB) {
      const int j = n_lead_o + (i - n_lead_n);
      const bool keep = i >= n_lead_n && j >= 0 && j < od.n_phases && od.phase[j].model == np_.model && od.phase[j].horizon == h;
      for (int k = 0; k < h; ++k) ent.push_back(RebCarryEntry{keep ? src->S.ph[j].reb_n + (size_t)k * ne * src->ldb : nullptr, dst_of(k), ne});
      continue;
    }
    const int ne_ = ns + h;
    int os = src_k0, sj = -1, sjs = 0, sje = 0;
    for (int j = 0; j < n_lead_o; ++j) {
      const CafePhase& op = od.phase[j];
      const int oe = os + op.horizon;
      bool same = true;
      for (int f = 0; f < 4; ++f) same = same && op.contact[f] == np_.contact[f];
      if (same && os <= ne_ && oe >= ns) { sj = j; sjs = os; sje = oe; break; }
      os = oe;
    }
    for (int k = 0; k < h; ++k) {
      const int a = ns + k;
      const unsigned char* sp = nullptr;
      if (sj >= 0 && sje > sjs) {
        if (a >= sjs && a < sje) sp = src->S.ph[sj].reb_n + (size_t)(a - sjs) * ne * src->ldb;
        else if (a >= sje) sp = src->S.ph[sj].reb_n + (size_t)(sje - sjs - 1) * ne * src->ldb;   // push_back(): params.push_back(params.back())
      }
      ent.push_back(RebCarryEntry{sp, dst_of(k), ne});
    }
    ns = ne_;
  }
  if (ent.empty()) return 0;
  RebCarryEntry* d_ent = nullptr;
  if (int rc = upload_table(owner, ent, &d_ent)) return rc;
  dim3 grid(64, (unsigned)ent.size());
  k_carry_reb<<<grid, 256, 0, owner->stream>>>(d_ent, src->ldb, owner->ldb, B, owner->d_reb_carry);
  CUDA_OK(cudaGetLastError());
  owner->reb_B = B;
  return 0;
}

// fills owner->d_guess (B packed records of sol_size doubles) on owner->stream
static int run_shift(CafeHandle* owner, const std::vector<ShiftEntry>& ent, int ldb_src, int B, long sol_size) {
  const size_t need = (size_t)B * (size_t)sol_size * sizeof(double);
  if (need > owner->guess_bytes) {
    cudaFree(owner->d_guess); owner->d_guess = nullptr; owner->guess_bytes = 0;
    CUDA_OK(cudaMalloc(&owner->d_guess, need + need / 8));
    owner->guess_bytes = need + need / 8;
  }
  cudaStream_t st = owner->stream;
  CUDA_OK(cudaMemsetAsync(owner->d_guess, 0, need, st));
  ShiftEntry* d_ent = nullptr;
  if (int rc = upload_table(owner, ent, &d_ent)) return rc;
  dim3 grid(64, (unsigned)ent.size());
  k_shift_guess<<<grid, 256, 0, st>>>(d_ent, (int)ent.size(), ldb_src, owner->ldb, B, sol_size, owner->d_guess);
  CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int cafe_gpu_shift_guess(CafeHandle* dst, CafeHandle* src, int src_k0, int dst_k0, int B) {
  if (!dst || !src || dst == src || B <= 0 || B > dst->max_batch || B > src->B || dst_k0 < src_k0) { cafe::set_last_error("bad argument"); return CAFE_ERR_ARG; }
  if (dst->device != src->device) { cafe::set_last_error("both solvers must live on the same device"); return CAFE_ERR_ARG; }
  CUDA_OK(cudaSetDevice(dst->device));
  std::vector<ShiftEntry> ent;
  long sol_size = 0;
  if (int rc = build_shift(src->S, src_k0, shift_dst_of(dst->S), dst_k0, ent, sol_size)) return rc;
  if (sol_size != cafe_solution_size(&dst->deck)) { cafe::set_last_error("internal: packed record size mismatch"); return CAFE_ERR_ARG; }
  // the previous solve ran on src's stream: order this after it, and later solves of dst after this
  CUDA_OK(cudaStreamSynchronize(src->stream));
  if (int rc = run_shift(dst, ent, src->ldb, B, sol_size)) return rc;
  if (int rc = run_al_carry(dst, src, src_k0, dst->deck, dst_k0, B)) return rc;
  if (int rc = run_reb_carry(dst, src, src_k0, dst->deck, dst_k0, B)) return rc;
  CUDA_OK(cudaStreamSynchronize(dst->stream));
  dst->guess_B = B;
  return 0;
}

// ---- the MPC update on ONE handle (MHPCLocomotion::update_mpc_if_needed -> MHPCProblem::update, MHPCLocomotion.cpp:91-150, MHPCProblem.cpp:252-397;
//      HKDMPCSolver::update -> HKDProblem::update): the deck re-cut `k_advance` knots later replaces the handle's deck, the previous solution
//      becomes the warm start (the rules of cafe_gpu_shift_guess), and the device arena, reference and mask buffers are re-used - a horizon
//      window only moves a few knots between phases from step to step, so nothing is allocated in steady state. Everything is queued on
//      the handle's stream: the shift reads the old layout, then the arena is cleared and laid out for the new deck. Per-problem
//      references (cafe_gpu_set_references) do not survive the update. B = 0: new deck, cold start.
extern "C" int cafe_gpu_update_deck(CafeHandle* H, const CafeDeck* deck, int k_advance, int B) {
  if (!H || !deck || k_advance < 0 || B < 0 || B > H->max_batch || (B > 0 && B > H->B)) { cafe::set_last_error("bad argument"); return CAFE_ERR_ARG; }
  bool all_hkd = true; int n_knots = 0;
  if (int rc = validate_deck(deck, all_hkd, n_knots)) return rc;
  if (all_hkd != (H->bwd_variant == 0)) { cafe::set_last_error("the new deck must be of the same family (HKD or whole-body / SRB) as the handle's"); return CAFE_ERR_UNSUPPORTED; }
  CUDA_OK(cudaSetDevice(H->device));
  if (B > 0) {
    // the new deck's reference records go to the device first: states of phases the old plan did not have are taken from them
    const size_t nref = (size_t)deck->n_records * CAFE_REF_W;
    if (nref > H->ref_cap) {
      CUDA_OK(cudaStreamSynchronize(H->stream));
      cudaFree(H->d_ref); H->d_ref = nullptr; H->ref_cap = 0;
      CUDA_OK(cudaMalloc(&H->d_ref, (nref + 8 * CAFE_REF_W) * sizeof(double)));
      H->ref_cap = nref + 8 * CAFE_REF_W;
    }
    CUDA_OK(cudaMemcpyAsync(H->d_ref, deck->ref, nref * sizeof(double), cudaMemcpyHostToDevice, H->stream));
    ShiftDst d; d.n_phases = deck->n_phases;
    for (int i = 0; i < deck->n_phases; ++i) {
      const CafePhase& p = deck->phase[i];
      d.ph[i] = ShiftDst::P{p.model, cafe_model_n(p.model), cafe_model_m(p.model), cafe_model_p(p.model), p.horizon,
                            {p.contact[0], p.contact[1], p.contact[2], p.contact[3]}, H->d_ref + (size_t)p.knot_offset * CAFE_REF_W, nullptr};
    }
    std::vector<ShiftEntry> ent;
    long sol_size = 0;
    if (int rc = build_shift(H->S, 0, d, k_advance, ent, sol_size)) return rc;
    if (sol_size != cafe_solution_size(deck)) { cafe::set_last_error("internal: packed record size mismatch"); return CAFE_ERR_ARG; }
    if (int rc = run_shift(H, ent, H->ldb, B, sol_size)) return rc;
    if (int rc = run_al_carry(H, H, 0, *deck, k_advance, B)) return rc;
    if (int rc = run_reb_carry(H, H, 0, *deck, k_advance, B)) return rc;
  } else {
    H->al_B = 0; H->reb_B = 0;
  }
  if (int rc = configure(H, deck, all_hkd, n_knots)) return rc;
  H->guess_B = B;
  return 0;
}

// planned state `knots_ahead` knots after the start of the plan (post-reset state at a phase boundary): out [B][n] on the host
extern "C" int cafe_gpu_get_planned_state(CafeHandle* H, int knots_ahead, double* out) {
  if (!H || !out || knots_ahead < 0) { cafe::set_last_error("bad argument"); return CAFE_ERR_ARG; }
  CUDA_OK(cudaSetDevice(H->device));
  int k = knots_ahead;
  for (int i = 0; i < H->S.n_phases; ++i) {
    const PhaseDev& ph = H->S.ph[i];
    if (k < ph.h || ph.model != H->S.ph[0].model || i == H->S.n_phases - 1) {
      if (k > ph.h || ph.n != H->S.ph[0].n) break;
      std::vector<PackSeg> segs{PackSeg{ph.Xbar + (size_t)k * ph.n * H->ldb, 1, ph.n, 0, 0, 0, 0}};
      return run_pack(H, segs, ph.n, 0, H->B, out);
    }
    k -= ph.h;
  }
  cafe::set_last_error("beyond the plan");
  return CAFE_ERR_ARG;
}

// ---- augmented-Lagrangian parameters across the solves of an MPC loop. al / out: host [B][n_phases][4][2] = (sigma, lambda) per touchdown-
//      constraint element (unused elements zero).
extern "C" int cafe_gpu_set_al_params(CafeHandle* H, const double* al, int B) {
  if (!H) { cafe::set_last_error("bad argument"); return CAFE_ERR_ARG; }
  if (!al) { H->al_B = 0; return 0; }
  if (B <= 0 || B > H->max_batch) { cafe::set_last_error("bad batch size for the augmented-Lagrangian parameter set"); return CAFE_ERR_ARG; }
  CUDA_OK(cudaSetDevice(H->device));
  const int np_ = H->S.n_phases;
  const size_t ldb = (size_t)H->ldb;
  std::vector<double> t((size_t)CAFE_MAX_PHASES * 8 * ldb, 0.0);
  for (int b = 0; b < B; ++b)
    for (int e = 0; e < np_ * 8; ++e) t[(size_t)e * ldb + b] = al[(size_t)b * np_ * 8 + e];
  CUDA_OK(cudaStreamSynchronize(H->stream));
  CUDA_OK(cudaMemcpy(H->d_al, t.data(), t.size() * sizeof(double), cudaMemcpyHostToDevice));
  H->al_B = B;
  return 0;
}

extern "C" int cafe_gpu_get_al_params(CafeHandle* H, double* out) {
  if (!H || !out || H->B <= 0) { cafe::set_last_error("bad argument"); return CAFE_ERR_ARG; }
  CUDA_OK(cudaSetDevice(H->device));
  CUDA_OK(cudaStreamSynchronize(H->stream));
  const int np_ = H->S.n_phases, B = H->B;
  const size_t ldb = (size_t)H->ldb;
  std::vector<double> sg(4 * ldb), lm(4 * ldb);
  for (int pi = 0; pi < np_; ++pi) {
    const PhaseDev& ph = H->S.ph[pi];
    CUDA_OK(cudaMemcpy(sg.data(), ph.al_sigma, 4 * ldb * sizeof(double), cudaMemcpyDeviceToHost));
    CUDA_OK(cudaMemcpy(lm.data(), ph.al_lambda, 4 * ldb * sizeof(double), cudaMemcpyDeviceToHost));
    for (int b = 0; b < B; ++b)
      for (int i = 0; i < 4; ++i) {
        const bool on = i < ph.n_td;
        out[((size_t)b * np_ + pi) * 8 + i * 2] = on ? sg[(size_t)i * ldb + b] : 0.0;
        out[((size_t)b * np_ + pi) * 8 + i * 2 + 1] = on ? lm[(size_t)i * ldb + b] : 0.0;
      }
  }
  return 0;
}

// ---- per-problem references with a shared contact schedule: every problem tracks its own records
extern "C" int cafe_gpu_set_references(CafeHandle* H, const double* refs, int B) {
  if (!H) { cafe::set_last_error("bad argument"); return CAFE_ERR_ARG; }
  CUDA_OK(cudaSetDevice(H->device));
  SolverDev& S = H->S;
  if (!refs) {  // back to the deck's shared records
    for (int i = 0; i < S.n_phases; ++i) S.ph[i].ref_pp = nullptr;
    H->ref_pp_B = 0;
    return 0;
  }
  if (B <= 0 || B > H->max_batch) { cafe::set_last_error("bad batch size for the reference set"); return CAFE_ERR_ARG; }
  const size_t nrec = (size_t)H->deck.n_records, W = CAFE_REF_W, ldb = (size_t)H->ldb;
  // the phase deck (horizons, contacts, touchdown feet) is shared: the contact flags of every record must be the deck's
  for (int b = 0; b < B; ++b)
    for (size_t r = 0; r < nrec; ++r)
      for (int f = 0; f < 4; ++f)
        if (refs[((size_t)b * nrec + r) * W + CAFE_REF_CONTACT + f] != H->ref_host[r * W + CAFE_REF_CONTACT + f]) {
          cafe::set_last_error("per-problem references must keep the deck's contact schedule");
          return CAFE_ERR_UNSUPPORTED;
        }
  std::vector<double> t(nrec * W * ldb, 0.0);  // batch-major transpose
  for (int b = 0; b < B; ++b)
    for (size_t e = 0; e < nrec * W; ++e) t[e * ldb + b] = refs[(size_t)b * nrec * W + e];
  if (t.size() > H->ref_pp_cap) {
    cudaFree(H->d_ref_pp); H->d_ref_pp = nullptr; H->ref_pp_cap = 0;
    CUDA_OK(cudaMalloc(&H->d_ref_pp, (t.size() + 8 * W * ldb) * sizeof(double)));
    H->ref_pp_cap = t.size() + 8 * W * ldb;
  }
  CUDA_OK(cudaMemcpy(H->d_ref_pp, t.data(), t.size() * sizeof(double), cudaMemcpyHostToDevice));
  for (int i = 0; i < S.n_phases; ++i) S.ph[i].ref_pp = H->d_ref_pp + (size_t)H->deck.phase[i].knot_offset * W * ldb;
  H->ref_pp_B = B;
  return 0;
}

// ---- MHPC_Command_lcmt record (lcmtypes/MHPC_Command_lcmt.lcm, filled on the host by MHPCLocomotion.cpp:236-281): emitted
//      as float32 straight from the device arrays for the first n_steps whole-body knots of the horizon
extern "C" long cafe_lcm_command_size(int n_steps) { return n_steps > 0 ? 1080L * n_steps : 0; }

extern "C" long cafe_hkd_lcm_command_size(int n_steps) { return n_steps > 0 ? 180L * n_steps : 0; }

static int lcm_impl(CafeHandle* H, int n_steps, float* out, float* dev_out, bool hkd = false) {
  if (!H || n_steps <= 0 || (!out && !dev_out)) { cafe::set_last_error("bad argument"); return CAFE_ERR_ARG; }
  CUDA_OK(cudaSetDevice(H->device));
  if (hkd) {
    // hkd_command_lcmt (lcmtypes/hkd_command_lcmt.lcm) as HKDMPCSolver::publish_mpc_cmd fills it (HKDMPC.cpp:243-290): per step the 24
    // controls, the first 12 states and the upper-left 12 x 12 block of K as [m][n]; steps run on across phase boundaries
    int knots = 0;
    for (int i = 0; i < H->S.n_phases && H->S.ph[i].model == CAFE_MODEL_HKD; ++i) knots += H->S.ph[i].h;
    if (H->S.ph[0].model != CAFE_MODEL_HKD || knots < n_steps) { cafe::set_last_error("not an HKD deck, or fewer knots than the requested command steps"); return CAFE_ERR_UNSUPPORTED; }
    const long N = n_steps, rec = 180L * N, oU = 0, oX = 24 * N, oFb = 36 * N;
    std::vector<PackSegF> segs;
    int s0 = 0;
    for (int i = 0; i < H->S.n_phases && s0 < n_steps; ++i) {
      const PhaseDev& ph = H->S.ph[i];
      const int g = std::min(ph.h, n_steps - s0);
      segs.push_back(PackSegF{ph.Ubar, 24, 0, 24, 0, g, oU + (long)s0 * 24, 0, 0, 0, 0, 0});
      segs.push_back(PackSegF{ph.Xbar, 24, 0, 12, 0, g, oX + (long)s0 * 12, 0, 0, 0, 0, 0});
      segs.push_back(PackSegF{ph.K, 576, 0, 144, 0, g, oFb + (long)s0 * 144, 0, 0, 0, 12, 24});
      s0 += g;
    }
    const size_t need = (size_t)H->B * rec * sizeof(float);
    float* dst = dev_out;
    if (!dst) {
      if (need > H->pack_bytes) {
        cudaFree(H->d_pack); H->d_pack = nullptr; H->pack_bytes = 0;
        CUDA_OK(cudaMalloc(&H->d_pack, need));
        H->pack_bytes = need;
      }
      dst = reinterpret_cast<float*>(H->d_pack);
    }
    PackSegF* d_segs = nullptr;
    if (int rc = upload_table(H, segs, &d_segs)) return rc;
    dim3 grid(592, (unsigned)segs.size());
    k_pack_lcm<<<grid, 256, 0, H->stream>>>(d_segs, (int)segs.size(), H->ldb, H->B, rec, dst);
    if (out) CUDA_OK(cudaMemcpyAsync(out, dst, need, cudaMemcpyDeviceToHost, H->stream));
    CUDA_OK(cudaStreamSynchronize(H->stream));
    CUDA_OK(cudaGetLastError());
    return 0;
  }
  int wb_knots = 0;
  for (int i = 0; i < H->S.n_phases && H->S.ph[i].model == CAFE_MODEL_WB; ++i) wb_knots += H->S.ph[i].h;
  if (wb_knots < n_steps) { cafe::set_last_error("the deck has fewer leading whole-body knots than the requested command steps"); return CAFE_ERR_UNSUPPORTED; }
  const long N = n_steps, rec = 1080L * N;
  // field bases, in the order of the LCM struct: torque eul pos qJ vWorld eulrate qJd GRF feedback Qu Quu Qux
  const long oTorque = 0, oEul = 12 * N, oPos = 15 * N, oQj = 18 * N, oVw = 30 * N, oEr = 33 * N, oQjd = 36 * N, oGrf = 48 * N, oFb = 60 * N, oQu = 492 * N,
             oQuu = 504 * N, oQux = 648 * N;
  std::vector<PackSegF> segs;
  int s0 = 0;
  for (int i = 0; i < H->S.n_phases && s0 < n_steps; ++i) {
    const PhaseDev& ph = H->S.ph[i];
    const int g = std::min(ph.h, n_steps - s0);
    auto add = [&](const double* src, int nc_src, int c0, int w, long base) { segs.push_back(PackSegF{src, nc_src, c0, w, 0, g, base + (long)s0 * w, 0, 0, 0, 0, 0}); };
    auto add_pm = [&](const double* src, int nc_src, int c0, int w, long base) { segs.push_back(PackSegF{src, nc_src, c0, w, 0, g, base + (long)s0 * w, cafe_dev::ld_mma(12), 12, ph.h, 0, 0}); };
    add(ph.Ubar, 12, 0, 12, oTorque);
    add(ph.Xbar, 36, 3, 3, oEul); add(ph.Xbar, 36, 0, 3, oPos); add(ph.Xbar, 36, 6, 12, oQj);
    add(ph.Xbar, 36, 18, 3, oVw); add(ph.Xbar, 36, 21, 3, oEr); add(ph.Xbar, 36, 24, 12, oQjd);
    add(ph.Y, 12, 0, 12, oGrf); add(ph.K, 432, 0, 432, oFb); add(ph.Qu, 12, 0, 12, oQu); add_pm(ph.Quu, 144, 0, 144, oQuu); add_pm(ph.Qux, 432, 0, 432, oQux);
    s0 += g;
  }
  const size_t need = (size_t)H->B * rec * sizeof(float);
  float* dst = dev_out;
  if (!dst) {
    if (need > H->pack_bytes) {
      cudaFree(H->d_pack); H->d_pack = nullptr; H->pack_bytes = 0;
      CUDA_OK(cudaMalloc(&H->d_pack, need));
      H->pack_bytes = need;
    }
    dst = reinterpret_cast<float*>(H->d_pack);
  }
  PackSegF* d_segs = nullptr;
  if (int rc = upload_table(H, segs, &d_segs)) return rc;
  dim3 grid(592, (unsigned)segs.size());
  k_pack_lcm<<<grid, 256, 0, H->stream>>>(d_segs, (int)segs.size(), H->ldb, H->B, rec, dst);
  if (out) CUDA_OK(cudaMemcpyAsync(out, dst, need, cudaMemcpyDeviceToHost, H->stream));
  CUDA_OK(cudaStreamSynchronize(H->stream));
  CUDA_OK(cudaGetLastError());
  return 0;
}
extern "C" int cafe_gpu_get_lcm_commands(CafeHandle* H, int n_steps, float* out) { return lcm_impl(H, n_steps, out, nullptr); }
extern "C" int cafe_gpu_get_lcm_commands_device(CafeHandle* H, int n_steps, float* out_dev) { return lcm_impl(H, n_steps, nullptr, out_dev); }
extern "C" int cafe_gpu_get_hkd_lcm_commands(CafeHandle* H, int n_steps, float* out) { return lcm_impl(H, n_steps, out, nullptr, true); }
extern "C" int cafe_gpu_get_hkd_lcm_commands_device(CafeHandle* H, int n_steps, float* out_dev) { return lcm_impl(H, n_steps, nullptr, out_dev, true); }

extern "C" long cafe_gpu_debug_get(CafeHandle* H, const char* name, int phase, int b, double* out) {
  if (!H || !name || !out || phase < 0 || phase >= H->S.n_phases || b < 0 || b >= H->B) { cafe::set_last_error("bad argument"); return CAFE_ERR_ARG; }
  if (cudaSetDevice(H->device) != cudaSuccess) return CAFE_ERR_CUDA;
  const PhaseDev& ph = H->S.ph[phase];
  const int n = ph.n, m = ph.m, p = ph.p, h = ph.h, nn = ph.n_next;
  const std::string nm(name);
  const double* src = nullptr;
  long knots = 0, nc = 0;
  auto set = [&](const double* s, long k, long c) { src = s; knots = k; nc = c; };
  if (nm == "X") set(ph.X, h + 1, n); else if (nm == "Xbar") set(ph.Xbar, h + 1, n); else if (nm == "U") set(ph.U, h, m);
  else if (nm == "Ubar") set(ph.Ubar, h, m); else if (nm == "Y") set(ph.Y, h, p); else if (nm == "Defect") set(ph.Defect, h + 1, n);
  else if (nm == "dX") set(ph.dX, h + 1, n); else if (nm == "dU") set(ph.dU, h, m); else if (nm == "G") set(ph.G, h + 1, n);
  else if (nm == "Qu") set(ph.Qu, h, m); else if (nm == "A") set(ph.A, h, n * n); else if (nm == "B") set(ph.Bm, h, n * m);
  else if (nm == "C") set(ph.C, h, p * n); else if (nm == "D") set(ph.D, h, p * m); else if (nm == "K") set(ph.K, h, m * n);
  else if (nm == "Quu") set(ph.Quu, h, m * m); else if (nm == "Qux") set(ph.Qux, h, m * n); else if (nm == "lx") set(ph.lx, h, n);
  else if (nm == "lu") set(ph.lu, h, m); else if (nm == "ly") set(ph.ly, h, p); else if (nm == "lxx") set(ph.lxx, h, n * n);
  else if (nm == "luu") set(ph.luu, h, m * m); else if (nm == "lyy") set(ph.lyy, h, p * p); else if (nm == "l") set(ph.lk, h + 1, 1);
  else if (nm == "Phix") set(ph.Phix, 1, n); else if (nm == "Phixx") set(ph.Phixx, 1, n * n); else if (nm == "Px") set(ph.Px, 1, (long)nn * n);
  else { cafe::set_last_error("unknown array name"); return CAFE_ERR_ARG; }
  if (nm == "Quu" || nm == "Qux") {
    // problem-major tiles [b][h][ld(m) x cols] written by the sweep
    const int ldm = cafe_dev::ld_mma(m), cols = nm == "Quu" ? m : n;
    std::vector<double> tl((size_t)h * ldm * cols);
    if (tl.empty()) return 0;
    if (cudaMemcpy(tl.data(), src + (size_t)b * h * ldm * cols, tl.size() * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) {
      cafe::set_last_error("cudaMemcpy failed"); return CAFE_ERR_CUDA;
    }
    for (int k = 0; k < h; ++k) for (int j = 0; j < cols; ++j) for (int i = 0; i < m; ++i) out[(size_t)k * m * cols + i + m * j] = tl[((size_t)k * cols + j) * ldm + i];
    return (long)knots * nc;
  }
  if (ph.model == CAFE_MODEL_WB && (nm == "A" || nm == "B" || nm == "C" || nm == "D")) {
    // the whole-body linearisation is stored problem-major in the sweep's tile layout (ABpm, CDpm); rows 0..17 of A are the
    // static [I, dt I] pattern kept in the batch-major array
    std::vector<double> ab((size_t)h * CAFE_WB_AB_TILE), cd((size_t)h * CAFE_WB_CD_TILE), a0;
    if (cudaMemcpy(ab.data(), ph.ABpm + (size_t)b * h * CAFE_WB_AB_TILE, ab.size() * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(cd.data(), ph.CDpm + (size_t)b * h * CAFE_WB_CD_TILE, cd.size() * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) {
      cafe::set_last_error("cudaMemcpy failed"); return CAFE_ERR_CUDA;
    }
    if (nm == "A") {
      a0.resize((size_t)h * 1296);
      if (cudaMemcpy2D(a0.data(), sizeof(double), ph.A + b, (size_t)H->ldb * sizeof(double), sizeof(double), (size_t)h * 1296, cudaMemcpyDeviceToHost) != cudaSuccess) {
        cafe::set_last_error("cudaMemcpy2D failed"); return CAFE_ERR_CUDA;
      }
    }
    for (int k = 0; k < h; ++k) {
      if (nm == "A") for (int j = 0; j < 36; ++j) for (int i = 0; i < 36; ++i)
        out[(size_t)k * 1296 + i + 36 * j] = i < 18 ? a0[(size_t)k * 1296 + i + 36 * j] : ab[(size_t)k * CAFE_WB_AB_TILE + (i - 18) + 20 * j];
      else if (nm == "B") for (int j = 0; j < 12; ++j) for (int i = 0; i < 36; ++i)
        out[(size_t)k * 432 + i + 36 * j] = i < 18 ? 0.0 : ab[(size_t)k * CAFE_WB_AB_TILE + (i - 18) + 20 * (36 + j)];
      else if (nm == "C") for (int j = 0; j < 36; ++j) for (int i = 0; i < 12; ++i) out[(size_t)k * 432 + i + 12 * j] = cd[(size_t)k * CAFE_WB_CD_TILE + i + 12 * j];
      else for (int j = 0; j < 12; ++j) for (int i = 0; i < 12; ++i) out[(size_t)k * 144 + i + 12 * j] = cd[(size_t)k * CAFE_WB_CD_TILE + i + 12 * (36 + j)];
    }
    return (long)knots * nc;
  }
  const long cnt = knots * nc;
  if (cnt == 0) return 0;
  if (cudaMemcpy2D(out, sizeof(double), src + b, (size_t)H->ldb * sizeof(double), sizeof(double), (size_t)cnt, cudaMemcpyDeviceToHost) != cudaSuccess) {
    cafe::set_last_error("cudaMemcpy2D failed"); return CAFE_ERR_CUDA;
  }
  return cnt;
}

extern "C" int cafe_gpu_measure_fp64_peak(int device, double* tflops) {
  if (!tflops) return CAFE_ERR_ARG;
  CUDA_OK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, device));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 16;
  double* d = nullptr;
  CUDA_OK(cudaMalloc(&d, (size_t)blocks * threads * sizeof(double)));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_fp64_peak<<<blocks, threads>>>(d, 1024);
  double best = 0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    k_fp64_peak<<<blocks, threads>>>(d, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fl = 2.0 * 8.0 * iters * (double)blocks * threads;
    const double tf = fl / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  // DMMA: the pipe the sweep GEMMs run on; the roofline denominator is the larger of the two
  k_fp64_mma_peak<<<blocks, threads>>>(d, 256);
  for (int rep = 0; rep < 5; ++rep) {
    const int it2 = 1 << 13;
    cudaEventRecord(e0);
    k_fp64_mma_peak<<<blocks, threads>>>(d, it2);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fl = 2.0 * 256.0 * 8.0 * it2 * (double)blocks * (threads / 32);
    const double tf = fl / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d);
  CUDA_OK(cudaGetLastError());
  *tflops = best;
  return 0;
}

// =====================================================================================================================
// Multi-GPU: problems are independent, so GPU g solves the contiguous slice [g ceil(B/G), min(B, (g+1) ceil(B/G)))
// with the deck replicated and NO traffic during the solve; the only collective is the final gather of the packed command
// records to the first GPU (NCCL send / recv in one group over NVLink / NVSwitch). Two ways in:
//   * one process, G GPUs        cafe_gpu_create_multi ... (ncclCommInitAll, one host thread per GPU during the solve)
//   * one process per GPU        cafe_gpu_nccl_unique_id + cafe_gpu_comm_init_rank + cafe_gpu_gather_commands (torchrun, MPI)
// NCCL is opened at run time (dlopen): the library loads, and single-GPU use works, on a box without it.
#include <dlfcn.h>
#include <thread>
#include <nccl.h>

namespace {
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    // CAFE_NCCL_LIB names a specific library; else whatever "libnccl.so.2" resolves to (a copy the process has already loaded, e.g.
    // the one bundled with PyTorch, wins: two NCCL builds with the same SONAME cannot coexist in one process)
    if (const char* e = getenv("CAFE_NCCL_LIB")) api.lib = dlopen(e, RTLD_NOW | RTLD_LOCAL);
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) { if (api.lib) break; api.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL); }
    if (api.lib) {
      bool ok = true;
      auto sym = [&](const char* n) { void* p = dlsym(api.lib, n); if (!p) ok = false; return p; };
      api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
      api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
      api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
      api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
      api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
      api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
      api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
      api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
      api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
      if (!ok) { dlclose(api.lib); api.lib = nullptr; }
    }
  }
  return api.lib ? &api : nullptr;
}
#define NCCL_OK(call)                                                                                      \
  do {                                                                                                     \
    ncclResult_t r_ = (call);                                                                              \
    if (r_ != ncclSuccess) { cafe::set_last_error(std::string(#call) + ": " + N->GetErrorString(r_)); return CAFE_ERR_CUDA; } \
  } while (0)
}  // namespace

struct CafeComm { ncclComm_t comm = nullptr; int nranks = 0, rank = 0; double* d_send = nullptr; size_t send_bytes = 0; };

extern "C" int cafe_gpu_shard_range(int B, int nranks, int rank, int* lo, int* hi) {
  if (B < 0 || nranks <= 0 || rank < 0 || rank >= nranks || !lo || !hi) { cafe::set_last_error("bad argument"); return CAFE_ERR_ARG; }
  const int per = (B + nranks - 1) / nranks;
  *lo = std::min(B, rank * per); *hi = std::min(B, (rank + 1) * per);
  return 0;
}

extern "C" int cafe_gpu_nccl_unique_id(char id[128]) {
  NcclApi* N = nccl_api();
  if (!N) { cafe::set_last_error("NCCL (libnccl.so.2) is not available"); return CAFE_ERR_UNSUPPORTED; }
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId u;
  NCCL_OK(N->GetUniqueId(&u));
  std::memcpy(id, &u, 128);
  return 0;
}

// the handle's communicator lives beside it (the handle struct itself stays what the single-GPU path needs)
#include <map>
#include <mutex>
namespace { std::map<CafeHandle*, CafeComm> g_comms; std::mutex g_comm_mu; }

extern "C" int cafe_gpu_comm_init_rank(CafeHandle* H, int nranks, int rank, const char id[128]) {
  NcclApi* N = nccl_api();
  if (!N) { cafe::set_last_error("NCCL (libnccl.so.2) is not available"); return CAFE_ERR_UNSUPPORTED; }
  if (!H || nranks <= 0 || rank < 0 || rank >= nranks || !id) { cafe::set_last_error("bad argument"); return CAFE_ERR_ARG; }
  CUDA_OK(cudaSetDevice(H->device));
  ncclUniqueId u;
  std::memcpy(&u, id, 128);
  CafeComm c; c.nranks = nranks; c.rank = rank;
  NCCL_OK(N->CommInitRank(&c.comm, nranks, u, rank));
  std::lock_guard<std::mutex> lk(g_comm_mu);
  g_comms[H] = c;
  return 0;
}

extern "C" int cafe_gpu_comm_destroy(CafeHandle* H) {
  std::lock_guard<std::mutex> lk(g_comm_mu);
  auto it = g_comms.find(H);
  if (it == g_comms.end()) return 0;
  NcclApi* N = nccl_api();
  if (N && it->second.comm) N->CommDestroy(it->second.comm);
  if (it->second.d_send) { cudaSetDevice(H->device); cudaFree(it->second.d_send); }
  g_comms.erase(it);
  return 0;
}

// Packs this rank's command records and gathers the ranks' slices to rank 0: rank r's records land at out_dev[r * per_rank ...] (device
// buffer of rank 0, [nranks * per_rank][cafe_command_size] doubles; ignored on the other ranks). Ragged tails are fine (a rank sends its
// own H->B records). Stream-ordered on the handle's stream; returns after the transfer has completed.
extern "C" int cafe_gpu_gather_commands(CafeHandle* H, int n_gain_knots, int per_rank, double* out_dev) {
  NcclApi* N = nccl_api();
  if (!N) { cafe::set_last_error("NCCL (libnccl.so.2) is not available"); return CAFE_ERR_UNSUPPORTED; }
  CafeComm c;
  { std::lock_guard<std::mutex> lk(g_comm_mu); auto it = g_comms.find(H); if (it == g_comms.end()) { cafe::set_last_error("no communicator: call cafe_gpu_comm_init_rank first"); return CAFE_ERR_ARG; } c = it->second; }
  if (per_rank < H->B || (c.rank == 0 && !out_dev)) { cafe::set_last_error("bad argument"); return CAFE_ERR_ARG; }
  CUDA_OK(cudaSetDevice(H->device));
  const size_t rec = (size_t)cafe_command_size(&H->deck, n_gain_knots);
  double* src = nullptr;
  if (c.rank == 0) src = out_dev;   // rank 0 packs straight into its slot of the gathered buffer
  else {
    const size_t need = (size_t)per_rank * rec * sizeof(double);   // a full slot: a ragged last shard is zero-padded
    if (need > c.send_bytes) {
      cudaFree(c.d_send); c.d_send = nullptr;
      CUDA_OK(cudaMalloc(&c.d_send, need)); c.send_bytes = need;
      std::lock_guard<std::mutex> lk(g_comm_mu); g_comms[H] = c;
    }
    src = c.d_send;
    if (H->B < per_rank) CUDA_OK(cudaMemsetAsync(src + (size_t)H->B * rec, 0, (size_t)(per_rank - H->B) * rec * sizeof(double), H->stream));
  }
  int rc = cafe_gpu_get_commands_device(H, n_gain_knots, src);
  if (rc) return rc;
  // equal slots of per_rank records: rank 0 need not know the other ranks' (possibly ragged) counts
  NCCL_OK(N->GroupStart());
  if (c.rank == 0) {
    for (int r = 1; r < c.nranks; ++r) NCCL_OK(N->Recv(out_dev + (size_t)r * per_rank * rec, (size_t)per_rank * rec, ncclDouble, r, c.comm, H->stream));
  } else {
    NCCL_OK(N->Send(src, (size_t)per_rank * rec, ncclDouble, 0, c.comm, H->stream));
  }
  NCCL_OK(N->GroupEnd());
  CUDA_OK(cudaStreamSynchronize(H->stream));
  return 0;
}

// The same gather without blocking the solver: this rank's records are packed on the solver's stream, the NCCL send / recv and - on rank 0 -
// the copy of ALL gathered records to `out_host` (page-locked, [nranks * per_rank][cafe_command_size]) run on the copy stream while the next solve
// proceeds. out_dev: rank 0's device buffer of the slot (the other ranks pass NULL). cafe_gpu_commands_wait(h, slot) waits for it.
extern "C" int cafe_gpu_gather_commands_async(CafeHandle* H, int n_gain_knots, int per_rank, double* out_dev, double* out_host, int slot) {
  NcclApi* N = nccl_api();
  if (!N) { cafe::set_last_error("NCCL (libnccl.so.2) is not available"); return CAFE_ERR_UNSUPPORTED; }
  CafeComm c;
  { std::lock_guard<std::mutex> lk(g_comm_mu); auto it = g_comms.find(H); if (it == g_comms.end()) { cafe::set_last_error("no communicator: call cafe_gpu_comm_init_rank first"); return CAFE_ERR_ARG; } c = it->second; }
  if (slot < 0 || slot > 1 || per_rank < H->B || (c.rank == 0 && (!out_dev || !out_host))) { cafe::set_last_error("bad argument"); return CAFE_ERR_ARG; }
  CUDA_OK(cudaSetDevice(H->device));
  std::vector<PackSeg> segs;
  const size_t rec = (size_t)command_segments(H, n_gain_knots, segs);
  double* src = out_dev;   // rank 0 packs straight into its slot of the gathered buffer
  if (c.rank != 0) {
    const size_t need = (size_t)per_rank * rec * sizeof(double);
    if (need > H->pack_async_bytes[slot]) {
      if (H->async_used[slot]) CUDA_OK(cudaEventSynchronize(H->ev_landed[slot]));
      cudaFree(H->d_pack_async[slot]); H->d_pack_async[slot] = nullptr; H->pack_async_bytes[slot] = 0;
      CUDA_OK(cudaMalloc(&H->d_pack_async[slot], need));
      H->pack_async_bytes[slot] = need;
      CUDA_OK(cudaMemset(H->d_pack_async[slot], 0, need));   // a ragged last shard stays zero-padded
    }
    src = H->d_pack_async[slot];
  }
  if (int rc = run_pack_nowait(H, segs, (long)rec, H->B, src, slot)) return rc;
  NCCL_OK(N->GroupStart());
  if (c.rank == 0) {
    for (int r = 1; r < c.nranks; ++r) NCCL_OK(N->Recv(out_dev + (size_t)r * per_rank * rec, (size_t)per_rank * rec, ncclDouble, r, c.comm, H->stream_copy));
  } else {
    NCCL_OK(N->Send(src, (size_t)per_rank * rec, ncclDouble, 0, c.comm, H->stream_copy));
  }
  NCCL_OK(N->GroupEnd());
  if (c.rank == 0) CUDA_OK(cudaMemcpyAsync(out_host, out_dev, (size_t)c.nranks * per_rank * rec * sizeof(double), cudaMemcpyDeviceToHost, H->stream_copy));
  CUDA_OK(cudaEventRecord(H->ev_landed[slot], H->stream_copy));
  return 0;
}

// ---- one process, several GPUs
struct CafeMulti {
  int ndev = 0, max_batch = 0, per = 0, B = 0;
  std::vector<CafeHandle*> h;
  std::vector<ncclComm_t> comms;
  std::vector<double*> d_send;     // per GPU > 0: packed records of its slice
  std::vector<size_t> send_bytes;
  double* d_all = nullptr; size_t all_bytes = 0;   // on GPU 0: the gathered records
  std::vector<int> lo, hi;
};

extern "C" int cafe_gpu_multi_destroy(CafeMulti* M) {
  if (!M) return 0;
  NcclApi* N = M->ndev > 1 ? nccl_api() : nullptr;
  for (size_t g = 0; g < M->h.size(); ++g) {
    if (M->h[g]) { cudaSetDevice(M->h[g]->device); if (g < M->d_send.size()) cudaFree(M->d_send[g]); if (g == 0) cudaFree(M->d_all); }
    if (N && g < M->comms.size() && M->comms[g]) N->CommDestroy(M->comms[g]);
    cafe_gpu_destroy(M->h[g]);
  }
  delete M;
  return 0;
}

extern "C" int cafe_gpu_create_multi(const CafeDeck* deck, int ndev, const int* devices, int max_batch, CafeMulti** out) {
  if (!deck || !out || ndev <= 0 || max_batch <= 0) { cafe::set_last_error("bad argument"); return CAFE_ERR_ARG; }
  NcclApi* N = ndev > 1 ? nccl_api() : nullptr;   // opened only when a collective is needed
  if (ndev > 1 && !N) { cafe::set_last_error("NCCL (libnccl.so.2) is not available"); return CAFE_ERR_UNSUPPORTED; }
  CafeMulti* M = new CafeMulti();
  M->ndev = ndev; M->max_batch = max_batch; M->per = (max_batch + ndev - 1) / ndev;
  M->h.assign(ndev, nullptr); M->comms.assign(ndev, nullptr); M->d_send.assign(ndev, nullptr); M->send_bytes.assign(ndev, 0);
  M->lo.assign(ndev, 0); M->hi.assign(ndev, 0);
  std::vector<int> devs(ndev);
  for (int g = 0; g < ndev; ++g) devs[g] = devices ? devices[g] : g;
  for (int g = 0; g < ndev; ++g) {
    int rc = cafe_gpu_create(deck, devs[g], M->per, &M->h[g]);
    if (rc) { cafe_gpu_multi_destroy(M); return rc; }
  }
  if (ndev > 1) {
    ncclResult_t r = N->CommInitAll(M->comms.data(), ndev, devs.data());
    if (r != ncclSuccess) { cafe::set_last_error(std::string("ncclCommInitAll: ") + N->GetErrorString(r)); cafe_gpu_multi_destroy(M); return CAFE_ERR_CUDA; }
  }
  *out = M;
  return 0;
}

extern "C" int cafe_gpu_multi_ndev(const CafeMulti* M) { return M ? M->ndev : 0; }
extern "C" CafeHandle* cafe_gpu_multi_handle(CafeMulti* M, int g) { return (M && g >= 0 && g < M->ndev) ? M->h[g] : nullptr; }

// x0: host [B][n0]. GPU g solves rows [lo_g, hi_g); one host thread per GPU (the tick loop polls one counter per tick and GPU)
extern "C" int cafe_gpu_multi_solve_batch(CafeMulti* M, const double* x0, int B, const CafeOptions* opt) {
  if (!M || !x0 || !opt || B <= 0 || B > M->max_batch) { cafe::set_last_error("bad argument"); return CAFE_ERR_ARG; }
  M->B = B;
  const int n0 = M->h[0]->S.ph[0].n;
  const int per = (B + M->ndev - 1) / M->ndev;
  std::vector<int> rc(M->ndev, 0);
  std::vector<std::thread> th;
  for (int g = 0; g < M->ndev; ++g) {
    M->lo[g] = std::min(B, g * per); M->hi[g] = std::min(B, (g + 1) * per);
    const int nb = M->hi[g] - M->lo[g];
    if (nb <= 0) { M->h[g]->B = 0; continue; }
    th.emplace_back([&, g, nb] { rc[g] = cafe_gpu_solve_batch(M->h[g], x0 + (size_t)M->lo[g] * n0, nb, opt); });
  }
  for (auto& t : th) t.join();
  for (int g = 0; g < M->ndev; ++g) if (rc[g]) return rc[g];
  return 0;
}

// the MPC update on every GPU's solver (cafe_gpu_update_deck): each keeps its slice of the previous solve as the warm start. B = the batch of
// the previous cafe_gpu_multi_solve_batch (the slices must not move), or 0 for a cold start on the new deck.
extern "C" int cafe_gpu_multi_update_deck(CafeMulti* M, const CafeDeck* deck, int k_advance, int B) {
  if (!M || !deck || (B != 0 && B != M->B)) { cafe::set_last_error("bad argument: B must be the previous batch size or 0"); return CAFE_ERR_ARG; }
  for (int g = 0; g < M->ndev; ++g) {
    const int nb = B > 0 ? M->hi[g] - M->lo[g] : 0;
    int rc = cafe_gpu_update_deck(M->h[g], deck, k_advance, nb > 0 ? nb : 0);
    if (rc) return rc;
  }
  return 0;
}

extern "C" int cafe_gpu_multi_get_info(CafeMulti* M, CafeInfo* info) {
  if (!M || !info) { cafe::set_last_error("bad argument"); return CAFE_ERR_ARG; }
  for (int g = 0; g < M->ndev; ++g) {
    if (M->hi[g] <= M->lo[g]) continue;
    int rc = cafe_gpu_get_info(M->h[g], info + M->lo[g]);
    if (rc) return rc;
  }
  return 0;
}

// packs every GPU's command records, gathers them on the first GPU (one NCCL group of sends / receives) and copies the B records to
// the host: cmd = host [B][cafe_command_size(deck, n_gain_knots)]
extern "C" int cafe_gpu_multi_get_commands(CafeMulti* M, int n_gain_knots, double* cmd) {
  if (!M || !cmd || n_gain_knots < 0 || M->B <= 0) { cafe::set_last_error("bad argument"); return CAFE_ERR_ARG; }
  NcclApi* N = M->ndev > 1 ? nccl_api() : nullptr;
  const size_t rec = (size_t)cafe_command_size(&M->h[0]->deck, n_gain_knots);
  CafeHandle* H0 = M->h[0];
  CUDA_OK(cudaSetDevice(H0->device));
  const size_t need = (size_t)M->B * rec * sizeof(double);
  if (need > M->all_bytes) { cudaFree(M->d_all); M->d_all = nullptr; CUDA_OK(cudaMalloc(&M->d_all, need)); M->all_bytes = need; }
  for (int g = 0; g < M->ndev; ++g) {
    const int nb = M->hi[g] - M->lo[g];
    if (nb <= 0) continue;
    CafeHandle* H = M->h[g];
    CUDA_OK(cudaSetDevice(H->device));
    double* dst = M->d_all + (size_t)M->lo[g] * rec;
    if (g > 0) {
      const size_t nbytes = (size_t)nb * rec * sizeof(double);
      if (nbytes > M->send_bytes[g]) { cudaFree(M->d_send[g]); M->d_send[g] = nullptr; CUDA_OK(cudaMalloc(&M->d_send[g], nbytes)); M->send_bytes[g] = nbytes; }
      dst = M->d_send[g];
    }
    int rc = cafe_gpu_get_commands_device(H, n_gain_knots, dst);
    if (rc) return rc;
  }
  if (M->ndev > 1) {
    NCCL_OK(N->GroupStart());
    for (int g = 1; g < M->ndev; ++g) {
      const int nb = M->hi[g] - M->lo[g];
      if (nb <= 0) continue;
      NCCL_OK(N->Send(M->d_send[g], (size_t)nb * rec, ncclDouble, 0, M->comms[g], M->h[g]->stream));
      NCCL_OK(N->Recv(M->d_all + (size_t)M->lo[g] * rec, (size_t)nb * rec, ncclDouble, g, M->comms[0], H0->stream));
    }
    NCCL_OK(N->GroupEnd());
  }
  CUDA_OK(cudaSetDevice(H0->device));
  CUDA_OK(cudaMemcpyAsync(cmd, M->d_all, need, cudaMemcpyDeviceToHost, H0->stream));
  CUDA_OK(cudaStreamSynchronize(H0->stream));
  for (int g = 1; g < M->ndev; ++g) { if (M->hi[g] > M->lo[g]) { CUDA_OK(cudaSetDevice(M->h[g]->device)); CUDA_OK(cudaStreamSynchronize(M->h[g]->stream)); } }
  return 0;
}
