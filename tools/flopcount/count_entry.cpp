/*
 * count_entry.cpp — MEASUREMENT TOOL. Compiled with `-include counted.hpp` together with the unchanged oracle/ sources
 * (tools/count_flops.py): runs the first DDP iteration of one problem stage by stage and phase by phase, the way
 * oracle::Solver::solve chains the phases (hsddp_oracle.cpp: hybrid_rollout, compute_cost, LQ_approximation,
 * backward_sweep, linear_rollout), and reports the operations the instrumented scalar tallied in each (phase, stage).
 * The reference's CasADi functions (separately compiled C in oracle/_ref) are interposed here: every call is tallied
 * per function and forwarded.
 */
#include "../../oracle/hsddp_oracle.hpp"
#include "../../oracle/casadi_ref.hpp"
#include <dlfcn.h>

using namespace oracle;
using flopcount::g_bucket;
using flopcount::g_cnt;

enum { ST_ROLL = 0, ST_COST, ST_LQ, ST_BWD, ST_LIN, N_STAGE };

static const char* kCasadiNames[] = {"hkinodyn", "hkinodyn_par", "compute_foot_position", "comp_foot_jacob_1", "comp_foot_jacob_2",
                                     "comp_foot_jacob_3", "comp_foot_jacob_4", "SRBDynamics", "SRBDynamicsDerivatives",
                                     "footVelPartialDq", "footAccPartialDq", "footAccPartialDv", "footForcePartialDq"};
constexpr int N_CASADI = sizeof(kCasadiNames) / sizeof(kCasadiNames[0]);
static long long g_calls[flopcount::NBUCKET][N_CASADI];
static void* g_ref = nullptr;

static casadi_fn_t real_fn(int id) {
  static casadi_fn_t cache[N_CASADI];
  if (!cache[id]) {
    if (!g_ref) throw std::runtime_error("flopcount: cafe_count_set_casadi_lib not called");
    cache[id] = (casadi_fn_t)dlsym(g_ref, kCasadiNames[id]);
    if (!cache[id]) throw std::runtime_error(std::string("flopcount: no symbol ") + kCasadiNames[id]);
  }
  return cache[id];
}
#define INTERPOSE(id, name) \
  extern "C" int name(const double** arg, double** res, casadi_int* iw, double* w, int mem) { ++g_calls[g_bucket][id]; return real_fn(id)(arg, res, iw, w, mem); }
INTERPOSE(0, hkinodyn)
INTERPOSE(1, hkinodyn_par)
INTERPOSE(2, compute_foot_position)
INTERPOSE(3, comp_foot_jacob_1)
INTERPOSE(4, comp_foot_jacob_2)
INTERPOSE(5, comp_foot_jacob_3)
INTERPOSE(6, comp_foot_jacob_4)
INTERPOSE(7, SRBDynamics)
INTERPOSE(8, SRBDynamicsDerivatives)
INTERPOSE(9, footVelPartialDq)
INTERPOSE(10, footAccPartialDq)
INTERPOSE(11, footAccPartialDv)
INTERPOSE(12, footForcePartialDq)

extern "C" int cafe_count_set_casadi_lib(const char* path) {
  g_ref = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  return g_ref ? 0 : -1;
}
extern "C" int cafe_count_n_stage() { return N_STAGE; }
extern "C" int cafe_count_n_kind() { return flopcount::NKIND; }
extern "C" int cafe_count_n_casadi() { return N_CASADI; }
extern "C" const char* cafe_count_casadi_name(int i) { return kCasadiNames[i]; }

/* ops[phase][stage][kind], calls[phase][stage][casadi function]; the last bucket row (phase = n_phases) holds what ran outside the
 * stage calls (set-up). Returns the number of phases, -1 on failure. */
extern "C" int cafe_count_first_iteration(const CafeDeck* deck, const CafeOptions* opt, const double* x0, long long* ops, long long* calls) {
  try {
    const CafeOptions& o = *opt;
    const int np = deck->n_phases;
    if ((np + 1) * N_STAGE > flopcount::NBUCKET) return -1;
    std::memset(g_cnt, 0, sizeof(g_cnt));
    std::memset(g_calls, 0, sizeof(g_calls));
    g_bucket = np * N_STAGE;
    Solver S;
    S.setup(deck);
    S.x0.assign(x0, x0 + S.phases[0]->n);
    auto at = [&](int i, int st) { g_bucket = i * N_STAGE + st; };
    /* MultiPhaseDDP::hybrid_rollout (Solver::hybrid_rollout): reset map of the previous phase belongs to that phase's rollout */
    Vec xinit = S.x0;
    for (int i = 0; i < np; ++i) {
      if (i > 0) { at(i - 1, ST_ROLL); xinit = S.phases[i - 1]->resetmap(S.phases[i - 1]->X.back()); }
      at(i, ST_ROLL);
      S.phases[i]->x_init = xinit;
      if (!S.phases[i]->hybrid_rollout(0.0, o.MS != 0)) return -1;
    }
    for (int i = 0; i < np; ++i) { at(i, ST_ROLL); S.phases[i]->update_nominal(); }
    for (int i = 0; i < np; ++i) { at(i, ST_COST); S.phases[i]->compute_cost(o); }
    for (int i = 0; i < np; ++i) { at(i, ST_LQ); S.phases[i]->LQ_approximation(o); }
    /* Solver::backward_sweep at the first regularisation (0): impact-aware hand-over counted with the phase that receives it */
    for (int i = np - 1; i >= 0; --i) {
      at(i, ST_BWD);
      const int xs = S.phases[i]->n;
      Vec Gp = zeros(xs);
      Mat Hp(xs, xs);
      if (i <= np - 2) {
        Mat Px = S.phases[i]->resetmap_partial(S.phases[i]->X.back());
        Gp = mtv(Px, S.phases[i + 1]->G[0]);
        Hp = mtm(Px, mm(S.phases[i + 1]->H[0], Px));
      }
      if (!S.phases[i]->backward_sweep(0.0, Gp, Hp)) return -2;
    }
    /* Solver::linear_rollout(1) */
    Vec dx_init = zeros(S.phases[0]->n);
    for (int i = 0; i < np; ++i) {
      at(i, ST_LIN);
      if (i > 0) {
        Mat Px = S.phases[i - 1]->resetmap_partial(S.phases[i - 1]->X.back());
        dx_init = mv(Px, S.phases[i - 1]->dX.back());
      }
      S.phases[i]->dx_init = dx_init;
      S.phases[i]->linear_rollout(1.0);
    }
    g_bucket = np * N_STAGE;
    std::memcpy(ops, g_cnt, sizeof(long long) * (size_t)(np + 1) * N_STAGE * flopcount::NKIND);
    for (int b = 0; b < (np + 1) * N_STAGE; ++b) for (int c = 0; c < N_CASADI; ++c) calls[(size_t)b * N_CASADI + c] = g_calls[b][c];
    return np;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "cafe_count_first_iteration: %s\n", e.what());
    return -1;
  }
}

/* One whole solve (oracle::Solver::solve, the loop of MultiPhaseDDP::solve) under the counters: ops[kind], calls[casadi function], counters = iterations,
 * line-search trials, regularisation steps. Returns 0, -1 on failure. */
extern "C" int cafe_count_solve(const CafeDeck* deck, const CafeOptions* opt, const double* x0, long long* ops, long long* calls, int* counters) {
  try {
    std::memset(g_cnt, 0, sizeof(g_cnt));
    std::memset(g_calls, 0, sizeof(g_calls));
    g_bucket = 1;
    Solver S;
    S.setup(deck);
    S.x0.assign(x0, x0 + S.phases[0]->n);
    g_bucket = 0;
    S.solve(*opt);
    g_bucket = 1;
    for (int k = 0; k < flopcount::NKIND; ++k) ops[k] = g_cnt[0][k];
    for (int c = 0; c < N_CASADI; ++c) calls[c] = g_calls[0][c];
    counters[0] = S.iter_; counters[1] = S.ls_iter_total_; counters[2] = S.reg_iter_total_;
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "cafe_count_solve: %s\n", e.what());
    return -1;
  }
}
