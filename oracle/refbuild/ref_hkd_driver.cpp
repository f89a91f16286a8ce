// ref_hkd_driver.cpp — TEST INFRASTRUCTURE ONLY (oracle/). Drives the REFERENCE's own HS-DDP solver on the HKD trot problem.
//
// What is linked here is the reference's code, compiled unchanged from /root/reference by oracle/refbuild/Makefile:
//   HSDDPSolver/source/{MultiPhaseDDP,SinglePhase,SinglePhaseInterface,TrajectoryManagement}.cpp, HSDDPSolver/common/HSDDP_Utils.cpp,
//   HKDMPC/HKD-TrajOpt/{HKDProblem,HKDCost,HKDConstraints,HKDReference}.cpp, Reference/QuadReference.cpp, common/casadi_interface.cpp
//   and the CasADi-generated C of the HKD model (oracle/_ref/hkd_*.o) — against the Eigen / Boost.PropertyTree / LCM stand-ins in
//   oracle/refbuild/shim (none of the three libraries exists in this image; see shim/eigen3/Eigen/cafe_eigen_shim.hpp for what
//   that means for rounding).
// This file only does what HKDMPCSolver<T>::initialize / update do around the solver (HKDMPC/HKDMPC.cpp:19-86, :89-160) minus LCM:
//   load the reference, build the HKDProblem, set x0 = compute_hkd_state(body, qJ), MultiPhaseDDP::solve; then n MPC updates
//   (HKDProblem::update, caps 2 x 1, x0 = the plan's own prediction + a caller-given nudge) — and writes down everything the
//   solver decided. The per-iteration record is taken in full precision by a decorator around every phase (SpyPhase forwards the
//   whole SinglePhaseBase interface to the reference's SinglePhase): the solver's own history buffers are float.
//
// usage (cwd must be a directory D with D/../HKDMPC/settings/{ddp_setting,constraint_params}.info, e.g. data/_run):
//   ref_hkd <quad_reference.csv> <in.txt> <out.txt>
//   in.txt : n_problems n_updates, then per problem 24 numbers (body 12, qJ 12) and n_updates x 3 numbers (position nudges)
#include <cstdio>
#include <fstream>
#include <memory>
#include <vector>

#include "HKDProblem.h"
#include "HSDDP_CompoundTypes.h"
#include "MultiPhaseDDP.h"

typedef double T;

// built without OpenMP (see Makefile): SinglePhase::LQ_approximation asks for its thread id inside a one-thread region
extern "C" int omp_get_thread_num(void) { return 0; }

struct Event { int type, phase; double a, b; };
// event types
enum { EV_ROLLOUT = 1, EV_COST = 2, EV_FEAS = 3, EV_LQ = 4, EV_BWD = 5, EV_LIN = 6, EV_ACCEPT = 7, EV_AL = 8, EV_REB = 9, EV_TCON = 10, EV_PCON = 11 };
static std::vector<Event> g_log;

class SpyPhase : public SinglePhaseBase<T> {
  std::shared_ptr<SinglePhaseBase<T>> p;
  int id;
 public:
  SpyPhase(std::shared_ptr<SinglePhaseBase<T>> p_, int id_) : p(p_), id(id_) {}
  void warmstart() override { p->warmstart(); }
  void initialization() override { p->initialization(); }
  void set_initial_condition(DVec<T>& x) override { p->set_initial_condition(x); }
  void set_initial_condition(DVec<T>& x, DVec<T>& xs) override { p->set_initial_condition(x, xs); }
  void set_initial_condition_dx(DVec<T>& dx) override { p->set_initial_condition_dx(dx); }
  void set_nominal_initial_condition(DVec<T>& x) override { p->set_nominal_initial_condition(x); }
  void linear_rollout(T eps, HSDDP_OPTION& o) override { p->linear_rollout(eps, o); T a, b; p->get_exp_cost_change(a, b); g_log.push_back({EV_LIN, id, a, b}); }
  bool hybrid_rollout(T eps, HSDDP_OPTION& o, bool last = false) override {
    const bool ok = p->hybrid_rollout(eps, o, last);
    g_log.push_back({EV_ROLLOUT, id, eps, ok ? 1.0 : 0.0});
    return ok;
  }
  void LQ_approximation(HSDDP_OPTION& o) override { p->LQ_approximation(o); g_log.push_back({EV_LQ, id, 0, 0}); }
  bool backward_sweep(T reg, DVec<T> G, DMat<T> H) override {
    const bool ok = p->backward_sweep(reg, G, H);
    g_log.push_back({EV_BWD, id, reg, ok ? 1.0 : 0.0});
    return ok;
  }
  DVec<T> resetmap(DVec<T>& x) override { return p->resetmap(x); }
  void resetmap_partial(DMat<T>& Px, DVec<T>& x) override { p->resetmap_partial(Px, x); }
  void get_value_approx(DVec<T>& G, DMat<T>& H) override { p->get_value_approx(G, H); }
  void get_exp_cost_change(T& a, T& b) override { p->get_exp_cost_change(a, b); }
  void get_terminal_state(DVec<T>& x) override { p->get_terminal_state(x); }
  void get_terminal_state(DVec<T>& x, DVec<T>& xs) override { p->get_terminal_state(x, xs); }
  void get_terminal_state_dx(DVec<T>& dx) override { p->get_terminal_state_dx(dx); }
  T get_actual_cost() override { const T c = p->get_actual_cost(); g_log.push_back({EV_COST, id, c, 0}); return c; }
  T get_max_tconstrs() override { const T c = p->get_max_tconstrs(); g_log.push_back({EV_TCON, id, c, 0}); return c; }
  T get_max_pconstrs() override { const T c = p->get_max_pconstrs(); g_log.push_back({EV_PCON, id, c, 0}); return c; }
  size_t get_state_dim() override { return p->get_state_dim(); }
  size_t get_control_dim() override { return p->get_control_dim(); }
  void update_AL_params(HSDDP_OPTION& o) override { p->update_AL_params(o); g_log.push_back({EV_AL, id, 0, 0}); }
  void update_REB_params(HSDDP_OPTION& o) override { p->update_REB_params(o); g_log.push_back({EV_REB, id, 0, 0}); }
  void update_nominal_trajectory() override { p->update_nominal_trajectory(); g_log.push_back({EV_ACCEPT, id, 0, 0}); }
  void empty_control() override { p->empty_control(); }
  void push_back_default() override { p->push_back_default(); }
  void pop_front() override { p->pop_front(); }
  void reset_params() override { p->reset_params(); }
  T measure_dynamics_feasibility(int norm_id) override { const T f = p->measure_dynamics_feasibility(norm_id); g_log.push_back({EV_FEAS, id, f, 0}); return f; }
  void update_SS_config(int n) override { p->update_SS_config(n); }
  void compute_cost(const HSDDP_OPTION& o) override { p->compute_cost(o); }
  void get_trajectory(std::vector<std::vector<float>>& x, std::vector<std::vector<float>>& u) override { p->get_trajectory(x, u); }
  void print() override { p->print(); }
};

template <class V> static void put_vec(FILE* f, const V& v) { for (int i = 0; i < (int)v.size(); ++i) fprintf(f, " %.17g", (double)v[i]); }
template <class M> static void put_mat(FILE* f, const M& m) { for (int j = 0; j < (int)m.cols(); ++j) for (int i = 0; i < (int)m.rows(); ++i) fprintf(f, " %.17g", (double)m(i, j)); }   // column-major

static void dump_problem(FILE* f, HKDProblemData<T>& pd, const char* tag) {
  fprintf(f, "%s n_phases %d\n", tag, pd.n_phases);
  for (int i = 0; i < pd.n_phases; ++i) {
    auto& tr = *pd.trajectory_ptrs[i];
    const int h = tr.horizon;
    fprintf(f, "phase %d horizon %d contact %d %d %d %d start %.9g end %.9g\n", i, h, pd.phase_contacts[i][0], pd.phase_contacts[i][1], pd.phase_contacts[i][2],
            pd.phase_contacts[i][3], (double)pd.phase_start_times[i], (double)pd.phase_end_times[i]);
    fprintf(f, "Xbar"); for (int k = 0; k <= h; ++k) put_vec(f, tr.Xbar[k]); fprintf(f, "\n");
    fprintf(f, "Ubar"); for (int k = 0; k < h; ++k) put_vec(f, tr.Ubar[k]); fprintf(f, "\n");
    fprintf(f, "K"); for (int k = 0; k < h; ++k) put_mat(f, tr.K[k]); fprintf(f, "\n");
    fprintf(f, "dU"); for (int k = 0; k < h; ++k) put_vec(f, tr.dU[k]); fprintf(f, "\n");
    fprintf(f, "G"); for (int k = 0; k <= h; ++k) put_vec(f, tr.G[k]); fprintf(f, "\n");
    fprintf(f, "Qu"); for (int k = 0; k < h; ++k) put_vec(f, tr.Qu[k]); fprintf(f, "\n");
    fprintf(f, "Quu"); for (int k = 0; k < h; ++k) put_mat(f, tr.Quu[k]); fprintf(f, "\n");
    fprintf(f, "Qux"); for (int k = 0; k < h; ++k) put_mat(f, tr.Qux[k]); fprintf(f, "\n");
    fprintf(f, "Defect"); for (int k = 0; k <= h; ++k) put_vec(f, tr.Defect_bar[k]); fprintf(f, "\n");
  }
}

static void run_solve(FILE* f, HKDProblemData<T>& pd, HSDDP_OPTION& opt, const VecM<T, 24>& xinit) {
  // what HKDMPCSolver::initialize / update do (HKDMPC.cpp:58-70, :136-150), with every phase behind a recording decorator
  MultiPhaseDDP<T> solver;
  deque<shared_ptr<SinglePhaseBase<T>>> multiple_phases;
  int id = 0;
  for (auto phase : pd.phase_ptrs) multiple_phases.push_back(std::make_shared<SpyPhase>(phase, id++));
  solver.set_multiPhaseProblem(multiple_phases);
  solver.set_initial_condition(xinit);
  g_log.clear();
  solver.solve(opt);
  int n_iters, n_ls, n_reg; float ms;
  solver.get_solver_info(n_iters, n_ls, n_reg, ms);
  std::vector<float> c, d, e, i;
  solver.get_solver_info(c, d, e, i);
  fprintf(f, "x0"); put_vec(f, xinit); fprintf(f, "\n");
  fprintf(f, "counters iter %d ls_iter_total %d reg_iter_total %d\n", n_iters, n_ls, n_reg);
  fprintf(f, "final cost %.17g feas %.17g tconstr %.17g pconstr %.17g\n", (double)solver.get_actual_cost(), (double)solver.get_dyn_infeasibility(),
          (double)solver.get_terminal_constraint_violation(), (double)solver.get_path_constraint_violation());
  fprintf(f, "float_cost_buffer %d", (int)c.size()); for (float v : c) fprintf(f, " %.9g", (double)v); fprintf(f, "\n");
  fprintf(f, "events %d\n", (int)g_log.size());
  for (const Event& ev : g_log) fprintf(f, "%d %d %.17g %.17g\n", ev.type, ev.phase, ev.a, ev.b);
  dump_problem(f, pd, "solution");
}

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: ref_hkd <quad_reference.csv> <in.txt> <out.txt>\n"); return 2; }
  std::ifstream in(argv[2]);
  int n_prob = 0, n_upd = 0;
  in >> n_prob >> n_upd;
  FILE* f = fopen(argv[3], "w");
  if (!in.good() || !f) { fprintf(stderr, "cannot open input / output\n"); return 2; }
  fprintf(f, "n_problems %d n_updates %d\n", n_prob, n_upd);
  for (int b = 0; b < n_prob; ++b) {
    Vec12<T> body, qJ, qdummy;
    for (int i = 0; i < 12; ++i) in >> body[i];
    for (int i = 0; i < 12; ++i) in >> qJ[i];
    std::vector<Vec3<T>> nudge(n_upd);
    for (int u = 0; u < n_upd; ++u) for (int i = 0; i < 3; ++i) in >> nudge[u][i];

    // HKDMPCSolver<T>::HKDMPCSolver + initialize (HKDMPC.h:27-32, HKDMPC.cpp:19-56)
    QuadReference quad_reference;
    quad_reference.load_top_level_data(argv[1], true);
    HSDDP_OPTION ddp_options;
    loadHSDDPSetting("../HKDMPC/settings/ddp_setting.info", ddp_options);
    HKDPlanConfig mpc_config;
    mpc_config.plan_duration = .6;
    mpc_config.nsteps_between_mpc = 2;
    mpc_config.timeStep = 0.01;
    HKDProblem<T> opt_problem;
    HKDProblemData<T> opt_problem_data;
    opt_problem.clear_problem_data();
    opt_problem_data.quad_ref_ptr = &quad_reference;
    opt_problem.set_problem_data(&opt_problem_data, mpc_config);
    opt_problem.initialization();

    Vec3<T> pos, eul;
    pos = body.segment(3, 3);
    eul = body.head(3);
    const auto& initial_contact = opt_problem_data.phase_contacts.front();
    compute_hkd_state(eul, pos, qJ, qdummy, initial_contact);
    VecM<T, 24> xinit;
    xinit << body, qdummy;

    fprintf(f, "problem %d\n", b);
    dump_problem(f, opt_problem_data, "guess");
    run_solve(f, opt_problem_data, ddp_options, xinit);

    for (int u = 0; u < n_upd; ++u) {
      // HKDMPCSolver<T>::update (HKDMPC.cpp:89-150): fewer iterations when re-solving
      ddp_options.max_AL_iter = 2;
      ddp_options.max_DDP_iter = 1;
      opt_problem.update();
      // the "measured" state: the plan's own prediction for the new start (front of the shifted nominal trajectory), position nudged
      xinit = opt_problem_data.trajectory_ptrs.front()->Xbar.front();
      xinit.segment(3, 3) += nudge[u];
      fprintf(f, "update %d\n", u);
      dump_problem(f, opt_problem_data, "guess");
      run_solve(f, opt_problem_data, ddp_options, xinit);
    }
  }
  fclose(f);
  return 0;
}
