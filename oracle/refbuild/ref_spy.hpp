// ref_spy.hpp — TEST INFRASTRUCTURE ONLY (oracle/): what the reference-solver drivers (ref_hkd_driver.cpp, ref_mhpc_driver.cpp) share.
// SpyPhase is a decorator around a phase of the reference: it forwards the whole SinglePhaseBase<T> interface to the reference's SinglePhase
// and writes down, in full precision, what MultiPhaseDDP<T>::solve asked for and got (the solver's own history buffers are float).
#pragma once
#include <cstdio>
#include <deque>
#include <memory>
#include <vector>
#include "HSDDP_CompoundTypes.h"
#include "MultiPhaseDDP.h"

typedef double T;

struct Event { int type, phase; double a, b; };
// event types
enum { EV_ROLLOUT = 1, EV_COST = 2, EV_FEAS = 3, EV_LQ = 4, EV_BWD = 5, EV_LIN = 6, EV_ACCEPT = 7, EV_AL = 8, EV_REB = 9, EV_TCON = 10, EV_PCON = 11, EV_BWD_DV = 12 };
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
    if (ok) { T a, b; p->get_exp_cost_change(a, b); g_log.push_back({EV_BWD_DV, id, a, b}); }   // what MS = false takes its expected cost change from
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


// one phase of a problem: horizon, stance, times, state / control dimensions, then the arrays of its Trajectory (matrices column-major)
template <class Traj, class Contact>
static void dump_traj(FILE* f, Traj& tr, int i, const Contact& contact, double start, double end) {
  const int h = tr.horizon;
  fprintf(f, "phase %d horizon %d contact %d %d %d %d start %.9g end %.9g n %d m %d\n", i, h, (int)contact[0], (int)contact[1], (int)contact[2], (int)contact[3],
          start, end, (int)tr.Xbar[0].size(), h > 0 ? (int)tr.Ubar[0].size() : 0);
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

// MultiPhaseDDP<T>::solve on the given phases, every phase behind a recording decorator; writes x0, the solver's counters and final
// figures and the event list
template <class XV>
static void solve_and_record(FILE* f, const std::deque<std::shared_ptr<SinglePhaseBase<T>>>& phases, HSDDP_OPTION& opt, const XV& xinit) {
  MultiPhaseDDP<T> solver;
  std::deque<std::shared_ptr<SinglePhaseBase<T>>> multiple_phases;
  int id = 0;
  for (auto phase : phases) multiple_phases.push_back(std::make_shared<SpyPhase>(phase, id++));
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
}
