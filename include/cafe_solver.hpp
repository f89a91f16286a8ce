// cafe_solver.hpp — header-only C++ host mirror of the reference's solver-facing API on top of the C ABI
// (include/cafe_gpu.h). Method names and call order follow the reference (file:line under /root/reference):
//   loadHSDDPSetting(filename, option)                    HSDDPSolver/common/HSDDP_CompoundTypes.h:57-82
//   HKDProblem / MHPCProblem ::initialization()           HKDMPC/HKD-TrajOpt/HKDProblem.cpp:15, MHPC/MHPC-Trajopt/MHPCProblem.cpp:13
//   MultiPhaseDDP::set_multiPhaseProblem / set_initial_condition / solve / get_solver_info / get_actual_cost ...
//                                                          HSDDPSolver/header/MultiPhaseDDP.h:33-93
// The reference solves ONE problem per MultiPhaseDDP object; this mirror solves a batch that shares the phase deck.
// Errors: the reference prints and returns void; here every failure throws cafe::Error (nothing crosses the C ABI as an exception).
#pragma once
#include <stdexcept>
#include <string>
#include <vector>
#include "cafe_gpu.h"

namespace cafe {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
inline void check(int rc) { if (rc != 0) throw Error(rc, cafe_last_error()); }

typedef CafeOptions HSDDP_OPTION;
inline void loadHSDDPSetting(const std::string& filename, HSDDP_OPTION& setting) { check(cafe_options_load(filename.c_str(), &setting)); }

class ProblemBase {
 public:
  ProblemBase() {}
  ProblemBase(const ProblemBase&) = delete;
  ProblemBase& operator=(const ProblemBase&) = delete;
  ~ProblemBase() { if (h_) cafe_deck_free(h_); }
  const CafeDeck* deck() const { return cafe_deck_get(h_); }
  // the deck stands for the problem after update() (MHPCProblem.cpp:252-372, HKDProblem.cpp:117-222; nsteps = shift in knots, 2 for HKD): a tail
  // phase that update has just opened has no shooting states yet. Returns the index of that phase, -1 if there is none.
  int mark_mpc_update(int nsteps) { int which = -1; check(cafe_deck_mark_mpc_update(h_, nsteps, &which)); return which; }
  int n_phases() const { return deck()->n_phases; }
 protected:
  CafeDeckHandle* h_ = nullptr;
};

struct HKDPlanConfig { float plan_duration = 0.6f; float timeStep = 0.01f; int nsteps_between_mpc = 2; };  // HKDMPC.cpp:26-28

class HKDProblem : public ProblemBase {
 public:
  // reference_csv: Reference/Data/<gait>/quad_reference.csv (loaded with reorder = true); constraint file: HKDMPC/settings/constraint_params.info
  void initialization(const std::string& reference_csv, const std::string& constraint_params_info, const HKDPlanConfig& cfg = HKDPlanConfig(), int k0 = 0) {
    if (h_) { cafe_deck_free(h_); h_ = nullptr; }
    check(cafe_deck_build_hkd(reference_csv.c_str(), constraint_params_info.c_str(), cfg.plan_duration, cfg.timeStep, cfg.nsteps_between_mpc, k0, &h_));
  }
  // compute_hkd_state (HKDModel.h:66-96) with the first phase's contact; body = [eul, pos, omega, vel]
  std::vector<double> initial_state(const double body[12], const double qJ[12]) const {
    std::vector<double> x0(24);
    check(cafe_hkd_state(body, qJ, deck()->phase[0].contact, x0.data()));
    return x0;
  }
};

class MHPCProblem : public ProblemBase {
 public:
  // mhpc_config_info: MHPC/settings/mhpc_config.info; settings_root: directory that plays the reference's "../"
  void initialization(const std::string& reference_csv, const std::string& mhpc_config_info, const std::string& settings_root, int k0 = 0) {
    if (h_) { cafe_deck_free(h_); h_ = nullptr; }
    check(cafe_deck_build_mhpc(reference_csv.c_str(), mhpc_config_info.c_str(), settings_root.c_str(), k0, &h_));
  }
};

// LocoProblem<T> : MHPCProblem<T> (MHPC/MHPC-Trajopt/Locomotion/LocoProblem.h:8-25): whole-body-only locomotion trajectory optimisation,
// torque-limit and GRF barriers only. loco_config_info: Locomotion/settings/loco_config.info
class LocoProblem : public MHPCProblem {
 public:
  void initialization(const std::string& reference_csv, const std::string& loco_config_info, const std::string& settings_root, int k0 = 0) {
    if (h_) { cafe_deck_free(h_); h_ = nullptr; }
    check(cafe_deck_build_loco(reference_csv.c_str(), loco_config_info.c_str(), settings_root.c_str(), k0, &h_));
  }
};

// The in-place barrel roll built by main() of MHPC/MHPC-Trajopt/BarrelRoll/BarrelRollTO.cpp:65-275 (no problem class in the reference)
class BarrelRollProblem : public ProblemBase {
 public:
  void initialization(const std::string& br_cost_weights_json, const std::string& br_constraint_params_info) {
    if (h_) { cafe_deck_free(h_); h_ = nullptr; }
    check(cafe_deck_build_barrel_to(br_cost_weights_json.c_str(), br_constraint_params_info.c_str(), &h_));
  }
  // interpolated initial state trajectories (BarrelRollTO.cpp:131-147) as packed guesses for MultiPhaseDDP::set_initial_guess
  std::vector<double> initial_guess(const double* x0, int B) const {
    std::vector<double> g((size_t)B * (size_t)cafe_solution_size(deck()));
    check(cafe_barrel_to_initial_guess(deck(), x0, B, g.data()));
    return g;
  }
};

class MultiPhaseDDP {
 public:
  MultiPhaseDDP() {}
  MultiPhaseDDP(const MultiPhaseDDP&) = delete;
  MultiPhaseDDP& operator=(const MultiPhaseDDP&) = delete;
  ~MultiPhaseDDP() { if (h_) cafe_gpu_destroy(h_); }

  void set_multiPhaseProblem(const ProblemBase& problem, int max_batch, int device = 0) {
    if (h_) { cafe_gpu_destroy(h_); h_ = nullptr; }
    deck_ = problem.deck();
    check(cafe_gpu_create(deck_, device, max_batch, &h_));
  }
  // x0: B rows of the first phase's state dimension
  void set_initial_condition(const std::vector<double>& x0, int B) { x0_ = x0; B_ = B; }
  void solve(HSDDP_OPTION& option) { check(cafe_gpu_solve_batch(h_, x0_.data(), B_, &option)); }
  // initial Xbar / Ubar / K per problem in the packed solution layout ([B][solution_size()]); an empty vector returns to the cold start
  void set_initial_guess(const std::vector<double>& guess) { check(cafe_gpu_set_initial_guess(h_, guess.empty() ? nullptr : guess.data(), B_)); }

  std::vector<CafeInfo> get_solver_info() const { std::vector<CafeInfo> v(B_); check(cafe_gpu_get_info(h_, v.data())); return v; }
  // the reference's scalar getters (header/MultiPhaseDDP.h:77-93), for problem b of the batch
  void get_solver_info(int b, int& n_iters, int& n_ls_iters, int& n_reg_iters, float& solve_time) const {
    const CafeInfo i = get_solver_info().at(b);
    n_iters = i.iter; n_ls_iters = i.ls_iter_total; n_reg_iters = i.reg_iter_total; solve_time = (float)solve_ms();  // batch time: the problems run together
  }
  double get_actual_cost(int b) const { return get_solver_info().at(b).cost; }
  double get_dyn_infeasibility(int b) const { return get_solver_info().at(b).feas; }
  double get_terminal_constraint_violation(int b) const { return get_solver_info().at(b).max_tconstr; }
  double get_path_constraint_violation(int b) const { return get_solver_info().at(b).max_pconstr; }
  // the four per-iteration buffers of get_solver_info(cost, dyn_feas, eqn_feas, ineq_feas) (float like the reference), problem b
  void get_solver_info(int b, std::vector<float>& cost, std::vector<float>& dyn_feas, std::vector<float>& eqn_feas, std::vector<float>& ineq_feas) const {
    const int cap = 256;
    const std::vector<double> h = get_history(cap);
    const int n = get_solver_info().at(b).n_hist;
    cost.clear(); dyn_feas.clear(); eqn_feas.clear(); ineq_feas.clear();
    for (int i = 0; i < n && i < cap; ++i) {
      const double* r = &h[((size_t)b * cap + i) * 4];
      cost.push_back((float)r[0]); dyn_feas.push_back((float)r[1]); eqn_feas.push_back((float)r[2]); ineq_feas.push_back((float)r[3]);
    }
  }
  // float32 MHPC_Command_lcmt fields for the first n_steps whole-body knots (what publish_mpc_cmd sends, MHPCLocomotion.cpp:236-281)
  std::vector<float> get_lcm_commands(int n_steps) const { std::vector<float> v((size_t)B_ * cafe_lcm_command_size(n_steps)); check(cafe_gpu_get_lcm_commands(h_, n_steps, v.data())); return v; }
  // float32 hkd_command_lcmt fields (hkd_controls, des_body_state, feedback) for n_steps knots (HKDMPCSolver::publish_mpc_cmd, HKDMPC.cpp:243-290)
  std::vector<float> get_hkd_lcm_commands(int n_steps) const { std::vector<float> v((size_t)B_ * cafe_hkd_lcm_command_size(n_steps)); check(cafe_gpu_get_hkd_lcm_commands(h_, n_steps, v.data())); return v; }
  // receding horizon (MHPCProblem::update, MHPCProblem.cpp:252-397): warm start of this solver (deck at start offset k0) from the plan held
  // by `prev` (deck at start offset prev_k0), shifted on the device; and the planned state `knots_ahead` knots into the plan, [B][n]
  void shift_guess_from(const MultiPhaseDDP& prev, int prev_k0, int k0) { check(cafe_gpu_shift_guess(h_, prev.h_, prev_k0, k0, B_)); }
  // the MPC update on this solver: `next` (the problem re-cut k_advance knots later) replaces the deck, the previous solution is the warm start
  void update_deck(const ProblemBase& next, int k_advance) { check(cafe_gpu_update_deck(h_, next.deck(), k_advance, B_)); }
  std::vector<double> planned_state(int knots_ahead, int n) const { std::vector<double> v((size_t)B_ * n); check(cafe_gpu_get_planned_state(h_, knots_ahead, v.data())); return v; }
  // cost / dynamics feasibility / terminal / path constraint buffers (get_solver_info(cost_out, ...), MultiPhaseDDP.cpp:554-563)
  std::vector<double> get_history(int cap) const { std::vector<double> v((size_t)B_ * cap * 4); check(cafe_gpu_get_history(h_, v.data(), cap)); return v; }
  long solution_size() const { return cafe_solution_size(deck_); }
  std::vector<double> get_solution(int b0, int nb) const { std::vector<double> v((size_t)nb * solution_size()); check(cafe_gpu_get_solution(h_, b0, nb, v.data())); return v; }
  long command_size(int n_gain_knots) const { return cafe_command_size(deck_, n_gain_knots); }
  std::vector<double> get_commands(int n_gain_knots) const { std::vector<double> v((size_t)B_ * command_size(n_gain_knots)); check(cafe_gpu_get_commands(h_, n_gain_knots, v.data())); return v; }
  double solve_ms() const { double ms = 0; check(cafe_gpu_get_solve_ms(h_, &ms)); return ms; }

 private:
  CafeHandle* h_ = nullptr;
  const CafeDeck* deck_ = nullptr;
  std::vector<double> x0_;
  int B_ = 0;
};

}  // namespace cafe
