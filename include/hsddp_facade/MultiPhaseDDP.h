// MultiPhaseDDP.h — MultiPhaseDDP<T> with the public interface of the reference (HSDDPSolver/header/MultiPhaseDDP.h:24-120):
//   set_multiPhaseProblem(deque<shared_ptr<SinglePhaseBase<T>>>), set_initial_condition(DVec<T>), solve(HSDDP_OPTION&, max_cputime),
//   get_actual_cost / get_dyn_infeasibility / get_path_constraint_violation / get_terminal_constraint_violation, both get_solver_info.
// solve() hands the phases' Xbar / Ubar / K to the CUDA solver as the starting guess (what the reference's solve starts from: its first
// hybrid_rollout(0) runs on the trajectories as the caller left them, MultiPhaseDDP.cpp:238), runs the batched HS-DDP path through the
// C ABI (include/cafe_gpu.h) with a batch of one, and writes Xbar, Ubar, Y, dU, K, Qu, Quu, Qux, G back into the phases' Trajectory
// deques (TrajectoryManagement.h:54-85), which is where MHPCLocomotion::publish_mpc_cmd / HKDMPCSolver::publish_mpc_cmd read them.
// Batch extension (not in the reference): set_initial_conditions(B states) + solve() solves B problems that share the phases' guess;
// problem 0 is written back, get_batch_*() return every problem's record.
// max_cputime is accepted and ignored: the reference polls a wall clock between the stages of an iteration (:287-377), which makes
// iteration counts machine-dependent; the GPU path runs to the iteration caps / convergence tests of HSDDP_OPTION.
// T must be double (the kernels are fp64).
#pragma once
#include <deque>
#include <memory>
#include <type_traits>
#include <vector>
#include "SinglePhase.h"

using std::deque;
using std::shared_ptr;
using std::vector;

template <typename T>
class MultiPhaseDDP {
  static_assert(std::is_same<T, double>::value, "the GPU path computes in double precision");

 private:
  deque<shared_ptr<SinglePhaseBase<T>>> phases;

 public:
  MultiPhaseDDP() {}
  MultiPhaseDDP(const MultiPhaseDDP&) = delete;
  MultiPhaseDDP& operator=(const MultiPhaseDDP&) = delete;
  ~MultiPhaseDDP() { release(); }

  void set_multiPhaseProblem(deque<shared_ptr<SinglePhaseBase<T>>> phases_in) {
    phases = phases_in;
    n_phases = (int)phases.size();
    if (n_phases == 0) throw std::invalid_argument("set_multiPhaseProblem: no phases");
    std::shared_ptr<cafe_facade::DeckOwner> d = phases.front()->cafe_deck;
    if (!d || !d->h) throw std::invalid_argument("set_multiPhaseProblem: phases not created by a problem builder of this library");
    if (d->deck()->n_phases != n_phases) throw std::invalid_argument("set_multiPhaseProblem: the deck has another number of phases");
    for (int i = 0; i < n_phases; ++i)
      if (phases[i]->cafe_deck != d || phases[i]->cafe_phase_index != i) throw std::invalid_argument("set_multiPhaseProblem: phases must be the consecutive phases of one problem");
    if (d != deck_) {   // another problem, or the next MPC window: its deck differs (horizons, references); solve() hands it to the
      deck_ = d;        // existing solver with cafe_gpu_update_deck (device buffers re-used) or creates the solver
      deck_changed_ = true;
    }
  }
  void set_initial_condition(DVec<T> x0_in) { x0 = x0_in; x0_batch.assign(x0.data(), x0.data() + x0.size()); B_ = 1; }
  // batch extension: B initial states, every problem starts from the phases' common guess
  void set_initial_conditions(const vector<DVec<T>>& x0s) {
    if (x0s.empty()) throw std::invalid_argument("set_initial_conditions: empty batch");
    x0 = x0s.front(); B_ = (int)x0s.size(); x0_batch.clear();
    for (const auto& v : x0s) x0_batch.insert(x0_batch.end(), v.data(), v.data() + v.size());
  }
  // device the solver lives on (before the first solve of a problem) and the batch capacity reserved at creation
  void set_device(int device, int max_batch = 1) { device_ = device; max_batch_ = max_batch; }

  void solve(HSDDP_OPTION& option, const float& max_cputime = 1e6) {
    (void)max_cputime;
    using cafe_facade::check;
    if (!deck_) throw std::logic_error("solve: set_multiPhaseProblem first");
    if (B_ <= 0) throw std::logic_error("solve: set_initial_condition first");
    if (h_ && deck_changed_) {
      // the next MPC window of the problem the solver has just solved (same lineage, same batch): the update on the device carries what lives there -
      // the relaxed-barrier update counts travel with the knots (ConstraintsBase.h:296-306); trajectories and AL parameters are handed over from the
      // phases below in any case. Anything else: new deck, cold start.
      const int adv = deck_->k0 - solved_k0_;
      const bool next_window = deck_->lineage == solved_lineage_ && adv >= 0 && B_ == solved_B_;
      if (!(next_window && cafe_gpu_update_deck(h_, deck_->deck(), adv, B_) == 0) && cafe_gpu_update_deck(h_, deck_->deck(), 0, 0) != 0) { cafe_gpu_destroy(h_); h_ = nullptr; }   // e.g. another model family
    }
    if (h_ && B_ > cap_) { cafe_gpu_destroy(h_); h_ = nullptr; }
    if (!h_) { cap_ = std::max(max_batch_, B_); check(cafe_gpu_create(deck_->deck(), device_, cap_, &h_)); }
    deck_changed_ = false;
    const long rec = cafe_solution_size(deck_->deck());
    std::vector<double> one(rec), guess((size_t)rec * B_);
    long off = 0;
    for (auto& ph : phases) { ph->cafe_pack_guess(one.data() + off); off += ph->cafe_record_size(); }
    if (off != rec) throw std::logic_error("solve: phase records do not add up to the deck's solution size");
    for (int b = 0; b < B_; ++b) std::copy(one.begin(), one.end(), guess.begin() + (size_t)b * rec);
    check(cafe_gpu_set_initial_guess(h_, guess.data(), B_));
    {
      // augmented-Lagrangian parameters: what the phases carry from the previous MPC step, else the deck's TD_AL values
      const CafeDeck* d = deck_->deck();
      std::vector<double> al((size_t)B_ * n_phases * 8, 0.0);
      for (int i = 0; i < n_phases; ++i)
        for (int e = 0; e < d->phase[i].n_td && e < 4; ++e) {
          const double sg = phases[i]->cafe_al_set ? phases[i]->cafe_al[2 * e] : d->phase[i].al_td.sigma;
          const double lm = phases[i]->cafe_al_set ? phases[i]->cafe_al[2 * e + 1] : d->phase[i].al_td.lambda;
          for (int b = 0; b < B_; ++b) { al[((size_t)b * n_phases + i) * 8 + 2 * e] = sg; al[((size_t)b * n_phases + i) * 8 + 2 * e + 1] = lm; }
        }
      check(cafe_gpu_set_al_params(h_, al.data(), B_));
    }
    const CafeOptions o = cafe_options_from_hsddp(option);
    check(cafe_gpu_solve_batch(h_, x0_batch.data(), B_, &o));
    solved_k0_ = deck_->k0; solved_lineage_ = deck_->lineage; solved_B_ = B_;
    info_.resize(B_);
    check(cafe_gpu_get_info(h_, info_.data()));
    check(cafe_gpu_get_solution(h_, 0, 1, one.data()));
    off = 0;
    for (auto& ph : phases) { ph->cafe_unpack_solution(one.data() + off); off += ph->cafe_record_size(); }
    {
      std::vector<double> al((size_t)B_ * n_phases * 8);
      check(cafe_gpu_get_al_params(h_, al.data()));
      for (int i = 0; i < n_phases; ++i) { for (int e = 0; e < 8; ++e) phases[i]->cafe_al[e] = al[(size_t)i * 8 + e]; phases[i]->cafe_al_set = true; }   // problem 0, like the trajectories
    }
    hist_.assign((size_t)B_ * HIST_CAP * 4, 0.0);
    check(cafe_gpu_get_history(h_, hist_.data(), HIST_CAP));
    double ms = 0; check(cafe_gpu_get_solve_ms(h_, &ms)); solve_time_ = (float)ms;
    actual_cost = info_[0].cost; feas = info_[0].feas; max_tconstr = info_[0].max_tconstr; max_pconstr = info_[0].max_pconstr;
    iter_ = info_[0].iter; ls_iter_total_ = info_[0].ls_iter_total; reg_iter_total_ += info_[0].reg_iter_total;   // never reset, like the reference (:218-219, :316)
  }

  T get_actual_cost() { return actual_cost; }
  T get_dyn_infeasibility() { return feas; }
  T get_path_constraint_violation() { return hist_at(0, info_.at(0).n_hist - 1, 3); }       // ineq_feas_buffer.back()
  T get_terminal_constraint_violation() { return hist_at(0, info_.at(0).n_hist - 1, 2); }   // eqn_feas_buffer.back()
  void get_solver_info(std::vector<float>& cost_out, std::vector<float>& dyn_feas_out, std::vector<float>& eqn_feas_out, std::vector<float>& ineq_feas_out) {
    cost_out.clear(); dyn_feas_out.clear(); eqn_feas_out.clear(); ineq_feas_out.clear();
    for (int i = 0; i < info_.at(0).n_hist && i < HIST_CAP; ++i) {
      cost_out.push_back((float)hist_at(0, i, 0)); dyn_feas_out.push_back((float)hist_at(0, i, 1));
      eqn_feas_out.push_back((float)hist_at(0, i, 2)); ineq_feas_out.push_back((float)hist_at(0, i, 3));
    }
  }
  void get_solver_info(int& n_iters, int& n_ls_iters, int& n_reg_iters, float& solve_time) {
    n_iters = iter_; n_ls_iters = ls_iter_total_; n_reg_iters = reg_iter_total_; solve_time = solve_time_;
  }

  // ---- batch extension
  const std::vector<CafeInfo>& get_batch_info() const { return info_; }
  std::vector<double> get_batch_solution(int b) const { std::vector<double> v(cafe_solution_size(deck_->deck())); cafe_facade::check(cafe_gpu_get_solution(h_, b, 1, v.data())); return v; }
  CafeHandle* cafe_handle() const { return h_; }

 private:
  enum { HIST_CAP = 320 };
  double hist_at(int b, int i, int c) const { return hist_[((size_t)b * HIST_CAP + (i < 0 ? 0 : i)) * 4 + c]; }
  void release() { if (h_) { cafe_gpu_destroy(h_); h_ = nullptr; } deck_.reset(); }

  int n_phases = 0;
  DVec<T> x0;
  std::vector<double> x0_batch;
  int B_ = 0, device_ = 0, max_batch_ = 1, cap_ = 0;
  CafeHandle* h_ = nullptr;
  bool deck_changed_ = false;
  int solved_k0_ = 0, solved_B_ = 0; long solved_lineage_ = -1;   // the deck of the last solve on h_
  std::shared_ptr<cafe_facade::DeckOwner> deck_;
  std::vector<CafeInfo> info_;
  std::vector<double> hist_;
  T actual_cost = 0, feas = 0, max_tconstr = 0, max_pconstr = 0;
  int iter_ = 0, ls_iter_total_ = 0, reg_iter_total_ = 0;
  float solve_time_ = 0;
};
