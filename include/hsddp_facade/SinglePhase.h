// SinglePhase.h — SinglePhase<T,xs,us,ys> (HSDDPSolver/header/SinglePhase.h:22-260): the object a problem builder creates per phase
// and hands to MultiPhaseDDP<T>::set_multiPhaseProblem. Same template parameters, constructor and set_trajectory / set_time_offset /
// update_SS_config / get_state_dim / get_control_dim / print as the reference. The callback setters (set_dynamics,
// set_dynamics_partial, set_resetmap(_partial), add_cost, add_pathConstraint, add_terminalConstraint) exist with the reference's
// names but REFUSE: a std::function cannot run inside a CUDA kernel, the GPU path evaluates the built-in HKD / whole-body / SRB
// models, costs and constraints selected by the phase deck (include/cafe_deck.h). A caller that needs another model has to add it
// to csrc/model_*.cuh; silently ignoring its callback would be worse than failing.
#pragma once
#include <cstdio>
#include <functional>
#include <numeric>
#include <stdexcept>
#include "SinglePhaseBase.h"
#include "TrajectoryManagement.h"

template <typename T, size_t xs, size_t us, size_t ys>
class SinglePhase : public SinglePhaseBase<T> {
 public:
  typedef VecM<T, xs> State;
  typedef VecM<T, us> Contrl;
  typedef VecM<T, ys> Output;
  typedef MatMN<T, xs, xs> StateMap;
  typedef MatMN<T, xs, us> ContrlMap;
  typedef MatMN<T, ys, xs> OutputMap;
  typedef MatMN<T, ys, us> DirectMap;

  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
  SinglePhase(int num_threads = 1) : num_threads_(num_threads) {}

  void set_trajectory(shared_ptr<Trajectory<T, xs, us, ys>> traj_) { traj = traj_; phase_horizon = traj->horizon; dt = traj->timeStep; }
  void set_time_offset(float t_offset_in) { t_offset = t_offset_in; }

  template <class F> void set_dynamics(F) { refuse("set_dynamics"); }
  template <class F> void set_dynamics_partial(const F&) { refuse("set_dynamics_partial"); }
  template <class F> void set_resetmap(F) { refuse("set_resetmap"); }
  template <class F> void set_resetmap_partial(F) { refuse("set_resetmap_partial"); }
  template <class P> void add_cost(P) { refuse("add_cost"); }
  template <class P> void add_pathConstraint(P) { refuse("add_pathConstraint"); }
  template <class P> void add_terminalConstraint(P) { refuse("add_terminalConstraint"); }

  void initialization() override {}
  size_t get_state_dim() override { return xs; }
  size_t get_control_dim() override { return us; }
  void empty_control() override { traj->zero_val_approx(); }
  // every knot a shooting state (ss_sz = horizon + 1) is what the solver runs; ss_sz = 0 marks a single-shooting phase (the tail
  // phase an MPC update has just opened, MHPCProblem.cpp:366-369) and must agree with the deck's CafePhase::single_shooting
  void update_SS_config(int ss_sz) override { SS_size = ss_sz; }
  int get_SS_size() const { return SS_size; }
  void get_trajectory(std::vector<std::vector<float>>& x_tau, std::vector<std::vector<float>>& u_tau) override {  // SinglePhase.cpp:536-553
    for (int k = 0; k < phase_horizon; ++k) {
      std::vector<float> x(xs), u(us);
      for (size_t i = 0; i < xs; ++i) x[i] = (float)traj->Xbar[k][i];
      for (size_t i = 0; i < us; ++i) u[i] = (float)traj->Ubar[k][i];
      x_tau.push_back(x); u_tau.push_back(u);
    }
  }
  void print() override {
    printf("Phase %d: horizon %d, dt %g, time offset %g, state %zu control %zu output %zu\n", this->cafe_phase_index, phase_horizon, (double)dt, (double)t_offset, xs, us, ys);
  }
  shared_ptr<Trajectory<T, xs, us, ys>> get_trajectory_ptr() const { return traj; }

  // ---- record <-> Trajectory (layout of cafe_solution_size, include/cafe_gpu.h): Xbar Ubar Y dU K Qu Quu Qux G
  long cafe_record_size() const override {
    const long h = phase_horizon;
    return (h + 1) * (long)xs + h * (long)us + h * (long)ys + h * (long)us + h * (long)(us * xs) + h * (long)us + h * (long)(us * us) + h * (long)(us * xs) + (h + 1) * (long)xs;
  }
  void cafe_pack_guess(double* r) const override {
    const int h = phase_horizon;
    for (long i = 0, n = cafe_record_size(); i < n; ++i) r[i] = 0.0;
    double* pX = r; double* pU = pX + (size_t)(h + 1) * xs; double* pK = pU + (size_t)h * us + (size_t)h * ys + (size_t)h * us;
    for (int k = 0; k <= h; ++k) for (size_t i = 0; i < xs; ++i) pX[k * xs + i] = (double)traj->Xbar[k][i];
    for (int k = 0; k < h; ++k) for (size_t i = 0; i < us; ++i) pU[k * us + i] = (double)traj->Ubar[k][i];
    for (int k = 0; k < h; ++k) for (size_t j = 0; j < xs; ++j) for (size_t i = 0; i < us; ++i) pK[(size_t)k * us * xs + i + us * j] = (double)traj->K[k](i, j);
  }
  void cafe_unpack_solution(const double* r) override {
    const int h = phase_horizon;
    const double* p = r;
    for (int k = 0; k <= h; ++k) for (size_t i = 0; i < xs; ++i) { traj->Xbar[k][i] = (T)p[k * xs + i]; traj->X[k][i] = traj->Xbar[k][i]; }
    p += (size_t)(h + 1) * xs;
    for (int k = 0; k < h; ++k) for (size_t i = 0; i < us; ++i) { traj->Ubar[k][i] = (T)p[k * us + i]; traj->U[k][i] = traj->Ubar[k][i]; }
    p += (size_t)h * us;
    for (int k = 0; k < h; ++k) for (size_t i = 0; i < ys; ++i) traj->Y[k][i] = (T)p[k * ys + i];
    p += (size_t)h * ys;
    for (int k = 0; k < h; ++k) for (size_t i = 0; i < us; ++i) traj->dU[k][i] = (T)p[k * us + i];
    p += (size_t)h * us;
    for (int k = 0; k < h; ++k) for (size_t j = 0; j < xs; ++j) for (size_t i = 0; i < us; ++i) traj->K[k](i, j) = (T)p[(size_t)k * us * xs + i + us * j];
    p += (size_t)h * us * xs;
    for (int k = 0; k < h; ++k) for (size_t i = 0; i < us; ++i) traj->Qu[k][i] = (T)p[k * us + i];
    p += (size_t)h * us;
    for (int k = 0; k < h; ++k) for (size_t j = 0; j < us; ++j) for (size_t i = 0; i < us; ++i) traj->Quu[k](i, j) = (T)p[(size_t)k * us * us + i + us * j];
    p += (size_t)h * us * us;
    for (int k = 0; k < h; ++k) for (size_t j = 0; j < xs; ++j) for (size_t i = 0; i < us; ++i) traj->Qux[k](i, j) = (T)p[(size_t)k * us * xs + i + us * j];
    p += (size_t)h * us * xs;
    for (int k = 0; k <= h; ++k) for (size_t i = 0; i < xs; ++i) traj->G[k][i] = (T)p[k * xs + i];
  }

 private:
  static void refuse(const char* what) {
    throw std::logic_error(std::string("SinglePhase::") + what + ": host callbacks cannot run on the GPU path; phases are bound to the built-in models by the problem builders (include/hsddp_facade/SinglePhase.h)");
  }
  shared_ptr<Trajectory<T, xs, us, ys>> traj;
  int phase_horizon = 0;
  T dt = 0;
  float t_offset = 0;
  int num_threads_ = 1;
  int SS_size = -1;
};
