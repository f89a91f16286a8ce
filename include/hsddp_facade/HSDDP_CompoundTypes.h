// HSDDP_CompoundTypes.h — HSDDP_OPTION and loadHSDDPSetting with the reference's fields, types and defaults
// (HSDDPSolver/common/HSDDP_CompoundTypes.h:13-82). The settings file is read by the C ABI (cafe_options_load, no Boost);
// cafe_options_from_hsddp() converts field by field into the plain-C CafeOptions the solver takes — the two structs do NOT share a
// layout (four 1-byte bools here, ints there), so never memcpy one onto the other.
#pragma once
#include <iostream>
#include <stdexcept>
#include <string>
#include "HSDDP_CPPTypes.h"
#include "../cafe_gpu.h"

struct HSDDP_OPTION {
  double alpha = 0.1;
  double gamma = 0.1;
  double update_penalty = 8;
  double update_relax = 0.1;
  double update_regularization = 2;
  double update_ReB = 7;
  int max_DDP_iter = 3;
  int max_AL_iter = 2;
  int max_DDP_iter_runtime = 1;
  int max_AL_iter_runtime = 2;
  double cost_thresh = 1e-03;
  double tconstr_thresh = 1e-03;
  double pconstr_thresh = 1e-03;
  double dynamics_feas_thresh = 1e-03;
  double merit_rho = 1e04;
  double merit_scale = 0.2;
  double merit_offset = 10;
  bool AL_active = 1;
  bool ReB_active = 1;
  bool smooth_active = 0;
  bool MS = true;
  int nsteps_per_node = 1;

  void print() {
    std::cout << "===================== HSDDP Setting =================== \n";
    std::cout << "Multiple Shooting \t" << MS << "\n";
    std::cout << "Number of integration time steps per node \t" << nsteps_per_node << "\n";
    std::cout << "Is Terminal Constraint active \t" << AL_active << "\n";
    std::cout << "Is path constraint active \t" << ReB_active << "\n";
    std::cout << "Maximum inner-loop iterations \t" << max_DDP_iter << "\n";
    std::cout << "Maximum outer-loop iterations \t" << max_AL_iter << "\n";
    std::cout << "Line search update param \t" << alpha << "\n";
    std::cout << "Merit function penalty \t" << merit_rho << "\n";
    std::cout << "Merit scale parameter \t" << merit_scale << "\n";
    std::cout << "Merit offset parameter \t" << merit_offset << "\n";
    std::cout << "Terminal constraint threshold \t" << tconstr_thresh << "\n";
    std::cout << "Path constraint threshold \t" << pconstr_thresh << "\n";
    std::cout << "Dynamics infeasibility threshold \t" << dynamics_feas_thresh << "\n";
    std::cout << "Cost convergence threshold \t" << cost_thresh << "\n\n";
  }
};

namespace cafe_facade {
struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
inline void check(int rc) { if (rc != 0) throw Error(rc, cafe_last_error()); }
}  // namespace cafe_facade

inline CafeOptions cafe_options_from_hsddp(const HSDDP_OPTION& s) {
  CafeOptions o;
  o.alpha = s.alpha; o.gamma = s.gamma; o.update_penalty = s.update_penalty; o.update_relax = s.update_relax;
  o.update_regularization = s.update_regularization; o.update_ReB = s.update_ReB;
  o.max_DDP_iter = s.max_DDP_iter; o.max_AL_iter = s.max_AL_iter; o.max_DDP_iter_runtime = s.max_DDP_iter_runtime; o.max_AL_iter_runtime = s.max_AL_iter_runtime;
  o.cost_thresh = s.cost_thresh; o.tconstr_thresh = s.tconstr_thresh; o.pconstr_thresh = s.pconstr_thresh; o.dynamics_feas_thresh = s.dynamics_feas_thresh;
  o.merit_rho = s.merit_rho; o.merit_scale = s.merit_scale; o.merit_offset = s.merit_offset;
  o.AL_active = s.AL_active ? 1 : 0; o.ReB_active = s.ReB_active ? 1 : 0; o.smooth_active = s.smooth_active ? 1 : 0; o.MS = s.MS ? 1 : 0;
  o.nsteps_per_node = s.nsteps_per_node;
  return o;
}
inline void cafe_options_to_hsddp(const CafeOptions& o, HSDDP_OPTION& s) {
  s.alpha = o.alpha; s.gamma = o.gamma; s.update_penalty = o.update_penalty; s.update_relax = o.update_relax;
  s.update_regularization = o.update_regularization; s.update_ReB = o.update_ReB;
  s.max_DDP_iter = o.max_DDP_iter; s.max_AL_iter = o.max_AL_iter; s.max_DDP_iter_runtime = o.max_DDP_iter_runtime; s.max_AL_iter_runtime = o.max_AL_iter_runtime;
  s.cost_thresh = o.cost_thresh; s.tconstr_thresh = o.tconstr_thresh; s.pconstr_thresh = o.pconstr_thresh; s.dynamics_feas_thresh = o.dynamics_feas_thresh;
  s.merit_rho = o.merit_rho; s.merit_scale = o.merit_scale; s.merit_offset = o.merit_offset;
  s.AL_active = o.AL_active != 0; s.ReB_active = o.ReB_active != 0; s.smooth_active = o.smooth_active != 0; s.MS = o.MS != 0;
  s.nsteps_per_node = o.nsteps_per_node;
}

// the keys the reference reads (update_regularization and smooth_active keep their defaults: the reference never loads them)
inline void loadHSDDPSetting(const std::string& filename, HSDDP_OPTION& setting) {
  std::cout << "********* loading HSDDP setting from file **********\n" << filename << "\n\n";
  const double update_regularization = setting.update_regularization;
  const bool smooth_active = setting.smooth_active;
  CafeOptions o = cafe_options_from_hsddp(setting);
  cafe_facade::check(cafe_options_load(filename.c_str(), &o));
  cafe_options_to_hsddp(o, setting);
  setting.update_regularization = update_regularization;
  setting.smooth_active = smooth_active;
}
