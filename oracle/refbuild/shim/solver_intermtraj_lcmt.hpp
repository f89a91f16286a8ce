// TEST INFRASTRUCTURE ONLY (oracle/): fields of lcmtypes/solver_intermtraj_lcmt.lcm as lcm-gen would emit them (no encoder).
#pragma once
#include <cstdint>
#include <vector>
struct solver_intermtraj_lcmt {
  int32_t tau_sz = 0, x_sz = 0, u_sz = 0;
  std::vector<std::vector<float>> x_tau, u_tau;
};
