// TEST INFRASTRUCTURE ONLY (oracle/): fields of lcmtypes/hkd_problem_data_lcm_t.lcm as lcm-gen would emit them (no encoder).
#pragma once
#include <cstdint>
#include <vector>
struct hkd_problem_data_lcm_t {
  int32_t n_timesteps = 0;
  std::vector<std::vector<float>> contacts, pos_r, eul_r, vel_r, omega_r, qdummy_r, pos, eul, vel, omega, qdummy;
  std::vector<float> times;
};
