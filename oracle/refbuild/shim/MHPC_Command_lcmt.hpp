// TEST INFRASTRUCTURE ONLY (oracle/): fields of lcmtypes/MHPC_Command_lcmt.lcm as lcm-gen would emit them (no encoder).
#pragma once
#include <cstdint>
#include <vector>
struct MHPC_Command_lcmt {
  int32_t N_mpcsteps = 0;
  std::vector<float> mpc_times;
  std::vector<std::vector<float>> torque, eul, pos, qJ, vWorld, eulrate, qJd, GRF, feedback, Qu, Quu, Qux, statusTimes;
  std::vector<std::vector<int32_t>> contacts;
};
