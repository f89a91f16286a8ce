// TEST INFRASTRUCTURE ONLY (oracle/): fields of lcmtypes/wbTraj_lcmt.lcm as lcm-gen would emit them (no encoder).
#pragma once
#include <cstdint>
#include <vector>
struct wbTraj_lcmt {
  int32_t sz = 0, wb_sz = 0;
  std::vector<double> time, defect;
  std::vector<std::vector<double>> pos, eul, vWorld, eulrate, qJ, qJd, torque, hg, dhg;
  std::vector<std::vector<int32_t>> contact;
};
