// TEST INFRASTRUCTURE ONLY (oracle/): LCM is not in this image; the solver only publishes debug trajectories through it.
#pragma once
#include <string>
namespace lcm {
class LCM {
 public:
  explicit LCM(const std::string& = "") {}
  bool good() const { return true; }
  template <class M> int publish(const std::string&, const M*) { return 0; }
  int handle() { return 0; }
  int handleTimeout(int) { return 0; }
};
}  // namespace lcm
