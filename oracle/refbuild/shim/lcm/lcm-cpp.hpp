// TEST INFRASTRUCTURE ONLY (oracle/): LCM is not in this image. publish() hands the message to an optional in-process listener (the reference's
// stand-alone programs publish their result as LCM messages: that is where ref_program_driver.cpp picks it up); nothing goes on a network.
#pragma once
#include <string>
#include <typeinfo>
#include <unistd.h>   // the real lcm-cpp.hpp brings it in (the reference calls sleep() without including it)
namespace lcm {
struct Listener { virtual void on_publish(const std::string& channel, const void* msg, const std::type_info& type) = 0; virtual ~Listener() {} };
inline Listener*& listener() { static Listener* l = nullptr; return l; }
class LCM {
 public:
  explicit LCM(const std::string& = "") {}
  bool good() const { return true; }
  template <class M> int publish(const std::string& channel, const M* msg) { if (listener()) listener()->on_publish(channel, msg, typeid(M)); return 0; }
  int handle() { return 0; }
  int handleTimeout(int) { return 0; }
};
}  // namespace lcm
