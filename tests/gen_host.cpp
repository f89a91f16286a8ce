// TEST INFRASTRUCTURE: host build of the re-emitted model functions (cafe_mpc_b200/csrc/gen/*.h) so that
// they can be compared on the CPU with the reference's own CasADi C (oracle/_ref).
#include <cstring>
#include <string>
#include "../cafe_mpc_b200/csrc/gen/hkd_gen.h"
#include "../cafe_mpc_b200/csrc/gen/srb_gen.h"

extern "C" int gen_eval(const char* name, const double* const* in, double* const* out) {
  const std::string n(name);
  auto o0 = [&](int i, double v) { out[0][i] = v; };
  auto o1 = [&](int i, double v) { out[1][i] = v; };
  using namespace cafe_gen_hkd;
  using namespace cafe_gen_srb;
  if (n == "hkinodyn") hkinodyn(in[0], in[1], in[2], in[3], o0);
  else if (n == "hkinodyn_par") hkinodyn_par(in[0], in[1], in[2], in[3], o0, o1);
  else if (n == "foot_position_1") foot_position_1(in[0], in[1], in[2], in[3], o0);
  else if (n == "foot_position_2") foot_position_2(in[0], in[1], in[2], in[3], o0);
  else if (n == "foot_position_3") foot_position_3(in[0], in[1], in[2], in[3], o0);
  else if (n == "foot_position_4") foot_position_4(in[0], in[1], in[2], in[3], o0);
  else if (n == "foot_jacobian_1") foot_jacobian_1(in[0], in[1], in[2], o0);
  else if (n == "foot_jacobian_2") foot_jacobian_2(in[0], in[1], in[2], o0);
  else if (n == "foot_jacobian_3") foot_jacobian_3(in[0], in[1], in[2], o0);
  else if (n == "foot_jacobian_4") foot_jacobian_4(in[0], in[1], in[2], o0);
  else if (n == "srb_dynamics") srb_dynamics(in[0], in[1], in[2], in[3], o0);
  else if (n == "srb_dynamics_derivatives") srb_dynamics_derivatives(in[0], in[1], in[2], in[3], o0, o1);
  else return -1;
  return 0;
}

// whole-body generated routines (only the ones instantiated here are compiled)
#include "../cafe_mpc_b200/csrc/gen/wb_gen.h"
extern "C" int gen_eval_wb(const char* name, const double* const* in, double* const* out) {
  const std::string n(name);
  auto o = [&](int k) { return [out, k](int i, double v) { out[k][i] = v; }; };
  using namespace cafe_gen_wb;
  if (n == "wb_terms") wb_terms(in[0], in[1], o(0), o(1), o(2), o(3), o(4), o(5));
  else if (n == "wb_feet") wb_feet(in[0], in[1], o(0), o(1), o(2));
  else if (n == "wb_kin_partials") wb_kin_partials(in[0], in[1], in[2], in[3], o(0), o(1), o(2), o(3));
  else if (n == "wb_footvel_partial") wb_footvel_partial(in[0], in[1], o(0));
  else return -1;
  return 0;
}
