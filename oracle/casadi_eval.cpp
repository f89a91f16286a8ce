/* casadi_eval.cpp — CPU ORACLE (test infrastructure only): flat access to the reference's
 * CasADi functions in oracle/_ref, used by tests to check the re-emitted device functions. */
#include "casadi_ref.hpp"
#include "oracle_api.h"
#include <string>
#include <map>

extern "C" int cafe_oracle_casadi_eval(const char* name, const double* const* in, double* const* out) {
  static const std::map<std::string, oracle::CasadiFn> table = {
      {"hkinodyn", CASADI_FN(hkinodyn)}, {"hkinodyn_par", CASADI_FN(hkinodyn_par)},
      {"compute_foot_position", CASADI_FN(compute_foot_position)},
      {"comp_foot_jacob_1", CASADI_FN(comp_foot_jacob_1)}, {"comp_foot_jacob_2", CASADI_FN(comp_foot_jacob_2)},
      {"comp_foot_jacob_3", CASADI_FN(comp_foot_jacob_3)}, {"comp_foot_jacob_4", CASADI_FN(comp_foot_jacob_4)},
      {"SRBDynamics", CASADI_FN(SRBDynamics)}, {"SRBDynamicsDerivatives", CASADI_FN(SRBDynamicsDerivatives)},
      {"footVelPartialDq", CASADI_FN(footVelPartialDq)}, {"footAccPartialDq", CASADI_FN(footAccPartialDq)},
      {"footAccPartialDv", CASADI_FN(footAccPartialDv)}, {"footForcePartialDq", CASADI_FN(footForcePartialDq)}};
  auto it = table.find(name);
  if (it == table.end()) return -1;
  casadi_int sz_arg = 0, sz_res = 0, sz_iw = 0, sz_w = 0;
  it->second.work(&sz_arg, &sz_res, &sz_iw, &sz_w);
  oracle::casadi_call(it->second, in, (int)sz_arg, out, (int)sz_res);
  return 0;
}
