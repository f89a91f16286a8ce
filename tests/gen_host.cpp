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
#include "../cafe_mpc_b200/csrc/wb_pieces.h"
extern "C" int gen_eval_wb(const char* name, const double* const* in, double* const* out) {
  const std::string n(name);
  auto o = [&](int k) { return [out, k](int i, double v) { out[k][i] = v; }; };
  using namespace cafe_gen_wb;
  if (n == "wb_terms") wb_terms(in[0], in[1], o(0), o(1), o(2), o(3), o(4), o(5));
  else if (n == "wb_terms_pieces") {  // trunk + four legs onto zero-initialised outputs, composed like the device wrapper does
    wb_terms_trunk(in[0], in[1], BiasDst{out[0]}, MassDst{out[1]});
    wb_terms_leg0(in[0], in[1], BiasDst{out[0]}, MassDst{out[1]}, PlainDst{out[2], 1}, PlainDst{out[3], 1}, PlainDst{out[4], 1}, PlainDst{out[5], 1});
    wb_terms_leg1(in[0], in[1], BiasDst{out[0]}, MassDst{out[1]}, PlainDst{out[2], 1}, PlainDst{out[3], 1}, PlainDst{out[4], 1}, PlainDst{out[5], 1});
    wb_terms_leg2(in[0], in[1], BiasDst{out[0]}, MassDst{out[1]}, PlainDst{out[2], 1}, PlainDst{out[3], 1}, PlainDst{out[4], 1}, PlainDst{out[5], 1});
    wb_terms_leg3(in[0], in[1], BiasDst{out[0]}, MassDst{out[1]}, PlainDst{out[2], 1}, PlainDst{out[3], 1}, PlainDst{out[4], 1}, PlainDst{out[5], 1});
  }
  else if (n == "wb_feet") wb_feet(in[0], in[1], o(0), o(1), o(2));
  else if (n == "wb_kin_partials") {  // the four per-foot pieces, composed like the device wrapper does
    wb_kin_partials_foot0(in[0], in[1], in[2], in[3], PlainDst{out[0], 1}, PlainDst{out[1], 1}, PlainDst{out[2], 1}, JtfDst<0>{out[3], 1});
    wb_kin_partials_foot1(in[0], in[1], in[2], in[3], PlainDst{out[0], 1}, PlainDst{out[1], 1}, PlainDst{out[2], 1}, JtfDst<1>{out[3], 1});
    wb_kin_partials_foot2(in[0], in[1], in[2], in[3], PlainDst{out[0], 1}, PlainDst{out[1], 1}, PlainDst{out[2], 1}, JtfDst<2>{out[3], 1});
    wb_kin_partials_foot3(in[0], in[1], in[2], in[3], PlainDst{out[0], 1}, PlainDst{out[1], 1}, PlainDst{out[2], 1}, JtfDst<3>{out[3], 1});
  } else if (n == "wb_rnea_derivs") {  // trunk + four legs
    wb_rnea_derivs_trunk(in[0], in[1], in[2], RneaDst<0>{out[0], 1}, RneaDst<0>{out[1], 1});
    wb_rnea_derivs_leg0(in[0], in[1], in[2], RneaDst<1>{out[0], 1}, RneaDst<1>{out[1], 1});
    wb_rnea_derivs_leg1(in[0], in[1], in[2], RneaDst<2>{out[0], 1}, RneaDst<2>{out[1], 1});
    wb_rnea_derivs_leg2(in[0], in[1], in[2], RneaDst<3>{out[0], 1}, RneaDst<3>{out[1], 1});
    wb_rnea_derivs_leg3(in[0], in[1], in[2], RneaDst<4>{out[0], 1}, RneaDst<4>{out[1], 1});
  }
  else if (n == "wb_footvel_partial") wb_footvel_partial(in[0], in[1], o(0));
  else return -1;
  return 0;
}
