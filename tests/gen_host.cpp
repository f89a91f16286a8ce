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

// leg-generic routines (gen/wb_leg_gen.h) composed into the same dense outputs through their compact-slot tables, the way the
// cooperative device kernels assemble them (wb_coop.cuh: wb_assemble_terms, k_wb_lq)
#include "../cafe_mpc_b200/csrc/gen/wb_leg_gen.h"
namespace {
const double kP[4][CAFE_WBL_NP] = CAFE_WBL_LEG_CONSTANTS;
const unsigned char kTmKind[] = CAFE_WBL_TM_KIND, kTmRow[] = CAFE_WBL_TM_ROW, kTmCol[] = CAFE_WBL_TM_COL;
const unsigned char kDpKind[] = CAFE_WBL_DP_KIND, kDpRow[] = CAFE_WBL_DP_ROW, kDpCol[] = CAFE_WBL_DP_COL;
inline int gidx(int f, int l) { return l < 6 ? l : 3 * f + l; }
struct Slot { double* p; void operator()(int i, double v) const { p[i] = v; } };
}
extern "C" int gen_eval_wbl(const char* name, const double* const* in, double* const* out) {
  const std::string n(name);
  using namespace cafe_gen_wbl;
  const double *q = in[0], *v = in[1];
  if (n == "wbl_terms") {   // out: nle[18], M[324 lower, ld 18], J[216, ld 12], gam[12], pf[12], vf[12] (zero-initialised by the caller)
    cafe_gen_wb::wb_terms_trunk(q, v, cafe_gen_wb::BiasDst{out[0]}, cafe_gen_wb::MassDst{out[1]});
    for (int f = 0; f < 4; ++f) {
      double ql[9], vl[9], tm[CAFE_WBL_TM_W] = {0};
      for (int i = 0; i < 9; ++i) { ql[i] = q[gidx(f, i)]; vl[i] = v[gidx(f, i)]; }
      wbl_terms_leg(ql, vl, kP[f], Slot{tm + CAFE_WBL_TM_NLE}, Slot{tm + CAFE_WBL_TM_M}, Slot{tm + CAFE_WBL_TM_J}, Slot{tm + CAFE_WBL_TM_GAM},
                    Slot{tm + CAFE_WBL_TM_PF}, Slot{tm + CAFE_WBL_TM_VF});
      for (int e = 0; e < CAFE_WBL_TM_W; ++e) {
        const int kind = kTmKind[e], r = kTmRow[e], c = kTmCol[e];
        if (kind == 0) { if (r < 6) out[0][gidx(f, r)] += tm[e]; else out[0][gidx(f, r)] = tm[e]; }
        else if (kind == 1) { const int idx = gidx(f, r) + 18 * gidx(f, c); if (r < 6 && c < 6) out[1][idx] += tm[e]; else out[1][idx] = tm[e]; }
        else if (kind == 2) out[2][(3 * f + r) + 12 * gidx(f, c)] = tm[e];
        else out[kind][3 * f + r] = tm[e];
      }
    }
  } else if (n == "wbl_derivs") {   // in: q, v, a, F[12]; out: dtau_dq[324], dtau_dv[324], dvq[216], daq[216], dav[216], djtf[324]
    const double *a = in[2], *F = in[3];
    cafe_gen_wb::wb_rnea_derivs_trunk(q, v, a, cafe_gen_wb::RneaDst<0>{out[0], 1}, cafe_gen_wb::RneaDst<0>{out[1], 1});
    for (int f = 0; f < 4; ++f) {
      double ql[9], vl[9], al[9], dp[CAFE_WBL_DP_W] = {0};
      for (int i = 0; i < 9; ++i) { ql[i] = q[gidx(f, i)]; vl[i] = v[gidx(f, i)]; al[i] = a[gidx(f, i)]; }
      Slot o{dp};
      wbl_rnea_q3(ql, vl, al, kP[f], o); wbl_rnea_q4(ql, vl, al, kP[f], o); wbl_rnea_q5(ql, vl, al, kP[f], o); wbl_rnea_q678(ql, vl, al, kP[f], o);
      wbl_rnea_v345(ql, vl, al, kP[f], o); wbl_rnea_v678(ql, vl, al, kP[f], o);
      wbl_kin_q345(ql, vl, al, F + 3 * f, kP[f], o); wbl_kin_q678(ql, vl, al, F + 3 * f, kP[f], o); wbl_kin_v345678(ql, vl, al, F + 3 * f, kP[f], o);
      for (int e = 0; e < CAFE_WBL_DP_W; ++e) {
        const int kind = kDpKind[e], r = kDpRow[e], c = kDpCol[e];
        if (kind <= 1) {
          const int idx = gidx(f, r) + 18 * gidx(f, c);
          const bool shared = r < 6 && c >= 3 && c <= 5, first = (f == 0 && r < 3);
          if (shared && !first) out[kind][idx] += dp[e]; else out[kind][idx] = dp[e];
        } else if (kind <= 4) {
          out[kind][(3 * f + r) + 12 * gidx(f, c)] = dp[e];
        } else {
          const int idx = gidx(f, r) + 18 * gidx(f, c);
          const bool shared = r >= 3 && r <= 5 && c >= 3 && c <= 5;
          if (shared && f != 0) out[5][idx] += dp[e]; else out[5][idx] = dp[e];
        }
      }
    }
  } else return -1;
  return 0;
}
