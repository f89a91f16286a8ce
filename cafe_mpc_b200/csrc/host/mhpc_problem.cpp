// mhpc_problem.cpp — phase deck for the MHPC (whole-body + single-rigid-body) problem; host-side mirror of
//   MHPCProblem<T>::initialization / prepare_initialization / initialize_parameters / initialize_multiPhaseProblem
//   / create_problem_one_phase / update_resetmap / add_tconstr_one_phase
//                                   /root/reference/MHPC/MHPC-Trajopt/MHPCProblem.cpp:13-250, :403-601
//   loadMHPCConfig                  MHPCProblem.h:67-83      loadCostWeights   MHPCCostUtil.h:10-143
//   load_reb_params / load_al_params   HSDDPSolver/header/ConstraintsBase.h:88-111
//   WBReference / SRBReference      MHPC/MHPC-Trajopt/MHPCReference.cpp:10-76
#include <cmath>
#include <cstring>
#include <stdexcept>
#include "info_reader.h"
#include "problem_builders.h"

namespace cafe {

void wb_lxx_pattern(const double* rec, unsigned long long out[21]) {
  for (int i = 0; i < 21; ++i) out[i] = 0;
  auto setbit = [&](int r, int c) { const int e = r + 36 * c; out[e >> 6] |= 1ULL << (e & 63); };
  for (int d = 0; d < 36; ++d) setbit(d, d);
  const int base[9] = {3, 4, 5, 18, 19, 20, 21, 22, 23};
  for (int a = 0; a < 9; ++a) for (int c = 0; c < 9; ++c) setbit(base[a], base[c]);
  for (int f = 0; f < 4; ++f) {
    int cols[15];
    for (int a = 0; a < 3; ++a) { cols[a] = 3 + a; cols[3 + a] = 6 + 3 * f + a; cols[12 + a] = 24 + 3 * f + a; }
    for (int a = 0; a < 6; ++a) cols[6 + a] = 18 + a;
    const int nc = rec[CAFE_REF_CONTACT + f] > 0 ? 6 : 15;
    for (int a = 0; a < nc; ++a) for (int c = 0; c < nc; ++c) setbit(cols[a], cols[c]);
  }
}


void loadMHPCConfig(const std::string& fname, MHPCConfig& c) {
  InfoFile pt(fname);
  c.plan_dur_wb = pt.num("config.plan_dur_wb");  // doubles (MHPCProblem.h:25-35): the whole-body and SRB time steps reach the dynamics,
  c.plan_dur_srb = pt.num("config.plan_dur_srb");  // the costs and the trajectories as 0.01 and 0.05 exactly, not float-rounded
  c.dt_mpc = (float)pt.num("config.dt_mpc");     // float (MHPCProblem.h:38)
  c.dt_wb = pt.num("config.dt_wb");
  c.dt_srb = pt.num("config.dt_srb");
  c.BG_alpha = (double)(float)pt.num("config.BG_alpha");
  c.num_threads = pt.integer("config.nthreads");
  c.referenceFileName = pt.str("config.referenceFile");
  c.costFileName = pt.str("config.costFile");
  c.constraintParamFileName = pt.str("config.constraintParamFile");
}

void MHPCProblem::set_problem_data(QuadReference* quad_ref, const MHPCConfig& config, const std::string& settings_root) {
  quad_reference = quad_ref;
  pconfig = config;
  root = settings_root;
  plan_dur_all = (float)(config.plan_dur_wb + config.plan_dur_srb);  // MHPCProblem.h:209 (a float member)
}

static CafeRebParam reb_params(const InfoFile& pt, const std::string& t) { return CafeRebParam{pt.num(t + "_ReB.delta"), pt.num(t + "_ReB.delta_min"), pt.num(t + "_ReB.eps")}; }

static void fill_wb_record(double* r, const QuadAugmentedState& s) {  // WBReference::get_reference_at_t
  for (int i = 0; i < 6; ++i) { r[CAFE_REF_XR + i] = s.body_state[i]; r[CAFE_REF_XR + 18 + i] = s.body_state[6 + i]; }
  for (int i = 0; i < 12; ++i) {
    r[CAFE_REF_XR + 6 + i] = s.qJ[i];
    r[CAFE_REF_XR + 24 + i] = s.qJd[i];
    r[CAFE_REF_UR + i] = s.torque[i];
    r[CAFE_REF_YR + i] = s.grf[i];
    r[CAFE_REF_PF + i] = s.foot_placements[i];
    r[CAFE_REF_VF + i] = s.foot_velocities[i];
    r[CAFE_REF_QJ + i] = s.qJ[i];
  }
  for (int i = 0; i < 3; ++i) r[CAFE_REF_PCOM + i] = s.body_state[i];
  for (int i = 0; i < 4; ++i) r[CAFE_REF_CONTACT + i] = (double)s.contact[i];
}
static void fill_srb_record(double* r, const QuadAugmentedState& s) {  // SRBReference::get_reference_at_t + MHPCFootStep
  for (int i = 0; i < 12; ++i) { r[CAFE_REF_XR + i] = s.body_state[i]; r[CAFE_REF_UR + i] = s.grf[i]; r[CAFE_REF_PF + i] = s.foot_placements[i]; r[CAFE_REF_VF + i] = s.foot_velocities[i]; r[CAFE_REF_QJ + i] = s.qJ[i]; }
  for (int i = 0; i < 3; ++i) r[CAFE_REF_PCOM + i] = s.body_state[i];
  for (int i = 0; i < 4; ++i) r[CAFE_REF_CONTACT + i] = (double)s.contact[i];
}

void MHPCProblem::initialization(DeckStorage& out) {
  quad_reference->initialize(plan_dur_all);
  CafeDeck& deck = out.deck;
  std::memset(&deck, 0, sizeof(deck));
  out.phase_start_times.clear();
  out.phase_end_times.clear();
  if (approx_leq_scalar(plan_dur_all, .0)) throw std::runtime_error("total planning horizon cannot be zero");

  /* ---- prepare_initialization: cut the WB phases along the reference contact schedule (MHPCProblem.cpp:93-136) */
  int n_wb = 0;
  if (pconfig.plan_dur_wb > 1e-5) {
    float start = 0.0f, end = 0.0f, t = 0.0f;
    int contact_prev[4], contact_cur[4];
    quad_reference->get_contact_at_t(contact_prev, t);
    while (approx_leq_scalar(t, pconfig.plan_dur_wb)) {
      quad_reference->get_contact_at_t(contact_cur, t);
      bool change = false;
      for (int l = 0; l < 4; ++l) change = change || (contact_cur[l] != contact_prev[l]);
      if (change || approx_eq_scalar(t, pconfig.plan_dur_wb)) {
        end = t;
        if (n_wb >= CAFE_MAX_PHASES - 1) throw std::runtime_error("too many phases");
        CafePhase& ph = deck.phase[n_wb];
        ph.model = CAFE_MODEL_WB;
        ph.horizon = (int)std::round((end - start) / pconfig.dt_wb);
        for (int l = 0; l < 4; ++l) ph.contact[l] = contact_prev[l];
        out.phase_start_times.push_back(start);
        out.phase_end_times.push_back(end);
        n_wb++;
        for (int l = 0; l < 4; ++l) contact_prev[l] = contact_cur[l];
        start = end;
      }
      t = (float)(t + pconfig.dt_wb);   // float t += double
    }
  }
  int n_srb = 0, srb_h = 0;
  float srb_start = 0;
  if (pconfig.plan_dur_srb > 1e-5) {
    srb_start = (float)pconfig.plan_dur_wb;
    srb_h = (int)std::round(pconfig.plan_dur_srb / pconfig.dt_srb);
    n_srb = srb_h > 0 ? 1 : 0;
  }
  deck.n_phases = n_wb + n_srb;
  deck.BG_alpha = pconfig.BG_alpha;
  deck.hip_yaw = 3.1415;  // urdf/mini_cheetah_simple_correctedInertia.urdf:79 (rpy="0.0 0.0 3.1415")

  /* ---- initialize_parameters (MHPCProblem.cpp:149-171) */
  InfoFile cpt(root + "/" + pconfig.constraintParamFileName);
  /* LocoProblem::initialize_parameters (LocoProblem.cpp:7-26) loads GRF, Torque and TD only; its settings file has no other block */
  const CafeRebParam grf = reb_params(cpt, "GRF"), torque = reb_params(cpt, "Torque");
  const CafeRebParam joint = loco ? CafeRebParam{1, 1, 0} : reb_params(cpt, "Joint"), minh = loco ? CafeRebParam{1, 1, 0} : reb_params(cpt, "MinHeight");
  CafeAlParam td{};
  td.sigma = cpt.num("TD_AL.sigma"); td.lambda = cpt.num("TD_AL.lambda"); td.sigma_max = cpt.num("TD_AL.sigma_max");
  JsonWeights cw(root + "/" + pconfig.costFileName);

  int rec = 0;
  for (int i = 0; i < n_wb; ++i) { deck.phase[i].knot_offset = rec; rec += deck.phase[i].horizon + 1; }
  if (n_srb) { CafePhase& s = deck.phase[n_wb]; s.model = CAFE_MODEL_SRB; s.horizon = srb_h; s.knot_offset = rec; rec += srb_h + 1; }
  deck.n_records = rec;
  out.ref.assign((size_t)rec * CAFE_REF_W, 0.0);

  /* ---- WB phases (MHPCProblem.cpp:176-217, :403-601) */
  for (int i = 0; i < n_wb; ++i) {
    CafePhase& ph = deck.phase[i];
    ph.dt = pconfig.dt_wb;   // Trajectory<T>(pconfig->dt_wb, h): T timeStep, the dt of dynamics and costs (MHPCProblem.cpp:183, :410)
    ph.t_offset = out.phase_start_times[i] - out.phase_start_times[0];
    ph.has_reset = 1;
    if (i < n_wb - 1) for (int l = 0; l < 4; ++l) ph.next_contact[l] = deck.phase[i + 1].contact[l];
    else quad_reference->get_contact_at_t(ph.next_contact, pconfig.plan_dur_wb + pconfig.dt_mpc);
    ph.next_model = (i < n_wb - 1) ? CAFE_MODEL_WB : (n_srb ? CAFE_MODEL_SRB : -1);
    ph.n_td = 0;
    for (int l = 0; l < 4; ++l) if (ph.contact[l] == 0 && ph.next_contact[l] == 1) ph.td_foot[ph.n_td++] = l;
    /* weights: [qw_qB, qw_qJ x4, qw_vB, qw_vJ x4 | rw x12 | qfw...] (MHPCCostUtil.h:22-80) */
    auto fill = [&](double* dst, const char* qB, const char* qJ, const char* vB, const char* vJ) {
      const auto &a = cw.vec(std::string("WB_Tracking_Cost.") + qB), &b = cw.vec(std::string("WB_Tracking_Cost.") + qJ),
                 &c = cw.vec(std::string("WB_Tracking_Cost.") + vB), &d = cw.vec(std::string("WB_Tracking_Cost.") + vJ);
      for (int j = 0; j < 6; ++j) { dst[j] = a.at(j); dst[18 + j] = c.at(j); }
      for (int l = 0; l < 4; ++l) for (int j = 0; j < 3; ++j) { dst[6 + 3 * l + j] = b.at(j); dst[24 + 3 * l + j] = d.at(j); }
    };
    fill(ph.q, "qw_qB", "qw_qJ", "qw_vB", "qw_vJ");
    fill(ph.qf, "qfw_qB", "qfw_qJ", "qfw_vB", "qfw_vJ");
    for (int j = 0; j < 12; ++j) ph.r[j] = cw.num("WB_Tracking_Cost.rw");
    for (int j = 0; j < 3; ++j) {
      ph.w_footreg[j] = cw.vec("WB_FootPlace_Reg.qw_per_foot").at(j);
      ph.w_swingpos[j] = cw.vec("Swing_Pos_Tracking.qw_per_foot").at(j);
      ph.w_swingvel[j] = cw.vec("Swing_Vel_Tracking.qw_per_foot").at(j);
    }
    ph.w_tdvel[0] = 0; ph.w_tdvel[1] = 0; ph.w_tdvel[2] = 1.0;  // TDVelocityPenalty::qFoot (MHPCCost.h:196)
    ph.reb_grf = grf; ph.reb_torque = torque; ph.reb_joint = joint; ph.reb_minheight = minh; ph.al_td = td;
    ph.mu = 0.6;              // MHPCConstraint.cpp:11
    ph.ground_height = 0;
    ph.h_min = 0.20;          // MHPCConstraint.h:148
    ph.torque_limit = 17.0;   // MHPCConstraint.cpp:77
    ph.no_joint_limit = ph.no_min_height = loco ? 1 : 0;  // LocoProblem.cpp:64-82: TorqueLimit + WBGRF only
    const double lb[3] = {-1.3, -5.0, -M_PI}, ub[3] = {1.3, 5.0, M_PI};  // MHPCConstraint.cpp:172-173
    for (int j = 0; j < 3; ++j) { ph.joint_lb[j] = lb[j]; ph.joint_ub[j] = ub[j]; }
    for (int k = 0; k <= ph.horizon; ++k) {
      float t_cost = (float)((double)ph.t_offset + (double)k * ph.dt);
      float t_init = (float)(out.phase_start_times[i] + k * pconfig.dt_wb);
      if (quad_reference->index_at_t(t_cost) != quad_reference->index_at_t(t_init)) throw std::runtime_error("reference index mismatch (WB)");
      fill_wb_record(&out.ref[(size_t)(ph.knot_offset + k) * CAFE_REF_W], *quad_reference->get_a_reference_ptr_at_t(t_cost));
    }
  }
  /* ---- SRB phase (MHPCProblem.cpp:219-249, :487-521) */
  if (n_srb) {
    CafePhase& ph = deck.phase[n_wb];
    ph.dt = pconfig.dt_srb;
    ph.t_offset = srb_start;
    ph.has_reset = 0;
    ph.next_model = -1;
    ph.n_td = 0;
    for (int l = 0; l < 4; ++l) { ph.contact[l] = 0; ph.next_contact[l] = 0; }
    const auto &a = cw.vec("SRB_Tracking_Cost.qw_qB"), &c = cw.vec("SRB_Tracking_Cost.qw_vB"), &af = cw.vec("SRB_Tracking_Cost.qfw_qB"), &cf = cw.vec("SRB_Tracking_Cost.qfw_vB");
    for (int j = 0; j < 6; ++j) { ph.q[j] = a.at(j); ph.q[6 + j] = c.at(j); ph.qf[j] = af.at(j); ph.qf[6 + j] = cf.at(j); }
    for (int j = 0; j < 12; ++j) ph.r[j] = cw.num("SRB_Tracking_Cost.rw");
    ph.reb_minheight = minh;
    ph.h_min = 0.18;  // MHPCConstraint.h:199
    for (int k = 0; k <= ph.horizon; ++k) {
      float t_cost = (float)((double)ph.t_offset + (double)k * ph.dt);
      float t_init = (float)(srb_start + k * pconfig.dt_srb);
      if (quad_reference->index_at_t(t_cost) != quad_reference->index_at_t(t_init)) throw std::runtime_error("reference index mismatch (SRB)");
      fill_srb_record(&out.ref[(size_t)(ph.knot_offset + k) * CAFE_REF_W], *quad_reference->get_a_reference_ptr_at_t(t_cost));
    }
  }
  deck.ref = out.ref.data();
}

}  // namespace cafe

/* ================= in-place barrel roll (MHPC/MHPC-Trajopt/BarrelRoll/BarrelRollTO.cpp:65-275) =================
 * Six hand-scheduled whole-body phases, no reference file: the tracking cost of phase i pulls towards a fixed desired state
 * xf_des[i] (set_reference_state, :171-174) with its own weight set (cost_phase_{i+1}), the barriers are the BarrelRoll:: copies
 * (torque, joint speed, joint, min height 0.13, GRF), the two flight phases end with a four-foot touchdown constraint (:232-241).
 * No foot regularisation / swing tracking / touchdown-velocity cost: those weights are zero in the deck (they add exact zeros). */
namespace cafe {

static const double kBarrelSwitch[7] = {0.0, 0.12, 0.33, 0.75, 0.90, 1.10, 1.25};  // BarrelRollTO.cpp:70
static const int kBarrelContact[6][4] = {{1, 1, 1, 1}, {0, 1, 0, 1}, {0, 0, 0, 0}, {1, 1, 1, 1}, {0, 0, 0, 0}, {1, 1, 1, 1}};  // :76-81

static void barrel_desired_states(double xd[6][36]) {  // load_desired_final_states, BarrelRollTO.cpp:277-339 (x = pos, eul, qJ, vWorld, euld, qJd)
  double pos[3] = {0, 0, 0}, eul[3] = {0, 0, 0}, vW[3] = {0, 0, 0}, euld[3] = {0, 0, 0}, qJ[12], qJd[12];
  for (int l = 0; l < 4; ++l) { qJ[3 * l] = 0; qJ[3 * l + 1] = -1.2; qJ[3 * l + 2] = 2.4; }
  for (int i = 0; i < 12; ++i) qJd[i] = 0;
  auto put = [&](int i) {
    for (int j = 0; j < 3; ++j) { xd[i][j] = pos[j]; xd[i][3 + j] = eul[j]; xd[i][18 + j] = vW[j]; xd[i][21 + j] = euld[j]; }
    for (int j = 0; j < 12; ++j) { xd[i][6 + j] = qJ[j]; xd[i][24 + j] = qJd[j]; }
  };
  pos[0] = 0; pos[1] = -0.15; pos[2] = 0.26; eul[0] = 0; eul[1] = 0; eul[2] = M_PI / 6; euld[2] = 3.0 * M_PI; vW[0] = 0; vW[1] = -1.0; vW[2] = 2.0;
  put(0);
  pos[0] = 0; pos[1] = -0.25; pos[2] = 0.33; eul[0] = 0; eul[1] = 0; eul[2] = 0.5 * M_PI; euld[0] = 0; euld[1] = 0; euld[2] = 3.0 * M_PI; vW[0] = 0; vW[1] = -1.2; vW[2] = 2.0;
  { const double q[12] = {M_PI / 6, -1.0, 2.0, -M_PI / 5, -0.5, 1.0, M_PI / 6, -1.0, 2.0, -M_PI / 5, -0.5, 1.0}; for (int j = 0; j < 12; ++j) qJ[j] = q[j]; }
  put(1);
  pos[0] = 0.0; pos[1] = -0.55; pos[2] = 0.22; eul[0] = 0; eul[1] = 0; eul[2] = 2.0 * M_PI; euld[0] = 0; euld[1] = 0; euld[2] = 3.0 * M_PI; vW[0] = 0.0; vW[1] = -1.5; vW[2] = -2.5;
  { const double q[12] = {0.3, -1.1, 2.2, -0.3, -1.1, 2.2, 0.3, -1.1, 2.2, -0.3, -1.1, 2.2}; for (int j = 0; j < 12; ++j) qJ[j] = q[j]; }
  put(2);
  pos[2] = 0.25; eul[2] = 2 * M_PI; euld[2] = 0; vW[0] = 0; vW[1] = 0; vW[2] = 0;
  put(3);
  for (int l = 0; l < 4; ++l) { qJ[3 * l] = 0; qJ[3 * l + 1] = -1.0; qJ[3 * l + 2] = 2.0; }
  put(4);
  put(5);
}

void build_barrel_to_deck(const std::string& cost_json, const std::string& constraint_info, DeckStorage& out) {
  CafeDeck& deck = out.deck;
  std::memset(&deck, 0, sizeof(deck));
  out.phase_start_times.clear(); out.phase_end_times.clear();
  const double dt = 0.01;  // BarrelRollTO.cpp:68
  deck.n_phases = 6;
  deck.BG_alpha = 10.0;    // :89
  deck.hip_yaw = 3.1415;
  InfoFile cpt(constraint_info);
  JsonWeights cw(cost_json);
  const CafeRebParam grf = reb_params(cpt, "GRF"), torque = reb_params(cpt, "Torque"), jv = reb_params(cpt, "JointVel"), joint = reb_params(cpt, "Joint"), minh = reb_params(cpt, "MinHeight");
  CafeAlParam td{};
  td.sigma = cpt.num("TD_AL.sigma"); td.lambda = cpt.num("TD_AL.lambda"); td.sigma_max = cpt.num("TD_AL.sigma_max");
  double xd[6][36];
  barrel_desired_states(xd);
  int rec = 0;
  for (int i = 0; i < 6; ++i) {
    CafePhase& ph = deck.phase[i];
    ph.model = CAFE_MODEL_WB;
    ph.horizon = (int)std::round((kBarrelSwitch[i + 1] - kBarrelSwitch[i]) / dt);  // :125
    ph.knot_offset = rec; rec += ph.horizon + 1;
    out.phase_start_times.push_back((float)kBarrelSwitch[i]); out.phase_end_times.push_back((float)kBarrelSwitch[i + 1]);
  }
  deck.n_records = rec;
  out.ref.assign((size_t)rec * CAFE_REF_W, 0.0);
  for (int i = 0; i < 6; ++i) {
    CafePhase& ph = deck.phase[i];
    ph.dt = dt;
    ph.t_offset = (float)kBarrelSwitch[i];  // set_time_offset(switching_times[i]), :129
    ph.has_reset = i < 5 ? 1 : 0;            // :157-169
    ph.next_model = i < 5 ? CAFE_MODEL_WB : -1;
    for (int l = 0; l < 4; ++l) { ph.contact[l] = kBarrelContact[i][l]; ph.next_contact[l] = kBarrelContact[i < 5 ? i + 1 : i][l]; }
    ph.n_td = 0;
    if (i == 2 || i == 4) for (int l = 0; l < 4; ++l) ph.td_foot[ph.n_td++] = l;  // TouchDown(Vec4<int>::Ones()), :232-241
    const std::string sec = "cost_phase_" + std::to_string(i + 1) + ".";  // load_cost_weights, :341-412
    auto fill = [&](double* dst, const char* qB, const char* qJ, const char* vB, const char* vJ) {
      const auto &a = cw.vec(sec + qB), &b = cw.vec(sec + qJ), &c = cw.vec(sec + vB), &d = cw.vec(sec + vJ);
      for (int j = 0; j < 6; ++j) { dst[j] = a.at(j); dst[18 + j] = c.at(j); }
      for (int l = 0; l < 4; ++l) for (int j = 0; j < 3; ++j) { dst[6 + 3 * l + j] = b.at(j); dst[24 + 3 * l + j] = d.at(j); }
    };
    fill(ph.q, "qw_qB", "qw_qJ", "qw_vB", "qw_vJ");
    fill(ph.qf, "qfw_qB", "qfw_qJ", "qfw_vB", "qfw_vJ");
    for (int j = 0; j < 12; ++j) ph.r[j] = cw.num(sec + "rw");
    ph.reb_grf = grf; ph.reb_torque = torque; ph.reb_joint = joint; ph.reb_minheight = minh; ph.al_td = td;
    ph.joint_speed_limit = 1; ph.reb_jointvel = jv; ph.jointvel_lb = -20.0; ph.jointvel_ub = 20.0;  // BarrelRollConstraints.h:71-72
    ph.mu = 0.6;             // BarrelRollConstraints.cpp:11
    ph.ground_height = 0;    // :236
    ph.h_min = 0.13;         // BarrelRollConstraints.h:147
    ph.torque_limit = 17.0;  // BarrelRollConstraints.cpp:77
    const double lb[3] = {-1.3, -5.0, -M_PI}, ub[3] = {1.3, 5.0, M_PI};  // :124-125
    for (int j = 0; j < 3; ++j) { ph.joint_lb[j] = lb[j]; ph.joint_ub[j] = ub[j]; }
    for (int k = 0; k <= ph.horizon; ++k) {
      double* r = &out.ref[(size_t)(ph.knot_offset + k) * CAFE_REF_W];
      for (int j = 0; j < 36; ++j) r[CAFE_REF_XR + j] = xd[i][j];  // set_reference_state(xf_des[i], 0), :173
      for (int l = 0; l < 4; ++l) r[CAFE_REF_CONTACT + l] = (double)ph.contact[l];
    }
  }
  deck.ref = out.ref.data();
}

// Initial state trajectory of BarrelRollTO.cpp:131-147: phase i interpolates linearly from the previous phase's desired state (phase 0:
// from the initial state) to its own over the phase duration, with the reference's float time arithmetic. Writes Xbar into packed
// solution records (cafe_solution_size doubles each; controls and gains stay zero).
void barrel_to_guess(const CafeDeck& deck, const double* x0, double* guess) {
  size_t off = 0;
  for (int i = 0; i < deck.n_phases; ++i) {
    const CafePhase& ph = deck.phase[i];
    const long n = 36, m = 12, p = 12, h = ph.horizon;
    const double* x1 = deck.ref + (size_t)ph.knot_offset * CAFE_REF_W + CAFE_REF_XR;
    const double* xa = i == 0 ? x0 : deck.ref + (size_t)deck.phase[i - 1].knot_offset * CAFE_REF_W + CAFE_REF_XR;
    float t = 0.0;
    const float t_dur = kBarrelSwitch[i + 1] - kBarrelSwitch[i];
    for (int k = 0; k <= h; ++k) {
      const float s = t / t_dur;
      for (int j = 0; j < 36; ++j) guess[off + (size_t)k * n + j] = xa[j] + (x1[j] - xa[j]) * s;  // lerp_eigen_vectors, :43-53
      t += 0.01;
    }
    off += (h + 1) * n + h * m + h * p + h * m + h * m * n + h * m + h * m * m + h * m * n + (h + 1) * n;
  }
}

}  // namespace cafe
