// hkd_problem.cpp — phase deck for the HKD trajectory optimisation; host-side mirror of
// HKDProblem<T>::initialization / create_problem_one_phase / add_tconstr_one_phase
// (/root/reference/HKDMPC/HKD-TrajOpt/HKDProblem.cpp:15-111, :224-311), HKDTrackingCost
// (HKDCost.h:8-38), HKDFootPlaceReg weights (HKDCost.h:52-75), HKDSinglePhaseReference
// (HKDReference.cpp:8-62) and loadConstrintParameters (HKDProblem.h:70-90).
#include <cmath>
#include <cstring>
#include <stdexcept>
#include "info_reader.h"
#include "problem_builders.h"
#include "../gen/hkd_gen.h"

namespace cafe {

void hkd_lq_patterns(unsigned long long out[36]) {
  for (int i = 0; i < 36; ++i) out[i] = 0;
  auto setbit = [&](int which, int e) { out[9 * which + (e >> 6)] |= 1ULL << (e & 63); };
  double x[24], u[24], c[4] = {1, 0, 1, 0};
  const double dt = 0.01;
  for (int i = 0; i < 24; ++i) { x[i] = 0.1 + 0.01 * i; u[i] = 0.2 - 0.01 * i; }
  cafe_gen_hkd::hkinodyn_par(x, u, &dt, c, [&](int i, double) { setbit(0, i); }, [&](int i, double) { setbit(1, i); });
  for (int d = 0; d < 24; ++d) setbit(2, d + 24 * d);
  for (int leg = 0; leg < 4; ++leg)
    for (int a = 0; a < 2; ++a) { setbit(2, (3 + a) + 24 * (12 + 3 * leg + a)); setbit(2, (12 + 3 * leg + a) + 24 * (3 + a)); }
  for (int leg = 0; leg < 4; ++leg)
    for (int r = 0; r < 3; ++r) for (int cc = 0; cc < 3; ++cc) setbit(3, (3 * leg + r) + 24 * (3 * leg + cc));
  for (int d = 12; d < 24; ++d) setbit(3, d + 24 * d);
}


void load_hsddp_setting(const std::string& fname, CafeOptions& o) {
  InfoFile pt(fname);
  std::memset(&o, 0, sizeof(o));
  o.alpha = pt.num("ddp.alpha");
  o.gamma = pt.num("ddp.gamma");
  o.update_penalty = pt.num("ddp.update_penalty");
  o.update_relax = pt.num("ddp.update_relax");
  o.update_ReB = pt.num("ddp.update_ReB");
  o.update_regularization = 2;  // never read from file by the reference (HSDDP_CompoundTypes.h:19,57-82)
  o.max_DDP_iter = pt.integer("ddp.max_DDP_iter");
  o.max_AL_iter = pt.integer("ddp.max_AL_iter");
  o.max_DDP_iter_runtime = pt.integer("ddp.max_DDP_iter_runtime");
  o.max_AL_iter_runtime = pt.integer("ddp.max_AL_iter_runtime");
  o.cost_thresh = pt.num("ddp.cost_thresh");
  o.tconstr_thresh = pt.num("ddp.tconstr_thresh");
  o.pconstr_thresh = pt.num("ddp.pconstr_thresh");
  o.dynamics_feas_thresh = pt.num("ddp.dynamics_feas_thresh");
  o.merit_rho = pt.num("ddp.merit_rho");
  o.merit_scale = pt.num("ddp.merit_scale");
  o.merit_offset = pt.num("ddp.merit_offset");
  o.AL_active = pt.boolean("ddp.AL_active");
  o.ReB_active = pt.boolean("ddp.ReB_active");
  o.smooth_active = 0;  // likewise not read from file
  o.MS = pt.boolean("ddp.MS");
  o.nsteps_per_node = pt.integer("ddp.nsteps_per_node");
}

void compute_hkd_state(const double eul[3], const double pos[3], const double qJ[12], double qdummy[12], const int contact[4]) {
  for (int l = 0; l < 4; ++l) {
    const double* ql = qJ + 3 * l;
    if (contact[l] == 0) { for (int a = 0; a < 3; ++a) qdummy[3 * l + a] = ql[a]; continue; }
    double pf[3] = {0, 0, 0};
    auto st = [&](int i, double v) { pf[i] = v; };
    switch (l) {
      case 0: cafe_gen_hkd::foot_position_1(pos, eul, ql, (const double*)nullptr, st); break;
      case 1: cafe_gen_hkd::foot_position_2(pos, eul, ql, (const double*)nullptr, st); break;
      case 2: cafe_gen_hkd::foot_position_3(pos, eul, ql, (const double*)nullptr, st); break;
      default: cafe_gen_hkd::foot_position_4(pos, eul, ql, (const double*)nullptr, st); break;
    }
    for (int a = 0; a < 3; ++a) qdummy[3 * l + a] = pf[a];
  }
}

void HKDProblem::set_problem_data(QuadReference* quad_ref, const HKDPlanConfig& config, const std::string& constraint_params_fname) {
  quad_ref_ptr = quad_ref;
  plan_duration = config.plan_duration;
  dt_sim = config.timeStep;
  nsteps_between_mpc = config.nsteps_between_mpc;
  dt_mpc = dt_sim * nsteps_between_mpc;
  InfoFile pt(constraint_params_fname);
  grf_reb_param = {pt.num("GRF_ReB.delta"), pt.num("GRF_ReB.delta_min"), pt.num("GRF_ReB.eps")};
  swing_reb_param = {pt.num("Swing_ReB.delta"), pt.num("Swing_ReB.delta_min"), pt.num("Swing_ReB.eps")};
  td_al_param.sigma = pt.num("TD_AL.sigma");
  td_al_param.lambda = pt.num("TD_AL.lambda");
  td_al_param.sigma_max = pt.num("TD_AL.sigma_max");
}

// HKDSinglePhaseReference::get_reference_at_t + the look-ups of HKDFootPlaceReg
static void fill_hkd_record(double* r, const QuadAugmentedState& s) {
  for (int i = 0; i < 3; ++i) {
    r[CAFE_REF_XR + i] = s.body_state[3 + i];
    r[CAFE_REF_XR + 3 + i] = s.body_state[i];
    r[CAFE_REF_XR + 6 + i] = s.body_state[9 + i];
    r[CAFE_REF_XR + 9 + i] = s.body_state[6 + i];
    r[CAFE_REF_PCOM + i] = s.body_state[i];
  }
  for (int leg = 0; leg < 4; ++leg)
    for (int a = 0; a < 3; ++a)
      r[CAFE_REF_XR + 12 + 3 * leg + a] = s.contact[leg] > 0 ? s.foot_placements[3 * leg + a] : s.qJ[3 * leg + a];
  for (int i = 0; i < 12; ++i) {
    r[CAFE_REF_UR + i] = s.grf[i];
    r[CAFE_REF_UR + 12 + i] = s.qJd[i];
    r[CAFE_REF_PF + i] = s.foot_placements[i];
    r[CAFE_REF_VF + i] = s.foot_velocities[i];
    r[CAFE_REF_QJ + i] = s.qJ[i];
  }
  for (int i = 0; i < 4; ++i) r[CAFE_REF_CONTACT + i] = (double)s.contact[i];
}

void HKDProblem::initialization(DeckStorage& out) {
  quad_ref_ptr->initialize(plan_duration);
  CafeDeck& deck = out.deck;
  std::memset(&deck, 0, sizeof(deck));
  out.phase_start_times.clear();
  out.phase_end_times.clear();

  int contact_prev[4], contact_cur[4];
  float phase_start_time = 0.0f, phase_end_time = 0.0f;
  int n_phases = 0;
  float t = 0.0f;
  quad_ref_ptr->get_contact_at_t(contact_prev, t);
  while (approx_leq_scalar(t, plan_duration)) {
    quad_ref_ptr->get_contact_at_t(contact_cur, t);
    bool change = false;
    for (int l = 0; l < 4; ++l) change = change || (contact_cur[l] != contact_prev[l]);
    if (change || approx_geq_scalar(t, plan_duration)) {
      phase_end_time = t;
      if (n_phases >= CAFE_MAX_PHASES) throw std::runtime_error("too many phases");
      CafePhase& ph = deck.phase[n_phases];
      ph.model = CAFE_MODEL_HKD;
      ph.horizon = (int)std::round((phase_end_time - phase_start_time) / dt_sim);
      for (int l = 0; l < 4; ++l) ph.contact[l] = contact_prev[l];
      out.phase_start_times.push_back(phase_start_time);
      out.phase_end_times.push_back(phase_end_time);
      n_phases++;
      for (int l = 0; l < 4; ++l) contact_prev[l] = contact_cur[l];
      phase_start_time = phase_end_time;
    }
    t += dt_sim;
  }
  deck.n_phases = n_phases;

  int rec = 0;
  for (int i = 0; i < n_phases; ++i) { deck.phase[i].knot_offset = rec; rec += deck.phase[i].horizon + 1; }
  deck.n_records = rec;
  out.ref.assign((size_t)rec * CAFE_REF_W, 0.0);

  for (int i = 0; i < n_phases; ++i) {
    CafePhase& ph = deck.phase[i];
    ph.dt = (double)dt_sim;  // Trajectory(dt_sim, h) and bind(..., (T)dt_sim): a float widened to double
    ph.t_offset = out.phase_start_times[i] - out.phase_start_times[0];
    ph.next_model = (i < n_phases - 1) ? CAFE_MODEL_HKD : -1;
    ph.has_reset = 1;
    if (i < n_phases - 1) for (int l = 0; l < 4; ++l) ph.next_contact[l] = deck.phase[i + 1].contact[l];
    else quad_ref_ptr->get_contact_at_t(ph.next_contact, plan_duration + dt_mpc);
    ph.n_td = 0;
    for (int l = 0; l < 4; ++l) if (ph.contact[l] == 0 && ph.next_contact[l] == 1) ph.td_foot[ph.n_td++] = l;

    /* HKDTrackingCost weights (HKDCost.h:11-36) */
    const double q_eul[3] = {1, 4, 4}, q_pos[3] = {1, 1, 30}, q_omega[3] = {1.0, 0.5, 0.2}, q_v[3] = {1, 1, 1};
    for (int a = 0; a < 3; ++a) { ph.q[a] = q_eul[a]; ph.q[3 + a] = q_pos[a]; ph.q[6 + a] = q_omega[a]; ph.q[9 + a] = q_v[a]; }
    for (int l = 0; l < 4; ++l) for (int a = 0; a < 3; ++a) ph.q[12 + 3 * l + a] = .1 * (1 - ph.contact[l]);
    const double scale[12] = {1, 1, 2, 1, 1, 20, 1.0, 0.2, 0.1, 1, 1, 1};
    for (int j = 0; j < 24; ++j) ph.qf[j] = (20 * (j < 12 ? scale[j] : .01 * 1.0)) * ph.q[j];
    for (int j = 0; j < 24; ++j) ph.r[j] = .1 * 1.0;
    /* HKDFootPlaceReg: Qfoot = diag(c,c,0) * 5 * 20 (HKDCost.h:56-71) */
    ph.w_footreg[0] = 1.0 * 5.0 * 20; ph.w_footreg[1] = 1.0 * 5.0 * 20; ph.w_footreg[2] = 0.0;
    ph.reb_grf = grf_reb_param;
    ph.al_td = td_al_param;
    ph.mu = 0.7;  // HKDConstraints.h:17
    ph.ground_height = 0;

    for (int k = 0; k <= ph.horizon; ++k) {
      /* time of the cost look-ups: float t = t_offset + k*dt with dt double (SinglePhase.cpp:243,255) */
      float t_cost = (float)((double)ph.t_offset + (double)k * ph.dt);
      /* time of the initial-guess look-up: float arithmetic (HKDProblem.cpp:88) */
      float t_init = out.phase_start_times[i] + k * dt_sim;
      int idx = quad_ref_ptr->index_at_t(t_cost);
      if (idx != quad_ref_ptr->index_at_t(t_init)) throw std::runtime_error("reference index mismatch between cost and initial-guess look-ups");
      fill_hkd_record(&out.ref[(size_t)(ph.knot_offset + k) * CAFE_REF_W], *quad_ref_ptr->get_a_reference_ptr_at_t(t_cost));
    }
  }
  deck.ref = out.ref.data();
  deck.BG_alpha = 0;
}

}  // namespace cafe
