// abi_host.cpp — host-only part of the C ABI (include/cafe_gpu.h): settings, reference ingestion,
// phase-deck building, result-layout sizes. No CUDA here.
#include <cstring>
#include <exception>
#include <string>
#include "../../../include/cafe_gpu.h"
#include "info_reader.h"
#include "problem_builders.h"

namespace cafe {
static thread_local std::string g_last_error;
void set_last_error(const std::string& s) { g_last_error = s; }
}  // namespace cafe

struct CafeDeckHandle {
  cafe::DeckStorage st;
  cafe::QuadReference ref;
};

extern "C" const char* cafe_last_error(void) { return cafe::g_last_error.c_str(); }

#define CAFE_TRY try {
#define CAFE_CATCH(code)                 \
  }                                      \
  catch (const std::exception& e) {      \
    cafe::set_last_error(e.what());      \
    return code;                         \
  }                                      \
  catch (...) {                          \
    cafe::set_last_error("unknown error"); \
    return code;                         \
  }

extern "C" int cafe_options_load(const char* fname, CafeOptions* out) {
  if (!fname || !out) { cafe::set_last_error("null argument"); return CAFE_ERR_ARG; }
  CAFE_TRY
  cafe::load_hsddp_setting(fname, *out);
  return 0;
  CAFE_CATCH(CAFE_ERR_IO)
}

// one key of a Boost-INFO settings file ("section.key"), for host code that mirrors the reference's load* helpers without Boost
extern "C" int cafe_info_get_number(const char* fname, const char* key, double* out) {
  if (!fname || !key || !out) { cafe::set_last_error("null argument"); return CAFE_ERR_ARG; }
  CAFE_TRY
  cafe::InfoFile pt(fname);
  *out = pt.num(key);
  return 0;
  CAFE_CATCH(CAFE_ERR_IO)
}
extern "C" int cafe_info_get_string(const char* fname, const char* key, char* out, int cap) {
  if (!fname || !key || !out || cap <= 0) { cafe::set_last_error("null argument"); return CAFE_ERR_ARG; }
  CAFE_TRY
  cafe::InfoFile pt(fname);
  const std::string v = pt.str(key);
  if ((int)v.size() + 1 > cap) { cafe::set_last_error("value longer than the buffer"); return CAFE_ERR_ARG; }
  std::memcpy(out, v.c_str(), v.size() + 1);
  return 0;
  CAFE_CATCH(CAFE_ERR_IO)
}

extern "C" int cafe_deck_build_hkd(const char* reference_csv, const char* constraint_params_info, float plan_duration,
                                   float time_step, int nsteps_between_mpc, int k0, CafeDeckHandle** out) {
  if (!reference_csv || !constraint_params_info || !out) { cafe::set_last_error("null argument"); return CAFE_ERR_ARG; }
  CAFE_TRY
  CafeDeckHandle* h = new CafeDeckHandle();
  try {
    h->ref.load_top_level_data(reference_csv, true, k0);
    cafe::HKDProblem prob;
    cafe::HKDPlanConfig cfg{plan_duration, time_step, nsteps_between_mpc};
    prob.set_problem_data(&h->ref, cfg, constraint_params_info);
    prob.initialization(h->st);
  } catch (...) { delete h; throw; }
  *out = h;
  return 0;
  CAFE_CATCH(CAFE_ERR_IO)
}

extern "C" int cafe_deck_build_mhpc(const char* reference_csv, const char* mhpc_config_info, const char* settings_root,
                                    int k0, CafeDeckHandle** out) {
  if (!reference_csv || !mhpc_config_info || !settings_root || !out) { cafe::set_last_error("null argument"); return CAFE_ERR_ARG; }
  CAFE_TRY
  CafeDeckHandle* h = new CafeDeckHandle();
  try {
    cafe::MHPCConfig cfg;
    cafe::loadMHPCConfig(mhpc_config_info, cfg);
    h->ref.load_top_level_data(reference_csv, false, k0);
    cafe::MHPCProblem prob;
    prob.set_problem_data(&h->ref, cfg, settings_root);
    prob.initialization(h->st);
  } catch (...) { delete h; throw; }
  *out = h;
  return 0;
  CAFE_CATCH(CAFE_ERR_IO)
}

extern "C" int cafe_deck_build_loco(const char* reference_csv, const char* loco_config_info, const char* settings_root,
                                    int k0, CafeDeckHandle** out) {
  if (!reference_csv || !loco_config_info || !settings_root || !out) { cafe::set_last_error("null argument"); return CAFE_ERR_ARG; }
  CAFE_TRY
  CafeDeckHandle* h = new CafeDeckHandle();
  try {
    cafe::MHPCConfig cfg;
    cafe::loadMHPCConfig(loco_config_info, cfg);
    h->ref.load_top_level_data(reference_csv, false, k0);
    cafe::MHPCProblem prob;
    prob.loco = true;
    prob.set_problem_data(&h->ref, cfg, settings_root);
    prob.initialization(h->st);
  } catch (...) { delete h; throw; }
  *out = h;
  return 0;
  CAFE_CATCH(CAFE_ERR_IO)
}

// MHPCProblem<T>::set_problem_data(pdata, pconfig) + initialization with an MHPCConfig the caller holds in memory (loaded by
// loadMHPCConfig and possibly edited, MHPCProblem.h:43-83, :198-212) instead of the name of its file
extern "C" int cafe_deck_build_mhpc_config(const char* reference_csv, const CafeMHPCConfig* c, const char* settings_root, int k0, int loco,
                                           CafeDeckHandle** out) {
  if (!reference_csv || !c || !c->costFileName || !c->constraintParamFileName || !settings_root || !out) { cafe::set_last_error("null argument"); return CAFE_ERR_ARG; }
  CAFE_TRY
  CafeDeckHandle* h = new CafeDeckHandle();
  try {
    cafe::MHPCConfig cfg;
    cfg.plan_dur_wb = c->plan_dur_wb; cfg.plan_dur_srb = c->plan_dur_srb; cfg.dt_mpc = c->dt_mpc;
    cfg.dt_wb = c->dt_wb; cfg.dt_srb = c->dt_srb; cfg.BG_alpha = (double)c->BG_alpha; cfg.num_threads = 1;
    cfg.costFileName = c->costFileName; cfg.constraintParamFileName = c->constraintParamFileName;
    h->ref.load_top_level_data(reference_csv, false, k0);
    cafe::MHPCProblem prob;
    prob.loco = loco != 0;
    prob.set_problem_data(&h->ref, cfg, settings_root);
    prob.initialization(h->st);
  } catch (...) { delete h; throw; }
  *out = h;
  return 0;
  CAFE_CATCH(CAFE_ERR_IO)
}

// phase_start_times / phase_end_times of the problem data (MHPCProblemData::wb_phase_start_times, HKDProblemData::phase_start_times),
// one pair per phase of the deck, seconds from the start of the plan
extern "C" int cafe_deck_phase_times(const CafeDeckHandle* h, float* start_times, float* end_times) {
  if (!h || !start_times || !end_times) { cafe::set_last_error("null argument"); return CAFE_ERR_ARG; }
  const CafeDeck& d = h->st.deck;
  float t = 0;
  for (int i = 0; i < d.n_phases; ++i) {
    const bool have = i < (int)h->st.phase_start_times.size() && i < (int)h->st.phase_end_times.size();
    start_times[i] = have ? h->st.phase_start_times[i] : t;
    end_times[i] = have ? h->st.phase_end_times[i] : t + (float)(d.phase[i].horizon * d.phase[i].dt);
    t = end_times[i];
  }
  return 0;
}

extern "C" int cafe_deck_build_barrel_to(const char* cost_weights_json, const char* constraint_params_info, CafeDeckHandle** out) {
  if (!cost_weights_json || !constraint_params_info || !out) { cafe::set_last_error("null argument"); return CAFE_ERR_ARG; }
  CAFE_TRY
  CafeDeckHandle* h = new CafeDeckHandle();
  try { cafe::build_barrel_to_deck(cost_weights_json, constraint_params_info, h->st); } catch (...) { delete h; throw; }
  *out = h;
  return 0;
  CAFE_CATCH(CAFE_ERR_IO)
}

extern "C" int cafe_barrel_to_initial_guess(const CafeDeck* deck, const double* x0, int B, double* guess) {
  if (!deck || !x0 || !guess || B <= 0 || deck->n_phases != 6) { cafe::set_last_error("bad argument"); return CAFE_ERR_ARG; }
  for (int i = 0; i < 6; ++i) if (deck->phase[i].model != CAFE_MODEL_WB) { cafe::set_last_error("not a barrel-roll deck"); return CAFE_ERR_ARG; }
  const long sz = cafe_solution_size(deck);
  std::memset(guess, 0, (size_t)B * sz * sizeof(double));
  for (int b = 0; b < B; ++b) cafe::barrel_to_guess(*deck, x0 + (size_t)b * 36, guess + (size_t)b * sz);
  return 0;
}

extern "C" int cafe_deck_mark_mpc_update(CafeDeckHandle* h, int nsteps, int* marked_phase) {
  if (!h || nsteps < 1 || !marked_phase) { cafe::set_last_error("bad argument"); return CAFE_ERR_ARG; }
  *marked_phase = -1;
  CafeDeck& d = h->st.deck;
  int last = -1;
  for (int i = 0; i < d.n_phases; ++i) if (d.phase[i].model != CAFE_MODEL_SRB) last = i;
  if (last < 0) return 0;
  // MHPCProblem.cpp:366-369: every phase but a tail phase not longer than the shift gets update_SS_config(h + 1); a tail phase that
  // short was opened by this very update (an older one has grown past nsteps), its SS_set is still empty (SinglePhase.cpp:34)
  if (d.phase[last].horizon <= nsteps) { d.phase[last].single_shooting = 1; *marked_phase = last; }
  // In the update path a phase receives its touchdown constraint (and, whole-body, the touchdown-velocity penalty that comes with it) only once the
  // contact change has reached the tail of the plan: add_tconstr_one_phase is called when the knot appended to the last phase carries a new contact
  // (MHPCProblem.cpp:346-350, HKDProblem.cpp:186-196), whereas initialization() attaches it from the look-ahead contact at plan end + dt_mpc
  // (:533-537). A tail phase whose terminal record still shows its own contact has therefore none yet; its reset map does use the look-ahead
  // contact (update_resetmap runs for every phase on every update, :357).
  {
    CafePhase& p = d.phase[last];
    const double* r = d.ref + ((size_t)p.knot_offset + p.horizon) * CAFE_REF_W + CAFE_REF_CONTACT;
    bool change = false;
    for (int l = 0; l < 4; ++l) change |= ((r[l] > 0) != (p.contact[l] > 0));
    if (!change) p.n_td = 0;
  }
  return 0;
}

extern "C" const CafeDeck* cafe_deck_get(const CafeDeckHandle* h) { return h ? &h->st.deck : nullptr; }
extern "C" void cafe_deck_free(CafeDeckHandle* h) { delete h; }

extern "C" int cafe_hkd_state(const double body[12], const double qJ[12], const int contact[4], double x0[24]) {
  if (!body || !qJ || !contact || !x0) { cafe::set_last_error("null argument"); return CAFE_ERR_ARG; }
  double qdummy[12];
  cafe::compute_hkd_state(body, body + 3, qJ, qdummy, contact);
  for (int i = 0; i < 12; ++i) { x0[i] = body[i]; x0[12 + i] = qdummy[i]; }
  return 0;
}

extern "C" int cafe_deck_lq_pattern(const CafeDeck* deck, int phase, int knot, int which, unsigned long long* out) {
  if (!deck || !out || phase < 0 || phase >= deck->n_phases || knot < 0 || knot >= deck->phase[phase].horizon || which < 0 || which > 3) {
    cafe::set_last_error("bad argument"); return CAFE_ERR_ARG;
  }
  const int model = deck->phase[phase].model;
  if (model == CAFE_MODEL_HKD) {
    unsigned long long hm[36];
    cafe::hkd_lq_patterns(hm);
    for (int i = 0; i < 9; ++i) out[i] = hm[9 * which + i];
    return 9;
  }
  if (model == CAFE_MODEL_WB && which == 2) {
    cafe::wb_lxx_pattern(deck->ref + ((size_t)deck->phase[phase].knot_offset + knot) * CAFE_REF_W, out);
    return 21;
  }
  cafe::set_last_error("no structural pattern is kept for this array of this model");
  return CAFE_ERR_UNSUPPORTED;
}

extern "C" long cafe_solution_size(const CafeDeck* deck) {
  long s = 0;
  for (int i = 0; i < deck->n_phases; ++i) {
    const CafePhase& ph = deck->phase[i];
    long n = cafe_model_n(ph.model), m = cafe_model_m(ph.model), p = cafe_model_p(ph.model), h = ph.horizon;
    s += (h + 1) * n + h * m + h * p + h * m + h * m * n + h * m + h * m * m + h * m * n + (h + 1) * n;
  }
  return s;
}

extern "C" long cafe_command_size(const CafeDeck* deck, int n_gain_knots) {
  long s = 0;
  int left = n_gain_knots;
  for (int i = 0; i < deck->n_phases; ++i) {
    const CafePhase& ph = deck->phase[i];
    long n = cafe_model_n(ph.model), m = cafe_model_m(ph.model), p = cafe_model_p(ph.model), h = ph.horizon;
    s += (h + 1) * n + h * m + h * p;
    long g = left < h ? left : h;
    s += g * (m * n + m + m * m + m * n);
    left -= (int)g;
  }
  return s;
}
