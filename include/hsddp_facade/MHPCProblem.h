// MHPCProblem.h — MHPCConfig / loadMHPCConfig / MHPCProblemData<T> / MHPCProblem<T> with the members and call sequence of the reference
// (MHPC/MHPC-Trajopt/MHPCProblem.h:21-289, MHPCProblem.cpp:13-397), so that testMHPCProblem.cpp:9-89 and MHPCLocomotion.cpp:20-150
// compile against this directory with only the include path changed:
//     problem.set_problem_data(&pdata, &config); problem.prepare_initialization(); problem.initialize_parameters();
//     problem.initialize_multiPhaseProblem();  ...  solver.set_multiPhaseProblem(phases); solver.solve(ddp_setting);  problem.update();
// What differs underneath: the phases are not wired with callbacks but bound to a plain-data phase deck built by
// cafe_deck_build_mhpc_config (C ABI), the numerics run in the CUDA solver. File names inside MHPCConfig are resolved like the
// reference does, relative to "../" of the working directory; set cafe_settings_root to run from somewhere else.
#pragma once
#include <cstdio>
#include <deque>
#include <iostream>
#include <memory>
#include <string>
#include <vector>
#include "ConstraintsBase.h"
#include "QuadReference.h"
#include "SRBM.h"
#include "SinglePhase.h"
#include "TrajectoryManagement.h"
#include "WBM.h"

using std::make_shared;
using std::vector;

struct MHPCConfig {
  double plan_dur_wb;
  double plan_dur_srb;
  double dt_wb;
  double dt_srb;
  float dt_mpc;
  float BG_alpha;
  int num_threads{1};
  std::string referenceFileName;
  std::string costFileName;
  std::string constraintParamFileName;
  void print() {
    std::cout << "===================== MHPC Config =================== \n";
    std::cout << "WB plan duration = \t" << plan_dur_wb << "\n";
    std::cout << "WB simulation timestep = \t" << dt_wb << "\n";
    std::cout << "SRB plan duration = \t" << plan_dur_srb << "\n";
    std::cout << "SRB simulation timestep = \t" << dt_srb << "\n";
    std::cout << "MPC updates every " << dt_mpc << " seconds" << "\n";
    std::cout << "Baumgart alpha (vel) " << BG_alpha << "\n";
    std::cout << "Reference " << referenceFileName << "\n";
    std::cout << "Cost file location " << costFileName << "\n";
    std::cout << "Constraint_param file location " << constraintParamFileName << "\n";
  }
};

// the keys of loadMHPCConfig (MHPCProblem.h:67-83), read by the library's INFO reader (cafe_info_get, no Boost)
inline void loadMHPCConfig(const std::string filename, MHPCConfig& config) {
  std::cout << "********* loading MHPC configuration from file *********\n" << filename << "\n\n";
  auto num = [&](const char* key) { double v = 0; cafe_facade::check(cafe_info_get_number(filename.c_str(), key, &v)); return v; };
  auto str = [&](const char* key) { char buf[512]; cafe_facade::check(cafe_info_get_string(filename.c_str(), key, buf, (int)sizeof buf)); return std::string(buf); };
  config.plan_dur_wb = num("config.plan_dur_wb");
  config.plan_dur_srb = num("config.plan_dur_srb");
  config.dt_mpc = (float)num("config.dt_mpc");
  config.dt_wb = num("config.dt_wb");
  config.dt_srb = num("config.dt_srb");
  config.BG_alpha = (float)num("config.BG_alpha");
  config.num_threads = (int)num("config.nthreads");
  config.referenceFileName = str("config.referenceFile");
  config.costFileName = str("config.costFile");
  config.constraintParamFileName = str("config.constraintParamFile");
}

template <typename T>
struct MHPCProblemData {
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
  std::deque<shared_ptr<Trajectory<T, WBM::xs, WBM::us, WBM::ys>>> wb_trajs;
  std::deque<shared_ptr<SinglePhase<T, WBM::xs, WBM::us, WBM::ys>>> wb_phases;
  shared_ptr<Trajectory<T, SRBM::xs, SRBM::us, SRBM::ys>> srb_traj = nullptr;
  shared_ptr<SinglePhase<T, SRBM::xs, SRBM::us, SRBM::ys>> srb_phase = nullptr;
  shared_ptr<QuadReference> quad_reference = nullptr;
  std::deque<int> wb_phase_horizons;
  std::deque<bool> wb_is_phase_reach_end;
  std::deque<float> wb_phase_start_times;
  std::deque<float> wb_phase_end_times;
  std::deque<VecM<double, 4>> wb_contact_durations;
  std::deque<VecM<int, 4>> wb_phase_contacts;
  int n_wb_phases = 0;
  float srb_start_time = 0;
  float srb_end_time = 0;
  int srb_phase_horizon = 0;
  int n_srb_phase = 0;

  void get_index(const int& k_cur, int& pidx, int& k_pidx) {   // MHPCProblem.h:110-126
    pidx = 0; k_pidx = 0;
    int s_i(0);
    for (int i(0); i < n_wb_phases; ++i) {
      int h = wb_phase_horizons[i];
      if (k_cur >= s_i && k_cur < s_i + h) { pidx = i; k_pidx = k_cur - s_i; break; }
      s_i += h;
    }
  }
  void clear() {
    wb_trajs.clear(); wb_phases.clear(); srb_traj = nullptr; srb_phase = nullptr; quad_reference = nullptr;
    wb_phase_horizons.clear(); wb_is_phase_reach_end.clear(); wb_phase_start_times.clear(); wb_phase_end_times.clear();
    wb_contact_durations.clear(); wb_phase_contacts.clear(); n_wb_phases = 0;
    srb_start_time = 0; srb_end_time = 0; srb_phase_horizon = 0; n_srb_phase = 0;
  }
};

template <typename T>
class MHPCProblem {
 public:
  typedef SinglePhase<T, WBM::xs, WBM::us, WBM::ys> WBPhase_T;
  typedef SinglePhase<T, SRBM::xs, SRBM::us, SRBM::ys> SRBPhase_T;
  typedef Trajectory<T, WBM::xs, WBM::us, WBM::ys> WBTraj_T;
  typedef Trajectory<T, SRBM::xs, SRBM::us, SRBM::ys> SRBTraj_T;
  typedef VecM<T, WBM::xs> WBState;
  typedef VecM<T, WBM::us> WBContrl;
  typedef VecM<T, WBM::ys> WBOutput;
  typedef VecM<T, SRBM::xs> SRBMState;
  typedef VecM<T, SRBM::us> SRBMContrl;

  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
  MHPCProblem() : pdata(nullptr), pconfig(nullptr), quad_reference(nullptr), plan_dur_all(0.0), wb_nsteps_between_mpc(0), srb_nsteps_between_mpc(0) {}
  virtual ~MHPCProblem() {}

  void set_problem_data(MHPCProblemData<T>* pdata_in, const MHPCConfig* pconfig_in) {
    pdata = pdata_in;
    pconfig = pconfig_in;
    quad_reference = pdata->quad_reference;
    plan_dur_all = pconfig->plan_dur_wb + pconfig->plan_dur_srb;
    wb_nsteps_between_mpc = (int)round(pconfig->dt_mpc / pconfig->dt_wb);
    srb_nsteps_between_mpc = (int)round(pconfig->dt_mpc / pconfig->dt_srb);
  }
  void clear_problem_data() { pdata->clear(); plan_dur_all = 0; wb_nsteps_between_mpc = 0; srb_nsteps_between_mpc = 0; deck_.reset(); }

  void initialization() { prepare_initialization(); initialize_parameters(); initialize_multiPhaseProblem(); }   // MHPCProblem.cpp:13-23

  // cuts the whole-body phases along the reference's contact schedule and sizes the SRB tail (MHPCProblem.cpp:82-146): here the deck
  // builder does it, the problem data are filled from the deck
  void prepare_initialization() {
    build_deck(quad_reference->cafe_k0(), 0);
    fill_problem_data();
  }
  // ReB / AL parameters from the constraint file (MHPCProblem.cpp:149-171): read by the deck builder; mirrored here from the deck
  virtual void initialize_parameters() {
    need_deck();
    const CafeDeck* d = deck_->deck();
    for (int i = 0; i < d->n_phases; ++i) if (d->phase[i].model == CAFE_MODEL_WB) {
      const CafePhase& p = d->phase[i];
      grf_reb_param = reb(p.reb_grf); torque_reb_param = reb(p.reb_torque); joint_reb_param = reb(p.reb_joint); minheight_reb_param = reb(p.reb_minheight);
      td_al_param.lambda = (T)p.al_td.lambda; td_al_param.sigma = (T)p.al_td.sigma; td_al_param.sigma_max = (T)p.al_td.sigma_max;
      break;
    }
  }
  // one Trajectory + SinglePhase per phase, the state trajectories initialised with the reference (MHPCProblem.cpp:173-250)
  void initialize_multiPhaseProblem() {
    need_deck();
    const CafeDeck* d = deck_->deck();
    pdata->wb_trajs.clear(); pdata->wb_phases.clear(); pdata->srb_traj = nullptr; pdata->srb_phase = nullptr;
    for (int i = 0; i < d->n_phases; ++i) {
      const CafePhase& p = d->phase[i];
      if (p.model == CAFE_MODEL_WB) {
        auto traj = make_shared<WBTraj_T>((T)pconfig->dt_wb, p.horizon);
        for (int k = 0; k <= p.horizon; ++k) {
          const double* r = d->ref + ((size_t)p.knot_offset + k) * CAFE_REF_W + CAFE_REF_XR;
          for (size_t c = 0; c < WBM::xs; ++c) { traj->X[k][c] = (T)r[c]; traj->Xbar[k][c] = (T)r[c]; }
        }
        auto phase = make_shared<WBPhase_T>(pconfig->num_threads);
        phase->set_trajectory(traj);
        bind(*phase, i);
        phase->set_time_offset(p.t_offset);
        phase->initialization();
        phase->update_SS_config(p.single_shooting ? 0 : p.horizon + 1);
        pdata->wb_trajs.push_back(traj); pdata->wb_phases.push_back(phase);
      } else {
        auto traj = make_shared<SRBTraj_T>((T)pconfig->dt_srb, p.horizon);
        for (int k = 0; k <= p.horizon; ++k) {
          const double* r = d->ref + ((size_t)p.knot_offset + k) * CAFE_REF_W + CAFE_REF_XR;
          for (size_t c = 0; c < SRBM::xs; ++c) traj->Xbar[k][c] = (T)r[c];
        }
        auto phase = make_shared<SRBPhase_T>(pconfig->num_threads);
        phase->set_trajectory(traj);
        bind(*phase, i);
        phase->set_time_offset(pdata->srb_start_time);
        phase->initialization();
        phase->update_SS_config(p.horizon + 1);
        pdata->srb_phase = phase; pdata->srb_traj = traj;
      }
    }
  }

  // receding horizon (MHPCProblem.cpp:252-397): the reference steps by dt_mpc, the whole-body plan loses its first nsteps knots and
  // grows by nsteps at the tail (a contact change there opens a new phase), the SRB plan keeps its data while dt_mpc < dt_srb. The
  // deck of the shifted window is rebuilt (same builder, start offset advanced) and marked as the product of an update (a tail phase
  // not longer than the shift has no shooting states yet, :366-369); the trajectories are carried over knot by knot:
  //   same stance + overlapping absolute knots  -> Xbar, Ubar, K of the old plan;   knots past the old end of the last phase -> its
  //   last state, zero control and gain (Trajectory::push_back_state, TrajectoryManagement.cpp:206-228);   a phase the old plan did
  //   not have -> a fresh zero trajectory (:324-326).
  void update() {
    need_deck();
    quad_reference->step(pconfig->dt_mpc);
    const int nsteps = (int)round(pconfig->dt_mpc / pconfig->dt_wb);
    struct Old { int s, e; int contact[4]; shared_ptr<WBTraj_T> traj; shared_ptr<WBPhase_T> phase; int n_td, td_foot[4]; };
    std::vector<Old> old;
    {
      const CafeDeck* d = deck_->deck();
      int s = deck_->k0;
      for (int i = 0, w = 0; i < d->n_phases; ++i) if (d->phase[i].model == CAFE_MODEL_WB) {
        Old o{s, s + d->phase[i].horizon, {0, 0, 0, 0}, pdata->wb_trajs[w], pdata->wb_phases[w], d->phase[i].n_td, {0, 0, 0, 0}};
        ++w;
        for (int f = 0; f < 4; ++f) { o.contact[f] = d->phase[i].contact[f]; o.td_foot[f] = d->phase[i].td_foot[f]; }
        old.push_back(o); s += d->phase[i].horizon;
      }
    }
    shared_ptr<SRBTraj_T> old_srb = pdata->srb_traj;
    build_deck(quad_reference->cafe_k0(), nsteps);
    fill_problem_data();
    initialize_multiPhaseProblem();
    const CafeDeck* d = deck_->deck();
    int s = deck_->k0;
    const int old_end = old.empty() ? 0 : old.back().e;
    for (int i = 0, w = 0; i < d->n_phases; ++i) {
      const CafePhase& p = d->phase[i];
      if (p.model != CAFE_MODEL_WB) {
        if (old_srb && old_srb->horizon == p.horizon) { pdata->srb_traj->Xbar = old_srb->Xbar; pdata->srb_traj->Ubar = old_srb->Ubar; pdata->srb_traj->K = old_srb->K; }
        continue;
      }
      WBTraj_T& nt = *pdata->wb_trajs[w];
      WBPhase_T& nphase = *pdata->wb_phases[w];
      ++w;
      const int e = s + p.horizon;
      const Old* src = nullptr;
      for (const Old& o : old) {
        bool same = true;
        for (int f = 0; f < 4; ++f) same = same && o.contact[f] == p.contact[f];
        if (same && o.s <= e && o.e >= s) { src = &o; break; }
      }
      const bool continues_last = src && src == &old.back();
      if (src && src->phase->cafe_al_set && src->n_td > 0 && src->n_td == p.n_td) {
        // the phase keeps its touchdown-constraint object, hence sigma / lambda (reset_params is empty, ConstraintsBase.h:367-374)
        bool feet = true;
        for (int f = 0; f < p.n_td; ++f) feet = feet && src->td_foot[f] == p.td_foot[f];
        if (feet) { for (int q = 0; q < 8; ++q) nphase.cafe_al[q] = src->phase->cafe_al[q]; nphase.cafe_al_set = true; }
      }
      for (int k = 0; k <= p.horizon; ++k) {
        const int a = s + k;
        if (src && src->s <= a && a <= src->e) nt.Xbar[k] = src->traj->Xbar[a - src->s];
        else if (continues_last && a > old_end) nt.Xbar[k] = old.back().traj->Xbar.back();
        else if (!src) nt.Xbar[k].setZero();
        nt.X[k] = nt.Xbar[k];
        if (k < p.horizon) {
          if (src && src->s <= a && a < src->e) { nt.Ubar[k] = src->traj->Ubar[a - src->s]; nt.K[k] = src->traj->K[a - src->s]; }
          else { nt.Ubar[k].setZero(); nt.K[k].setZero(); }
          nt.U[k] = nt.Ubar[k];
        }
      }
      s = e;
    }
  }
  void update_WB_plan() {}    // both halves of the update run inside update() above
  void update_SRB_plan() {}
  int get_num_control_steps() { return (int)round(pconfig->dt_mpc / pconfig->dt_wb); }

  void pretty_print() {   // MHPCProblem.cpp:604-640
    printf("************Whole-Body Plan*************\n");
    for (int i = 0; i < pdata->n_wb_phases; i++) {
      printf("phase %d: contact [%d %d %d %d], horizon %d, start %.3f, end %.3f\n", i, pdata->wb_phase_contacts[i][0], pdata->wb_phase_contacts[i][1],
             pdata->wb_phase_contacts[i][2], pdata->wb_phase_contacts[i][3], pdata->wb_phase_horizons[i], pdata->wb_phase_start_times[i], pdata->wb_phase_end_times[i]);
    }
    printf("************SRB Plan*************\n");
    printf("horizon %d, start %.3f, end %.3f\n", pdata->srb_phase_horizon, pdata->srb_start_time, pdata->srb_end_time);
  }

  // ---- binding
  std::shared_ptr<cafe_facade::DeckOwner> cafe_deck() const { return deck_; }
  std::string cafe_settings_root = "..";   // the reference opens "../" + costFileName etc. (MHPCProblem.cpp:151, :421)

 public:
  MHPCProblemData<T>* pdata;
  const MHPCConfig* pconfig;
  shared_ptr<QuadReference> quad_reference;
  float plan_dur_all;
  int wb_nsteps_between_mpc;
  int srb_nsteps_between_mpc;
  REB_Param_Struct<T> grf_reb_param;
  REB_Param_Struct<T> torque_reb_param;
  REB_Param_Struct<T> jointspeed_reb_param;
  REB_Param_Struct<T> joint_reb_param;
  REB_Param_Struct<T> minheight_reb_param;
  AL_Param_Struct<T> td_al_param;

 protected:
  virtual int cafe_loco() const { return 0; }

 private:
  static REB_Param_Struct<T> reb(const CafeRebParam& p) { REB_Param_Struct<T> r; r.delta = (T)p.delta; r.delta_min = (T)p.delta_min; r.eps = (T)p.eps; return r; }
  void need_deck() { if (!deck_) prepare_initialization(); }
  template <class Phase> void bind(Phase& ph, int i) { ph.cafe_deck = deck_; ph.cafe_phase_index = i; }
  void build_deck(int k0, int mark_nsteps) {
    if (!pdata || !pconfig || !quad_reference) throw std::logic_error("MHPCProblem: set_problem_data first");
    CafeMHPCConfig c;
    c.plan_dur_wb = pconfig->plan_dur_wb; c.plan_dur_srb = pconfig->plan_dur_srb; c.dt_wb = pconfig->dt_wb; c.dt_srb = pconfig->dt_srb;
    c.dt_mpc = pconfig->dt_mpc; c.BG_alpha = pconfig->BG_alpha;
    c.costFileName = pconfig->costFileName.c_str(); c.constraintParamFileName = pconfig->constraintParamFileName.c_str();
    auto owner = std::make_shared<cafe_facade::DeckOwner>();
    cafe_facade::check(cafe_deck_build_mhpc_config(quad_reference->cafe_file().c_str(), &c, cafe_settings_root.c_str(), k0, cafe_loco(), &owner->h));
    owner->k0 = k0;
    owner->lineage = (mark_nsteps > 0 && deck_) ? deck_->lineage : cafe_facade::DeckOwner::next_lineage();
    if (mark_nsteps > 0) { int which = -1; cafe_facade::check(cafe_deck_mark_mpc_update(owner->h, mark_nsteps, &which)); }
    deck_ = owner;
  }
  void fill_problem_data() {
    const CafeDeck* d = deck_->deck();
    std::vector<float> ts(d->n_phases), te(d->n_phases);
    cafe_facade::check(cafe_deck_phase_times(deck_->h, ts.data(), te.data()));
    pdata->wb_phase_horizons.clear(); pdata->wb_is_phase_reach_end.clear(); pdata->wb_phase_start_times.clear(); pdata->wb_phase_end_times.clear();
    pdata->wb_contact_durations.clear(); pdata->wb_phase_contacts.clear(); pdata->n_wb_phases = 0; pdata->n_srb_phase = 0; pdata->srb_phase_horizon = 0;
    for (int i = 0; i < d->n_phases; ++i) {
      const CafePhase& p = d->phase[i];
      if (p.model == CAFE_MODEL_WB) {
        pdata->wb_phase_horizons.push_back(p.horizon);
        pdata->wb_phase_start_times.push_back(ts[i]); pdata->wb_phase_end_times.push_back(te[i]);
        VecM<int, 4> c; for (int f = 0; f < 4; ++f) c[f] = p.contact[f];
        pdata->wb_phase_contacts.push_back(c);
        pdata->wb_contact_durations.push_back(VecM<double, 4>());
        pdata->wb_is_phase_reach_end.push_back(p.n_td > 0 || i + 1 < d->n_phases);
        pdata->n_wb_phases++;
      } else {
        pdata->srb_phase_horizon = p.horizon; pdata->n_srb_phase = 1;
        pdata->srb_start_time = ts[i]; pdata->srb_end_time = te[i];
      }
    }
  }
  std::shared_ptr<cafe_facade::DeckOwner> deck_;
};
