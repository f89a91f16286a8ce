// HKDProblem.h — HKDPlanConfig / HKDProblemData<T> / HKDProblem<T> with the members and call sequence of the reference
// (HKDMPC/HKD-TrajOpt/HKDProblem.h:20-168, HKDProblem.cpp:15-222), as used by HKDMPCSolver (HKDMPC/HKDMPC.cpp:20-140):
//     pdata.quad_ref_ptr = &quad_ref; problem.set_problem_data(&pdata, config); problem.initialization();
//     solver.set_multiPhaseProblem(phases); solver.solve(ddp_setting);  problem.update();
// The phases are bound to a plain-data phase deck built by cafe_deck_build_hkd (C ABI); the constraint parameters come from
// "../HKDMPC/settings/constraint_params.info" like in the reference (HKDProblem.cpp:71) unless cafe_constraint_params says otherwise.
#pragma once
#include <cstdio>
#include <deque>
#include <memory>
#include <string>
#include <vector>
#include "ConstraintsBase.h"
#include "QuadReference.h"
#include "SinglePhase.h"
#include "TrajectoryManagement.h"

using std::make_shared;

struct HKDPlanConfig {
  float plan_duration;
  float timeStep;
  int nsteps_between_mpc;
};

template <typename T>
struct HKDProblemData {
  QuadReference* quad_ref_ptr = nullptr;
  std::deque<shared_ptr<Trajectory<T, 24, 24, 0>>> trajectory_ptrs;
  std::deque<shared_ptr<SinglePhase<T, 24, 24, 0>>> phase_ptrs;
  std::deque<int> phase_horizons;
  std::deque<bool> is_phase_reach_end;
  std::deque<float> phase_start_times;
  std::deque<float> phase_end_times;
  std::deque<VecM<double, 4>> contact_durations;
  std::deque<VecM<int, 4>> phase_contacts;
  int n_phases = 0;
  void clear() {
    trajectory_ptrs.clear(); phase_ptrs.clear(); phase_horizons.clear(); is_phase_reach_end.clear(); phase_start_times.clear();
    phase_end_times.clear(); phase_contacts.clear(); contact_durations.clear(); n_phases = 0; quad_ref_ptr = nullptr;
  }
};

template <typename T>
class HKDProblem {
 public:
  const static size_t xs = 24;
  const static size_t us = 24;
  const static size_t ys = 0;
  typedef Trajectory<T, 24, 24, 0> Traj_T;
  typedef SinglePhase<T, 24, 24, 0> Phase_T;

  HKDProblem() { plan_duration = 0; dt_sim = 0; nsteps_between_mpc = 0; dt_mpc = 0; }

  void set_problem_data(HKDProblemData<T>* pdata_in, const HKDPlanConfig& config) {
    plan_duration = config.plan_duration;
    dt_sim = config.timeStep;
    nsteps_between_mpc = config.nsteps_between_mpc;
    dt_mpc = dt_sim * nsteps_between_mpc;
    pdata = pdata_in;
    quad_ref_ptr = pdata_in->quad_ref_ptr;
  }

  void initialization() {   // HKDProblem.cpp:15-111
    printf("Initializing HKDProblem ... \n\n");
    build_deck(quad_ref_ptr->cafe_k0(), false);
    fill();
    printf("Finished initializing HKDProbelm! \n\n");
  }

  // receding horizon (HKDProblem.cpp:117-222): the same front / back bookkeeping as MHPCProblem::update on the 24-state phases, plus
  // the quirk Ubar[0] = 0 of the front trajectory (:220); the tail phase opened by this update (horizon <= nsteps) has no shooting states
  void update() {
    if (!deck_) throw std::logic_error("HKDProblem::update before initialization");
    quad_ref_ptr->step(dt_mpc);
    struct Old { int s, e; int contact[4]; shared_ptr<Traj_T> traj; shared_ptr<Phase_T> phase; int n_td, td_foot[4]; };
    std::vector<Old> old;
    {
      const CafeDeck* d = deck_->deck();
      int s = deck_->k0;
      for (int i = 0; i < d->n_phases; ++i) {
        Old o{s, s + d->phase[i].horizon, {0, 0, 0, 0}, pdata->trajectory_ptrs[i], pdata->phase_ptrs[i], d->phase[i].n_td, {0, 0, 0, 0}};
        for (int f = 0; f < 4; ++f) { o.contact[f] = d->phase[i].contact[f]; o.td_foot[f] = d->phase[i].td_foot[f]; }
        old.push_back(o); s += d->phase[i].horizon;
      }
    }
    build_deck(quad_ref_ptr->cafe_k0(), true);
    fill();
    const CafeDeck* d = deck_->deck();
    int s = deck_->k0;
    const int old_end = old.back().e;
    for (int i = 0; i < d->n_phases; ++i) {
      const CafePhase& p = d->phase[i];
      Traj_T& nt = *pdata->trajectory_ptrs[i];
      const int e = s + p.horizon;
      const Old* src = nullptr;
      for (const Old& o : old) {
        bool same = true;
        for (int f = 0; f < 4; ++f) same = same && o.contact[f] == p.contact[f];
        if (same && o.s <= e && o.e >= s) { src = &o; break; }
      }
      const bool continues_last = src && src == &old.back();
      if (src && src->phase->cafe_al_set && src->n_td > 0 && src->n_td == p.n_td) {
        // the phase keeps its TouchDownConstraint object, hence sigma / lambda (reset_params is empty, ConstraintsBase.h:367-374)
        bool feet = true;
        for (int f = 0; f < p.n_td; ++f) feet = feet && src->td_foot[f] == p.td_foot[f];
        if (feet) { for (int q = 0; q < 8; ++q) pdata->phase_ptrs[i]->cafe_al[q] = src->phase->cafe_al[q]; pdata->phase_ptrs[i]->cafe_al_set = true; }
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
    if (!pdata->trajectory_ptrs.empty() && pdata->trajectory_ptrs.front()->horizon > 0) pdata->trajectory_ptrs.front()->Ubar[0].setZero();   // :220
  }

  void clear_problem_data() { if (pdata != nullptr) pdata->clear(); deck_.reset(); }
  void pretty_print() {
    for (int i = 0; i < pdata->n_phases; i++)
      printf("phase %d: contact [%d %d %d %d], horizon %d, start %.3f, end %.3f\n", i, pdata->phase_contacts[i][0], pdata->phase_contacts[i][1],
             pdata->phase_contacts[i][2], pdata->phase_contacts[i][3], pdata->phase_horizons[i], pdata->phase_start_times[i], pdata->phase_end_times[i]);
  }
  void lcm_publish() {}      // the problem-data LCM message is telemetry outside the solve path
  void reset_lcm_data() {}

  // ---- binding
  std::shared_ptr<cafe_facade::DeckOwner> cafe_deck() const { return deck_; }
  std::string cafe_constraint_params = "../HKDMPC/settings/constraint_params.info";

 private:
  void build_deck(int k0, bool mpc_update) {
    if (!pdata || !quad_ref_ptr) throw std::logic_error("HKDProblem: set_problem_data first");
    auto owner = std::make_shared<cafe_facade::DeckOwner>();
    cafe_facade::check(cafe_deck_build_hkd(quad_ref_ptr->cafe_file().c_str(), cafe_constraint_params.c_str(), plan_duration, dt_sim, nsteps_between_mpc, k0, &owner->h));
    owner->k0 = k0;
    owner->lineage = (mpc_update && deck_) ? deck_->lineage : cafe_facade::DeckOwner::next_lineage();
    if (mpc_update) { int which = -1; cafe_facade::check(cafe_deck_mark_mpc_update(owner->h, nsteps_between_mpc, &which)); }
    deck_ = owner;
  }
  void fill() {
    const CafeDeck* d = deck_->deck();
    std::vector<float> ts(d->n_phases), te(d->n_phases);
    cafe_facade::check(cafe_deck_phase_times(deck_->h, ts.data(), te.data()));
    QuadReference* q = pdata->quad_ref_ptr;
    pdata->clear();
    pdata->quad_ref_ptr = q;
    for (int i = 0; i < d->n_phases; ++i) {
      const CafePhase& p = d->phase[i];
      pdata->phase_horizons.push_back(p.horizon);
      pdata->phase_start_times.push_back(ts[i]); pdata->phase_end_times.push_back(te[i]);
      VecM<int, 4> c; for (int f = 0; f < 4; ++f) c[f] = p.contact[f];
      pdata->phase_contacts.push_back(c);
      pdata->contact_durations.push_back(VecM<double, 4>());
      pdata->is_phase_reach_end.push_back(false);   // HKDProblem.cpp:60: contact_prev != contact_prev, always false
      auto traj = make_shared<Traj_T>((T)dt_sim, p.horizon);
      for (int k = 0; k <= p.horizon; ++k) {
        const double* r = d->ref + ((size_t)p.knot_offset + k) * CAFE_REF_W + CAFE_REF_XR;
        for (size_t cc = 0; cc < xs; ++cc) { traj->X[k][cc] = (T)r[cc]; traj->Xbar[k][cc] = (T)r[cc]; }
      }
      auto phase = make_shared<Phase_T>();
      phase->set_trajectory(traj);
      phase->cafe_deck = deck_; phase->cafe_phase_index = i;
      phase->set_time_offset(p.t_offset);
      phase->initialization();
      phase->update_SS_config(p.single_shooting ? 0 : p.horizon + 1);
      pdata->trajectory_ptrs.push_back(traj); pdata->phase_ptrs.push_back(phase);
    }
    pdata->n_phases = d->n_phases;
  }

  HKDProblemData<T>* pdata = nullptr;
  QuadReference* quad_ref_ptr = nullptr;
  std::shared_ptr<cafe_facade::DeckOwner> deck_;
  float plan_duration, dt_sim, dt_mpc;
  int nsteps_between_mpc;
};
