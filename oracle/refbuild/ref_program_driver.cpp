// ref_program_driver.cpp — TEST INFRASTRUCTURE ONLY (oracle/). Runs one of the reference's stand-alone trajectory-optimisation PROGRAMS unchanged
// and writes down what its solver did:
//   MHPC/MHPC-Trajopt/Locomotion/Loco_TO.cpp        (LocoProblem: whole-body-only locomotion TO, 1 s flypace plan)
//   MHPC/MHPC-Trajopt/BarrelRoll/BarrelRollTO.cpp   (in-place barrel roll: six hand-scheduled phases built inside its main())
// Their whole problem set-up lives in main(), so the source file is compiled and linked as it is - its main() is the program's main().
// This file adds no entry point, only two hooks that see inside without touching the source (set up by a static initialiser):
//   * the linker wraps MultiPhaseDDP<double>::solve (-Wl,--wrap=<mangled name>, oracle/refbuild/Makefile): __wrap_... below puts the
//     recording decorator of ref_spy.hpp around every phase the program handed to its solver, calls the real solve, and dumps counters,
//     events and the phases' trajectories (this translation unit reads private members: "#define private public" around the reference's
//     headers - same layout, the reference's own objects are compiled without it);
//   * the LCM stand-in hands the messages the program publishes afterwards to the listener below (not needed for the record, kept as a
//     cross-check: the published wbTraj_lcmt carries the final trajectory in double precision).
// usage (cwd = a directory D with D/../MHPC/..., e.g. data/_run):  REF_OUT=<out.txt> REF_HIP_YAW=<yaw> ref_loco|ref_barrel_to
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <typeinfo>
#include <vector>
#include <functional>
#include <algorithm>
#include <numeric>
#include <chrono>
#include <random>
#include <unordered_map>
#include <tuple>
#include <utility>
#include <iomanip>
#include <cmath>
#include <cassert>
#include <cstring>
#include <eigen3/Eigen/Dense>
#include <boost/property_tree/ptree.hpp>
#include <boost/property_tree/info_parser.hpp>
#include <boost/property_tree/json_parser.hpp>
#include <lcm/lcm-cpp.hpp>
#include "cafe_pinocchio_shim.hpp"

#define private public
#define protected public
#include "SinglePhase.h"
#include "MultiPhaseDDP.h"
#undef private
#undef protected
#include "PinocchioInteface.h"
#include "wbTraj_lcmt.hpp"

#include "ref_spy.hpp"

template <typename TT>
void buildPinModelFromURDF(const std::string&, pinocchio::ModelTpl<TT>& mc_model) {
  const char* y = getenv("REF_HIP_YAW");
  mc_model.set_hip_yaw(y ? atof(y) : 3.1415);
}
template void buildPinModelFromURDF<double>(const std::string&, pinocchio::ModelTpl<double>&);

static FILE* g_out = nullptr;

#define SOLVE_SYM _ZN13MultiPhaseDDPIdE5solveER12HSDDP_OPTIONRKf
#define CAT2(a, b) a##b
#define CAT(a, b) CAT2(a, b)
extern "C" void CAT(__real_, SOLVE_SYM)(MultiPhaseDDP<double>*, HSDDP_OPTION&, const float&);
extern "C" void CAT(__wrap_, SOLVE_SYM)(MultiPhaseDDP<double>* self, HSDDP_OPTION& opt, const float& max_cputime) {
  std::deque<std::shared_ptr<SinglePhaseBase<T>>> orig = self->phases, spied;
  int id = 0;
  for (auto& p : orig) spied.push_back(std::make_shared<SpyPhase>(p, id++));
  self->phases = spied;
  g_log.clear();
  CAT(__real_, SOLVE_SYM)(self, opt, max_cputime);
  self->phases = orig;
  FILE* f = g_out;
  fprintf(f, "n_problems 1 n_updates 0\nproblem 0\n");
  fprintf(f, "guess n_phases 0\n");
  fprintf(f, "x0"); put_vec(f, self->x0); fprintf(f, "\n");
  fprintf(f, "counters iter %d ls_iter_total %d reg_iter_total %d\n", self->iter_, self->ls_iter_total_, self->reg_iter_total_);
  fprintf(f, "final cost %.17g feas %.17g tconstr %.17g pconstr %.17g\n", (double)self->actual_cost, (double)self->feas, (double)self->max_tconstr, (double)self->max_pconstr);
  fprintf(f, "float_cost_buffer %d", (int)self->cost_buffer.size()); for (float v : self->cost_buffer) fprintf(f, " %.9g", (double)v); fprintf(f, "\n");
  fprintf(f, "events %d\n", (int)g_log.size());
  for (const Event& ev : g_log) fprintf(f, "%d %d %.17g %.17g\n", ev.type, ev.phase, ev.a, ev.b);
  fprintf(f, "solution n_phases %d\n", (int)orig.size());
  const int none[4] = {0, 0, 0, 0};
  for (int i = 0; i < (int)orig.size(); ++i) {
    if (auto* wb = dynamic_cast<SinglePhase<double, 36, 12, 12>*>(orig[i].get())) dump_traj(f, *wb->traj, i, none, (double)wb->t_offset, (double)wb->t_offset + wb->phase_horizon * wb->dt);
    else if (auto* srb = dynamic_cast<SinglePhase<double, 12, 12, 0>*>(orig[i].get())) dump_traj(f, *srb->traj, i, none, (double)srb->t_offset, (double)srb->t_offset + srb->phase_horizon * srb->dt);
    else { fprintf(stderr, "unknown phase type\n"); exit(3); }
  }
  fflush(f);
}

struct Capture : lcm::Listener {
  void on_publish(const std::string& channel, const void* msg, const std::type_info& type) override {
    if (type != typeid(wbTraj_lcmt)) return;
    const wbTraj_lcmt& m = *static_cast<const wbTraj_lcmt*>(msg);
    fprintf(g_out, "published %s sz %d\n", channel.c_str(), (int)m.sz);
    for (int k = 0; k < m.sz; ++k) {
      fprintf(g_out, "wbtraj");
      for (double v : m.pos[k]) fprintf(g_out, " %.17g", v);
      for (double v : m.eul[k]) fprintf(g_out, " %.17g", v);
      for (double v : m.qJ[k]) fprintf(g_out, " %.17g", v);
      for (double v : m.vWorld[k]) fprintf(g_out, " %.17g", v);
      for (double v : m.eulrate[k]) fprintf(g_out, " %.17g", v);
      for (double v : m.qJd[k]) fprintf(g_out, " %.17g", v);
      for (double v : m.torque[k]) fprintf(g_out, " %.17g", v);
      fprintf(g_out, "\n");
    }
  }
};

static void close_out() { if (g_out) fclose(g_out); g_out = nullptr; }
struct Install {
  Install() {
    const char* path = getenv("REF_OUT");
    g_out = fopen(path ? path : "ref_program_out.txt", "w");
    if (!g_out) { fprintf(stderr, "cannot open REF_OUT\n"); exit(2); }
    static Capture cap;
    lcm::listener() = &cap;
    atexit(close_out);
  }
};
static Install g_install;
