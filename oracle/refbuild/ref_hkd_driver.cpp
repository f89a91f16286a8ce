// ref_hkd_driver.cpp — TEST INFRASTRUCTURE ONLY (oracle/). Drives the REFERENCE's own HS-DDP solver on the HKD trot problem.
//
// What is linked here is the reference's code, compiled unchanged from /root/reference by oracle/refbuild/Makefile:
//   HSDDPSolver/source/{MultiPhaseDDP,SinglePhase,SinglePhaseInterface,TrajectoryManagement}.cpp, HSDDPSolver/common/HSDDP_Utils.cpp,
//   HKDMPC/HKD-TrajOpt/{HKDProblem,HKDCost,HKDConstraints,HKDReference}.cpp, Reference/QuadReference.cpp, common/casadi_interface.cpp
//   and the CasADi-generated C of the HKD model (oracle/_ref/hkd_*.o) — against the Eigen / Boost.PropertyTree / LCM stand-ins in
//   oracle/refbuild/shim (none of the three libraries exists in this image; see shim/eigen3/Eigen/cafe_eigen_shim.hpp for what
//   that means for rounding).
// This file only does what HKDMPCSolver<T>::initialize / update do around the solver (HKDMPC/HKDMPC.cpp:19-86, :89-160) minus LCM:
//   load the reference, build the HKDProblem, set x0 = compute_hkd_state(body, qJ), MultiPhaseDDP::solve; then n MPC updates
//   (HKDProblem::update, caps 2 x 1, x0 = the plan's own prediction + a caller-given nudge) — and writes down everything the
//   solver decided. The per-iteration record is taken in full precision by a decorator around every phase (SpyPhase forwards the
//   whole SinglePhaseBase interface to the reference's SinglePhase): the solver's own history buffers are float.
//
// usage (cwd must be a directory D with D/../HKDMPC/settings/{ddp_setting,constraint_params}.info, e.g. data/_run):
//   ref_hkd <quad_reference.csv> <in.txt> <out.txt>
//   in.txt : n_problems n_updates, then per problem 24 numbers (body 12, qJ 12) and n_updates x 3 numbers (position nudges)
#include <cstdio>
#include <fstream>
#include <memory>
#include <vector>

#include "HKDProblem.h"

#include "ref_spy.hpp"

static void dump_problem(FILE* f, HKDProblemData<T>& pd, const char* tag) {
  fprintf(f, "%s n_phases %d\n", tag, pd.n_phases);
  for (int i = 0; i < pd.n_phases; ++i) dump_traj(f, *pd.trajectory_ptrs[i], i, pd.phase_contacts[i], (double)pd.phase_start_times[i], (double)pd.phase_end_times[i]);
}

static void run_solve(FILE* f, HKDProblemData<T>& pd, HSDDP_OPTION& opt, const VecM<T, 24>& xinit) {
  // what HKDMPCSolver::initialize / update do (HKDMPC.cpp:58-70, :136-150)
  std::deque<std::shared_ptr<SinglePhaseBase<T>>> phases;
  for (auto phase : pd.phase_ptrs) phases.push_back(phase);
  solve_and_record(f, phases, opt, xinit);
  dump_problem(f, pd, "solution");
}

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: ref_hkd <quad_reference.csv> <in.txt> <out.txt>\n"); return 2; }
  std::ifstream in(argv[2]);
  int n_prob = 0, n_upd = 0;
  in >> n_prob >> n_upd;
  FILE* f = fopen(argv[3], "w");
  if (!in.good() || !f) { fprintf(stderr, "cannot open input / output\n"); return 2; }
  fprintf(f, "n_problems %d n_updates %d\n", n_prob, n_upd);
  for (int b = 0; b < n_prob; ++b) {
    Vec12<T> body, qJ, qdummy;
    for (int i = 0; i < 12; ++i) in >> body[i];
    for (int i = 0; i < 12; ++i) in >> qJ[i];
    std::vector<Vec3<T>> nudge(n_upd);
    for (int u = 0; u < n_upd; ++u) for (int i = 0; i < 3; ++i) in >> nudge[u][i];

    // HKDMPCSolver<T>::HKDMPCSolver + initialize (HKDMPC.h:27-32, HKDMPC.cpp:19-56)
    QuadReference quad_reference;
    quad_reference.load_top_level_data(argv[1], true);
    HSDDP_OPTION ddp_options;
    loadHSDDPSetting("../HKDMPC/settings/ddp_setting.info", ddp_options);
    HKDPlanConfig mpc_config;
    mpc_config.plan_duration = .6;
    mpc_config.nsteps_between_mpc = 2;
    mpc_config.timeStep = 0.01;
    HKDProblem<T> opt_problem;
    HKDProblemData<T> opt_problem_data;
    opt_problem.clear_problem_data();
    opt_problem_data.quad_ref_ptr = &quad_reference;
    opt_problem.set_problem_data(&opt_problem_data, mpc_config);
    opt_problem.initialization();

    Vec3<T> pos, eul;
    pos = body.segment(3, 3);
    eul = body.head(3);
    const auto& initial_contact = opt_problem_data.phase_contacts.front();
    compute_hkd_state(eul, pos, qJ, qdummy, initial_contact);
    VecM<T, 24> xinit;
    xinit << body, qdummy;

    fprintf(f, "problem %d\n", b);
    dump_problem(f, opt_problem_data, "guess");
    run_solve(f, opt_problem_data, ddp_options, xinit);

    for (int u = 0; u < n_upd; ++u) {
      // HKDMPCSolver<T>::update (HKDMPC.cpp:89-150): fewer iterations when re-solving
      ddp_options.max_AL_iter = 2;
      ddp_options.max_DDP_iter = 1;
      opt_problem.update();
      // the "measured" state: the plan's own prediction for the new start (front of the shifted nominal trajectory), position nudged
      xinit = opt_problem_data.trajectory_ptrs.front()->Xbar.front();
      xinit.segment(3, 3) += nudge[u];
      fprintf(f, "update %d\n", u);
      dump_problem(f, opt_problem_data, "guess");
      run_solve(f, opt_problem_data, ddp_options, xinit);
    }
  }
  fclose(f);
  return 0;
}
