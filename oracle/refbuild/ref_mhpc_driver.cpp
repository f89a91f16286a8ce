// ref_mhpc_driver.cpp — TEST INFRASTRUCTURE ONLY (oracle/). Drives the REFERENCE's own MHPC problem (whole-body + single-rigid-body phases)
// through the reference's own HS-DDP solver.
//
// Linked here, compiled unchanged from /root/reference by oracle/refbuild/Makefile: the solver layer (HSDDPSolver/...), MHPC/MHPC-Trajopt/
// {MHPCProblem,WBM,MHPCCost,MHPCConstraint,MHPCReset,MHPCReference}.cpp (+ their headers: SRBM.h, MHPCFootStep.h, MHPCCostUtil.h ...),
// Reference/QuadReference.cpp, common/casadi_interface.cpp and the CasADi-generated C (SRB dynamics, whole-body kinematic partials) - against
// the Eigen / Boost.PropertyTree / LCM stand-ins and, for the rigid-body algorithms WBM.cpp calls, the Pinocchio stand-in
// shim/pinocchio/cafe_pinocchio_shim.hpp (the oracle's own recursions behind Pinocchio's names; read its header for what that pins and what not).
// MHPC-Trajopt/PinocchioInteface.cpp (URDF parsing through pinocchio::urdf / urdfdom) is the one file replaced: buildPinModelFromURDF below.
//
// This file does what MHPCLocomotion<T>::initialize / update do around the solver (MHPC/MHPCLocomotion.cpp:20-88, :91-150) minus LCM, with the
// set-up of MHPC-Trajopt/test/testMHPCProblem.cpp: load config / settings / reference, MHPCProblem::initialization, solve from x0, then n MPC
// updates (MHPCProblem::update, run-time caps, x0 = the plan's own prediction + a caller-given nudge). The wall-clock limit MHPCLocomotion passes
// to the re-solves (0.9 dt_mpc) is not applied: results would depend on the machine.
//
// usage (cwd must be a directory D with D/../MHPC/settings/..., e.g. data/_run):
//   ref_mhpc <quad_reference.csv> <hip_yaw> <in.txt> <out.txt> [mhpc_config.info]
//   in.txt : n_problems n_updates, then per problem 36 numbers (x0) and n_updates x 36 numbers (state nudges)
//   the optional last argument replaces "../MHPC/settings/mhpc_config.info" (MHPCLocomotion.cpp:24), e.g. the barrel-roll configuration
#include <cstdio>
#include <fstream>
#include <memory>
#include <vector>

#include "MHPCProblem.h"
#include "QuadReference.h"

#include "ref_spy.hpp"

static double g_hip_yaw = 3.1415;
// replaces MHPC-Trajopt/PinocchioInteface.cpp: the stand-in model is the fixed Mini-Cheetah tree of that file (PX PY PZ RZ RY RX + four legs)
// with the inertial data of urdf/mini_cheetah_simple_correctedInertia.urdf (oracle/wb_dynamics.hpp)
template <typename TT>
void buildPinModelFromURDF(const std::string& urdf_filename, pinocchio::ModelTpl<TT>& mc_model) {
  (void)urdf_filename;
  mc_model.set_hip_yaw(g_hip_yaw);
}
template void buildPinModelFromURDF<double>(const std::string&, pinocchio::ModelTpl<double>&);

static void dump_problem(FILE* f, MHPCProblemData<T>& pd, const char* tag) {
  const bool srb = pd.srb_phase.get() != nullptr;
  fprintf(f, "%s n_phases %d\n", tag, pd.n_wb_phases + (srb ? 1 : 0));
  for (int i = 0; i < pd.n_wb_phases; ++i) dump_traj(f, *pd.wb_trajs[i], i, pd.wb_phase_contacts[i], (double)pd.wb_phase_start_times[i], (double)pd.wb_phase_end_times[i]);
  const int none[4] = {0, 0, 0, 0};
  if (srb) dump_traj(f, *pd.srb_traj, pd.n_wb_phases, none, (double)pd.srb_start_time, (double)pd.srb_end_time);
}

static void run_solve(FILE* f, MHPCProblemData<T>& pd, HSDDP_OPTION& opt, const VecM<T, 36>& xinit) {
  std::deque<std::shared_ptr<SinglePhaseBase<T>>> phases;
  for (const auto& phase : pd.wb_phases) phases.push_back(phase);
  if (pd.srb_phase.get() != nullptr) phases.push_back(pd.srb_phase);
  solve_and_record(f, phases, opt, xinit);
  dump_problem(f, pd, "solution");
}

int main(int argc, char** argv) {
  if (argc < 5) { fprintf(stderr, "usage: ref_mhpc <quad_reference.csv> <hip_yaw> <in.txt> <out.txt> [mhpc_config.info]\n"); return 2; }
  g_hip_yaw = atof(argv[2]);
  std::ifstream in(argv[3]);
  int n_prob = 0, n_upd = 0;
  in >> n_prob >> n_upd;
  FILE* f = fopen(argv[4], "w");
  if (!in.good() || !f) { fprintf(stderr, "cannot open input / output\n"); return 2; }
  fprintf(f, "n_problems %d n_updates %d\n", n_prob, n_upd);
  for (int b = 0; b < n_prob; ++b) {
    VecM<T, 36> xinit;
    for (int i = 0; i < 36; ++i) in >> xinit[i];
    std::vector<VecM<T, 36>> nudge(n_upd);
    for (int u = 0; u < n_upd; ++u) for (int i = 0; i < 36; ++i) in >> nudge[u][i];

    // MHPCLocomotion<T>::initialize (MHPCLocomotion.cpp:20-66)
    MHPCConfig mpc_config;
    loadMHPCConfig(argc > 5 ? argv[5] : "../MHPC/settings/mhpc_config.info", mpc_config);
    HSDDP_OPTION ddp_setting;
    loadHSDDPSetting("../MHPC/settings/ddp_setting.info", ddp_setting);
    MHPCProblem<T> opt_problem;
    MHPCProblemData<T> opt_problem_data;
    opt_problem_data.quad_reference = std::make_shared<QuadReference>();
    opt_problem_data.quad_reference->load_top_level_data(argv[1], false);
    opt_problem.set_problem_data(&opt_problem_data, &mpc_config);
    opt_problem.initialization();

    fprintf(f, "problem %d\n", b);
    dump_problem(f, opt_problem_data, "guess");
    run_solve(f, opt_problem_data, ddp_setting, xinit);

    // run-time DDP setting when re-solving DDP in MPC (MHPCLocomotion.cpp:85-87)
    ddp_setting.max_AL_iter = ddp_setting.max_AL_iter_runtime;
    ddp_setting.max_DDP_iter = ddp_setting.max_DDP_iter_runtime;
    for (int u = 0; u < n_upd; ++u) {
      opt_problem.update();
      // the "measured" state: the plan's own prediction for the new start (front of the shifted nominal trajectory) + the nudge
      xinit = opt_problem_data.wb_trajs.front()->Xbar.front();
      xinit += nudge[u];
      fprintf(f, "update %d\n", u);
      dump_problem(f, opt_problem_data, "guess");
      run_solve(f, opt_problem_data, ddp_setting, xinit);
    }
  }
  fclose(f);
  return 0;
}
