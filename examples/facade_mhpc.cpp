// facade_mhpc.cpp — the reference's own set-up / solve / MPC-update sequence (MHPC-Trajopt/test/testMHPCProblem.cpp:9-89,
// MHPCLocomotion.cpp:20-150; HKDMPC.cpp:20-140 for the HKD part) compiled against include/hsddp_facade/ — reference class names,
// reference method signatures, the GPU path underneath. Run from a directory whose parent holds MHPC/, HKDMPC/ and Reference/
// (the reference's "../" convention), e.g. data/run. Prints one JSON line per solve; tests/test_gpu_facade.py compares them
// with the same solves made through the C ABI from Python.
// build: g++ -std=c++17 -Iinclude/hsddp_facade -Iinclude examples/facade_mhpc.cpp -Lcafe_mpc_b200 -lcafe_gpu -Wl,-rpath,$PWD/cafe_mpc_b200 -o facade_mhpc
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "HKDProblem.h"
#include "MHPCProblem.h"
#include "MultiPhaseDDP.h"
#include "QuadReference.h"

template <class Solver>
static void report(const char* tag, Solver& solver, const std::vector<double>& first_u, const std::vector<double>& x_end) {
  int n_iters, n_ls, n_reg; float ms;
  solver.get_solver_info(n_iters, n_ls, n_reg, ms);
  std::vector<float> cost, dyn, eqn, ineq;
  solver.get_solver_info(cost, dyn, eqn, ineq);
  std::printf("{\"tag\": \"%s\", \"iter\": %d, \"ls\": %d, \"reg_total\": %d, \"n_hist\": %zu, \"cost\": %.17g, \"feas\": %.17g, \"tconstr\": %.17g, \"pconstr\": %.17g, \"u0\": [",
              tag, n_iters, n_ls, n_reg, cost.size(), (double)solver.get_actual_cost(), (double)solver.get_dyn_infeasibility(),
              (double)solver.get_terminal_constraint_violation(), (double)solver.get_path_constraint_violation());
  for (size_t i = 0; i < first_u.size(); ++i) std::printf("%s%.17g", i ? ", " : "", first_u[i]);
  std::printf("], \"x_end\": [");
  for (size_t i = 0; i < x_end.size(); ++i) std::printf("%s%.17g", i ? ", " : "", x_end[i]);
  std::printf("]}\n");
}

template <class V> static std::vector<double> as_std(const V& v) { std::vector<double> o(v.size()); for (int i = 0; i < (int)v.size(); ++i) o[i] = v[i]; return o; }

static int run_mhpc(int n_updates, int start_row) {
  MHPCConfig config;
  loadMHPCConfig("../MHPC/settings/mhpc_config.info", config);
  config.referenceFileName = "trot/heuristic";   // the shipped config names a reference that is not in the repository ("bound")
  std::shared_ptr<QuadReference> quad_ref = std::make_shared<QuadReference>();
  std::string file("../Reference/Data/");
  file.append(config.referenceFileName);
  file.append("/quad_reference.csv");
  quad_ref->load_top_level_data(file, false);
  if (start_row > 0) quad_ref->step(0.01f * start_row);   // begin the plan start_row reference rows into the file

  MHPCProblem<double> problem;
  MHPCProblemData<double> pdata;
  pdata.quad_reference = quad_ref;
  problem.set_problem_data(&pdata, &config);
  problem.prepare_initialization();
  problem.initialize_parameters();
  problem.initialize_multiPhaseProblem();

  MultiPhaseDDP<double> solver;
  auto collect = [&]() {
    std::deque<shared_ptr<SinglePhaseBase<double>>> phases;
    for (auto phase : pdata.wb_phases) phases.push_back(phase);
    if (pdata.srb_phase.get() != nullptr) phases.push_back(pdata.srb_phase);
    return phases;
  };

  VecM<double, WBM::xs> xinit;
  VecM<double, 3> pos, eul, vel, eulrate;
  VecM<double, 12> qJ, qJd;
  pos.setZero(); eul.setZero(); vel.setZero(); eulrate.setZero(); qJd.setZero();
  qJ = Vec3<double>(0, -0.8, 1.6).replicate<4, 1>();
  pos[2] = 0.2486;
  xinit << pos, eul, qJ, vel, eulrate, qJd;

  HSDDP_OPTION ddp_setting;
  loadHSDDPSetting("../MHPC/settings/ddp_setting.info", ddp_setting);

  for (auto& tau_i : pdata.wb_trajs) for (int k = 0; k < tau_i->size() - 1; k++) tau_i->Ubar[k].setConstant(.0);
  if (pdata.srb_phase.get() != nullptr) for (int k = 0; k < pdata.srb_traj->size() - 1; k++) pdata.srb_traj->Ubar[k].setZero();

  solver.set_initial_condition(xinit);
  solver.set_multiPhaseProblem(collect());
  solver.solve(ddp_setting);
  report("mhpc_initial", solver, as_std(pdata.wb_trajs.front()->Ubar[0]), as_std(pdata.wb_trajs.back()->Xbar.back()));

  // MPC updates with the run-time caps (MHPCLocomotion.cpp:91-150): the plan's own prediction stands in for the simulator state
  ddp_setting.max_AL_iter = ddp_setting.max_AL_iter_runtime;
  ddp_setting.max_DDP_iter = ddp_setting.max_DDP_iter_runtime;
  for (int loop = 0; loop < n_updates; ++loop) {
    // the state the plan predicts after one MPC period (nsteps knots ahead, across a phase boundary if need be)
    int pidx, k_pidx;
    pdata.get_index(problem.get_num_control_steps(), pidx, k_pidx);
    VecM<double, WBM::xs> xnext = pdata.wb_trajs[pidx]->Xbar[k_pidx];
    problem.update();
    solver.set_initial_condition(xnext);
    solver.set_multiPhaseProblem(collect());
    solver.solve(ddp_setting);
    char tag[32]; std::snprintf(tag, sizeof tag, "mhpc_update_%d", loop + 1);
    report(tag, solver, as_std(pdata.wb_trajs.front()->Ubar[0]), as_std(pdata.wb_trajs.back()->Xbar.back()));
  }
  return 0;
}

static int run_hkd(int n_updates, int start_row) {
  QuadReference quad_ref;
  quad_ref.load_top_level_data("../Reference/Data/trot/heuristic/quad_reference.csv", true);
  if (start_row > 0) quad_ref.step(0.01f * start_row);
  HKDPlanConfig config{0.6f, 0.01f, 2};   // HKDMPC.cpp:26-28
  HKDProblemData<double> pdata;
  pdata.quad_ref_ptr = &quad_ref;
  HKDProblem<double> problem;
  problem.set_problem_data(&pdata, config);
  problem.initialization();

  HSDDP_OPTION ddp_setting;
  loadHSDDPSetting("../HKDMPC/settings/ddp_setting.info", ddp_setting);
  MultiPhaseDDP<double> solver;
  auto collect = [&]() { std::deque<shared_ptr<SinglePhaseBase<double>>> phases; for (auto p : pdata.phase_ptrs) phases.push_back(p); return phases; };
  VecM<double, 24> x0 = pdata.trajectory_ptrs.front()->Xbar.front();
  x0[5] += 0.01; x0[9] += 0.05;   // a little off the reference: height and forward speed
  solver.set_initial_condition(x0);
  solver.set_multiPhaseProblem(collect());
  solver.solve(ddp_setting);
  report("hkd_initial", solver, as_std(pdata.trajectory_ptrs.front()->Ubar[0]), as_std(pdata.trajectory_ptrs.back()->Xbar.back()));
  ddp_setting.max_AL_iter = ddp_setting.max_AL_iter_runtime;
  ddp_setting.max_DDP_iter = ddp_setting.max_DDP_iter_runtime;
  for (int loop = 0; loop < n_updates; ++loop) {
    VecM<double, 24> xnext = pdata.trajectory_ptrs.front()->horizon > config.nsteps_between_mpc ? pdata.trajectory_ptrs.front()->Xbar[config.nsteps_between_mpc]
                                                                                                  : pdata.trajectory_ptrs[1]->Xbar[config.nsteps_between_mpc - pdata.trajectory_ptrs.front()->horizon];
    problem.update();
    solver.set_initial_condition(xnext);
    solver.set_multiPhaseProblem(collect());
    solver.solve(ddp_setting);
    char tag[32]; std::snprintf(tag, sizeof tag, "hkd_update_%d", loop + 1);
    report(tag, solver, as_std(pdata.trajectory_ptrs.front()->Ubar[0]), as_std(pdata.trajectory_ptrs.back()->Xbar.back()));
  }
  return 0;
}

int main(int argc, char** argv) {
  const int n_updates = argc > 2 ? std::atoi(argv[2]) : 2;
  const int start_row = argc > 3 ? std::atoi(argv[3]) : 0;
  try {
    if (argc > 1 && !std::strcmp(argv[1], "hkd")) return run_hkd(n_updates, start_row);
    return run_mhpc(n_updates, start_row);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
}
