// mhpc_batch.cpp — the three-call pattern of the reference (set_initial_condition -> set_multiPhaseProblem -> solve,
// MHPC/MHPCLocomotion.cpp:52-66, MHPC-Trajopt/test/testMHPCProblem.cpp:59-89) on a batch, through include/cafe_solver.hpp.
// build:  g++ -std=c++17 -Iinclude examples/mhpc_batch.cpp -Lcafe_mpc_b200 -lcafe_gpu -Wl,-rpath,$PWD/cafe_mpc_b200 -o mhpc_batch
// run  :  ./mhpc_batch data 256
#include <cstdio>
#include <cstdlib>
#include <string>
#include "cafe_solver.hpp"

int main(int argc, char** argv) {
  const std::string root = argc > 1 ? argv[1] : "data";
  const int B = argc > 2 ? std::atoi(argv[2]) : 64;
  try {
    cafe::HSDDP_OPTION ddp_setting;
    cafe::loadHSDDPSetting(root + "/MHPC/settings/ddp_setting.info", ddp_setting);
    cafe::MHPCProblem problem;
    problem.initialization(root + "/Reference/Data/trot/heuristic/quad_reference.csv", root + "/MHPC/settings/mhpc_config.info", root);
    // x0: pos (0,0,0.2183), qJ (0,-1,2) x4 (Loco_TO.cpp:49-55), small deterministic offsets per problem
    std::vector<double> x0((size_t)B * 36, 0.0);
    for (int b = 0; b < B; ++b) {
      double* x = &x0[(size_t)b * 36];
      x[2] = 0.2183 + 0.0001 * (b % 50);
      for (int l = 0; l < 4; ++l) { x[6 + 3 * l] = 0; x[7 + 3 * l] = -1.0; x[8 + 3 * l] = 2.0; }
      x[18] = 0.002 * (b % 17);
    }
    cafe::MultiPhaseDDP solver;
    solver.set_initial_condition(x0, B);
    solver.set_multiPhaseProblem(problem, B);
    solver.solve(ddp_setting);
    const auto info = solver.get_solver_info();
    std::printf("solved %d MHPC problems in %.2f ms (device); problem 0: %d DDP iterations, cost %.9f, feasibility %.3e\n", B, solver.solve_ms(),
                info[0].iter, info[0].cost, info[0].feas);
    return 0;
  } catch (const cafe::Error& e) {
    std::fprintf(stderr, "cafe error %d: %s\n", e.code, e.what());
    return 1;
  }
}
