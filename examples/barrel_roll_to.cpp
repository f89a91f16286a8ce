// barrel_roll_to.cpp — main() of MHPC/MHPC-Trajopt/BarrelRoll/BarrelRollTO.cpp:65-275 on a batch of perturbed initial states, through
// include/cafe_solver.hpp: build the six-phase deck, start from the interpolated state trajectory, solve with br_ddp_setting.info.
// build:  g++ -std=c++17 -Iinclude examples/barrel_roll_to.cpp -Lcafe_mpc_b200 -lcafe_gpu -Wl,-rpath,$PWD/cafe_mpc_b200 -o barrel_roll_to
// run  :  ./barrel_roll_to data 64
#include <cstdio>
#include <cstdlib>
#include <string>
#include "cafe_solver.hpp"

int main(int argc, char** argv) {
  const std::string root = argc > 1 ? argv[1] : "data";
  const int B = argc > 2 ? std::atoi(argv[2]) : 16;
  const std::string dir = root + "/MHPC/MHPC-Trajopt/BarrelRoll/setting/";
  try {
    cafe::HSDDP_OPTION ddp_setting;
    cafe::loadHSDDPSetting(dir + "br_ddp_setting.info", ddp_setting);
    cafe::BarrelRollProblem problem;
    problem.initialization(dir + "br_cost_weights.JSON", dir + "br_constraint_params.info");
    // xinit of BarrelRollTO.cpp:96-112 (pos z 0.2183, qJ (0,-1,2) x4), small deterministic offsets per problem
    std::vector<double> x0((size_t)B * 36, 0.0);
    for (int b = 0; b < B; ++b) {
      double* x = &x0[(size_t)b * 36];
      x[2] = 0.2183 + 0.0002 * (b % 25);
      for (int l = 0; l < 4; ++l) { x[6 + 3 * l] = 0; x[7 + 3 * l] = -1.0; x[8 + 3 * l] = 2.0; }
      x[1] = 0.001 * (b % 7);
    }
    cafe::MultiPhaseDDP solver;
    solver.set_initial_condition(x0, B);
    solver.set_multiPhaseProblem(problem, B);
    solver.set_initial_guess(problem.initial_guess(x0.data(), B));
    solver.solve(ddp_setting);
    const auto info = solver.get_solver_info();
    std::printf("solved %d barrel-roll problems in %.1f ms (device); problem 0: %d DDP iterations, %d outer iterations, cost %.6f, touchdown violation %.2e\n",
                B, solver.solve_ms(), info[0].iter, info[0].outer_iter, info[0].cost, info[0].max_tconstr);
    return 0;
  } catch (const cafe::Error& e) {
    std::fprintf(stderr, "cafe error %d: %s\n", e.code, e.what());
    return 1;
  }
}
