// mhpc_problem.cpp — phase deck for the MHPC (whole-body + single-rigid-body) problem; host-side mirror of
// MHPCProblem<T>::initialization (/root/reference/MHPC/MHPC-Trajopt/MHPCProblem.cpp:13-250).
#include <stdexcept>
#include "info_reader.h"
#include "problem_builders.h"

namespace cafe {

void loadMHPCConfig(const std::string& fname, MHPCConfig& c) {  // MHPCProblem.h:67-83
  InfoFile pt(fname);
  c.plan_dur_wb = (float)pt.num("config.plan_dur_wb");
  c.plan_dur_srb = (float)pt.num("config.plan_dur_srb");
  c.dt_mpc = (float)pt.num("config.dt_mpc");
  c.dt_wb = (float)pt.num("config.dt_wb");
  c.dt_srb = (float)pt.num("config.dt_srb");
  c.BG_alpha = pt.num("config.BG_alpha");
  c.num_threads = pt.integer("config.nthreads");
  c.referenceFileName = pt.str("config.referenceFile");
  c.costFileName = pt.str("config.costFile");
  c.constraintParamFileName = pt.str("config.constraintParamFile");
}

void MHPCProblem::set_problem_data(QuadReference* quad_ref, const MHPCConfig& config, const std::string& settings_root) {
  quad_reference = quad_ref;
  pconfig = config;
  root = settings_root;
  plan_dur_all = config.plan_dur_wb + config.plan_dur_srb;
}

void MHPCProblem::initialization(DeckStorage&) { throw std::runtime_error("MHPC deck builder not implemented yet"); }

}  // namespace cafe
