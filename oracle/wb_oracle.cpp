/* wb_oracle.cpp — CPU ORACLE (test infrastructure only): whole-body phase. Placeholder until the WB model lands. */
#include "hsddp_oracle.hpp"
namespace oracle { std::unique_ptr<Phase> make_wb_phase(double) { return nullptr; } }
