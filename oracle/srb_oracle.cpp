/* srb_oracle.cpp — CPU ORACLE (test infrastructure only): single-rigid-body phase. Placeholder until the SRB model lands. */
#include "hsddp_oracle.hpp"
namespace oracle { std::unique_ptr<Phase> make_srb_phase() { return nullptr; } }
