// HSDDP_CPPTypes.h — same name as HSDDPSolver/common/HSDDP_CPPTypes.h; the aliases live in cppTypes.h
#pragma once
#include "cppTypes.h"
