// SRBM.h — dimensions of the single-rigid-body model (MHPC/MHPC-Trajopt/SRBM.h:13-15); the model itself is csrc/model_srb.cuh
#pragma once
#include <cstddef>
namespace SRBM { const size_t xs = 12; const size_t us = 12; const size_t ys = 0; }
