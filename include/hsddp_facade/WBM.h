// WBM.h — dimensions of the whole-body model (MHPC/MHPC-Trajopt/WBM.h:13-16); the model itself is csrc/model_wb.cuh
#pragma once
#include <cstddef>
namespace WBM { const size_t xs = 36; const size_t us = 12; const size_t ys = 12; }
