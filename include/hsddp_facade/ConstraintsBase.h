// ConstraintsBase.h — the two parameter structs of HSDDPSolver/header/ConstraintsBase.h:58-86 that problem set-up code touches.
// The constraint classes themselves (PathConstraintBase, TerminalConstraintBase, ConstraintContainer) run inside the CUDA kernels
// (csrc/model_*.cuh, k_select) and have no host objects here.
#pragma once
#include <cmath>
template <typename T>
struct AL_Param_Struct {
  T lambda = 0;
  T sigma = 0;
  T sigma_max = 0;
  void update_penalty(T beta) { sigma *= beta; }
  void update_Lagrange(T h) { lambda += h * sigma; }
};
template <typename T>
struct REB_Param_Struct {
  T delta = 0.1;
  T delta_min = 0.01;
  T eps = 1;
  void update_relax(T beta) { delta *= beta; delta = std::fmax(delta, delta_min); }
  void update_weight(T beta) { eps *= beta; }
};
