// TrajectoryManagement.h — Trajectory<T,xs,us,ys> with the members and sizes of the reference
// (HSDDPSolver/header/TrajectoryManagement.h:22-85, source/TrajectoryManagement.cpp:5-40, :130-228). Host-side container only: the
// solver reads Xbar / Ubar / K from it before a solve (the guess the reference's solve would start from) and writes the solution
// back after it (MultiPhaseDDP.h of this directory). The LQ arrays A, B, C, D, H and the cost data stay on the device.
#pragma once
#include <deque>
#include <memory>
#include <vector>
#include "HSDDP_CPPTypes.h"
#include "HSDDP_CompoundTypes.h"

using std::deque;
using std::shared_ptr;
using std::vector;

template <typename T, size_t xs, size_t us, size_t ys>
class Trajectory {
 public:
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
  Trajectory() {}
  Trajectory(T timeStep_, int horizon_) { create_data(timeStep_, horizon_); }
  void create_data(T timeStep_, int horizon_) {
    timeStep = timeStep_; horizon = horizon_; duration = timeStep * horizon;
    Xbar.assign(horizon + 1, VecM<T, xs>::Zero()); X.assign(horizon + 1, VecM<T, xs>::Zero());
    Ubar.assign(horizon, VecM<T, us>::Zero()); U.assign(horizon, VecM<T, us>::Zero()); Y.assign(horizon, VecM<T, ys>::Zero());
    Xsim.assign(horizon + 1, VecM<T, xs>::Zero()); Defect_bar.assign(horizon + 1, VecM<T, xs>::Zero()); Defect.assign(horizon + 1, VecM<T, xs>::Zero());
    V.assign(horizon + 1, 0); dV.assign(horizon + 1, 0);
    dU.assign(horizon, VecM<T, us>::Zero()); Qu.assign(horizon, VecM<T, us>::Zero());
    Quu.assign(horizon, MatMN<T, us, us>::Zero()); Qux.assign(horizon, MatMN<T, us, xs>::Zero());
    G.assign(horizon + 1, VecM<T, xs>::Zero()); K.assign(horizon + 1, MatMN<T, us, xs>::Zero()); dX.assign(horizon + 1, VecM<T, xs>::Zero());
  }
  void zero_all() {
    for (auto* d : {&Xbar, &X, &Xsim, &Defect_bar, &Defect, &G, &dX}) for (auto& m : *d) m.setZero();
    for (auto* d : {&Ubar, &U, &dU, &Qu}) for (auto& m : *d) m.setZero();
    for (auto& m : Y) m.setZero();
    zero_val_approx();
  }
  void zero_val_approx() {
    for (auto& v : V) v = 0;
    for (auto& v : dV) v = 0;
    for (auto* d : {&dU, &Qu}) for (auto& m : *d) m.setZero();
    for (auto* d : {&G, &dX}) for (auto& m : *d) m.setZero();
    for (auto* d : {&Qux, &K}) for (auto& m : *d) m.setZero();
    for (auto& m : Quu) m.setZero();
  }
  void clear() {
    Xbar.clear(); X.clear(); Ubar.clear(); U.clear(); Y.clear(); Xsim.clear(); Defect_bar.clear(); Defect.clear();
    V.clear(); dV.clear(); dU.clear(); Qu.clear(); Quu.clear(); Qux.clear(); G.clear(); K.clear(); dX.clear();
    horizon = 0; timeStep = 0; duration = 0;
  }
  void update_nominal_vals() { Xbar = X; Ubar = U; Defect_bar = Defect; }
  int size() { return (int)Xbar.size(); }

 public:
  T duration = 0;
  T timeStep = 0;
  int horizon = 0;  // the state trajectory holds horizon + 1 knots
  deque<VecM<T, xs>> Xbar, X;
  deque<VecM<T, us>> Ubar, U;
  deque<VecM<T, ys>> Y;
  deque<VecM<T, xs>> Xsim, Defect_bar, Defect;
  deque<T> V, dV;
  deque<VecM<T, us>> dU, Qu;
  deque<MatMN<T, us, us>> Quu;
  deque<MatMN<T, us, xs>> Qux;
  deque<VecM<T, xs>> G;
  deque<MatMN<T, us, xs>> K;
  deque<VecM<T, xs>> dX;
};
