/*
 * hsddp_oracle.hpp — CPU ORACLE (test infrastructure, never shipped, never on the
 * product path): plain-C++17 restatement of the reference's HS-DDP solver.
 *
 *   MultiPhaseDDP<T>::solve / line_search / backward_sweep(_regularized) / linear_rollout /
 *   hybrid_rollout                 /root/reference/HSDDPSolver/source/MultiPhaseDDP.cpp:12-447
 *   SinglePhase<T,xs,us,ys>::*     /root/reference/HSDDPSolver/source/SinglePhase.cpp:145-450
 *   Trajectory                     /root/reference/HSDDPSolver/source/TrajectoryManagement.cpp:5-38,122-127,231-259
 *   PathConstraintBase / TerminalConstraintBase / ConstraintContainer
 *                                  /root/reference/HSDDPSolver/header/ConstraintsBase.h:114-587
 *   CostContainer semantics        /root/reference/HSDDPSolver/source/SinglePhaseInterface.cpp:136-181
 *
 * Eigen is not available in this image; dense column-major helpers below stand in
 * for it, including a restatement of Eigen 3.3's pivoted LDLT (used by the
 * reference for the positive-definiteness test and for Quu^-1, SinglePhase.cpp:366-375).
 *
 * PARITY STATUS: pinned. The reference's own solver and problem code (HSDDPSolver,
 * HKD-TrajOpt, MHPC-Trajopt, Loco_TO.cpp, BarrelRollTO.cpp) compiles unchanged from
 * /root/reference against stand-ins for Eigen / Boost / LCM / Pinocchio (oracle/refbuild,
 * binaries in oracle/_ref); its records are committed (tests/golden/ref_*.npz) and this
 * restatement reproduces them decision for decision, the rest at 1e-9
 * (tests/test_cpu_reference_solver.py). The model math is pinned by the reference's CasADi C
 * compiled unchanged into oracle/_ref; NOT pinned by those records: the rigid-body algorithms
 * (oracle/wb_dynamics.hpp), which also stand behind the Pinocchio names in that build - they
 * are held to the reference's known answers (test/testKKTDynamics.cpp) and its CasADi partials.
 */
#pragma once
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>
#include <algorithm>
#include "../include/cafe_deck.h"

namespace oracle {

typedef std::vector<double> Vec;

struct Mat {  // column-major dense
  int r = 0, c = 0;
  std::vector<double> a;
  Mat() {}
  Mat(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_, 0.0) {}
  double& operator()(int i, int j) { return a[(size_t)i + (size_t)r * j]; }
  double operator()(int i, int j) const { return a[(size_t)i + (size_t)r * j]; }
  void zero() { std::fill(a.begin(), a.end(), 0.0); }
  void identity() { zero(); for (int i = 0; i < std::min(r, c); ++i) (*this)(i, i) = 1.0; }
};

inline Vec zeros(int n) { return Vec((size_t)n, 0.0); }
inline double dot(const Vec& a, const Vec& b) { double s = 0; for (size_t i = 0; i < a.size(); ++i) s += a[i] * b[i]; return s; }
/* y = A x */
inline Vec mv(const Mat& A, const Vec& x) { Vec y(A.r, 0.0); for (int j = 0; j < A.c; ++j) for (int i = 0; i < A.r; ++i) y[i] += A(i, j) * x[j]; return y; }
/* y = A^T x */
inline Vec mtv(const Mat& A, const Vec& x) { Vec y(A.c, 0.0); for (int j = 0; j < A.c; ++j) { double s = 0; for (int i = 0; i < A.r; ++i) s += A(i, j) * x[i]; y[j] = s; } return y; }
inline Mat mm(const Mat& A, const Mat& B) { Mat C(A.r, B.c); for (int j = 0; j < B.c; ++j) for (int k = 0; k < A.c; ++k) { double b = B(k, j); if (b == 0.0) continue; for (int i = 0; i < A.r; ++i) C(i, j) += A(i, k) * b; } return C; }
inline Mat mtm(const Mat& A, const Mat& B) { Mat C(A.c, B.c); for (int j = 0; j < B.c; ++j) for (int i = 0; i < A.c; ++i) { double s = 0; for (int k = 0; k < A.r; ++k) s += A(k, i) * B(k, j); C(i, j) = s; } return C; }
inline Mat transpose(const Mat& A) { Mat T(A.c, A.r); for (int j = 0; j < A.c; ++j) for (int i = 0; i < A.r; ++i) T(j, i) = A(i, j); return T; }
inline void axpy(Vec& y, double a, const Vec& x) { for (size_t i = 0; i < y.size(); ++i) y[i] += a * x[i]; }
inline void madd(Mat& Y, double a, const Mat& X) { for (size_t i = 0; i < Y.a.size(); ++i) Y.a[i] += a * X.a[i]; }
/* Y += a * x x^T */
inline void rank1(Mat& Y, double a, const Vec& x) { int n = (int)x.size(); for (int j = 0; j < n; ++j) { double xj = a * x[j]; if (xj == 0.0) continue; for (int i = 0; i < n; ++i) Y(i, j) += x[i] * xj; } }

/* Restatement of Eigen 3.3 LDLT<MatrixXd, Lower> (in-place, diagonal pivoting on the largest
 * |diagonal| of the trailing block, sign tracking) as used through `Chol<T>` =
 * Eigen::LDLT<DMat<T>> (HSDDP_CPPTypes.h:64). */
struct PivLDLT {
  int n = 0;
  Mat m;                 // L (unit lower) and D on the diagonal
  std::vector<int> tr;   // transpositions
  int sign = 0;          // 0 ZeroSign, +1 PositiveSemiDef, -1 NegativeSemiDef, 2 Indefinite
  double min_pivot = 0;
  void compute(const Mat& A);
  bool isPositive() const { return sign == 1 || sign == 0; }
  Mat solveIdentity() const;  // (P^T L^-T D^+ L^-1 P) * I
};

struct RCost {
  double l = 0;
  Vec lx, lu, ly;
  Mat lxx, luu, lyy;  // lux is identically zero for every cost in the reference (SinglePhaseInterface.cpp:47)
  void init(int n, int m, int p) { lx = zeros(n); lu = zeros(m); ly = zeros(p); lxx = Mat(n, n); luu = Mat(m, m); lyy = Mat(p, p); }
  void zero() { l = 0; std::fill(lx.begin(), lx.end(), 0.0); std::fill(lu.begin(), lu.end(), 0.0); std::fill(ly.begin(), ly.end(), 0.0); lxx.zero(); luu.zero(); lyy.zero(); }
};
struct TCost {
  double Phi = 0;
  Vec Phix;
  Mat Phixx;
  void init(int n) { Phix = zeros(n); Phixx = Mat(n, n); }
  void zero() { Phi = 0; std::fill(Phix.begin(), Phix.end(), 0.0); Phixx.zero(); }
};

struct IneqData { double g = 0; Vec gx, gu, gy; };
struct PathConstraint {  // PathConstraintBase, ConstraintsBase.h:114-318 (g_zz are identically zero in shipped constraints)
  int kind = 0;          // model-specific id
  int size = 0, len = 0;
  std::vector<std::vector<IneqData>> data;      // [k][i]
  std::vector<std::vector<CafeRebParam>> params; // [k][i]
  double max_violation = 0;
  void create(int size_, int len_, int n, int m, int p, const CafeRebParam& init);
  void update_max_violation(int k);
  double reb_cost(int k) const;
  void reb_partials(int k, Vec& gu_, Vec& gx_, Vec& gy_, Mat& hu_, Mat& hx_, Mat& hy_) const;
  void update_params(double thresh, double beta_relax, double beta_weight);
};
struct TermData { double h = 0; Vec hx; };
struct TermConstraint {  // TerminalConstraintBase, ConstraintsBase.h:320-429
  int size = 0;
  std::vector<TermData> data;
  std::vector<CafeAlParam> params;
  double max_violation = 0;
  void create(int size_, int n, const CafeAlParam& init);
  void update_max_violation();
  double al_cost() const;
  void al_partials(Vec& grad, Mat& hess) const;
  void update_params(double thresh, double beta);
};

/* One phase = SinglePhase + Trajectory + its model callbacks. */
class Phase {
 public:
  const CafePhase* ph = nullptr;
  const double* ref = nullptr;  // records of this phase, [h+1][CAFE_REF_W]
  int n = 0, m = 0, p = 0, h = 0;
  double dt = 0;
  std::vector<Vec> Xbar, X, Xsim, Defect, Defect_bar, dX, G, Ubar, U, dU, Qu, Y;
  std::vector<Mat> A, B, C, D, H, K, Quu, Qux;
  std::vector<RCost> rcost;
  TCost tcost;
  std::vector<PathConstraint> pcon;
  std::vector<TermConstraint> tcon;
  Vec x_init, dx_init;
  double actual_cost = 0, dV_1 = 0, dV_2 = 0;
  double min_pivot = 1e300;  // smallest LDLT pivot seen in successful sweeps (tie monitoring)

  virtual ~Phase() {}
  void allocate(const CafePhase* ph_, const double* ref_);
  virtual void build_model() {}  // create constraint objects etc. once the phase data are known
  const double* rec(int k) const { return ref + (size_t)k * CAFE_REF_W; }

  /* model callbacks (the reference's std::function / virtual objects) */
  virtual void dynamics(Vec& xnext, Vec& y, const Vec& x, const Vec& u, int k) = 0;
  virtual void dynamics_partial(Mat& A_, Mat& B_, Mat& C_, Mat& D_, const Vec& x, const Vec& u, int k) = 0;
  virtual void running_cost(RCost& rc, const Vec& x, const Vec& u, const Vec& y, int k) = 0;      // sets rc.l (sum over cost objects)
  virtual void running_cost_par(RCost& rc, const Vec& x, const Vec& u, const Vec& y, int k) = 0;  // accumulates partials
  virtual void terminal_cost(TCost& tc, const Vec& x) = 0;
  virtual void terminal_cost_par(TCost& tc, const Vec& x) = 0;
  virtual void path_constraints(const Vec& x, const Vec& u, const Vec& y, int k) = 0;
  virtual void path_constraints_par(const Vec& x, const Vec& u, const Vec& y, int k) = 0;
  virtual void terminal_constraints(const Vec& x) = 0;
  virtual void terminal_constraints_par(const Vec& x) = 0;
  virtual Vec resetmap(const Vec& x) = 0;
  virtual Mat resetmap_partial(const Vec& x) = 0;  // rows = next state dim, cols = n

  /* SinglePhase methods */
  void linear_rollout(double eps);
  bool hybrid_rollout(double eps, bool MS);
  void compute_cost(const CafeOptions& o);
  void LQ_approximation(const CafeOptions& o);
  bool backward_sweep(double reg, const Vec& Gprime, const Mat& Hprime);
  void update_nominal();
  double defect_sq() const;
  double max_pconstr() const;
  double max_tconstr() const;
};

std::unique_ptr<Phase> make_hkd_phase();
std::unique_ptr<Phase> make_srb_phase();
std::unique_ptr<Phase> make_wb_phase(double BG_alpha, double hip_yaw);

/* MultiPhaseDDP */
class Solver {
 public:
  std::vector<std::unique_ptr<Phase>> phases;
  Vec x0;
  int iter_ = 0, ls_iter_total_ = 0, reg_iter_total_ = 0, iter_ou = 0;
  double actual_cost = 0, merit = 0, feas = 0, dV_1 = 0, dV_2 = 0;
  double max_tconstr_prev = 0, max_pconstr_prev = 0, max_tconstr = 0, max_pconstr = 0, merit_rho = 0;
  double last_eps = 0;
  bool reg_failed = false;
  std::vector<double> hist;   // 4 per push
  std::vector<double> trace;  // CAFE_TRACE_W per iteration

  void setup(const CafeDeck* deck);
  void solve(const CafeOptions& o);
  void linear_rollout(double eps);
  bool hybrid_rollout(double eps, const CafeOptions& o);
  std::pair<bool, int> line_search(const CafeOptions& o);
  void compute_cost(const CafeOptions& o);
  void LQ_approximation(const CafeOptions& o);
  bool backward_sweep(double reg);
  std::pair<bool, int> backward_sweep_regularized(double& reg, const CafeOptions& o);
  void update_nominal();
  double measure_dynamics_feasibility();
  void push_hist();
};

}  // namespace oracle
