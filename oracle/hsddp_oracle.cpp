/*
 * hsddp_oracle.cpp — CPU ORACLE (test infrastructure only; see hsddp_oracle.hpp).
 * Every function cites the reference lines it restates.
 */
#include "hsddp_oracle.hpp"
#include "oracle_api.h"
#include <cassert>
#include <stdexcept>
#include <string>

namespace oracle {

/* ---------------- Eigen 3.3 LDLT restatement (Eigen/src/Cholesky/LDLT.h, ldlt_inplace<Lower>::unblocked) */
void PivLDLT::compute(const Mat& Ain) {
  n = Ain.r;
  m = Ain;
  tr.assign(n, 0);
  sign = 0;
  min_pivot = 1e300;
  if (n <= 1) {
    if (n == 1) { double v = m(0, 0); sign = v > 0 ? 1 : (v < 0 ? -1 : 0); min_pivot = v; }
    return;
  }
  Vec temp(n, 0.0);
  bool found_zero_pivot = false;
  for (int k = 0; k < n; ++k) {
    /* largest |diagonal| of the trailing block; first index on ties (maxCoeff) */
    int piv = k;
    double best = std::fabs(m(k, k));
    for (int i = k + 1; i < n; ++i) { double v = std::fabs(m(i, i)); if (v > best) { best = v; piv = i; } }
    tr[k] = piv;
    if (piv != k) {
      /* symmetric row/column swap acting on the lower triangle only */
      int s = n - piv - 1;
      for (int j = 0; j < k; ++j) std::swap(m(k, j), m(piv, j));
      for (int i = 0; i < s; ++i) std::swap(m(piv + 1 + i, k), m(piv + 1 + i, piv));
      std::swap(m(k, k), m(piv, piv));
      for (int i = k + 1; i < piv; ++i) std::swap(m(i, k), m(piv, i));
    }
    int rs = n - k - 1;
    if (k > 0) {
      for (int j = 0; j < k; ++j) temp[j] = m(j, j) * m(k, j);
      double s = 0;
      for (int j = 0; j < k; ++j) s += m(k, j) * temp[j];
      m(k, k) -= s;
      for (int i = 0; i < rs; ++i) {
        double t = 0;
        for (int j = 0; j < k; ++j) t += m(k + 1 + i, j) * temp[j];
        m(k + 1 + i, k) -= t;
      }
    }
    double akk = m(k, k);
    bool pivot_is_valid = std::fabs(akk) > 0.0;
    if (k == 0 && !pivot_is_valid) { sign = 0; for (int j = 0; j < n; ++j) tr[j] = j; min_pivot = 0; return; }
    if (rs > 0 && pivot_is_valid) for (int i = 0; i < rs; ++i) m(k + 1 + i, k) /= akk;
    if (!pivot_is_valid) found_zero_pivot = true;
    (void)found_zero_pivot;
    if (sign == 1) { if (akk < 0) sign = 2; }
    else if (sign == -1) { if (akk > 0) sign = 2; }
    else if (sign == 0) { if (akk > 0) sign = 1; else if (akk < 0) sign = -1; }
    min_pivot = std::min(min_pivot, akk);
  }
}

Mat PivLDLT::solveIdentity() const {
  /* LDLT::_solve_impl: dst = P b; L^-1; D^-1 (pseudo); L^-T; P^T */
  Mat X(n, n);
  X.identity();
  for (int k = 0; k < n; ++k) if (tr[k] != k) for (int j = 0; j < n; ++j) std::swap(X(k, j), X(tr[k], j));
  for (int j = 0; j < n; ++j) {
    for (int i = 0; i < n; ++i) { double s = X(i, j); for (int l = 0; l < i; ++l) s -= m(i, l) * X(l, j); X(i, j) = s; }
    const double tol = 1.0 / 1.7976931348623157e308;
    for (int i = 0; i < n; ++i) { double d = m(i, i); if (std::fabs(d) > tol) X(i, j) /= d; else X(i, j) = 0; }
    for (int i = n - 1; i >= 0; --i) { double s = X(i, j); for (int l = i + 1; l < n; ++l) s -= m(l, i) * X(l, j); X(i, j) = s; }
  }
  for (int k = n - 1; k >= 0; --k) if (tr[k] != k) for (int j = 0; j < n; ++j) std::swap(X(k, j), X(tr[k], j));
  return X;
}

/* ---------------- PathConstraintBase (ConstraintsBase.h:114-318) */
void PathConstraint::create(int size_, int len_, int n, int m, int p, const CafeRebParam& init) {
  size = size_; len = len_;
  data.assign(len, std::vector<IneqData>(size));
  for (auto& dk : data) for (auto& d : dk) { d.gx = zeros(n); d.gu = zeros(m); d.gy = zeros(p); }
  params.assign(len, std::vector<CafeRebParam>(size, init));
  max_violation = 0;
}
void PathConstraint::update_max_violation(int k) {  // :217-228
  if (k == 0) max_violation = 0;
  double mk = 0;
  for (auto& c : data[k]) mk = std::min(mk, c.g);
  max_violation = std::min(max_violation, mk);
}
double PathConstraint::reb_cost(int k) const {  // :230-248
  double cost = 0;
  for (int i = 0; i < size; ++i) {
    double g = data[k][i].g, delta = params[k][i].delta, eps = params[k][i].eps, barr;
    if (g > delta) barr = -std::log(g);
    else { barr = .5 * (((g - 2 * delta) / delta) * ((g - 2 * delta) / delta) - 1); barr -= std::log(delta); }
    cost += eps * barr;
  }
  return cost;
}
void PathConstraint::reb_partials(int k, Vec& gu_, Vec& gx_, Vec& gy_, Mat& hu_, Mat& hx_, Mat& hy_) const {  // :250-289
  std::fill(gu_.begin(), gu_.end(), 0.0); std::fill(gx_.begin(), gx_.end(), 0.0); std::fill(gy_.begin(), gy_.end(), 0.0);
  hu_.zero(); hx_.zero(); hy_.zero();
  for (int i = 0; i < size; ++i) {
    const IneqData& d = data[k][i];
    double g = d.g, delta = params[k][i].delta, eps = params[k][i].eps, bd, bdd;
    if (g > delta) { bd = -1.0 / g; bdd = std::pow(g, -2); }
    else { bd = (g - 2 * delta) / delta / delta; bdd = std::pow(delta, -2); }
    axpy(gu_, eps * bd, d.gu); axpy(gx_, eps * bd, d.gx); axpy(gy_, eps * bd, d.gy);
    rank1(hu_, eps * bdd, d.gu); rank1(hx_, eps * bdd, d.gx); rank1(hy_, eps * bdd, d.gy);
  }
}
void PathConstraint::update_params(double thresh, double beta_relax, double beta_weight) {  // :194-209
  for (int k = 0; k < len; ++k) for (int i = 0; i < size; ++i) {
    if (data[k][i].g > -thresh) continue;
    params[k][i].eps *= beta_weight;
    params[k][i].delta *= beta_relax;
    params[k][i].delta = std::fmax(params[k][i].delta, params[k][i].delta_min);
  }
}

/* ---------------- TerminalConstraintBase (ConstraintsBase.h:320-429) */
void TermConstraint::create(int size_, int n, const CafeAlParam& init) {
  size = size_;
  data.assign(size, TermData());
  for (auto& d : data) d.hx = zeros(n);
  params.assign(size, init);
  max_violation = 0;
}
void TermConstraint::update_max_violation() { max_violation = 0; for (auto& c : data) max_violation = std::max(max_violation, std::fabs(c.h)); }  // :392-399
double TermConstraint::al_cost() const {  // :400-411
  double c = 0;
  for (int i = 0; i < size; ++i) { double s = params[i].sigma, l = params[i].lambda, h = data[i].h; c += 0.5 * s * h * h; c += l * h; }
  return c;
}
void TermConstraint::al_partials(Vec& grad, Mat& hess) const {  // :412-425, note the (sigma*(1+h)+lambda) Hessian factor
  std::fill(grad.begin(), grad.end(), 0.0); hess.zero();
  for (int i = 0; i < size; ++i) {
    double s = params[i].sigma, l = params[i].lambda, h = data[i].h;
    axpy(grad, s * h + l, data[i].hx);
    rank1(hess, s * (1 + h) + l, data[i].hx);
  }
}
void TermConstraint::update_params(double thresh, double beta) {  // :375-391
  for (int i = 0; i < size; ++i) {
    if (std::fabs(data[i].h) < thresh) continue;
    if (std::fabs(data[i].h) > 0.005) { params[i].sigma *= beta; params[i].sigma = std::min(params[i].sigma, params[i].sigma_max); }
    else params[i].lambda += data[i].h * params[i].sigma;
  }
}

/* ---------------- Trajectory::create_data (TrajectoryManagement.cpp:5-38) */
void Phase::allocate(const CafePhase* ph_, const double* ref_) {
  ph = ph_; ref = ref_;
  n = cafe_model_n(ph->model); m = cafe_model_m(ph->model); p = cafe_model_p(ph->model);
  h = ph->horizon; dt = ph->dt;
  auto vz = [](int cnt, int len) { return std::vector<Vec>(cnt, zeros(len)); };
  Xbar = vz(h + 1, n); X = vz(h + 1, n); Xsim = vz(h + 1, n); Defect = vz(h + 1, n); Defect_bar = vz(h + 1, n); dX = vz(h + 1, n); G = vz(h + 1, n);
  Ubar = vz(h, m); U = vz(h, m); dU = vz(h, m); Qu = vz(h, m); Y = vz(h, p);
  A.assign(h + 1, Mat(n, n)); B.assign(h, Mat(n, m)); C.assign(h, Mat(p, n)); D.assign(h, Mat(p, m));
  H.assign(h + 1, Mat(n, n)); K.assign(h + 1, Mat(m, n)); Quu.assign(h, Mat(m, m)); Qux.assign(h, Mat(m, n));
  rcost.assign(h, RCost());
  for (auto& r : rcost) r.init(n, m, p);
  tcost.init(n);
  x_init = zeros(n); dx_init = zeros(n);
  /* initial guess: X and Xbar from the state reference
     (HKDProblem.cpp:86-91, MHPCProblem.cpp:187-192, :226-230) */
  for (int k = 0; k <= h; ++k) for (int i = 0; i < n; ++i) { Xbar[k][i] = rec(k)[CAFE_REF_XR + i]; X[k][i] = Xbar[k][i]; }
}

/* SinglePhase::linear_rollout (SinglePhase.cpp:145-178) */
void Phase::linear_rollout(double eps) {
  dV_1 = 0; dV_2 = 0;
  for (int i = 0; i < n; ++i) dX[0][i] = dx_init[i] + eps * Defect[0][i];
  for (int k = 0; k < h; ++k) {
    const RCost& rc = rcost[k];
    const Vec& dxk = dX[k];
    Vec duk = mv(K[k], dxk);
    for (int i = 0; i < m; ++i) duk[i] = eps * dU[k][i] + duk[i];
    Vec a = mv(A[k], dxk), b = mv(B[k], duk);
    for (int i = 0; i < n; ++i) dX[k + 1][i] = a[i] + b[i] + eps * Defect[k + 1][i];
    dV_1 += dot(rc.lx, dxk) + dot(rc.lu, duk);
    dV_2 += dot(dxk, mv(rc.lxx, dxk));
    dV_2 += dot(duk, mv(rc.luu, duk));
    /* + duk^T lux dxk with lux == 0 */
  }
  const Vec& dxk = dX[h];
  dV_1 += dot(tcost.Phix, dxk);
  dV_2 += dot(dxk, mv(tcost.Phixx, dxk));
}

/* SinglePhase::hybrid_rollout (SinglePhase.cpp:182-233); every knot 0..h is a shooting state
 * (update_SS_config(h+1): HKDProblem.cpp:105, MHPCProblem.cpp:209,243). */
bool Phase::hybrid_rollout(double eps, bool MS) {
  Xsim[0] = x_init;
  /* empty SS_set: the freshly opened tail phase of an MPC update (MHPCProblem.cpp:366-369), or every phase when MS is off
   * (MultiPhaseDDP::hybrid_rollout calls update_SS_config(0) on each phase, MultiPhaseDDP.cpp:65-68) */
  const bool ss = ph->single_shooting != 0 || !MS;
  /* SS_set.front()==0 */
  if (!ss) for (int i = 0; i < n; ++i) X[0][i] = Xbar[0][i] + eps * dX[0][i];
  else X[0] = x_init;
  for (int k = 0; k < h; ++k) {
    Vec dx(n);
    for (int i = 0; i < n; ++i) dx[i] = X[k][i] - Xbar[k][i];
    Vec kd = mv(K[k], dx);
    for (int i = 0; i < m; ++i) U[k][i] = Ubar[k][i] + eps * dU[k][i] + kd[i];
    dynamics(Xsim[k + 1], Y[k], X[k], U[k], k);
    double nrm = 0;
    for (int i = 0; i < n; ++i) nrm += Xsim[k + 1][i] * Xsim[k + 1][i];
    if (std::sqrt(nrm) > 1e6) return false;
    if (MS && !ss) { for (int i = 0; i < n; ++i) X[k + 1][i] = Xbar[k + 1][i] + eps * dX[k + 1][i]; }
    else X[k + 1] = Xsim[k + 1];
    path_constraints(X[k], U[k], Y[k], k);
  }
  terminal_constraints(X[h]);
  for (int k = 0; k <= h; ++k) for (int i = 0; i < n; ++i) Defect[k][i] = Xsim[k][i] - X[k][i];  // TrajectoryManagement.cpp:231-238
  return true;
}

/* SinglePhase::compute_cost (SinglePhase.cpp:236-262, :394-402, :426-435) */
void Phase::compute_cost(const CafeOptions& o) {
  actual_cost = 0;
  for (int k = 0; k < h; ++k) {
    rcost[k].zero();
    running_cost(rcost[k], X[k], U[k], Y[k], k);
    if (o.ReB_active) for (auto& pc : pcon) rcost[k].l += dt * pc.reb_cost(k);
    actual_cost += rcost[k].l;
  }
  tcost.zero();
  terminal_cost(tcost, X[h]);
  if (o.AL_active) for (auto& tc : tcon) tcost.Phi += tc.al_cost();
  actual_cost += tcost.Phi;
}

/* SinglePhase::LQ_approximation (SinglePhase.cpp:265-320, :405-418, :438-450) */
void Phase::LQ_approximation(const CafeOptions& o) {
  for (int k = 0; k < h; ++k) dynamics_partial(A[k], B[k], C[k], D[k], X[k], U[k], k);
  Vec gu = zeros(m), gx = zeros(n), gy = zeros(p);
  Mat hu(m, m), hx(n, n), hy(p, p);
  for (int k = 0; k < h; ++k) {
    running_cost_par(rcost[k], X[k], U[k], Y[k], k);
    if (o.ReB_active) {
      path_constraints_par(X[k], U[k], Y[k], k);
      for (auto& pc : pcon) {
        pc.reb_partials(k, gu, gx, gy, hu, hx, hy);
        axpy(rcost[k].lu, dt, gu); axpy(rcost[k].lx, dt, gx); axpy(rcost[k].ly, dt, gy);
        madd(rcost[k].luu, dt, hu); madd(rcost[k].lxx, dt, hx); madd(rcost[k].lyy, dt, hy);
      }
    }
  }
  terminal_cost_par(tcost, X[h]);
  if (o.AL_active) {
    terminal_constraints_par(X[h]);
    Vec g = zeros(n); Mat hh(n, n);
    for (auto& tc : tcon) { tc.al_partials(g, hh); axpy(tcost.Phix, 1.0, g); madd(tcost.Phixx, 1.0, hh); }
  }
}

/* SinglePhase::backward_sweep (SinglePhase.cpp:323-391) */
bool Phase::backward_sweep(double reg, const Vec& Gprime, const Mat& Hprime) {
  for (int i = 0; i < n; ++i) G[h][i] = tcost.Phix[i] + Gprime[i];
  for (size_t i = 0; i < H[h].a.size(); ++i) H[h].a[i] = tcost.Phixx.a[i] + Hprime.a[i];
  dV_1 = 0; dV_2 = 0;
  PivLDLT chol;
  for (int k = h - 1; k >= 0; --k) {
    const RCost& rc = rcost[k];
    const Mat& Hn = H[k + 1];
    Vec Gn = G[k + 1];
    Vec hd = mv(Hn, Defect[k + 1]);
    for (int i = 0; i < n; ++i) Gn[i] += hd[i];
    Vec Qx = mtv(A[k], Gn);
    for (int i = 0; i < n; ++i) Qx[i] += rc.lx[i];
    Vec qu = mtv(B[k], Gn);
    for (int i = 0; i < m; ++i) qu[i] += rc.lu[i];
    Mat HA = mm(Hn, A[k]), HB = mm(Hn, B[k]);
    Mat Qxx = mtm(A[k], HA); madd(Qxx, 1.0, rc.lxx);
    Mat quu = mtm(B[k], HB); madd(quu, 1.0, rc.luu);
    Mat qux = mtm(B[k], HA);
    if (p > 0) {
      Vec cl = mtv(C[k], rc.ly), dl = mtv(D[k], rc.ly);
      for (int i = 0; i < n; ++i) Qx[i] += cl[i];
      for (int i = 0; i < m; ++i) qu[i] += dl[i];
      Mat SC = mm(rc.lyy, C[k]), SD = mm(rc.lyy, D[k]);
      madd(Qxx, 1.0, mtm(C[k], SC));
      madd(quu, 1.0, mtm(D[k], SD));
      madd(qux, 1.0, mtm(D[k], SC));
    }
    for (int i = 0; i < n; ++i) Qxx(i, i) += reg;
    for (int i = 0; i < m; ++i) quu(i, i) += reg;
    Qu[k] = qu; Quu[k] = quu; Qux[k] = qux;
    Mat shifted = quu;
    for (int i = 0; i < m; ++i) shifted(i, i) -= 1e-9;
    chol.compute(shifted);
    if (!chol.isPositive()) return false;
    min_pivot = std::min(min_pivot, chol.min_pivot);
    Mat Quu_inv = chol.solveIdentity();
    Mat QxxT = transpose(Qxx);
    for (size_t i = 0; i < Qxx.a.size(); ++i) Qxx.a[i] = (Qxx.a[i] + QxxT.a[i]) / 2;
    Vec du = mv(Quu_inv, qu);
    for (int i = 0; i < m; ++i) dU[k][i] = -du[i];
    Mat Kk = mm(Quu_inv, qux);
    for (auto& v : Kk.a) v = -v;
    K[k] = Kk;
    /* G = Qx - Qux^T Quu_inv Qu ; H = Qxx - Qux^T Quu_inv Qux */
    Vec qd = mtv(qux, du);
    for (int i = 0; i < n; ++i) G[k][i] = Qx[i] - qd[i];
    Mat QiQ = mm(Quu_inv, qux);
    Mat t = mtm(qux, QiQ);
    for (size_t i = 0; i < Qxx.a.size(); ++i) H[k].a[i] = Qxx.a[i] - t.a[i];
    double dV_k = -dot(qu, dU[k]);
    dV_1 -= dV_k;
    dV_2 += dV_k;
  }
  Vec hd = mv(H[0], Defect[0]);
  for (int i = 0; i < n; ++i) G[0][i] += hd[i];
  return true;
}

void Phase::update_nominal() { Xbar = X; Ubar = U; Defect_bar = Defect; }  // TrajectoryManagement.cpp:122-127
double Phase::defect_sq() const { double s = 0; for (auto& d : Defect) for (double v : d) s += v * v; return s; }  // :240-259
double Phase::max_pconstr() const { double v = 0; for (auto& c : pcon) v = std::min(v, c.max_violation); return v; }  // ConstraintsBase.h:494-502
double Phase::max_tconstr() const { double v = 0; for (auto& c : tcon) v = std::max(v, c.max_violation); return v; }  // :503-511

/* ---------------- MultiPhaseDDP */
void Solver::setup(const CafeDeck* deck) {
  phases.clear();
  for (int i = 0; i < deck->n_phases; ++i) {
    const CafePhase* ph = &deck->phase[i];
    std::unique_ptr<Phase> P;
    if (ph->model == CAFE_MODEL_HKD) P = make_hkd_phase();
    else if (ph->model == CAFE_MODEL_SRB) P = make_srb_phase();
    else P = make_wb_phase(deck->BG_alpha, deck->hip_yaw);
    if (!P) throw std::runtime_error("oracle: model not available");
    P->allocate(ph, deck->ref + (size_t)ph->knot_offset * CAFE_REF_W);
    P->build_model();
    phases.push_back(std::move(P));
  }
}

/* MultiPhaseDDP::linear_rollout (MultiPhaseDDP.cpp:12-42) */
void Solver::linear_rollout(double eps) {
  dV_1 = 0; dV_2 = 0;
  Vec dx_init = zeros(phases[0]->n);
  for (size_t i = 0; i < phases.size(); ++i) {
    if (i > 0) {
      Mat Px = phases[i - 1]->resetmap_partial(phases[i - 1]->X.back());
      dx_init = mv(Px, phases[i - 1]->dX.back());
    }
    phases[i]->dx_init = dx_init;
    phases[i]->linear_rollout(eps);
    dV_1 += phases[i]->dV_1;
    dV_2 += phases[i]->dV_2;
  }
}

/* MultiPhaseDDP::hybrid_rollout (MultiPhaseDDP.cpp:49-92) */
bool Solver::hybrid_rollout(double eps, const CafeOptions& o) {
  actual_cost = 0; max_pconstr = 0; max_tconstr = 0;
  Vec xinit = x0;
  for (size_t i = 0; i < phases.size(); ++i) {
    if (i > 0) xinit = phases[i - 1]->resetmap(phases[i - 1]->X.back());
    phases[i]->x_init = xinit;
    if (!phases[i]->hybrid_rollout(eps, o.MS != 0)) return false;
    max_pconstr = std::min(max_pconstr, phases[i]->max_pconstr());
    max_tconstr = std::max(max_tconstr, phases[i]->max_tconstr());
  }
  return true;
}

void Solver::compute_cost(const CafeOptions& o) { actual_cost = 0; for (auto& P : phases) { P->compute_cost(o); actual_cost += P->actual_cost; } }  // :449-458
void Solver::LQ_approximation(const CafeOptions& o) { for (auto& P : phases) P->LQ_approximation(o); }                                               // :460-467
void Solver::update_nominal() { for (auto& P : phases) P->update_nominal(); }                                                                         // :527-534
double Solver::measure_dynamics_feasibility() { double f = 0; for (auto& P : phases) f += P->defect_sq(); return std::sqrt(f); }                      // :536-552

/* MultiPhaseDDP::line_search (MultiPhaseDDP.cpp:95-133) */
std::pair<bool, int> Solver::line_search(const CafeOptions& o) {
  double eps = 1, merit_prev = merit, feas_prev = feas;
  bool success = false;
  int iter = 0;
  last_eps = 0;
  while (eps > 1e-3) {
    iter++;
    bool rollout_success = hybrid_rollout(eps, o);
    compute_cost(o);
    feas = measure_dynamics_feasibility();
    merit = actual_cost + merit_rho * feas;
    double exp_cost_change = eps * dV_1 + 0.5 * eps * eps * dV_2;
    double exp_merit_change = exp_cost_change - eps * merit_rho * feas_prev;
    if ((merit <= merit_prev + o.gamma * exp_merit_change) && rollout_success) { success = true; last_eps = eps; break; }
    eps *= o.alpha;
  }
  return std::make_pair(success, iter);
}

/* MultiPhaseDDP::backward_sweep (MultiPhaseDDP.cpp:174-213) + impact_aware_step (:499-503) */
bool Solver::backward_sweep(double reg) {
  dV_1 = 0; dV_2 = 0;
  int np = (int)phases.size();
  for (int i = np - 1; i >= 0; --i) {
    int xs = phases[i]->n;
    Vec Gp = zeros(xs);
    Mat Hp(xs, xs);
    if (i <= np - 2) {
      Mat Px = phases[i]->resetmap_partial(phases[i]->X.back());  // (n_next x n_i)
      Gp = mtv(Px, phases[i + 1]->G[0]);
      Hp = mtm(Px, mm(phases[i + 1]->H[0], Px));
    }
    if (!phases[i]->backward_sweep(reg, Gp, Hp)) return false;
    dV_1 += phases[i]->dV_1;
    dV_2 += phases[i]->dV_2;
  }
  return true;
}

/* MultiPhaseDDP::backward_sweep_regularized (MultiPhaseDDP.cpp:136-165) */
std::pair<bool, int> Solver::backward_sweep_regularized(double& reg, const CafeOptions& o) {
  bool success = false;
  int iter = 0;
  while (!success) {
    iter++;
    success = backward_sweep(reg);
    if (success) break;
    reg = std::max(reg * o.update_regularization, 1e-03);
    if (reg > 1e2) break;
  }
  reg = reg / 20;
  if (reg < 1e-06) reg = 0;
  return std::make_pair(success, iter);
}

void Solver::push_hist() { hist.push_back(actual_cost); hist.push_back(feas); hist.push_back(max_tconstr); hist.push_back(max_pconstr); }

/* MultiPhaseDDP::solve (MultiPhaseDDP.cpp:216-447), wall-clock exits removed (max_cputime = 1e6 never fires) */
void Solver::solve(const CafeOptions& o) {
  iter_ = 0; ls_iter_total_ = 0; reg_iter_total_ = 0; iter_ou = 0;
  int iter_in = 0;
  double cost_prev = 0;
  reg_failed = false;
  hist.clear(); trace.clear();
  hybrid_rollout(0, o);
  update_nominal();
  compute_cost(o);
  feas = measure_dynamics_feasibility();
  push_hist();
  double regularization = 0;
  while (iter_ou < o.max_AL_iter) {
    iter_ou++;
    max_tconstr_prev = max_tconstr;
    max_pconstr_prev = max_pconstr;
    regularization = 0;
    iter_in = 0;
    while (iter_in < o.max_DDP_iter) {
      compute_cost(o);
      feas = measure_dynamics_feasibility();
      iter_in++;
      iter_++;
      double tr[CAFE_TRACE_W] = {0};
      tr[0] = actual_cost; tr[1] = feas;
      LQ_approximation(o);
      auto br = backward_sweep_regularized(regularization, o);
      reg_iter_total_ += br.second;
      tr[5] = regularization; tr[6] = br.second;
      if (!br.first) { reg_failed = true; trace.insert(trace.end(), tr, tr + CAFE_TRACE_W); return; }
      if (o.MS) linear_rollout(1.0);
      double dV_abs = std::fabs(dV_1 + 0.5 * dV_2);
      merit_rho = (feas > o.dynamics_feas_thresh) ? dV_abs / ((1 - o.merit_scale) * feas) + o.merit_offset : 0;
      merit = actual_cost + merit_rho * feas;
      cost_prev = actual_cost;
      double merit_prev = merit;
      tr[2] = dV_1; tr[3] = dV_2; tr[4] = merit_rho;
      if ((dV_abs < o.cost_thresh) && (feas <= o.dynamics_feas_thresh)) {
        tr[10] = actual_cost; tr[11] = feas;
        trace.insert(trace.end(), tr, tr + CAFE_TRACE_W);
        break;
      }
      auto ls = line_search(o);
      ls_iter_total_ += ls.second;
      if (ls.first) update_nominal();
      else { actual_cost = cost_prev; merit = merit_prev; }
      tr[7] = ls.second; tr[8] = ls.first ? 1 : 0; tr[9] = last_eps; tr[10] = actual_cost; tr[11] = feas;
      trace.insert(trace.end(), tr, tr + CAFE_TRACE_W);
      if ((std::fabs((cost_prev - actual_cost) / cost_prev) < o.cost_thresh) && (feas <= o.dynamics_feas_thresh)) break;
      push_hist();
    }
    if (max_tconstr < o.tconstr_thresh && std::fabs(max_pconstr) < o.pconstr_thresh && feas <= o.dynamics_feas_thresh) break;
    if (std::fabs(max_tconstr - max_tconstr_prev) < 0.0001 && std::fabs(max_pconstr - max_pconstr_prev) < 0.0001 && feas <= o.dynamics_feas_thresh) break;
    if (o.AL_active) for (auto& P : phases) for (auto& tc : P->tcon) tc.update_params(o.tconstr_thresh, o.update_penalty);
    if (o.ReB_active) for (auto& P : phases) for (auto& pc : P->pcon) pc.update_params(o.pconstr_thresh, o.update_relax, o.update_ReB);
    if (iter_ou >= o.max_AL_iter) break;
  }
}

}  // namespace oracle

/* ---------------- C API */
using namespace oracle;

/* TEST HOOK: the Eigen-3.3 LDLT restatement on its own. A: n x n column-major (lower part used). Returns isPositive();
 * inv receives solve(Identity) (what SinglePhase::backward_sweep uses as Quu_inv), pivot the smallest |D| entry. */
extern "C" int cafe_oracle_ldlt(const double* A, int n, double* inv, double* min_pivot) {
  oracle::Mat M(n, n);
  for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) M(i, j) = A[i + n * j];
  oracle::PivLDLT f;
  f.compute(M);
  if (inv) { oracle::Mat X = f.solveIdentity(); for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) inv[i + n * j] = X(i, j); }
  if (min_pivot) *min_pivot = f.min_pivot;
  return f.isPositive() ? 1 : 0;
}


extern "C" long cafe_oracle_solution_size(const CafeDeck* deck) {
  long s = 0;
  for (int i = 0; i < deck->n_phases; ++i) {
    const CafePhase& ph = deck->phase[i];
    long n = cafe_model_n(ph.model), m = cafe_model_m(ph.model), p = cafe_model_p(ph.model), h = ph.horizon;
    s += (h + 1) * n + h * m + h * p + h * m + h * m * n + h * m + h * m * m + h * m * n + (h + 1) * n;
  }
  return s;
}

static std::unique_ptr<Solver> g_last;  // state of the most recent solve, for per-knot parity checks

/* Copy an internal per-knot array of the last solve: vectors as [k][i], matrices column-major per knot. */
extern "C" long cafe_oracle_get(const char* name, int phase, double* out) {
  if (!g_last || phase < 0 || phase >= (int)g_last->phases.size()) return -1;
  Phase& P = *g_last->phases[phase];
  std::string nm(name);
  double* w = out;
  auto putv = [&](const std::vector<Vec>& V) { for (auto& v : V) for (double x : v) *w++ = x; };
  auto putm = [&](const std::vector<Mat>& M, int cnt) { for (int k = 0; k < cnt; ++k) for (double x : M[k].a) *w++ = x; };
  if (nm == "X") putv(P.X); else if (nm == "Xbar") putv(P.Xbar); else if (nm == "U") putv(P.U); else if (nm == "Ubar") putv(P.Ubar);
  else if (nm == "Y") putv(P.Y); else if (nm == "Defect") putv(P.Defect); else if (nm == "dX") putv(P.dX); else if (nm == "dU") putv(P.dU);
  else if (nm == "G") putv(P.G); else if (nm == "Qu") putv(P.Qu);
  else if (nm == "A") putm(P.A, P.h); else if (nm == "B") putm(P.B, P.h); else if (nm == "C") putm(P.C, P.h); else if (nm == "D") putm(P.D, P.h);
  else if (nm == "K") putm(P.K, P.h); else if (nm == "Quu") putm(P.Quu, P.h); else if (nm == "Qux") putm(P.Qux, P.h); else if (nm == "H") putm(P.H, P.h + 1);
  else if (nm == "lx") { for (auto& r : P.rcost) for (double x : r.lx) *w++ = x; }
  else if (nm == "lu") { for (auto& r : P.rcost) for (double x : r.lu) *w++ = x; }
  else if (nm == "ly") { for (auto& r : P.rcost) for (double x : r.ly) *w++ = x; }
  else if (nm == "lxx") { for (auto& r : P.rcost) for (double x : r.lxx.a) *w++ = x; }
  else if (nm == "luu") { for (auto& r : P.rcost) for (double x : r.luu.a) *w++ = x; }
  else if (nm == "lyy") { for (auto& r : P.rcost) for (double x : r.lyy.a) *w++ = x; }
  else if (nm == "l") { for (auto& r : P.rcost) *w++ = r.l; *w++ = P.tcost.Phi; }
  else if (nm == "Phix") { for (double x : P.tcost.Phix) *w++ = x; }
  else if (nm == "Phixx") { for (double x : P.tcost.Phixx.a) *w++ = x; }
  else if (nm == "Px") { Mat Px = P.resetmap_partial(P.X.back()); for (double x : Px.a) *w++ = x; }
  else return -1;
  return (long)(w - out);
}

/* dynamics and their partials of phase `phase` at knot k for a given (x, u): xnext[n], y[p], A[n*n], B[n*m], C[p*n], D[p*m]
 * (column-major); any output may be NULL. For the finite-difference recipe of the reference (test/testKKTDynamics.cpp:39-93). */
extern "C" int cafe_oracle_dynamics(const CafeDeck* deck, int phase, int k, const double* x, const double* u, double* xnext, double* y,
                                    double* A, double* B, double* C, double* D) {
  try {
    Solver S;
    S.setup(deck);
    Phase& P = *S.phases.at(phase);
    Vec xv(x, x + P.n), uv(u, u + P.m), xn(P.n, 0.0), yv(P.p, 0.0);
    P.dynamics(xn, yv, xv, uv, k);
    if (xnext) std::memcpy(xnext, xn.data(), sizeof(double) * P.n);
    if (y && P.p) std::memcpy(y, yv.data(), sizeof(double) * P.p);
    if (A || B || C || D) {
      Mat Am(P.n, P.n), Bm(P.n, P.m), Cm(P.p, P.n), Dm(P.p, P.m);
      P.dynamics_partial(Am, Bm, Cm, Dm, xv, uv, k);
      if (A) std::memcpy(A, Am.a.data(), sizeof(double) * Am.a.size());
      if (B) std::memcpy(B, Bm.a.data(), sizeof(double) * Bm.a.size());
      if (C && P.p) std::memcpy(C, Cm.a.data(), sizeof(double) * Cm.a.size());
      if (D && P.p) std::memcpy(D, Dm.a.data(), sizeof(double) * Dm.a.size());
    }
    return 0;
  } catch (const std::exception& e) { std::fprintf(stderr, "cafe_oracle_dynamics: %s\n", e.what()); return -1; }
}

/* reset map and its Jacobian at the end of phase `phase`: xnext[n_next], Px[n_next*n] column-major */
extern "C" int cafe_oracle_resetmap(const CafeDeck* deck, int phase, const double* x, double* xnext, double* Px) {
  try {
    Solver S;
    S.setup(deck);
    Phase& P = *S.phases.at(phase);
    Vec xv(x, x + P.n);
    Vec xn = P.resetmap(xv);
    if (xnext) std::memcpy(xnext, xn.data(), sizeof(double) * xn.size());
    if (Px) { Mat m = P.resetmap_partial(xv); std::memcpy(Px, m.a.data(), sizeof(double) * m.a.size()); }
    return (int)xn.size();
  } catch (const std::exception& e) { std::fprintf(stderr, "cafe_oracle_resetmap: %s\n", e.what()); return -1; }
}

/* guess (optional): initial Xbar / Ubar / K in the packed solution layout (cafe_solution_size doubles; the other arrays of the record
 * are ignored). This is the state MHPCProblem::update leaves behind for the re-solve: MultiPhaseDDP::solve starts with
 * hybrid_rollout(eps = 0), i.e. U = Ubar + K (X - Xbar) around whatever the trajectories hold (MultiPhaseDDP.cpp:238). */
static void apply_guess(Solver& S, const double* guess) {
  const double* r = guess;
  auto getv = [&](std::vector<Vec>& V, int cnt, bool use) { for (int k = 0; k < cnt; ++k) for (double& v : V[k]) { if (use) v = *r; ++r; } };
  auto getm = [&](std::vector<Mat>& M, int cnt, bool use) { for (int k = 0; k < cnt; ++k) for (double& v : M[k].a) { if (use) v = *r; ++r; } };
  for (auto& P : S.phases) {
    getv(P->Xbar, P->h + 1, true); getv(P->Ubar, P->h, true); getv(P->Y, P->h, false); getv(P->dU, P->h, false); getm(P->K, P->h, true);
    getv(P->Qu, P->h, false); getm(P->Quu, P->h, false); getm(P->Qux, P->h, false); getv(P->G, P->h + 1, false);
    P->X = P->Xbar; P->U = P->Ubar;
  }
}

/* al_in / al_out: [n_phases][4][2] = (sigma, lambda) of every touchdown-constraint element. The reference's MPC loop never resets them:
 * TerminalConstraintBase::reset_params is an empty function (ConstraintsBase.h:367-374), so what update_params left behind in one solve is
 * what the next solve after HKDProblem::update / MHPCProblem::update starts from. al_in = NULL: the deck's initial values. */
/* Relaxed-barrier parameters across the solves of an MPC loop. reb_in / reb_out: for every phase, for every knot k < h, for every element of
 * every path constraint of the phase (in their order): (delta, eps) - cafe_oracle_reb_ne gives the elements per knot of every phase. The reference
 * keeps them with the knots: PathConstraintBase::pop_front / push_back (ConstraintsBase.h:296-306; an appended knot copies the last knot's
 * values), reset_params is empty (:191-193). */
extern "C" int cafe_oracle_reb_ne(const CafeDeck* deck, int* ne) {
  try {
    Solver S;
    S.setup(deck);
    for (size_t i = 0; i < S.phases.size(); ++i) { int n = 0; for (auto& pc : S.phases[i]->pcon) n += pc.size; ne[i] = n; }
    return 0;
  } catch (const std::exception& e) { std::fprintf(stderr, "cafe_oracle_reb_ne: %s\n", e.what()); return -1; }
}
static void reb_io(Solver& S, const double* in, double* out) {
  size_t o = 0;
  for (auto& P : S.phases)
    for (int k = 0; k < P->h; ++k)
      for (auto& pc : P->pcon)
        for (int e = 0; e < pc.size; ++e, o += 2) {
          if (in) { pc.params[k][e].delta = in[o]; pc.params[k][e].eps = in[o + 1]; }
          if (out) { out[o] = pc.params[k][e].delta; out[o + 1] = pc.params[k][e].eps; }
        }
}
/* the values a fresh deck starts from */
extern "C" int cafe_oracle_reb_init(const CafeDeck* deck, double* out) {
  try { Solver S; S.setup(deck); reb_io(S, nullptr, out); return 0; }
  catch (const std::exception& e) { std::fprintf(stderr, "cafe_oracle_reb_init: %s\n", e.what()); return -1; }
}

extern "C" int cafe_oracle_solve_carry(const CafeDeck* deck, const CafeOptions* opt, const double* x0, const double* guess, const double* al_in,
                                       double* al_out, const double* reb_in, double* reb_out, CafeInfo* info, double* hist, int hist_cap,
                                       double* trace, int trace_cap, double* sol) {
  try {
    g_last.reset(new Solver());
    Solver& S = *g_last;
    S.setup(deck);
    if (guess) apply_guess(S, guess);
    if (al_in)
      for (size_t i = 0; i < S.phases.size(); ++i)
        for (auto& tc : S.phases[i]->tcon)
          for (int e = 0; e < tc.size && e < 4; ++e) { tc.params[e].sigma = al_in[(i * 4 + e) * 2]; tc.params[e].lambda = al_in[(i * 4 + e) * 2 + 1]; }
    if (reb_in) reb_io(S, reb_in, nullptr);
    S.x0.assign(x0, x0 + S.phases[0]->n);
    S.solve(*opt);
    if (reb_out) reb_io(S, nullptr, reb_out);
    if (al_out)
      for (size_t i = 0; i < S.phases.size(); ++i) {
        for (int e = 0; e < 8; ++e) al_out[i * 8 + e] = 0;
        for (auto& tc : S.phases[i]->tcon)
          for (int e = 0; e < tc.size && e < 4; ++e) { al_out[(i * 4 + e) * 2] = tc.params[e].sigma; al_out[(i * 4 + e) * 2 + 1] = tc.params[e].lambda; }
      }
    if (info) {
      info->status = S.reg_failed ? CAFE_STATUS_REG_FAIL : CAFE_STATUS_OK;
      info->iter = S.iter_; info->ls_iter_total = S.ls_iter_total_; info->reg_iter_total = S.reg_iter_total_;
      info->outer_iter = S.iter_ou; info->n_hist = (int)(S.hist.size() / 4);
      info->cost = S.actual_cost; info->feas = S.feas; info->max_tconstr = S.max_tconstr; info->max_pconstr = S.max_pconstr;
    }
    if (hist) { size_t nh = std::min<size_t>(S.hist.size(), (size_t)hist_cap * 4); std::memcpy(hist, S.hist.data(), nh * sizeof(double)); }
    if (trace) { size_t nt = std::min<size_t>(S.trace.size(), (size_t)trace_cap * CAFE_TRACE_W); std::memcpy(trace, S.trace.data(), nt * sizeof(double)); }
    if (sol) {
      double* w = sol;
      auto putv = [&](const std::vector<Vec>& V, int cnt) { for (int k = 0; k < cnt; ++k) for (double v : V[k]) *w++ = v; };
      auto putm = [&](const std::vector<Mat>& M, int cnt) { for (int k = 0; k < cnt; ++k) for (double v : M[k].a) *w++ = v; };
      for (auto& P : S.phases) {
        putv(P->Xbar, P->h + 1); putv(P->Ubar, P->h); putv(P->Y, P->h); putv(P->dU, P->h); putm(P->K, P->h);
        putv(P->Qu, P->h); putm(P->Quu, P->h); putm(P->Qux, P->h); putv(P->G, P->h + 1);
      }
    }
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "cafe_oracle_solve: %s\n", e.what());
    return -1;
  }
}

extern "C" int cafe_oracle_solve_al(const CafeDeck* deck, const CafeOptions* opt, const double* x0, const double* guess, const double* al_in,
                                    double* al_out, CafeInfo* info, double* hist, int hist_cap, double* trace, int trace_cap, double* sol) {
  return cafe_oracle_solve_carry(deck, opt, x0, guess, al_in, al_out, nullptr, nullptr, info, hist, hist_cap, trace, trace_cap, sol);
}

extern "C" int cafe_oracle_solve_warm(const CafeDeck* deck, const CafeOptions* opt, const double* x0, const double* guess, CafeInfo* info,
                                      double* hist, int hist_cap, double* trace, int trace_cap, double* sol) {
  return cafe_oracle_solve_al(deck, opt, x0, guess, nullptr, nullptr, info, hist, hist_cap, trace, trace_cap, sol);
}

extern "C" int cafe_oracle_solve(const CafeDeck* deck, const CafeOptions* opt, const double* x0, CafeInfo* info,
                                 double* hist, int hist_cap, double* trace, int trace_cap, double* sol) {
  return cafe_oracle_solve_warm(deck, opt, x0, nullptr, info, hist, hist_cap, trace, trace_cap, sol);
}
