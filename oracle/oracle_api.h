/*
 * oracle_api.h — C entry points of the CPU ORACLE.
 *
 * TEST INFRASTRUCTURE ONLY. The oracle is a plain-C++ restatement of the reference's
 * HS-DDP solver and models; it exists to CHECK the CUDA path (tests/, smoke(),
 * bench.py's cpu_baseline / --impl reference leg). Nothing in the product library
 * links, loads or calls it.
 */
#ifndef CAFE_ORACLE_API_H
#define CAFE_ORACLE_API_H
#include "../include/cafe_deck.h"
#ifdef __cplusplus
extern "C" {
#endif

/* number of doubles of one packed solution (see cafe_solution_layout in cafe_gpu.h):
 * per phase: Xbar[(h+1)n] Ubar[h m] Y[h p] dU[h m] K[h m n] Qu[h m] Quu[h m m] Qux[h m n] G[(h+1) n] */
long cafe_oracle_solution_size(const CafeDeck* deck);
/* test hook: the Eigen-3.3 pivoted LDLT restatement alone (isPositive, solve(Identity), smallest pivot) */
int cafe_oracle_ldlt(const double* A, int n, double* inv, double* min_pivot);

/* per-iteration trace record (doubles), one per executed DDP iteration */
#define CAFE_TRACE_W 12
/* 0 cost_at_start 1 feas_at_start 2 dV_1 3 dV_2 4 merit_rho 5 reg_after 6 reg_iters 7 ls_iters
 * 8 ls_success 9 eps_accepted(0 if none) 10 cost_after 11 feas_after */

/* Solve one problem. hist: [hist_cap][4] = cost, feas, max_tconstr, max_pconstr.
 * trace: [trace_cap][CAFE_TRACE_W]. sol: packed solution or NULL. Returns 0. */
int cafe_oracle_solve(const CafeDeck* deck, const CafeOptions* opt, const double* x0,
                      CafeInfo* info, double* hist, int hist_cap,
                      double* trace, int trace_cap, double* sol);
/* Same, started from an initial guess: Xbar / Ubar / K taken from `guess` (packed solution layout, other arrays ignored; NULL =
 * the cold start above). Restates the re-solve after MHPCProblem::update: the solver begins with hybrid_rollout(eps = 0) around
 * whatever the trajectories hold (MultiPhaseDDP.cpp:238). */
int cafe_oracle_solve_warm(const CafeDeck* deck, const CafeOptions* opt, const double* x0, const double* guess,
                           CafeInfo* info, double* hist, int hist_cap, double* trace, int trace_cap, double* sol);
/* Same, with the augmented-Lagrangian parameters carried between the solves of an MPC loop: al_in / al_out [n_phases][4][2] = (sigma,
 * lambda) per touchdown-constraint element (NULL = the deck's initial values / not wanted). The reference never resets them between MPC
 * steps: TerminalConstraintBase::reset_params is empty (ConstraintsBase.h:367-374), called from HKDProblem.cpp:208 / MHPCProblem.cpp:363. */
int cafe_oracle_solve_al(const CafeDeck* deck, const CafeOptions* opt, const double* x0, const double* guess, const double* al_in,
                         double* al_out, CafeInfo* info, double* hist, int hist_cap, double* trace, int trace_cap, double* sol);
/* ... and the relaxed-barrier parameters, which the reference keeps with the knots of a phase (PathConstraintBase::pop_front / push_back,
 * ConstraintsBase.h:296-306: an appended knot copies the LAST knot's values; reset_params empty, :191-193). reb_in / reb_out: per phase, per knot
 * k < h, per element of the phase's path constraints in their order: (delta, eps); cafe_oracle_reb_ne: elements per knot of every phase;
 * cafe_oracle_reb_init: the values a fresh deck starts from. Only matters when update_relax / update_ReB differ from the shipped 1 / 1. */
int cafe_oracle_reb_ne(const CafeDeck* deck, int* ne);
int cafe_oracle_reb_init(const CafeDeck* deck, double* out);
int cafe_oracle_solve_carry(const CafeDeck* deck, const CafeOptions* opt, const double* x0, const double* guess, const double* al_in, double* al_out,
                            const double* reb_in, double* reb_out, CafeInfo* info, double* hist, int hist_cap, double* trace, int trace_cap, double* sol);

/* Internal per-knot array of the most recent cafe_oracle_solve (names: X Xbar U Ubar Y Defect dX dU G Qu A B C D K
 * Quu Qux H lx lu ly lxx luu lyy l Phix Phixx Px). Returns the number of doubles written or -1. */
long cafe_oracle_get(const char* name, int phase, double* out);

/* Dynamics (and partials) of one phase at knot k for a given (x, u); reset map (and Jacobian) at a phase end. */
int cafe_oracle_dynamics(const CafeDeck* deck, int phase, int k, const double* x, const double* u, double* xnext, double* y,
                         double* A, double* B, double* C, double* D);
int cafe_oracle_resetmap(const CafeDeck* deck, int phase, const double* x, double* xnext, double* Px);

/* Whole-body continuous-time KKT contact dynamics (WBM::dynamics_continuousTime), for the known-answer test
 * of the reference (test/testKKTDynamics.cpp:95-121). */
int cafe_oracle_wb_dynamics(double hip_yaw, double BG_alpha, const double* q, const double* v, const double* u,
                            const int* contact, double* qdd, double* grf);

/* Reference CasADi functions (from oracle/_ref) behind a flat signature, for the
 * "re-emitted device functions == reference generated code" tests. Returns 0, or
 * -1 if the name is unknown. in/out are dense column-major buffers. */
int cafe_oracle_casadi_eval(const char* name, const double* const* in, double* const* out);

#ifdef __cplusplus
}
#endif
#endif
