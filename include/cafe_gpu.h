/*
 * cafe_gpu.h — C ABI of the B200 batched HS-DDP path (libcafe_gpu.so).
 *
 * Drop-in boundary: these entry points are what a binding of the reference's
 * solver API would call. Each one names the reference interface it stands behind:
 *
 *   cafe_options_load        loadHSDDPSetting             HSDDPSolver/common/HSDDP_CompoundTypes.h:57-82
 *   cafe_deck_build_hkd      HKDProblem<T>::initialization HKDMPC/HKD-TrajOpt/HKDProblem.cpp:15-111
 *                            (+ QuadReference::load_top_level_data / initialize,
 *                               Reference/QuadReference.cpp:6-31,134-356)
 *   cafe_deck_build_mhpc     MHPCProblem<T>::initialization MHPC/MHPC-Trajopt/MHPCProblem.cpp:13-250
 *   cafe_deck_build_loco     LocoProblem<T> (initialize_parameters, create_problem_one_phase)
 *                                                         MHPC/MHPC-Trajopt/Locomotion/LocoProblem.cpp:7-84
 *   cafe_deck_build_barrel_to / cafe_barrel_to_initial_guess   main() of MHPC/MHPC-Trajopt/BarrelRoll/BarrelRollTO.cpp:65-275
 *   cafe_hkd_state           compute_hkd_state            HKDMPC/HKD-TrajOpt/HKDModel.h:66-96
 *   cafe_gpu_create          MultiPhaseDDP<T>::set_multiPhaseProblem  HSDDPSolver/header/MultiPhaseDDP.h:33-42
 *   cafe_gpu_solve_batch     MultiPhaseDDP<T>::set_initial_condition + solve
 *                                                          HSDDPSolver/header/MultiPhaseDDP.h:44-46,
 *                                                          HSDDPSolver/source/MultiPhaseDDP.cpp:216-447
 *   cafe_gpu_get_info        MultiPhaseDDP<T>::get_solver_info (both) HSDDPSolver/header/MultiPhaseDDP.h:84-93,
 *                                                          MultiPhaseDDP.cpp:554-563
 *   cafe_gpu_get_solution    reads of Trajectory::Xbar/Ubar/Y/K/dU/Qu/Quu/Qux/G after solve
 *                                                          HSDDPSolver/header/TrajectoryManagement.h:54-85,
 *                                                          MHPC/MHPCLocomotion.cpp:236-281
 *
 * Plain pointers and sizes only; every function returns 0 on success and a negative code on
 * error (cafe_last_error() holds the text); nothing throws across the boundary. A handle is not
 * thread-safe: one handle per host thread per GPU. There is NO CPU fallback: without a CUDA device
 * cafe_gpu_create fails with CAFE_ERR_CUDA.
 */
#ifndef CAFE_GPU_H
#define CAFE_GPU_H
#include "cafe_deck.h"
#ifdef __cplusplus
extern "C" {
#endif

#define CAFE_ERR_ARG -1
#define CAFE_ERR_CUDA -2
#define CAFE_ERR_IO -3
#define CAFE_ERR_UNSUPPORTED -4

typedef struct CafeDeckHandle CafeDeckHandle;
typedef struct CafeHandle CafeHandle;

const char* cafe_last_error(void);

/* ---- host-side problem setup ---- */
int cafe_options_load(const char* ddp_setting_info, CafeOptions* out);
/* one key ("section.key") of a Boost-INFO settings file, for host code that mirrors the reference's load* helpers
 * (loadMHPCConfig, MHPCProblem.h:67-83) without Boost */
int cafe_info_get_number(const char* info_file, const char* key, double* out);
int cafe_info_get_string(const char* info_file, const char* key, char* out, int cap);
/* reorder legs = true (HKD convention), k0 = number of leading reference samples to drop */
int cafe_deck_build_hkd(const char* reference_csv, const char* constraint_params_info, float plan_duration,
                        float time_step, int nsteps_between_mpc, int k0, CafeDeckHandle** out);
/* mhpc_config_info is MHPC/settings/mhpc_config.info; its costFile / constraintParamFile entries are
 * resolved relative to settings_root (the reference resolves them relative to "../"). reference_csv
 * overrides the file's referenceFile entry. */
int cafe_deck_build_mhpc(const char* reference_csv, const char* mhpc_config_info, const char* settings_root,
                         int k0, CafeDeckHandle** out);
/* LocoProblem<T> (MHPC/MHPC-Trajopt/Locomotion/LocoProblem.cpp:7-84, driver Loco_TO.cpp:16-82): whole-body-only locomotion
 * trajectory optimisation. Same arguments as cafe_deck_build_mhpc with Locomotion/settings/loco_config.info; the whole-body
 * phases carry the torque-limit and GRF barriers only (no joint-limit / min-height barrier), touchdown constraints as in MHPC. */
int cafe_deck_build_loco(const char* reference_csv, const char* loco_config_info, const char* settings_root,
                         int k0, CafeDeckHandle** out);
/* The same two builders from an MHPCConfig held in memory (the struct loadMHPCConfig fills, MHPCProblem.h:43-83; field types as there),
 * i.e. MHPCProblem<T>::set_problem_data(pdata, pconfig) + prepare_initialization + initialize_parameters + initialize_multiPhaseProblem
 * (MHPCProblem.h:198-212, MHPCProblem.cpp:13-250). loco != 0: LocoProblem's constraint set. Used by include/hsddp_facade/MHPCProblem.h. */
typedef struct {
  double plan_dur_wb, plan_dur_srb, dt_wb, dt_srb;
  float dt_mpc, BG_alpha;
  const char* costFileName;            /* relative to settings_root, as in mhpc_config.info */
  const char* constraintParamFileName;
} CafeMHPCConfig;
int cafe_deck_build_mhpc_config(const char* reference_csv, const CafeMHPCConfig* config, const char* settings_root, int k0, int loco,
                                CafeDeckHandle** out);
/* start / end time of every phase of the deck in seconds from the start of the plan (MHPCProblemData::wb_phase_start_times /
 * wb_phase_end_times, srb_start_time / srb_end_time; HKDProblemData::phase_start_times / phase_end_times): [n_phases] each */
int cafe_deck_phase_times(const CafeDeckHandle* h, float* start_times, float* end_times);
/* In-place barrel roll (MHPC/MHPC-Trajopt/BarrelRoll/BarrelRollTO.cpp:65-275): six hand-scheduled whole-body phases
 * (stance, right pair, flight, stance, flight, stance; switching times :70), per-phase weight sets from br_cost_weights.JSON, fixed
 * desired states (:277-339), the BarrelRoll:: barriers incl. the joint-speed limit (br_constraint_params.info), four-foot touchdown
 * constraints at the end of both flight phases. No reference file. */
int cafe_deck_build_barrel_to(const char* cost_weights_json, const char* constraint_params_info, CafeDeckHandle** out);
/* The state trajectory BarrelRollTO.cpp:131-147 starts from (linear interpolation between the desired states; phase 0 from x0) as
 * packed guesses for cafe_gpu_set_initial_guess: x0 = host [B][36], guess = host [B][cafe_solution_size(deck)] (controls, gains zero). */
int cafe_barrel_to_initial_guess(const CafeDeck* deck, const double* x0, int B, double* guess);
/* Marks a deck built at a later start offset as the product of MHPCProblem::update (MHPCProblem.cpp:252-372) rather than of
 * initialization(): a tail whole-body phase not longer than the shift (nsteps = round(dt_mpc / dt_wb)) was opened by that update and has
 * no shooting states yet (:366-369) -> CafePhase::single_shooting. *marked_phase = index of that phase, -1 if there is none. */
int cafe_deck_mark_mpc_update(CafeDeckHandle* h, int nsteps, int* marked_phase);
const CafeDeck* cafe_deck_get(const CafeDeckHandle* h);
void cafe_deck_free(CafeDeckHandle* h);
/* Structural non-zero pattern the backward sweep assumes for one LQ array of a running knot, as a bit mask (bit i + rows * j):
 * which = 0 A, 1 B, 2 lxx, 3 luu. HKD phases: all four (576 bits, 9 words; A / B are the CCS patterns of the generated
 * hkinodyn_par, HKDModel.h:33-61); whole-body phases: lxx only (1296 bits, 21 words; MHPCCost.cpp's cost objects + the ReB terms).
 * Returns the number of 64-bit words written (out holds at least 21), < 0 on error. Entries outside the pattern are never read
 * by the sweep: tests check the oracle's arrays against these masks. */
int cafe_deck_lq_pattern(const CafeDeck* deck, int phase, int knot, int which, unsigned long long* out);
/* x0 = [body(12) = eul,pos,omega,vel ; qdummy(12)] from joint angles */
int cafe_hkd_state(const double body[12], const double qJ[12], const int contact[4], double x0[24]);

/* packed solution of ONE problem, in doubles; per phase, in order:
 *   Xbar[(h+1) n]  Ubar[h m]  Y[h p]  dU[h m]  K[h m n]  Qu[h m]  Quu[h m m]  Qux[h m n]  G[(h+1) n]
 * vectors knot-major, matrices column-major per knot (the reference's Eigen layout). */
long cafe_solution_size(const CafeDeck* deck);
/* compact "command" record, what the MPC loop consumes (MHPCLocomotion.cpp:236-281):
 * per phase Xbar, Ubar, Y for all knots, then K, Qu, Quu, Qux for the first n_gain_knots knots
 * of the whole horizon. */
long cafe_command_size(const CafeDeck* deck, int n_gain_knots);

/* ---- GPU solver ---- */
int cafe_gpu_create(const CafeDeck* deck, int device, int max_batch, CafeHandle** out);
int cafe_gpu_destroy(CafeHandle* h);
/* Optional per-problem references (different velocity commands / targets per problem) on the deck's shared phase schedule:
 * refs = host [B][deck->n_records][CAFE_REF_W], same record layout as CafeDeck.ref (what WBReference / SRBReference /
 * HKDSinglePhaseReference hand to the costs, MHPCReference.cpp:10-76, HKDReference.cpp:8-62); contact flags must equal the deck's.
 * Also replaces the cold-start guess Xbar = reference. refs = NULL returns to the shared records. */
int cafe_gpu_set_references(CafeHandle* h, const double* refs, int B);
/* Warm start (receding-horizon re-solves): initial Xbar / Ubar / K per problem in the packed solution layout, host
 * [B][cafe_solution_size] (the other arrays of the record are ignored). It stands for the trajectories MHPCProblem::update leaves
 * behind (MHPCProblem.cpp:252-397): MultiPhaseDDP::solve begins with hybrid_rollout(eps = 0), U = Ubar + K (X - Xbar), around
 * them (MultiPhaseDDP.cpp:238). Stays in force for the following solves; guess = NULL returns to the cold start. */
int cafe_gpu_set_initial_guess(CafeHandle* h, const double* guess, int B);
/* Augmented-Lagrangian parameters across the solves of an MPC loop. The reference keeps a phase's TouchDownConstraint object - and the
 * sigma / lambda its update_params left behind (ConstraintsBase.h:375-392) - for as long as the phase lives: reset_params(), which
 * HKDProblem::update / MHPCProblem::update call for every phase (HKDProblem.cpp:208, MHPCProblem.cpp:363), is an empty function
 * (ConstraintsBase.h:367-374). al / out = host [B][n_phases][4][2] = (sigma, lambda) per touchdown-constraint element, unused entries zero.
 * set: the following solves start from these values instead of the deck's TD_AL values (al = NULL returns to the deck's);
 * get: the values the last solve left behind. cafe_gpu_update_deck and cafe_gpu_shift_guess carry them over on the device by themselves
 * (a phase that continues an old phase with the same touchdown feet inherits, every other constraint starts from the deck's values), and with
 * them the relaxed-barrier parameters when the last solve could change them (update_relax / update_ReB != 1): the per-(knot, element) update
 * counts travel with the knots like PathConstraintBase::pop_front / push_back moves them (ConstraintsBase.h:296-306: a knot appended at the tail
 * copies the last knot's values). */
int cafe_gpu_set_al_params(CafeHandle* h, const double* al, int B);
int cafe_gpu_get_al_params(CafeHandle* h, double* out);
/* The same warm start without leaving the device: the guess of `dst` (deck at start offset dst_k0 of the reference file) is built from
 * the solution held by `src` (deck at start offset src_k0 <= dst_k0, same device, solved with at least B problems) by the shift
 * MHPCProblem::update applies to the trajectories (MHPCProblem.cpp:252-397; Trajectory::pop_front / push_back_state,
 * TrajectoryManagement.cpp:130-228): popped knots vanish, a phase continues the old phase with the same stance, knots past the old
 * plan repeat its last state with zero control and gain, a new phase starts from the reference, the reduced-order tail is kept. */
int cafe_gpu_shift_guess(CafeHandle* dst, CafeHandle* src, int src_k0, int dst_k0, int B);
/* Planned state `knots_ahead` knots after the start of the plan, out [B][n] (host): the state an MPC loop hands to the next solve
 * (MHPCLocomotion::update takes it from the simulator; a closed-loop Monte-Carlo without one uses the plan's own prediction). */
int cafe_gpu_get_planned_state(CafeHandle* h, int knots_ahead, double* out);
/* The MPC update on ONE solver (MHPCLocomotion::update_mpc_if_needed -> MHPCProblem::update, MHPC/MHPCLocomotion.cpp:91-150,
 * MHPC/MHPC-Trajopt/MHPCProblem.cpp:252-397; HKDMPCSolver::update -> HKDProblem::update, HKDMPC/HKD-TrajOpt/HKDProblem.cpp:117-222):
 * `new_deck` (the problem re-cut k_advance knots later, cafe_deck_build_* + cafe_deck_mark_mpc_update) replaces the solver's deck, the
 * previous solution of the first B problems becomes the warm start by the rules of cafe_gpu_shift_guess, and the device arena,
 * reference and mask buffers of the handle are re-used (nothing is allocated in steady state). B = 0: new deck, cold start.
 * Per-problem references do not survive the update (set them again). */
int cafe_gpu_update_deck(CafeHandle* h, const CafeDeck* new_deck, int k_advance, int B);
/* x0: host [B][n0] row per problem. Runs every problem of the batch to its own termination. */
int cafe_gpu_solve_batch(CafeHandle* h, const double* x0, int B, const CafeOptions* opt);
/* same with x0 already resident on the device, layout [n0][ldb] (component-major), ldb >= B */
int cafe_gpu_solve_batch_device(CafeHandle* h, const double* x0_dev, int ldb, int B, const CafeOptions* opt);
int cafe_gpu_get_info(CafeHandle* h, CafeInfo* info /*[B]*/);
/* hist: [B][hist_cap][4] = cost, feas, max_tconstr, max_pconstr per pushed entry */
int cafe_gpu_get_history(CafeHandle* h, double* hist, int hist_cap);
/* trace: [B][trace_cap][12], same record as the oracle's CAFE_TRACE_W */
int cafe_gpu_get_trace(CafeHandle* h, double* trace, int trace_cap);
int cafe_gpu_get_solution(CafeHandle* h, int b0, int nb, double* sol /*[nb][cafe_solution_size]*/);
int cafe_gpu_get_commands(CafeHandle* h, int n_gain_knots, double* cmd /*[B][cafe_command_size]*/);
/* same, packed into a caller-owned DEVICE buffer (source of the final NCCL gather in multi-GPU runs) */
int cafe_gpu_get_commands_device(CafeHandle* h, int n_gain_knots, double* cmd_dev);
/* Collection overlapped with the next solve (a two-deep pipeline: solve i + 1 runs while the records of solve i travel). The records are packed on
 * the solver's stream - a following cafe_gpu_solve_batch is ordered behind the pack and may be issued at once - and copied to `cmd` (page-locked host
 * memory, [B][cafe_command_size]) by a copy stream. slot = 0 / 1; cafe_gpu_commands_wait(h, slot) blocks until that slot's records have landed.
 * Same bytes as cafe_gpu_get_commands. */
int cafe_gpu_get_commands_async(CafeHandle* h, int n_gain_knots, double* cmd_pinned, int slot);
int cafe_gpu_commands_wait(CafeHandle* h, int slot);
/* Wire-format step after the path: the per-problem part of MHPC_Command_lcmt (lcmtypes/MHPC_Command_lcmt.lcm), which
 * MHPCLocomotion::publish_mpc_cmd fills on the host from Xbar/Ubar/Y/K/Qu/Quu/Qux with cast<float>() (MHPCLocomotion.cpp:236-281).
 * Emitted as float32 directly from the device arrays for the first n_steps whole-body knots; per problem, in the struct's order:
 *   torque[N][12] eul[N][3] pos[N][3] qJ[N][12] vWorld[N][3] eulrate[N][3] qJd[N][12] GRF[N][12] feedback[N][432] Qu[N][12]
 *   Quu[N][144] Qux[N][432]          (matrices column-major = Eigen .data() order; N = n_steps; 1080 N floats)
 * The deck-level fields (mpc_times, contacts, statusTimes) are the same for every problem and stay with the caller. */
long cafe_lcm_command_size(int n_steps);
int cafe_gpu_get_lcm_commands(CafeHandle* h, int n_steps, float* out /*[B][cafe_lcm_command_size]*/);
int cafe_gpu_get_lcm_commands_device(CafeHandle* h, int n_steps, float* out_dev);
/* The same for the HKD application: the per-problem part of hkd_command_lcmt (lcmtypes/hkd_command_lcmt.lcm) as
 * HKDMPCSolver::publish_mpc_cmd fills it (HKDMPC/HKDMPC.cpp:243-290; N = nsteps_between_mpc + 7 there), float32, per problem:
 *   hkd_controls[N][24] (Ubar)  des_body_state[N][12] (first 12 states of Xbar)  feedback[N][12][12] (K(m, n), m major)   = 180 N floats
 * Steps run on across phase boundaries (s >= horizon -> next phase). mpc_times, contacts, statusTimes, foot_placement stay with the caller. */
long cafe_hkd_lcm_command_size(int n_steps);
int cafe_gpu_get_hkd_lcm_commands(CafeHandle* h, int n_steps, float* out /*[B][cafe_hkd_lcm_command_size]*/);
int cafe_gpu_get_hkd_lcm_commands_device(CafeHandle* h, int n_steps, float* out_dev);
/* ---- multi-GPU: the batch is cut into contiguous slices, GPU / rank g solves rows
 * [g ceil(B/G), min(B, (g+1) ceil(B/G))) with the deck replicated and no traffic during the solve; the one collective is the final gather of
 * the packed command records to the first GPU / rank 0 (NCCL send / recv in one group; NCCL is opened with dlopen at first use).
 * Stands behind "one MultiPhaseDDP per problem" of the reference (HSDDPSolver/header/MultiPhaseDDP.h:31-93): the batch is ours. */
typedef struct CafeMulti CafeMulti;
int cafe_gpu_shard_range(int B, int nranks, int rank, int* lo, int* hi);
/* (a) one process, ndev GPUs of one box: ncclCommInitAll; devices = NULL means 0..ndev-1; max_batch is the GLOBAL batch */
int cafe_gpu_create_multi(const CafeDeck* deck, int ndev, const int* devices, int max_batch, CafeMulti** out);
int cafe_gpu_multi_destroy(CafeMulti* m);
int cafe_gpu_multi_ndev(const CafeMulti* m);
CafeHandle* cafe_gpu_multi_handle(CafeMulti* m, int g); /* the per-GPU solver (histories, traces, debug reads) */
int cafe_gpu_multi_solve_batch(CafeMulti* m, const double* x0 /*host [B][n0]*/, int B, const CafeOptions* opt);
int cafe_gpu_multi_update_deck(CafeMulti* m, const CafeDeck* new_deck, int k_advance, int B); /* cafe_gpu_update_deck on every GPU's slice */
int cafe_gpu_multi_get_info(CafeMulti* m, CafeInfo* info /*[B]*/);
int cafe_gpu_multi_get_commands(CafeMulti* m, int n_gain_knots, double* cmd /*host [B][cafe_command_size]*/);
/* (b) one process per GPU (torchrun, MPI): rank 0 makes the id, the launcher hands it to every rank */
int cafe_gpu_nccl_unique_id(char id[128]);
int cafe_gpu_comm_init_rank(CafeHandle* h, int nranks, int rank, const char id[128]);
int cafe_gpu_comm_destroy(CafeHandle* h);
/* packs this rank's records and gathers all ranks' slices on rank 0: rank r's records land at out_dev + r * per_rank records (device
 * buffer of rank 0 with nranks * per_rank records; per_rank >= the local batch, equal on all ranks; out_dev is ignored elsewhere) */
int cafe_gpu_gather_commands(CafeHandle* h, int n_gain_knots, int per_rank, double* out_dev);
/* ... and without blocking the solver: pack on the solver's stream, NCCL send / recv and (rank 0) the copy of all gathered records to out_host
 * (page-locked, [nranks * per_rank][cafe_command_size]) on the copy stream; out_dev = rank 0's device buffer of this slot (NULL on the other ranks);
 * cafe_gpu_commands_wait(h, slot) waits. */
int cafe_gpu_gather_commands_async(CafeHandle* h, int n_gain_knots, int per_rank, double* out_dev, double* out_host_pinned, int slot);

/* device-time breakdown of the last solve, ms per kernel family, and launch counts */
/* timing slots of cafe_gpu_get_timing: one per kernel family */
#define CAFE_K_ROLL 0      /* k_roll: trial states / controls of every knot; SRB, HKD and terminal knots completely */
#define CAFE_K_SELECT 1    /* k_ls_scan, k_select, k_compact */
#define CAFE_K_ACCEPT 2    /* k_accept */
#define CAFE_K_LQ 3        /* k_lq: SRB / HKD knots, terminal knots (incl. impact-map Jacobians) */
#define CAFE_K_BWD 4       /* k_bwd2: backward sweep + linear rollout */
#define CAFE_K_MISC 5      /* initialisation, warm-start unpack */
#define CAFE_K_WB_TERMS 6  /* k_wb_terms: leg-parallel rigid-body terms of the whole-body trial knots */
#define CAFE_K_WB_FWD 7    /* k_wb_fwd: cooperative KKT contact dynamics, x+, GRF, cost, defects */
#define CAFE_K_WB_DERIVS 8 /* k_wb_derivs: leg-parallel RNEA derivatives and foot kinematic partials */
#define CAFE_K_WB_SENS 9   /* k_wb_sens: cooperative KKT sensitivities (A, B, C, D tiles) */
#define CAFE_K_WB_COST 10  /* k_wb_cost: cooperative cost / barrier partials and running cost */
#define CAFE_NKERNELS 11
int cafe_gpu_get_timing(CafeHandle* h, double ms[CAFE_NKERNELS], long launches[CAFE_NKERNELS], int* ticks);
/* work items launched per slot during the last solve: (problem, knot[, step size]) triples of the knot kernels, problems of the sweep */
int cafe_gpu_get_units(CafeHandle* h, double units[CAFE_NKERNELS]);
/* device time (CUDA events on the solver's stream) of the last cafe_gpu_solve_batch*, in ms */
int cafe_gpu_get_solve_ms(CafeHandle* h, double* ms);
/* enable per-kernel CUDA-event timing (costs a few us per launch) */
int cafe_gpu_set_profiling(CafeHandle* h, int on);

/* raw read-back of an internal per-knot array of problem b (same names/layout as the oracle's cafe_oracle_get:
 * vectors [k][i], matrices column-major per knot); for parity tests. Returns doubles written or a negative error. */
long cafe_gpu_debug_get(CafeHandle* h, const char* name, int phase, int b, double* out);

/* fp64 FMA peak microbenchmark on the current device (TFLOP/s), for the roofline denominator */
int cafe_gpu_measure_fp64_peak(int device, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* CAFE_GPU_H */
