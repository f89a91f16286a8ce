/*
 * cafe_deck.h — plain-data "phase deck": everything the batched HS-DDP path needs
 * to know about ONE multi-phase problem (all problems of a batch share a deck and
 * differ only in their initial state x0).
 *
 * The deck is what the reference's problem builders produce implicitly as a deque
 * of SinglePhase objects wired with std::function callbacks:
 *   HKDProblem<T>::initialization        /root/reference/HKDMPC/HKD-TrajOpt/HKDProblem.cpp:15-111
 *   MHPCProblem<T>::initialization       /root/reference/MHPC/MHPC-Trajopt/MHPCProblem.cpp:13-250
 * Here the callbacks are replaced by a model id + plain numbers, so that the same
 * bytes can be handed to the CUDA path and to the CPU oracle.
 *
 * All reference look-ups that the reference performs with `float` time arithmetic
 * (QuadReference::get_a_reference_ptr_at_t, Reference/QuadReference.cpp:70-85) are
 * resolved on the host when the deck is built; the per-knot records below hold the
 * looked-up (float-rounded, then widened) values.
 */
#ifndef CAFE_DECK_H
#define CAFE_DECK_H

#ifdef __cplusplus
extern "C" {
#endif

#define CAFE_MODEL_HKD 0 /* hybrid kinodynamic, (n,m,p) = (24,24,0)   HKDModel.h:12-14 */
#define CAFE_MODEL_WB  1 /* whole body,         (n,m,p) = (36,12,12)  WBM.h:13-15      */
#define CAFE_MODEL_SRB 2 /* single rigid body,  (n,m,p) = (12,12,0)   SRBM.h:13-15     */

#define CAFE_MAX_PHASES 16
#define CAFE_MAX_N 36
#define CAFE_MAX_M 24
#define CAFE_MAX_P 12

/* Per-knot reference record (doubles). One record per knot k = 0..h of every phase
 * (record h is the terminal knot). Offsets into the record: */
#define CAFE_REF_XR      0   /* [36] state reference  (get_reference_at_t)                 */
#define CAFE_REF_UR      36  /* [24] control reference                                     */
#define CAFE_REF_YR      60  /* [12] output reference (WB: GRF reference)                  */
#define CAFE_REF_PF      72  /* [12] reference foot placements (world), app leg order      */
#define CAFE_REF_PCOM    84  /* [3]  reference CoM position                                */
#define CAFE_REF_VF      87  /* [12] reference foot velocities (world)                     */
#define CAFE_REF_CONTACT 99  /* [4]  reference contact flags at this knot's time (0/1)     */
#define CAFE_REF_QJ      103 /* [12] reference joint angles                                */
#define CAFE_REF_W       120 /* record width                                               */

/* Relaxed-barrier parameters (REB_Param_Struct, HSDDPSolver/header/ConstraintsBase.h:73-86) */
typedef struct {
  double delta, delta_min, eps;
} CafeRebParam;

/* Augmented-Lagrangian parameters (AL_Param_Struct, ConstraintsBase.h:58-70) */
typedef struct {
  double lambda, sigma, sigma_max;
} CafeAlParam;

typedef struct {
  int model;       /* CAFE_MODEL_*                                                        */
  int horizon;     /* h: number of controls; states are k = 0..h                          */
  int knot_offset; /* first record of this phase in CafeDeck.ref (h+1 records)            */
  int next_model;  /* model of the following phase, -1 for the last phase                 */
  double dt;       /* step handed to dynamics and costs (HKD: (double)(float)0.01,
                      HKDProblem.cpp:226,83 — dt_sim is a float)                          */
  float t_offset;  /* SinglePhase::set_time_offset (SinglePhase.h:199); informational      */
  int contact[4];      /* phase contact status, app leg order                             */
  int next_contact[4]; /* contact of next phase (last phase: reference at plan_dur+dt_mpc,
                          HKDProblem.cpp:284-287, MHPCProblem.cpp:534-537)                 */
  int has_reset;   /* a reset map is attached (HKD/WB: yes, SRB: no)                      */
  int n_td;        /* number of touchdown terminal constraints                            */
  int td_foot[4];  /* feet (app order) with contact 0 -> 1                                */
  /* QuadraticTrackingCost weights (diag), SinglePhaseInterface.cpp:6-18 */
  double q[CAFE_MAX_N], r[CAFE_MAX_M], qf[CAFE_MAX_N];
  /* foot cost weights per axis (WB: JSON qw_per_foot; HKD: fixed, HKDCost.h:56-71) */
  double w_footreg[3], w_swingpos[3], w_swingvel[3], w_tdvel[3];
  /* path-constraint ReB parameters at initialisation */
  CafeRebParam reb_grf, reb_torque, reb_joint, reb_minheight;
  /* terminal-constraint AL parameters at initialisation */
  CafeAlParam al_td;
  double mu;            /* friction coefficient (HKD 0.7, WB 0.6)                          */
  double ground_height; /* touchdown ground height (0)                                     */
  double h_min;         /* minimum body height (WB 0.20, SRB 0.18; MHPCConstraint.h:148,199) */
  double torque_limit;  /* WB joint torque bound (17; MHPCConstraint.cpp:77)               */
  double joint_lb[3], joint_ub[3]; /* WB joint limits per leg (MHPCConstraint.cpp:172-175) */
  /* WB path-constraint set. 0/0 = MHPCProblem (torque, joint, min height, GRF; MHPCProblem.cpp:436-481);
   * 1/1 = LocoProblem (torque and GRF only; Locomotion/LocoProblem.cpp:64-82) */
  int no_joint_limit, no_min_height;
  /* BarrelRoll::JointSpeedLimit (BarrelRoll/BarrelRollConstraints.cpp:151-193, .h:71-72): qJd - lb >= 0, -qJd + ub >= 0 on the twelve
   * joint rates, attached between the torque and the joint-limit barrier (BarrelRollTO.cpp:190-198). 0 = absent (MHPC, Loco). */
  int joint_speed_limit;
  CafeRebParam reb_jointvel;
  double jointvel_lb, jointvel_ub;
  /* 1 = the phase has no shooting states (SS_set empty): hybrid_rollout integrates it sequentially from the state the previous phase
   * hands over, X[k+1] = Xsim[k+1], zero defects (SinglePhase.cpp:187-221). The reference leaves the tail whole-body phase like that
   * in the MPC update that opens it, until it is longer than the shift (MHPCProblem.cpp:366-369); 0 = every knot is a shooting
   * state (update_SS_config(h + 1)). Supported when the previous phase (if any) has the same model. */
  int single_shooting;
} CafePhase;

typedef struct {
  int n_phases;
  int n_records; /* total knot records = sum(h_i + 1)                                      */
  CafePhase phase[CAFE_MAX_PHASES];
  const double* ref; /* [n_records][CAFE_REF_W], owned by whoever built the deck           */
  double BG_alpha;   /* WB Baumgarte gain (mhpc_config.info:8)                              */
  double hip_yaw;    /* yaw of the hip-pitch joint placement as loaded from the URDF (3.1415,
                        urdf/mini_cheetah_simple_correctedInertia.urdf:79); the kinematic partials that
                        replace the reference's CasADi code always use pi */
} CafeDeck;

/* Mirror of HSDDP_OPTION, field for field
 * (/root/reference/HSDDPSolver/common/HSDDP_CompoundTypes.h:13-36). */
typedef struct {
  double alpha, gamma, update_penalty, update_relax, update_regularization, update_ReB;
  int max_DDP_iter, max_AL_iter, max_DDP_iter_runtime, max_AL_iter_runtime;
  double cost_thresh, tconstr_thresh, pconstr_thresh, dynamics_feas_thresh;
  double merit_rho, merit_scale, merit_offset;
  int AL_active, ReB_active, smooth_active, MS, nsteps_per_node;
} CafeOptions;

/* Per-problem solve record. */
#define CAFE_STATUS_OK 0
#define CAFE_STATUS_REG_FAIL 1      /* regularisation exceeded 1e2 (MultiPhaseDDP.cpp:150-155) */
#define CAFE_STATUS_DIVERGED 2      /* the state kept after the line search came from a rollout
                                       with |Xsim| > 1e6 (SinglePhase.cpp:205-208)             */
typedef struct {
  int status;
  int iter;           /* iter_          (MultiPhaseDDP.cpp:286)                            */
  int ls_iter_total;  /* ls_iter_total_ (:355)                                             */
  int reg_iter_total; /* reg_iter_total_(:316)                                             */
  int outer_iter;     /* iter_ou at exit                                                   */
  int n_hist;         /* entries pushed to cost_buffer etc. (:258-261, :382-385)           */
  double cost, feas, max_tconstr, max_pconstr; /* values at exit (:427-430)                */
} CafeInfo;

static inline int cafe_model_n(int model) { return model == CAFE_MODEL_HKD ? 24 : model == CAFE_MODEL_WB ? 36 : 12; }
static inline int cafe_model_m(int model) { return model == CAFE_MODEL_HKD ? 24 : 12; }
static inline int cafe_model_p(int model) { return model == CAFE_MODEL_WB ? 12 : 0; }

#ifdef __cplusplus
}
#endif
#endif /* CAFE_DECK_H */
