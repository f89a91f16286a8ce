"""ctypes mirrors of the POD structs in include/cafe_deck.h (keep in sync with the header)."""
import ctypes as C

CAFE_MAX_PHASES = 16
CAFE_MAX_N, CAFE_MAX_M, CAFE_MAX_P = 36, 24, 12
CAFE_REF_W = 120
CAFE_TRACE_W = 12
CAFE_NKERNELS = 11
MODEL_HKD, MODEL_WB, MODEL_SRB = 0, 1, 2
MODEL_DIMS = {MODEL_HKD: (24, 24, 0), MODEL_WB: (36, 12, 12), MODEL_SRB: (12, 12, 0)}


class RebParam(C.Structure):
    _fields_ = [("delta", C.c_double), ("delta_min", C.c_double), ("eps", C.c_double)]


class AlParam(C.Structure):
    _fields_ = [("lambda_", C.c_double), ("sigma", C.c_double), ("sigma_max", C.c_double)]


class Phase(C.Structure):
    _fields_ = [
        ("model", C.c_int), ("horizon", C.c_int), ("knot_offset", C.c_int), ("next_model", C.c_int),
        ("dt", C.c_double), ("t_offset", C.c_float),
        ("contact", C.c_int * 4), ("next_contact", C.c_int * 4),
        ("has_reset", C.c_int), ("n_td", C.c_int), ("td_foot", C.c_int * 4),
        ("q", C.c_double * CAFE_MAX_N), ("r", C.c_double * CAFE_MAX_M), ("qf", C.c_double * CAFE_MAX_N),
        ("w_footreg", C.c_double * 3), ("w_swingpos", C.c_double * 3), ("w_swingvel", C.c_double * 3),
        ("w_tdvel", C.c_double * 3),
        ("reb_grf", RebParam), ("reb_torque", RebParam), ("reb_joint", RebParam), ("reb_minheight", RebParam),
        ("al_td", AlParam), ("mu", C.c_double), ("ground_height", C.c_double),
        ("h_min", C.c_double), ("torque_limit", C.c_double), ("joint_lb", C.c_double * 3), ("joint_ub", C.c_double * 3),
        ("no_joint_limit", C.c_int), ("no_min_height", C.c_int),
        ("joint_speed_limit", C.c_int), ("reb_jointvel", RebParam), ("jointvel_lb", C.c_double), ("jointvel_ub", C.c_double), ("single_shooting", C.c_int),
    ]


class Deck(C.Structure):
    _fields_ = [
        ("n_phases", C.c_int), ("n_records", C.c_int), ("phase", Phase * CAFE_MAX_PHASES),
        ("ref", C.POINTER(C.c_double)), ("BG_alpha", C.c_double), ("hip_yaw", C.c_double),
    ]


class Options(C.Structure):
    _fields_ = [
        ("alpha", C.c_double), ("gamma", C.c_double), ("update_penalty", C.c_double), ("update_relax", C.c_double),
        ("update_regularization", C.c_double), ("update_ReB", C.c_double),
        ("max_DDP_iter", C.c_int), ("max_AL_iter", C.c_int), ("max_DDP_iter_runtime", C.c_int),
        ("max_AL_iter_runtime", C.c_int),
        ("cost_thresh", C.c_double), ("tconstr_thresh", C.c_double), ("pconstr_thresh", C.c_double),
        ("dynamics_feas_thresh", C.c_double),
        ("merit_rho", C.c_double), ("merit_scale", C.c_double), ("merit_offset", C.c_double),
        ("AL_active", C.c_int), ("ReB_active", C.c_int), ("smooth_active", C.c_int), ("MS", C.c_int),
        ("nsteps_per_node", C.c_int),
    ]


class Info(C.Structure):
    _fields_ = [
        ("status", C.c_int), ("iter", C.c_int), ("ls_iter_total", C.c_int), ("reg_iter_total", C.c_int),
        ("outer_iter", C.c_int), ("n_hist", C.c_int),
        ("cost", C.c_double), ("feas", C.c_double), ("max_tconstr", C.c_double), ("max_pconstr", C.c_double),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}
CAFE_STATUS_OK, CAFE_STATUS_REG_FAIL, CAFE_STATUS_DIVERGED = 0, 1, 2
