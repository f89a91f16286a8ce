"""TEST INFRASTRUCTURE: a second, independent reader of the reference's input files for the MHPC problem. It restates, in numpy with the
reference's float / double types, what the reference's own set-up code does with quad_reference.csv, mhpc_config.info,
cost_weights_*.JSON and constraint_params_*.info, and yields the value every field of the phase deck must have:
    QuadReference::load_top_level_data / reorder_body_states / initialize / get_a_reference_ptr_at_t   Reference/QuadReference.cpp:6-31, 70-85, 134-371
    MHPCProblem::prepare_initialization / initialize_parameters / initialize_multiPhaseProblem         MHPC/MHPC-Trajopt/MHPCProblem.cpp:76-250
    create_problem_one_phase / update_resetmap / add_tconstr_one_phase                                :403-601
    WBReference / SRBReference::get_reference_at_t                                                     MHPCReference.cpp:10-76
    loadCostWeights                                                                                   MHPCCostUtil.h:10-143
    SinglePhase: the time a cost object is asked for at knot k is (float)(t_offset + k * dt)           SinglePhase.cpp:243, :254, :298, :311
It shares no code with csrc/host/ (the deck builders it checks) or with oracle/."""
import json
import math
import re

import numpy as np

f32 = np.float32
KEYS = ("body_state", "jnt_angle", "jnt_vel", "foot_placements", "foot_velocities", "grf", "torque", "contact", "status_dur")


def read_reference_csv(path, k0=0):
    """rows of the top-level reference; values parsed as float (std::stof) and widened; body_state re-ordered to [pos, eul, vWorld, eulrate]"""
    rows, cur, dt = [], {}, None
    with open(path) as f:
        lines = [l.rstrip("\n") for l in f]
    i = 0
    while i < len(lines):
        key = lines[i].strip()
        if key == "dt":
            dt = f32(lines[i + 1]); i += 2; continue
        hit = [k for k in KEYS if k in key]
        if hit:
            vals = lines[i + 1].split()
            cur[hit[0]] = np.array([int(v) for v in vals]) if hit[0] == "contact" else np.array([f32(v) for v in vals], dtype=f32).astype(np.float64)
            if hit[0] == "status_dur":
                b = cur["body_state"]
                cur["body_state"] = np.concatenate([b[3:6], b[0:3], b[9:12], b[6:9]])
                rows.append(cur); cur = {}
            i += 2; continue
        i += 1
    return dt, rows[k0:]


def read_info(path):
    """Boost-INFO subset of the settings files: `section { key value ... }`"""
    out, sec = {}, None
    for line in open(path):
        line = line.split(";")[0].strip()
        if not line or line == "{":
            continue
        if line == "}":
            sec = None; continue
        parts = line.split()
        if len(parts) == 1:
            sec = parts[0]
        elif sec is not None:
            out[sec + "." + parts[0]] = parts[1]
    return out


def approx_eq(a, b):          # HSDDP_Utils.h:46-56: float tol, float err of the (double) difference
    return f32(abs(float(a) - float(b))) <= f32(1e-6)


def approx_leq(a, b):
    return float(a) < float(b) or approx_eq(a, b)


class Reference:
    def __init__(self, path, k0, plan_horizon):
        self.dt, rows = read_reference_csv(path, k0)
        self.sz = int(round(float(f32(plan_horizon) / self.dt))) + 1            # QuadReference.cpp:17 (float / float)
        self.rows = rows[: self.sz + 1]

    def index(self, t):                                                          # :70-85, t is a float parameter
        t = f32(t)
        k = int(math.floor(float(t / self.dt)))
        if float(t - f32(k) * self.dt) > 0.5 * float(self.dt):
            k += 1
        return min(k, self.sz - 1)

    def at(self, t):
        return self.rows[self.index(t)]


def expected_mhpc_deck(csv, config_info, settings_root, k0=0):
    """list of per-phase dicts with the expected CafePhase fields and the expected reference records, from the input files alone"""
    cfg = read_info(config_info)
    plan_wb, plan_srb = float(cfg["config.plan_dur_wb"]), float(cfg["config.plan_dur_srb"])        # doubles (MHPCProblem.h:25-35)
    dt_wb, dt_srb = float(cfg["config.dt_wb"]), float(cfg["config.dt_srb"])
    dt_mpc = f32(cfg["config.dt_mpc"])                                                                  # float (:38)
    plan_all = f32(plan_wb + plan_srb)                                                                  # float member (:272)
    ref = Reference(csv, k0, plan_all)
    # ---- prepare_initialization (MHPCProblem.cpp:88-146)
    phases = []
    if plan_wb > 1e-5:
        start, t = f32(0), f32(0)
        prev = ref.at(t)["contact"]
        while approx_leq(t, plan_wb):
            cur = ref.at(t)["contact"]
            if np.any(cur != prev) or approx_eq(t, plan_wb):
                end = t
                phases.append({"model": 1, "horizon": int(round(float(end - start) / dt_wb)), "contact": tuple(int(c) for c in prev), "start": start, "end": end, "dt": dt_wb})
                prev, start = cur, end
            t = f32(float(t) + dt_wb)                                                                   # float += double
    n_wb = len(phases)
    if plan_srb > 1e-5:
        phases.append({"model": 2, "horizon": int(round(plan_srb / dt_srb)), "contact": (0, 0, 0, 0), "start": f32(plan_wb), "end": plan_all, "dt": dt_srb})
    # ---- initialize_parameters: constraint file and cost weights (:149-171, MHPCCostUtil.h)
    con = read_info(settings_root + "/" + cfg["config.constraintParamFile"])
    reb = lambda name: tuple(float(con["%s_ReB.%s" % (name, k)]) for k in ("delta", "delta_min", "eps"))
    w = json.load(open(settings_root + "/" + cfg["config.costFile"]))
    wt, st = w["WB_Tracking_Cost"], w["SRB_Tracking_Cost"]
    wbq = lambda p: list(wt[p + "_qB"]) + list(wt[p + "_qJ"]) * 4 + list(wt[p + "_vB"]) + list(wt[p + "_vJ"]) * 4
    for i, ph in enumerate(phases):
        h, dt = ph["horizon"], ph["dt"]
        if ph["model"] == 1:
            ph["t_offset"] = f32(ph["start"] - phases[0]["start"])                                       # :205
            nxt = phases[i + 1]["contact"] if i < n_wb - 1 else tuple(int(c) for c in ref.at(f32(plan_wb + float(dt_mpc)))["contact"])   # :533-537
            ph["next_contact"] = nxt
            ph["td_foot"] = [l for l in range(4) if ph["contact"][l] == 0 and nxt[l] == 1]               # :580-586
            ph["next_model"] = 1 if i < n_wb - 1 else (2 if len(phases) > n_wb else -1)                  # :543-547
            ph["q"], ph["qf"], ph["r"] = wbq("qw"), wbq("qfw"), [float(wt["rw"])] * 12
            ph["w_footreg"], ph["w_swingpos"], ph["w_swingvel"] = (list(map(float, w[k]["qw_per_foot"])) for k in ("WB_FootPlace_Reg", "Swing_Pos_Tracking", "Swing_Vel_Tracking"))
            ph["reb"] = {"reb_grf": reb("GRF"), "reb_torque": reb("Torque"), "reb_joint": reb("Joint"), "reb_minheight": reb("MinHeight")}
            ph["al_td"] = (float(con["TD_AL.lambda"]), float(con["TD_AL.sigma"]), float(con["TD_AL.sigma_max"]))
        else:
            ph["t_offset"] = ph["start"]                                                                # :241
            ph["next_contact"], ph["td_foot"], ph["next_model"] = (0, 0, 0, 0), [], -1
            ph["q"] = list(st["qw_qB"]) + list(st["qw_vB"]); ph["qf"] = list(st["qfw_qB"]) + list(st["qfw_vB"]); ph["r"] = [float(st["rw"])] * 12
            ph["reb"] = {"reb_minheight": reb("MinHeight")}
        # ---- the record a cost / reference object sees at knot k: time (float)(t_offset + k dt), SinglePhase.cpp:243-311
        recs = []
        for k in range(h + 1):
            s = ref.at(f32(float(ph["t_offset"]) + k * dt))
            r = {"pf": s["foot_placements"], "vf": s["foot_velocities"], "pcom": s["body_state"][:3], "contact": s["contact"], "qJ": s["jnt_angle"]}
            if ph["model"] == 1:                                                                         # WBReference (MHPCReference.cpp:10-41)
                r["xr"] = np.concatenate([s["body_state"][:6], s["jnt_angle"], s["body_state"][6:], s["jnt_vel"]]); r["ur"] = s["torque"]; r["yr"] = s["grf"]
            else:                                                                                        # SRBReference (:49-76)
                r["xr"] = s["body_state"]; r["ur"] = s["grf"]; r["yr"] = np.zeros(0)
            recs.append(r)
        ph["records"] = recs
    return phases
