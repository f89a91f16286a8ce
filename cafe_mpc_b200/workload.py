"""Synthetic perturbed-initial-state batches: x0[b][j] = x0_nom[j] + s_j (2u - 1),
u = SplitMix64(seed = 0xCAFE, counter = b*64 + j) mapped to [0,1) with 53 bits."""
import numpy as np

MASK = (1 << 64) - 1


def splitmix64(counter, seed=0xCAFE):
    z = (seed + (counter + 1) * 0x9E3779B97F4A7C15) & MASK
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK
    return z ^ (z >> 31)


def uniform(b, j):
    return (splitmix64(b * 64 + j) >> 11) * (1.0 / (1 << 53))


HKD_NOMINAL_BODY = np.array([0, 0, 0, 0, 0, 0.2486, 0, 0, 0, 0, 0, 0], dtype=np.float64)  # HKDMPC.cpp:46
HKD_NOMINAL_QJ = np.array([0, -0.8, 1.6] * 4, dtype=np.float64)                            # HKDMPC.cpp:47
# scales: eul 0.05 rad, pos x,y 0.02 m, z 0.01 m, omega 0.1 rad/s, vel 0.1 m/s, joint angles 0.05 rad
HKD_BODY_SCALE = np.array([0.05] * 3 + [0.02, 0.02, 0.01] + [0.1] * 3 + [0.1] * 3)
HKD_QJ_SCALE = np.full(12, 0.05)


def hkd_batch(problem, B, perturb=True):
    """Perturbed HKD initial states: joint angles are perturbed BEFORE compute_hkd_state."""
    x0 = np.zeros((B, 24))
    for b in range(B):
        body = HKD_NOMINAL_BODY.copy()
        qJ = HKD_NOMINAL_QJ.copy()
        if perturb and b > 0:  # problem 0 is the nominal problem of HKDMPCSolver::initialize
            for j in range(12):
                body[j] += HKD_BODY_SCALE[j] * (2 * uniform(b, j) - 1)
            for j in range(12):
                qJ[j] += HKD_QJ_SCALE[j] * (2 * uniform(b, 12 + j) - 1)
        x0[b] = problem.initial_state(body, qJ)
    return x0


# MHPC whole-body nominal state (Loco_TO.cpp:49-55): pos (0,0,0.2183), qJ (0,-1.0,2.0) x4, rest 0
MHPC_NOMINAL = np.zeros(36)
MHPC_NOMINAL[2] = 0.2183
MHPC_NOMINAL[6:18] = [0, -1.0, 2.0] * 4
# scales: pos x,y 0.02, z 0.01; euler 0.05; joints 0.05; base lin vel 0.1; euler rate 0.1; joint vel 0.2
MHPC_SCALE = np.array([0.02, 0.02, 0.01] + [0.05] * 3 + [0.05] * 12 + [0.1] * 3 + [0.1] * 3 + [0.2] * 12)


def mhpc_batch(B, perturb=True):
    x0 = np.tile(MHPC_NOMINAL, (B, 1))
    if perturb:
        for b in range(1, B):  # problem 0 is the nominal problem
            for j in range(36):
                x0[b, j] += MHPC_SCALE[j] * (2 * uniform(b, j) - 1)
    return x0


# ---- MHPC running barrel roll (BASELINE config 4; Reference/Data/running_br, cost_weights_barrel.JSON,
# constraint_params_barrel.info). Start offsets: k0 = 0 (stance -> diagonal pair -> flight) and k0 = 205 (mid-roll flight,
# 4-foot landing impact with touchdown constraints, SURVEY.md §8 phase table).
import os as _os

_DATA = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "data")
BARREL_CSV = _os.path.join(_DATA, "Reference/Data/running_br/quad_reference.csv")
BARREL_CONFIG = _os.path.join(_DATA, "MHPC/settings/mhpc_config_barrel.info")
BARREL_K0_IMPACT = 205


def reference_state(problem):
    """Whole-body reference state (36) at the first knot of the deck = the state the tracked motion has at the start offset."""
    deck = problem.deck.contents
    return np.array([deck.ref[i] for i in range(36)], dtype=np.float64)


def barrel_batch(problem, B, perturb=True):
    """x0 = reference state at the start offset + the same SplitMix64 perturbation table as mhpc_batch."""
    x0 = np.tile(reference_state(problem), (B, 1))
    if perturb:
        for b in range(1, B):
            for j in range(36):
                x0[b, j] += MHPC_SCALE[j] * (2 * uniform(b, j) - 1)
    return x0


# ---- per-problem references on a shared phase schedule: every problem is commanded a different forward
# speed. Record layout: include/cafe_deck.h (CAFE_REF_XR 0, PF 72, PCOM 84, W 120).
def speed_command_references(problem, B, max_dv=0.2):
    """Returns [B, n_records, 120]: problem b tracks the deck's reference shifted by a constant extra forward speed
    dv_b = max_dv (2u - 1), u = SplitMix64 counter b*64 + 40 (problem 0 keeps the deck's reference): x position += dv t,
    forward velocity += dv, reference foot placements and CoM x += dv t. Contact flags (the phase schedule) are untouched."""
    base = problem.reference_records()
    d = problem.deck.contents
    refs = np.tile(base, (B, 1, 1))
    t_of, vx_of = np.zeros(d.n_records), np.zeros(d.n_records, dtype=int)
    for i in range(d.n_phases):
        ph = d.phase[i]
        for k in range(ph.horizon + 1):
            t_of[ph.knot_offset + k] = ph.t_offset + k * ph.dt
            vx_of[ph.knot_offset + k] = {0: 9, 1: 18, 2: 6}[ph.model]   # index of the forward velocity in the state: HKD, WB, SRB
    px_of = np.array([{0: 3, 1: 0, 2: 0}[d.phase[i].model] for i in range(d.n_phases) for _ in range(d.phase[i].horizon + 1)])
    for b in range(1, B):
        dv = max_dv * (2 * uniform(b, 40) - 1)
        for r in range(d.n_records):
            refs[b, r, px_of[r]] += dv * t_of[r]
            refs[b, r, vx_of[r]] += dv
            refs[b, r, 72:84:3] += dv * t_of[r]
            refs[b, r, 84] += dv * t_of[r]
    return refs


# ---- LocoProblem (MHPC/MHPC-Trajopt/Locomotion): whole-body-only locomotion trajectory optimisation, 1 s flypace plan.
# Loco_TO.cpp:49-55 starts it from MHPC_NOMINAL; mhpc_batch is its perturbed batch.
LOCO_DDP_SETTING = _os.path.join(_DATA, "MHPC/MHPC-Trajopt/Locomotion/settings/loco_ddp_setting.info")
BARREL_TO_DDP_SETTING = _os.path.join(_DATA, "MHPC/MHPC-Trajopt/BarrelRoll/setting/br_ddp_setting.info")   # in-place barrel roll (BarrelRollTO.cpp)
